nvidia-smi -L | wc -l
echo "== 2-rank replica test"; timeout 600 python -m pytest tests/test_replica_gpu.py -x -q 2>&1 | tail -2
echo "== N=8 default bench"
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/r02_bench_default_n8.json 2> gpurun_out/r02_bench_default_n8.err
tail -c 1500 gpurun_out/r02_bench_default_n8.json; tail -5 gpurun_out/r02_bench_default_n8.err
echo "== N=8 reference arm (rank 0 only; bounded)"; echo skipped
