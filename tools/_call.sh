echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -x -q 2>&1 | tail -3
echo "== trace"; IKD_LIB_PATH=$PWD/ikd-tree_b200/variants/libikd_b200_trace.so python bench.py --workload scanloop --no-cpu-baseline --steps 4 --warmup 3 2>&1 >/dev/null | grep "refit trace" | head -6
echo "== scanloop"; tools/sweep_env.sh IKD_DUMMY 0 0
echo "== phases"; IKD_PHASES=1 python bench.py --workload scanloop --no-cpu-baseline 2>&1 >/dev/null | grep "ikd phases" | tail -22
