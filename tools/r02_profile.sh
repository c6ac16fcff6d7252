#!/bin/bash
# Round-2 profiling pass (one GPU): launch list of the default bench command, ncu --set full of the dominant kernels.
# usage (on the GPU box): tools/r02_profile.sh [tag]      outputs under gpurun_out/
tag=${1:-r02}
names=$(grep -ho "[a-z_0-9]*_kernel\b" ikd-tree_b200/csrc/*.cu | sort -u | tr '\n' '|' | sed 's/|$//')
# 1. launch list: the default bench command (extras: nested scan loop + c3; c5 = 1000 scans is left out of the ncu pass)
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^($names)$" --csv --log-file gpurun_out/${tag}_default_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extras scan_loop,c3 > gpurun_out/${tag}_default_launches.stdout 2> gpurun_out/${tag}_default_launches.stderr
# 2. the dominant kernel of the default bench: one 100M-query launch of knn_reg_persist_kernel<5> on the 100M-point map
ncu --set full --clock-control none --import-source on -k regex:knn_reg_persist -s 3 -c 1 -f -o gpurun_out/${tag}_knn_large_100M \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extras none > /dev/null 2> gpurun_out/${tag}_knn_large_ncu.stderr
# 3. every major kernel once (1M-point map)
ncu --set full --clock-control none --import-source on -k "regex:^($names)$" -f -o gpurun_out/${tag}_all_kernels \
    python tools/gpu_all_kernels.py > /dev/null 2> gpurun_out/${tag}_all_kernels.stderr
# 4. range search at c3 size (10M points, 100k queries): first call of each kind
ncu --set full --clock-control none --import-source on -k "regex:range_" -c 12 -f -o gpurun_out/${tag}_range_c3 \
    python tools/gpu_range_profile.py > /dev/null 2> gpurun_out/${tag}_range_c3.stderr
ls -la gpurun_out/
