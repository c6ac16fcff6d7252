#!/bin/bash
# ncu launch list (per-launch gpu__time_duration, cold cache, serialised) of this library's kernels in a short default
# bench run. The kernel filter names OUR kernels: the bench's synthetic LiDAR world is generated with torch ops (thousands
# of launches) that must not be profiled. usage: tools/ncu_launch_list.sh OUT.csv [steps] [warmup]
out=${1:-gpurun_out/r01_scanloop_launches.csv}; steps=${2:-4}; warm=${3:-3}
names=$(grep -ho "[a-z_0-9]*_kernel\b" ikd-tree_b200/csrc/*.cu | sort -u | tr '\n' '|' | sed 's/|$//')
exec ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^($names)$" --csv --log-file "$out" \
    python bench.py --steps $steps --warmup $warm --no-cpu-baseline
