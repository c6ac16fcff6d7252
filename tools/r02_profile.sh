#!/bin/bash
# Round-2 profiling pass (one GPU). Reports are reduced to text / CSV on the box (gpurun_out/ is capped at 64 MiB).
# usage (on the GPU box): tools/r02_profile.sh TAG [launches] [knn100m] [all] [range]
tag=${1:-r02}; shift
what="$*"; [ -z "$what" ] && what="launches knn100m all range"
names=$(grep -ho "[a-z_0-9]*_kernel\b" ikd-tree_b200/csrc/*.cu | sort -u | tr '\n' '|' | sed 's/|$//')
mkdir -p gpurun_out /tmp/ncu
for w in $what; do
  t0=$(date +%s)
  case $w in
  launches)
    # launch list of the default bench command (extras: nested scan loop + c3; c5 = 1000 scans is left out of the ncu pass)
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^($names)$" --csv --log-file gpurun_out/${tag}_default_launches.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extras scan_loop,c3 > gpurun_out/${tag}_default_launches.stdout 2> /tmp/ncu/launches.stderr
    python tools/launch_summary.py gpurun_out/${tag}_default_launches.csv > gpurun_out/${tag}_default_launch_summary.txt 2>&1
    gzip -f gpurun_out/${tag}_default_launches.csv ;;
  knn100m)
    # the dominant kernel of the default bench: one 100M-query launch of knn_reg_persist_kernel<5> on the 100M-point map
    timeout 600 ncu --set full --clock-control none -k regex:knn_reg_persist -s 3 -c 1 -f -o /tmp/ncu/knn100m \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extras none > /dev/null 2> /tmp/ncu/knn100m.stderr
    python tools/ncu_summary.py /tmp/ncu/knn100m.ncu-rep gpurun_out/${tag}_knn_large_ncu_summary.json "100M-point map, 100M queries, one launch, k=5" > gpurun_out/${tag}_knn_large_100M_ncu_summary.txt 2>&1
    ncu -i /tmp/ncu/knn100m.ncu-rep --page details > gpurun_out/${tag}_knn_large_100M_ncu_details.txt 2>&1 ;;
  all)
    timeout 600 ncu --set full --clock-control none -k "regex:^($names)$" -f -o /tmp/ncu/all \
        python tools/gpu_all_kernels.py > /dev/null 2> /tmp/ncu/all.stderr
    python tools/ncu_kernel_table.py /tmp/ncu/all.ncu-rep > gpurun_out/${tag}_all_kernels_ncu.txt 2>&1 ;;
  range)
    timeout 600 ncu --set full --clock-control none -k "regex:range_" -c 12 -f -o /tmp/ncu/range \
        python tools/gpu_range_profile.py > /dev/null 2> /tmp/ncu/range.stderr
    python tools/ncu_kernel_table.py /tmp/ncu/range.ncu-rep > gpurun_out/${tag}_range_c3_ncu.txt 2>&1
    python tools/ncu_summary.py /tmp/ncu/range.ncu-rep > gpurun_out/${tag}_range_c3_ncu_summary.txt 2>&1 ;;
  esac
  echo "[profile] $w: $(( $(date +%s) - t0 )) s"
done
tail -3 /tmp/ncu/*.stderr | tail -30
du -sh gpurun_out
