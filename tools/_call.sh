mkdir -p /tmp/ncu
timeout 600 ncu --set full --clock-control none -k "regex:^(refit_kernel|mark_kernel|vox_decide_linked_kernel|descend_link_kernel|collect_viol_kernel|flatten_kernel|insert_group_kernel)$" -s 21 -c 21 -f -o /tmp/ncu/upd \
    python bench.py --workload scanloop --no-cpu-baseline --steps 4 --warmup 3 > /dev/null 2> /tmp/ncu/upd.stderr
python tools/ncu_stalls.py /tmp/ncu/upd.ncu-rep > gpurun_out/r02_update_kernels_stalls.txt 2>&1
python tools/ncu_kernel_table.py /tmp/ncu/upd.ncu-rep > gpurun_out/r02_update_kernels_ncu.txt 2>&1
cat gpurun_out/r02_update_kernels_stalls.txt
