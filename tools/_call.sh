nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/lat_probe tools/lat_probe.cu && /tmp/lat_probe 1572864
mkdir -p /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:range_collect" -c 2 -f -o /tmp/ncu/rc python tools/gpu_range_profile.py > /dev/null 2> /tmp/ncu/rc.stderr
ncu -i /tmp/ncu/rc.ncu-rep --page details > gpurun_out/r02b_range_collect_details.txt 2>&1
ncu -i /tmp/ncu/rc.ncu-rep --page source --csv > /tmp/ncu/rc_source.csv 2>/dev/null; python tools/ncu_sass_hotspots.py /tmp/ncu/rc_source.csv > gpurun_out/r02b_range_collect_hotspots.txt 2>&1 || true
ls -la /tmp/ncu gpurun_out | tail -8
