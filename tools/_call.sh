nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/lat_probe tools/lat_probe.cu && /tmp/lat_probe 1572864
echo "== parity (knn)"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "knn" 2>&1 | tail -2
lb() { python bench.py --no-cpu-baseline --extras none --steps 3 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('queries', d['config']['queries'], 'value %.3fG' % (d['value']/1e9), 'kernel_ms %.2f' % r['kernel_ms_mean'], 'frac %.3f' % r['frac'], 'e2e %.3fG' % (d['e2e']['value']/1e9))
"; }
echo "== pops3 (default)"; lb; lb --queries 12500000
for v in pops1 pops2 pops4; do echo "== $v"; IKD_LIB_PATH=$PWD/ikd-tree_b200/variants/libikd_b200_$v.so lb; IKD_LIB_PATH=$PWD/ikd-tree_b200/variants/libikd_b200_$v.so lb --queries 12500000; done
