nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/lat_probe tools/lat_probe.cu && /tmp/lat_probe 1572864 | grep -v " 0.0 cycles"
lb() { python bench.py --no-cpu-baseline --extras none --steps 3 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('queries', d['config']['queries'], 'value %.3fG' % (d['value']/1e9), 'kernel_ms %.2f' % r['kernel_ms_mean'], 'frac %.3f' % r['frac'], 'e2e %.3fG' % (d['e2e']['value']/1e9), 'e2e_ms %.1f' % d['e2e']['ms_per_step'])
"; }
echo "== default (refill 8, ramp)"; lb; lb --queries 12500000
echo "== no ramp"; IKD_KNN_NO_RAMP=1 lb
for c in 4194304 8388608; do echo "== chunk $c ramp"; IKD_KNN_CHUNK=$c lb; done
for v in refill4 refill12 refill16; do echo "== $v"; IKD_LIB_PATH=$PWD/ikd-tree_b200/variants/libikd_b200_$v.so lb; IKD_LIB_PATH=$PWD/ikd-tree_b200/variants/libikd_b200_$v.so lb --queries 12500000; done
