echo "== parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
echo "== largebatch LDG.256 (default)"; python bench.py --no-cpu-baseline --extras c3 --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value %.3fG' % (d['value']/1e9), 'kernel_ms %.2f' % r['kernel_ms_mean'], 'frac %.3f' % r['frac'], 'e2e %.3fG' % (d['e2e']['value']/1e9))
c=d['c3_range_search']
for k in ('box','radius'): print(k, 'device_s %.5f' % c[k]['device_s'], 'frac %.3f' % c[k]['roofline']['frac'], 'results', c[k]['results'])
"
echo "== largebatch LDG.128"; IKD_LIB_PATH=$PWD/ikd-tree_b200/variants/libikd_b200_ldg128.so python bench.py --no-cpu-baseline --extras c3 --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value %.3fG' % (d['value']/1e9), 'kernel_ms %.2f' % r['kernel_ms_mean'], 'frac %.3f' % r['frac'], 'e2e %.3fG' % (d['e2e']['value']/1e9))
c=d['c3_range_search']
for k in ('box','radius'): print(k, 'device_s %.5f' % c[k]['device_s'], 'frac %.3f' % c[k]['roofline']['frac'], 'results', c[k]['results'])
"
echo "== scanloop LDG.256"; tools/sweep_env.sh IKD_DUMMY 0 0
echo "== scanloop LDG.128"; IKD_LIB_PATH=$PWD/ikd-tree_b200/variants/libikd_b200_ldg128.so tools/sweep_env.sh IKD_DUMMY 0 0
tools/r02_profile.sh r02 range all
