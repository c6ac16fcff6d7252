echo "== parity knn"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -x -q -k "knn or degenerate or golden or sequences" 2>&1 | tail -2
k32() { python bench.py --no-cpu-baseline --extras k32 --steps 2 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['k32']
print('k32 value %.4fG' % (r['value']/1e9), 'frac %.3f' % r['roofline']['frac'], 'visits %.1f' % r['roofline']['visits_per_query'])
"; }
echo "== heap batch 8 (default)"; k32
for v in 4 16 32; do echo "== heap batch $v"; IKD_LIB_PATH=$PWD/ikd-tree_b200/variants/libikd_b200_hb$v.so k32; done
