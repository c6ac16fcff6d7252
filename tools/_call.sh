lb() { python bench.py --no-cpu-baseline --extras none --steps 3 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('queries', d['config']['queries'], 'value %.3fG' % (d['value']/1e9), 'kernel_ms %.2f' % r['kernel_ms_mean'], 'frac %.3f' % r['frac'], 'e2e %.3fG' % (d['e2e']['value']/1e9))
"; }
echo "== off"; lb; lb --queries 12500000
for mb in 32 64 96; do echo "== persist $mb MB"; IKD_L2_PERSIST_MB=$mb lb; IKD_L2_PERSIST_MB=$mb lb --queries 12500000; done
