#!/bin/bash
# per-GPU large-batch rate at the shard sizes of N = 1, 2, 4, 8 (100M-point map, one GPU): usage tools/shard_density.sh [env...]
for q in 100000000 50000000 25000000 12500000; do
  python bench.py --queries $q --no-cpu-baseline --extras none --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('queries', d['config']['queries'], 'value %.3fG' % (d['value']/1e9), 'ms %.2f' % d['ms_per_step'], 'kernel_ms %.2f' % r['kernel_ms_mean'], 'frac %.3f' % r['frac'], 'e2e %.3fG' % (d['e2e']['value']/1e9))
"
done
