#!/bin/bash
# usage: tools/sweep_env.sh VAR v1 v2 ... : scan-loop bench (no CPU baseline) once per value, prints the headline numbers
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --workload scanloop --no-cpu-baseline --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$var=$v', 'value %.2fM' % (d['value']/1e6), 'ms/step %.4f' % d['ms_per_step'], 'p50 %.4f' % d['scan_p50_ms'], 'knn %.4f' % d['knn_ms_per_step'], 'add %.4f' % d['add_points_ms_per_step'], 'add_p50 %.4f' % d['add_points_p50_ms'], 'e2e %.2fM' % (d['e2e']['value']/1e6), 'e2e_p50 %.4f' % d['e2e']['scan_p50_ms'], 'async', d['tree_stats']['rebuilds_async'], 'launches', d['gpu_launches'], 'parity', d['parity']['ok'])
"
done
