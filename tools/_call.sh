echo "== gpu tests"; (time timeout 900 python -m pytest tests -m gpu -x -q) 2>&1 | tail -6
echo "== default bench"; (time python bench.py) > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r02_bench_default_n1.err; tail -3 gpurun_out/r02_bench_default_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_default_n1.json').read().strip().splitlines()[-1])
print('value %.3fG'%(d['value']/1e9),'frac %.3f'%d['roofline']['frac'],'e2e %.3fG'%(d['e2e']['value']/1e9),'cpu %.3fM'%(d['cpu_baseline']['value']/1e6))
s=d['scan_loop']; print('scan value %.1fM'%(s['value']/1e6),'ms',round(s['ms_per_step'],4),'add',round(s['add_points_ms_per_step'],4),'e2e %.1fM'%(s['e2e']['value']/1e6))
c=d['c3_range_search']
for k in ('box','radius'): print(k,'device_s %.5f'%c[k]['device_s'],'frac %.3f'%c[k]['roofline']['frac'])
c5=d['c5_streaming']; print('c5 upd',c5['update_ms'],'add',c5['add_points_ms'],'longest',c5['rebuilds']['longest_ms'])
PY
