"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import collections, csv, re, sys
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))[skip:]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in rows:
    name = row['Kernel Name']
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    short = re.sub(r'^void ', '', name)
    short = re.sub(r'\(anonymous namespace\)::', '', short)
    short = re.sub(r'cub::CUB_\w+::', 'cub::', short)
    short = re.sub(r'\(.*', '', short)
    short = re.sub(r'<.*', '', short) if short.startswith('cub::') else short
    agg[short][0] += 1
    agg[short][1] += v
tot = sum(v[1] for v in agg.values())
print(f"launches {len(rows)}  total {tot:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:5d}  mean {v[1] / v[0]:8.2f} us  {k}")
