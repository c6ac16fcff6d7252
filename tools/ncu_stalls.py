"""Per-kernel warp-stall breakdown (cycles per issued instruction, by reason) from an ncu --set full report."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h = r[0]
stall = [(i, re.sub(r"smsp__average_warps?_issue_stalled_|_per_issue_active.ratio|smsp__average_warp_latency_issue_stalled_", "", m))
         for i, m in enumerate(h) if "issue_stalled" in m and m.endswith("per_issue_active.ratio")]
extra = [m for m in ("gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                     "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum") if m in h]
agg = collections.OrderedDict()
for row in r[2:]:
    name = re.sub(r"\(.*", "", re.sub(r"^void ", "", row[h.index("Kernel Name")])).replace("unnamed>::", "")
    a = agg.setdefault(name, [0, collections.Counter(), collections.Counter()])
    a[0] += 1
    for i, s in stall:
        try: a[1][s] += float(row[i].replace(",", ""))
        except ValueError: pass
    for m in extra:
        try: a[2][m] += float(row[h.index(m)].replace(",", ""))
        except ValueError: pass
for name, (n, st, ex) in agg.items():
    print(f"{name}  (n={n})  " + "  ".join(f"{m.split('.')[0].split('__')[-1]}={ex[m]/n:.1f}" for m in extra))
    tot = sum(st.values())
    for s, v in st.most_common(7):
        print(f"    {s:32s} {v/n:8.2f} cycles per issued instruction ({100*v/tot:4.1f}%)")
