"""Per-kernel table (mean over launches) of key ncu metrics from an --set full report."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, units = r[0], r[1]
cols = [("gpu__time_duration.sum", "time_us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("l1tex__t_sector_hit_rate.pct", "L1hit%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst")]
sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "second": 1e6}
agg = collections.OrderedDict()
for row in r[2:]:
    name = row[h.index("Kernel Name")]
    name = re.sub(r"\(.*", "", re.sub(r"^void ", "", name)).replace("unnamed>::", "")
    vals = []
    for m, _ in cols:
        i = h.index(m)
        v = float(row[i].replace(",", "")) * sc.get(units[i], 1)
        vals.append(v)
    a = agg.setdefault(name, [0, [0.0] * len(cols)])
    a[0] += 1
    a[1] = [x + y for x, y in zip(a[1], vals)]
print(f"{'kernel':58s} {'n':>4s} " + " ".join(f"{c[1]:>9s}" for c in cols))
for name, (n, s) in agg.items():
    m = [x / n for x in s]
    fmt = []
    for (metric, label), v in zip(cols, m):
        fmt.append(f"{v/1e6:8.2f}M" if label.startswith("dram_") else f"{v:9.1f}")
    print(f"{name[:58]:58s} {n:4d} " + " ".join(fmt))
