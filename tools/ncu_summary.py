"""Key metrics of every kernel launch in an .ncu-rep (ncu --set full), as text; optional JSON with DRAM bytes/launch."""
import csv, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum']
sc = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
tot = []
for row in r[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:72s} {row[i][:150]} {units[i]}")
    print()
    ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    tot.append(float(row[ir]) * sc[units[ir]] + float(row[iw]) * sc[units[iw]])
if len(sys.argv) > 2:
    kname = r[2][hdr.index('Kernel Name')] if len(r) > 2 else ""
    json.dump({"kernel": kname, "config": sys.argv[3] if len(sys.argv) > 3 else "", "dram_bytes_per_launch": sum(tot) / len(tot),
               "launches": len(tot), "source": rep}, open(sys.argv[2], "w"))
