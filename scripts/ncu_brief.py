"""python scripts/ncu_brief.py file.ncu-rep ... : key metrics + stall reasons of each report, side by side"""
import csv, io, subprocess, sys
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
cols = []
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, vals = rows[0], rows[2]
    d = dict(zip(hdr, vals))
    cols.append(d)
stalls = sorted({h for d in cols for h in d if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h})
for w in want + stalls:
    name = w.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", "")
    vals = [d.get(w, "") for d in cols]
    try:
        if all(float(v or 0) < 0.05 for v in vals) and w in stalls: continue
    except ValueError: pass
    print("%-62s" % name, *["%14s" % v[:14] for v in vals])
