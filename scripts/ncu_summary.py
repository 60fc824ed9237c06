"""Summarise .ncu-rep files (read here on the CPU box).
    python scripts/ncu_summary.py report.ncu-rep profiles/name.txt      # one report -> text summary for profiles/
    python scripts/ncu_summary.py --brief a.ncu-rep b.ncu-rep ...        # key metrics and stall reasons side by side"""
import csv, io, subprocess, sys


def brief(reps):
    want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'sm__cycles_active.avg',
            'sm__cycles_active.max', 'sm__cycles_elapsed.max', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
            'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
    cols = []
    for rep in reps:
        rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
        cols.append(dict(zip(rows[0], rows[2])))
    stalls = sorted({h for d in cols for h in d if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                     and "not_issued" not in h})
    for w in want + stalls:
        vals = [d.get(w, "") for d in cols]
        try:
            if w in stalls and all(float(v or 0) < 0.05 for v in vals):
                continue
        except ValueError:
            pass
        print("%-60s" % w.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", ""), *["%14s" % v[:14] for v in vals])


if sys.argv[1] == "--brief":
    brief(sys.argv[2:])
    sys.exit(0)
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_active.max",
        "sm__cycles_elapsed.max"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
lines = [f"# ncu summary of {rep.split('/')[-1]} (ncu --set full --clock-control none --import-source on)"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        lines.append(f"{w} [{units[i]}] = {vals[i]}")
lines.append("# warp stall reasons (cycles stalled per issued instruction)")
pairs = sorted(((float(vals[hdr.index(h)] or 0), h) for h in stall), reverse=True)
for v, h in pairs[:10]:
    lines.append(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} = {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    h2, data = rows[1], rows[2:]
    ia, isamp, isrc = h2.index("Instructions Executed"), h2.index("# Samples"), h2.index("Source")
    tot = sum(int(r[isamp]) for r in data) or 1
    lines.append(f"# hottest SASS instructions by stall samples (total samples {tot})")
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:12]:
        lines.append(f"{100 * int(r[isamp]) / tot:5.1f}%  exec={r[ia]:>10s}  {r[isrc].strip()[:90]}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
