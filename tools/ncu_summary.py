"""Condense an .ncu-rep (ncu --set full) into the handful of numbers profiles/README.md cites.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_fp64.sum",
        "sm__inst_executed_pipe_xu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in KEEP:
        if k in d:
            print(f"{k} = {d[k]} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[h]
iS = hdr.index("# Samples")
stall = {k: hdr.index(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
agg = {k: 0 for k in stall}
lines = []
for r in rows[h + 1:]:
    if not r or r[0] == "" or len(r) <= iS:
        continue
    try:
        n, ln = int(r[iS]), int(r[0])
    except ValueError:
        continue
    for k, i in stall.items():
        try:
            agg[k] += int(r[i])
        except ValueError:
            pass
    if n:
        lines.append((n, ln, r[1].strip()[:110]))
tot = sum(n for n, _, _ in lines) or 1
print("\nwarp-stall samples by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
print("top source lines by samples:")
for n, ln, txt in sorted(lines, reverse=True)[:14]:
    print(f"  {100 * n / tot:5.1f}%  L{ln}: {txt}")
