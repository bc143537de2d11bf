#!/usr/bin/env python
"""Summarise an ncu report here (no GPU): python profiles/ncu_summarise.py rep.ncu-rep [kernel-regex] [launch#]
Prints key raw metrics, the stall-reason totals and the opcode mix from the SASS source page."""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "k_"
idx = sys.argv[3] if len(sys.argv) > 3 else "1"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("==", r[hdr.index("ID")], name[:90])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:70s} {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{kre}:{idx}"],
                     capture_output=True, text=True).stdout
lines = src.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i, l in enumerate(lines) if i > start and l.startswith('"Kernel Name"')), len(lines))
rd = csv.DictReader(io.StringIO("\n".join(lines[start:end])))
stall = collections.Counter(); ops = collections.Counter(); tot = 0; execd = 0
for row in rd:
    try:
        n = int(row["Instructions Executed"] or 0)
    except ValueError:
        continue
    execd += n
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", row["Source"])
    op = m.group(2) if m else "?"
    ops[op] += n
    for k, v in row.items():
        if k.startswith("stall_") and "Not Issued" not in k and v:
            stall[k] += int(v); tot += int(v)
print("-- stall samples (all):")
for k, v in stall.most_common(10):
    print(f"   {k:28s} {100.0 * v / max(tot, 1):5.1f} %")
print(f"-- opcode mix of {execd:.3e} warp instructions:")
for k, v in ops.most_common(28):
    print(f"   {k:12s} {100.0 * v / max(execd, 1):5.1f} %")
