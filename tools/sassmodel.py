#!/usr/bin/env python
"""Crude single-warp, in-order issue model of one loop of a kernel in libmxadv.so.

usage: sassmodel.py <lib.so> <kernel-substring> <loop-start-hex> <loop-end-hex>

For every instruction of the loop body (straight-line: branches are ignored, predicated
instructions are assumed to execute): issue cycle = max(previous issue + 1, operands ready,
pipe free).  Latencies / issue rates are the ones measured on B200 with tools/ubench (fp64 add /
mul / fma: 8 cycles dependent latency, one warp instruction per 2 cycles; ALU and FMA-lite pipes
one per 2 cycles; LDS ~30, SHFL ~25 cycles).  Global loads are taken as prefetched (40 cycles).
Prints the modelled cycles for one trip of one warp, the issue-slot count and the fp64-pipe
cycles: the first is what a warp needs alone, the other two bound what several warps can share.
"""
import re
import subprocess
import sys
import tempfile
import os
import glob

so, key, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
cub = [c for c in glob.glob(d + "/*.cubin") if "fused" in c][0]
elf = subprocess.run(["cuobjdump", "-elf", cub], capture_output=True, text=True).stdout
fn = sorted(set(re.findall(r"_ZN[0-9A-Za-z_]*" + re.escape(key) + r"[0-9A-Za-z_]*", elf)), key=len)[0]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, cub], capture_output=True, text=True).stdout
ins = []
for l in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m and lo <= int(m.group(1), 16) <= hi:
        ins.append(m.group(2).strip())

LAT = {"D": 8, "LDS": 30, "SHFL": 25, "LDG": 40, "LDL": 30, "LDC": 20, "S2R": 20}


def regs(tok, wide):
    out = []
    for m in re.finditer(r"\b(U?R)(\d+)\b", tok):
        n = int(m.group(2))
        out.append((m.group(1), n))
        if wide:
            out.append((m.group(1), n + 1))
    for m in re.finditer(r"\b(U?P)(\d)\b", tok):
        out.append((m.group(1), int(m.group(2))))
    return out


ready = {}
t = 0
pipe_free = {"fp64": 0, "alu": 0, "fma": 0, "lsu": 0}
n_f64 = 0
for text in ins:
    parts = text.split(None, 1)
    guard = None
    if parts[0].startswith("@"):
        guard = parts[0][1:].lstrip("!")
        parts = parts[1].split(None, 1)
    op = parts[0]
    args = parts[1] if len(parts) > 1 else ""
    ops = [a.strip() for a in args.split(",")] if args else []
    base = op.split(".")[0]
    wide = base.startswith("D") and base not in ("DEPBAR",) or ".64" in op or base == "IMAD" and ".WIDE" in op
    is_store = base in ("STS", "STG", "STL", "BAR", "BRA", "BSSY", "BSYNC", "EXIT", "DEPBAR", "LDGDEPBAR",
                        "WARPSYNC", "NOP", "BRX", "SYNCS", "UTMALDG", "ARRIVES", "UCGABAR_ARV", "MEMBAR",
                        "FENCE", "ERRBAR", "CCTL", "REDUX", "RED", "ATOMS", "USETMAXREG")
    dsts, srcs = [], []
    if not is_store and ops:
        ndst = 2 if base in ("DSETP", "ISETP", "PLOP3", "UISETP", "UPLOP3", "FSETP") else 1
        if base in ("SHFL", "VOTE", "LEA", "IADD3", "UIADD3", "ULEA") and len(ops) > 1 and re.match(r"^U?P[T\d]$", ops[0]):
            ndst = 2
        if base in ("IADD3", "LEA", "UIADD3", "ULEA"):
            # Rd, [P], [P], srcs...
            k = 1
            while k < len(ops) and re.match(r"^U?P[T\d]$", ops[k]):
                k += 1
            ndst = k
        for o in ops[:ndst]:
            dsts += regs(o, wide and base.startswith("D") and base != "DSETP" or (".64" in op) or (".WIDE" in op))
        for o in ops[ndst:]:
            srcs += regs(o, base.startswith("D") or (".64" in o))
    else:
        for o in ops:
            srcs += regs(o, ".64" in op or ".64" in o)
    if guard:
        srcs += regs(guard, False)
    if base.startswith("D") and base != "DEPBAR":
        pipe, lat = "fp64", LAT["D"]
        n_f64 += 1
    elif base in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "LDGSTS"):
        pipe, lat = "lsu", LAT.get(base, 4)
    elif base in ("IMAD", "FMUL", "FFMA", "FADD"):
        pipe, lat = "fma", 5
    elif base in ("LDC", "LDCU", "S2R", "S2UR"):
        pipe, lat = "alu", LAT.get(base if base in LAT else "LDC", 20)
    else:
        pipe, lat = "alu", 5
    issue = max(t + 1, pipe_free[pipe], max([ready.get(r, 0) for r in srcs] + [0]))
    pipe_free[pipe] = issue + (2 if pipe in ("fp64", "alu", "fma") else 2)
    for r in dsts:
        ready[r] = issue + lat
    t = issue
print(f"{fn[-40:]} loop {lo:#x}..{hi:#x}: {len(ins)} instr, fp64 {n_f64}; single-warp in-order model: {t} cycles "
      f"({t / max(len(ins), 1):.2f} cyc/instr); fp64 pipe {2 * n_f64} cycles")
