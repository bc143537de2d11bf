#!/usr/bin/env python
"""Static instruction statistics of the plane loop of a streaming kernel (non-BC body = the last
inner loop in the function).  usage: loopstat.py <lib.so> <kernel-substring> [nth-loop-from-end]"""
import collections, re, subprocess, sys, tempfile, os, glob
so, key = sys.argv[1], sys.argv[2]
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 0
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
cub = [c for c in glob.glob(d + "/*.cubin") if "fused" in c][0]
elf = subprocess.run(["cuobjdump", "-elf", cub], capture_output=True, text=True).stdout
fn = sorted(set(re.findall(r"_ZN[0-9A-Za-z_]*" + re.escape(key) + r"[0-9A-Za-z_]*", elf)), key=len)[0]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, cub], capture_output=True, text=True).stdout
ins = []
for l in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
back = []
for a, t in ins:
    m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        back.append((a, int(m.group(1), 16)))
back.sort(reverse=True)
# inner plane loops: backward branches spanning > 0x2000 bytes, innermost = smallest span containing 5 BAR
loops = []
for a, tgt in back:
    body = [t for (x, t) in ins if tgt <= x <= a]
    nbar = sum(1 for t in body if t.split()[-2 if t.startswith('@') else 0].startswith('BAR') or 'BAR.SYNC' in t)
    if nbar == 5:
        loops.append((a, tgt, body))
a, tgt, body = loops[nth]
ops = collections.Counter()
for t in body:
    p = t.split()
    op = p[1] if p[0].startswith('@') else p[0]
    ops[op] += 1
tot = len(body)
f64 = sum(v for k, v in ops.items() if k.startswith(('DADD', 'DMUL', 'DFMA', 'DSETP')))
print(f"{fn[-40:]}: loop 0x{tgt:x}..0x{a:x}: {tot} instr, fp64 {f64}, FSEL {ops['FSEL']}, other {tot - f64 - ops['FSEL']}")
if '-v' in sys.argv:
    for k, v in ops.most_common(40):
        print(f"   {k:24s}{v}")
