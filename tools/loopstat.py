#!/usr/bin/env python
"""Static instruction statistics of the loops of one kernel in libmxadv.so.
usage: loopstat.py <lib.so> <kernel-substring> [-v] [--min N]
Lists every loop (backward branch) with at least --min instructions (default 300): size, fp64-pipe
count, selects, local-memory (spill) accesses, barriers; -v adds the opcode histogram."""
import collections, re, subprocess, sys, tempfile, os, glob
so, key = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[sys.argv.index('--min') + 1]) if '--min' in sys.argv else 300
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
def opof(t):
    p = t.split()
    return p[1] if p[0].startswith('@') else p[0]
loops = []
for a, t in ins:
    m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        tgt = int(m.group(1), 16)
        body = [x for x in ins if tgt <= x[0] <= a]
        if len(body) >= minlen:
            loops.append((tgt, a, body))
print(f"{fn[-46:]}: {len(ins)} instructions")
for tgt, a, body in sorted(loops):
    ops = collections.Counter(opof(t) for _, t in body)
    f64 = sum(v for k, v in ops.items() if k.startswith(('DADD', 'DMUL', 'DFMA', 'DSETP')))
    sel = sum(v for k, v in ops.items() if k.startswith(('FSEL', 'SEL')))
    loc = sum(v for k, v in ops.items() if k.startswith(('LDL', 'STL')))
    bar = sum(v for k, v in ops.items() if k.startswith('BAR'))
    sm = sum(v for k, v in ops.items() if k.startswith(('LDS', 'STS')))
    print(f"  loop 0x{tgt:x}..0x{a:x}: {len(body)} instr, fp64 {f64}, sel {sel}, lds/sts {sm}, local {loc}, bar {bar}")
    if '-v' in sys.argv:
        for k, v in ops.most_common(28):
            print(f"       {k:24s}{v}")
