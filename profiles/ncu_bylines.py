#!/usr/bin/env python
"""Attribute executed SASS instructions / stall samples of one kernel to source lines.
usage: ncu_bylines.py <source-page.csv from ncu --page source --csv> <nvdisasm --print-line-info dump> <mangled-kernel-substring>"""
import re, collections, csv, io, sys
srccsv, sass, key = sys.argv[1:4]
cur = None; fn = None; off2line = {}
for l in open(sass):
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: fn = m.group(1)
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2)))
    if fn and key in fn:
        m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
        if m: off2line[int(m.group(1), 16)] = cur
lines = open(srccsv).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i, l in enumerate(lines) if i > start and l.startswith('"Kernel Name"')), len(lines))
rd = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
def addr(r):
    a = r['Address']; return int(a, 16) if a.startswith('0x') else int(a)
base = addr(rd[0])
byline = collections.Counter(); bra = collections.Counter(); mov = collections.Counter(); samp = collections.Counter(); f64 = collections.Counter()
tot = 0
for r in rd:
    ln = off2line.get(addr(r) - base)
    n = int(r['Instructions Executed'] or 0); tot += n
    byline[ln] += n
    samp[ln] += int(r['# Samples'] or 0)
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r['Source'])
    op = m.group(2) if m else '?'
    if op.startswith(('BRA', 'BSSY', 'BSYNC')): bra[ln] += n
    if op.startswith(('IMAD.MOV', 'MOV')): mov[ln] += n
    if op.startswith(('DADD', 'DMUL', 'DFMA', 'DSETP')): f64[ln] += n
print("total warp instr", tot)
st = sum(samp.values())
print("line | %instr | %fp64 | %branch | %mov | %samples")
for k, v in byline.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 40):
    print(k, f"{100*v/tot:5.1f}% f64 {100*f64[k]/tot:4.1f}% bra {100*bra[k]/tot:4.1f}% mov {100*mov[k]/tot:4.1f}% samp {100*samp[k]/st:4.1f}%")
