#!/usr/bin/env python
"""Samples and executed instructions of one kernel launch in an .ncu-rep, bucketed by SASS instruction index ranges.
usage: python tools/ncu_regions.py file.ncu-rep <kernel regex> <section index> name:lo-hi [name:lo-hi ...]   (lo/hi = instruction indices, 0-based, inclusive)"""
import csv, subprocess, sys, io, collections
rep, rx, idx = sys.argv[1], sys.argv[2], int(sys.argv[3])
regions = []
for a in sys.argv[4:]:
    n, r = a.split(":"); lo, hi = r.split("-"); regions.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(out)))
starts = [n for n, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
rows = allrows[starts[idx]:starts[idx + 1]]
print(rows[0][1][:120])
hdr = rows[1]; H = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[H['# Samples']]) for r in data)
print("instructions", len(data), "samples", tot)
if not regions:
    # print a compact listing: index, samples, executed, source
    for n, r in enumerate(data):
        print(n, r[H['# Samples']], r[H['Instructions Executed']], r[H['Source']].strip()[:60])
for name, lo, hi in regions:
    sub = data[lo:hi + 1]
    s = sum(int(r[H['# Samples']]) for r in sub); ex = sum(int(r[H['Instructions Executed']]) for r in sub)
    agg = collections.Counter()
    for r in sub:
        for st in stalls:
            agg[st[6:]] += int(r[H[st]])
    top = ", ".join(f"{k} {100*v/max(s,1):.0f}%" for k, v in agg.most_common(6))
    print(f"{name:10s} idx {lo:5d}-{hi:5d}  samples {s:7d} ({100*s/tot:5.1f}%)  warp-instr executed {ex/1e6:8.1f}M  samples/Minstr {s/max(ex/1e6,1e-9):6.1f}   {top}")
