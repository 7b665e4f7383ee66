#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel launch in an .ncu-rep (source page).
usage: python tools/ncu_hot.py file.ncu-rep <kernel regex> <launch index among matches> [topN]"""
import csv, subprocess, sys, io, collections
rep, rx, idx = sys.argv[1], sys.argv[2], int(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(out)))
starts = [n for n, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
rows = allrows[starts[idx]:starts[idx + 1]]
print(rows[0][1][:150])
hdr = rows[1]
H = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[H['# Samples']]) for r in data)
print("instructions:", len(data), "samples:", tot)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.Counter()
for r in data:
    for s in stalls:
        agg[s] += int(r[H[s]])
print("stall totals:", {k: v for k, v in agg.most_common(10)})
byop = collections.Counter(); byop_n = collections.Counter()
for r in data:
    op = r[H['Source']].split()[0] if not r[H['Source']].strip().startswith('@') else r[H['Source']].split()[1]
    byop[op] += int(r[H['# Samples']]); byop_n[op] += int(r[H['Instructions Executed']])
print("samples by opcode:", [(k, v, byop_n[k]) for k, v in byop.most_common(14)])
srt = sorted(data, key=lambda r: -int(r[H['# Samples']]))[:top]
for r in srt:
    st = {s[6:]: int(r[H[s]]) for s in stalls if int(r[H[s]]) > 0}
    st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
    print(f"{int(r[H['# Samples']]):6d} {r[H['Instructions Executed']]:>9s} {r[H['Source']].strip()[:70]:70s} {st}")
