#!/usr/bin/env python
"""Static opcode mix of one kernel in an object file (cuobjdump -sass).
usage: python tools/sass_mix.py obj.o <substring of mangled kernel name> """
import subprocess, sys, collections, re
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur = None; c = collections.Counter(); n = 0
for ln in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", ln)
    if m:
        cur = m.group(1); continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if m:
            c[m.group(2)] += 1; n += 1
fp64 = sum(v for k, v in c.items() if k.split('.')[0] in ("DFMA", "DMUL", "DADD", "DSETP"))
print("total", n, "fp64-pipe", fp64, f"({100*fp64/max(n,1):.1f}%)")
print([(k, v) for k, v in c.most_common(28)])
