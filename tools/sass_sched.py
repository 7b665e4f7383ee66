#!/usr/bin/env python
"""Static schedule of a kernel's SASS: decode the control word of every instruction (stall count, yield,
scoreboard set / wait) from `cuobjdump -sass` and print per-opcode totals for an address range.
usage: python tools/sass_sched.py obj.o <substring of mangled name> [lo_addr hi_addr] [--list]"""
import subprocess, sys, re, collections

def parse(obj, pat):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur = None; ins = []; pending = None
    for ln in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", ln)
        if m:
            cur = m.group(1); continue
        if not (cur and pat in cur):
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", ln)
        if m:
            pending = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16)]
            continue
        m = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", ln)
        if m and pending:
            hi = int(m.group(1), 16)
            stall = (hi >> 41) & 0xf; yld = (hi >> 45) & 1; wbar = (hi >> 46) & 7; rbar = (hi >> 49) & 7; wait = (hi >> 52) & 0x3f
            ins.append(dict(addr=pending[0], text=pending[1], stall=stall, yld=yld, wbar=wbar, rbar=rbar, wait=wait))
            pending = None
    return ins

if __name__ == "__main__":
    obj, pat = sys.argv[1], sys.argv[2]
    args = [a for a in sys.argv[3:] if not a.startswith("--")]
    lo = int(args[0], 16) if len(args) > 0 else 0
    hi = int(args[1], 16) if len(args) > 1 else 1 << 60
    ins = [i for i in parse(obj, pat) if lo <= i["addr"] <= hi]
    tot = collections.Counter(); cnt = collections.Counter()
    for i in ins:
        op = re.sub(r"^@!?U?P\d+\s+", "", i["text"]).split()[0].split(".")[0]
        tot[op] += i["stall"]; cnt[op] += 1
        if "--list" in sys.argv:
            print(f"{i['addr']:06x} st={i['stall']:2d} y={i['yld']} wb={i['wbar']} rb={i['rbar']} wait={i['wait']:06b}  {i['text']}")
    n = len(ins); s = sum(tot.values())
    fp = sum(cnt[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(f"instructions {n}  sum(stall) {s}  avg {s/max(n,1):.2f}  fp64 {fp} ({100*fp/max(n,1):.1f}%)")
    for op, c in cnt.most_common(20):
        print(f"  {op:12s} n={c:4d} stall_sum={tot[op]:5d} avg={tot[op]/c:.2f}")
