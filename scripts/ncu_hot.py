#!/usr/bin/env python
"""Top stall lines of one kernel from an ncu report (SASS view).  usage: ncu_hot.py rep kernel-regex [N]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
lines = out.split("\n")
# may contain several kernels; take the first block
blocks, cur = [], []
for l in lines:
    if l.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = [l]
    elif cur: cur.append(l)
if cur: blocks.append(cur)
b = blocks[min(int(sys.argv[4]) if len(sys.argv) > 4 else 0, len(blocks) - 1)] if blocks else sys.exit("no such kernel")
print(b[0][:200])
rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
hdr = rows[0]
si = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in rows[1:] if len(r) > si)
print("total samples", tot, "instructions", sum(int(r[ie] or 0) for r in rows[1:] if len(r) > ie))
agg = {}
for r in rows[1:]:
    if len(r) <= si: continue
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
top = sorted((r for r in rows[1:] if len(r) > si), key=lambda r: -int(r[si] or 0))[:N]
for r in top:
    st = sorted(((hdr[i], int(r[i] or 0)) for i in stall_cols), key=lambda kv: -kv[1])[:2]
    print(f"{int(r[si]):6d} {100*int(r[si])/max(tot,1):5.1f}% ex={r[ie]:>8} {r[src][:90]:90s} {st}")
