#!/usr/bin/env python
"""Per-CUDA-source-line stall samples of one kernel.  usage: ncu_lines.py rep kernel-regex [N]"""
import csv, io, subprocess, sys, os
rep, kre = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fpath, hdr, acc, tot, seen_fn = "", None, {}, 0, None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        if seen_fn is None: seen_fn = r[1]
        cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed"); continue
    if hdr is None or cur_fn != seen_fn: continue
    if r[0].strip().isdigit() and len(r) > si and r[si].isdigit():
        key = (fpath, int(r[0]), r[1].strip()[:100])
        a = acc.setdefault(key, [0, 0])
        a[0] += int(r[si]); a[1] += int(r[ie]) if r[ie].isdigit() else 0
        tot += int(r[si])
print(seen_fn, "total samples", tot)
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0])[:N]:
    print(f"{v[0]:6d} {100*v[0]/max(tot,1):5.1f}% ex={v[1]:>9} {k[0]}:{k[1]:<4} {k[2]}")
