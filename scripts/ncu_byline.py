#!/usr/bin/env python
"""Instructions executed per CUDA source line (sorted by line) of one kernel.  usage: ncu_byline.py rep kernel-regex [minM]"""
import csv, io, subprocess, sys, os
rep, kre = sys.argv[1], sys.argv[2]
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
    if r[0].strip().isdigit() and len(r) > si and r[ie].isdigit():
        key = (fpath, int(r[0]))
        a = acc.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += int(r[si]) if r[si].isdigit() else 0; a[1] += int(r[ie]); tot += int(r[ie])
print(seen_fn, "total warp-instr", tot)
thr = float(sys.argv[3]) * 1e6 if len(sys.argv) > 3 else tot * 0.004
for k, v in sorted(acc.items()):
    if v[1] >= thr: print(f"{v[1]/1e6:8.2f}M smp={v[0]:5d} {k[0]}:{k[1]:<4} {v[2]}")
