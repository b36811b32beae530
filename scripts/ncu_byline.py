#!/usr/bin/env python
"""Instructions executed / stall samples per CUDA source line of one kernel (needs -lineinfo + --import-source).
usage: ncu_byline.py rep kernel-regex [N]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows, cur_file, hdr = [], "", None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].isdigit():
        def num(x):
            try:
                return int(float(x))
            except ValueError:
                return 0
        rows.append((cur_file, int(r[0]), r[1], num(r[hdr.index("# Samples")]), num(r[hdr.index("Instructions Executed")])))
ti = sum(r[4] for r in rows); ts = sum(r[3] for r in rows)
print("total warp instructions", ti, "samples", ts)
for f, ln, src, smp, ie in sorted(rows, key=lambda r: -r[4])[:N]:
    print(f"{100*ie/max(ti,1):5.1f}% inst {100*smp/max(ts,1):5.1f}% smp  {f}:{ln}  {src.strip()[:110]}")
