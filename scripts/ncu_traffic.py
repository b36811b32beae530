#!/usr/bin/env python
"""DRAM bytes per launch of every kernel in one or more `ncu --set full` reports -> profiles/ncu_traffic.json
usage: ncu_traffic.py <source note> rep [rep ...] > profiles/ncu_traffic.json"""
import collections, csv, io, json, subprocess, sys
note, reps = sys.argv[1], sys.argv[2:]
acc = collections.OrderedDict()
for rep in reps:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    ki = hdr.index("Kernel Name"); ri = hdr.index("dram__bytes_read.sum"); wi = hdr.index("dram__bytes_write.sum")
    units = rows[1]
    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    for r in rows[2:]:
        if len(r) <= wi:
            continue
        name = r[ki].split("(")[0].split("<")[0].split("::")[-1].replace("void ", "").strip()
        a = acc.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
json.dump({k: {"dram_bytes_per_launch": int(v[1] / v[0]), "captures": v[0], "source": note} for k, v in acc.items()}, sys.stdout, indent=1)
