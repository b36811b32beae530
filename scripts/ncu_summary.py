#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/x_launches.csv  > profiles/x_launches.md
  python scripts/ncu_summary.py full     gpurun_out/x.ncu-rep       > profiles/x_full.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    acc = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        us = v / 1000.0 if u in ("ns", "nsecond") else (v * 1000.0 if u in ("ms", "msecond") else v)
        nm = r[ki].split("(")[0]
        a = acc.setdefault(nm, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in acc.values())
    print(f"# ncu launch list: {path}\n\n(cold-cache, serialised launches -- compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | avg us | share |\n|---|--:|--:|--:|--:|")
    for nm, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"| {nm} | {c} | {t:.1f} | {t / c:.2f} | {100 * t / tot:.1f}% |")
    print(f"| total | {sum(a[0] for a in acc.values())} | {tot:.1f} | | |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary: {path}\n")
    for r in rows[2:]:
        print(f"## {r[hdr.index('Kernel Name')]}  (launch id {r[0]})\n")
        print("| metric | value | unit |\n|---|--:|---|")
        for k in KEYS:
            if k in hdr:
                print(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
