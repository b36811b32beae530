#!/usr/bin/env python
"""One codec, device resident, a few encode+decode repetitions of the 88 MB DCT stream: the command
to run under ncu.  usage: prof_one.py <codec> [reps] [enc|dec|both] [nbytes]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
import numpy as np, torch
import hsrle_b200 as hs
from common import gen_dct

name = sys.argv[1] if len(sys.argv) > 1 else "rle8_multi"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
what = sys.argv[3] if len(sys.argv) > 3 else "both"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 88473600
dev = torch.device("cuda:0")
data = gen_dct(n)
cap = n + n // 256 + 512
t_in = [torch.from_numpy(data).to(dev) for _ in range(2)]
t_out = torch.empty(cap, dtype=torch.uint8, device=dev)
t_dec = torch.empty(n + 128, dtype=torch.uint8, device=dev)
ws = torch.empty(max(hs.compress_workspace_size(name, n), hs.decompress_workspace_size(name, cap, n)), dtype=torch.uint8, device=dev)
res = torch.zeros(16, dtype=torch.int32, device=dev)
sp = torch.cuda.current_stream().cuda_stream
hs.compress_device_async(name, t_in[0], t_out, ws, res[:8], sp)
torch.cuda.synchronize()
r = int(res[0].item())
assert r > 0, res
hs.decompress_device_async(name, t_out, r, t_dec, n, ws, res[8:], sp)   # warm-up (lazy module load)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
if what in ("enc", "both"):
    for k in range(reps):
        hs.compress_device_async(name, t_in[k % 2], t_out, ws, res[:8], sp)
ev[1].record()
if what in ("dec", "both"):
    for k in range(reps):
        hs.decompress_device_async(name, t_out, r, t_dec, n, ws, res[8:], sp)
ev[2].record()
torch.cuda.synchronize()
import ctypes
buf = ctypes.create_string_buffer(8192)
hs.lib.hsrle_timing_begin()
if what in ("enc", "both"):
    hs.compress_device_async(name, t_in[1], t_out, ws, res[:8], sp)
if what in ("dec", "both"):
    hs.decompress_device_async(name, t_out, r, t_dec, n, ws, res[8:], sp)
hs.lib.hsrle_timing_end(buf, 8192)
kt = " ".join(f"{p.split(':')[0][2:]}={1e3*float(p.split(':')[2]):.0f}" for p in buf.value.decode().split(";") if p)
print(name, "n", n, "clen", r, "enc us", 1e3 * ev[0].elapsed_time(ev[1]) / reps, "dec us", 1e3 * ev[1].elapsed_time(ev[2]) / reps, "| kernel us:", kt, "| res", res.tolist())
if what in ("dec", "both"):
    assert torch.equal(t_dec[:n], t_in[0])
