#!/usr/bin/env python
"""Workspace-poison check: encode/decode must not depend on what the caller's workspace held before.
usage: debug_poison.py codec log2n [kind]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
import numpy as np, torch
import hsrle_b200 as hs
from common import gen_dct

name = sys.argv[1]; n = 1 << int(sys.argv[2]); kind = sys.argv[3] if len(sys.argv) > 3 else "run_mixed"
dev = torch.device("cuda:0")
if kind == "dct":
    t_in = torch.from_numpy(gen_dct(n)).to(dev)
else:
    g = torch.Generator(device=dev); g.manual_seed(77)
    starts = torch.rand(n, device=dev, generator=g) < (1.0 / 37.0)
    seg = torch.cumsum(starts.to(torch.int32), 0)
    t_in = (seg.to(torch.int64) * 2654435761 % 251).to(torch.uint8)
    del starts, seg
cap = n + n // 256 + 512
ws = torch.empty(max(hs.compress_workspace_size(name, n), hs.decompress_workspace_size(name, cap, n)), dtype=torch.uint8, device=dev)
res = torch.zeros(16, dtype=torch.int32, device=dev)
sp = torch.cuda.current_stream().cuda_stream
outs = []
for fill in (0x00, 0xFF, 0x01, None):
    if fill is not None: ws.fill_(fill)
    t_out = torch.zeros(cap, dtype=torch.uint8, device=dev)
    hs.compress_device_async(name, t_in, t_out, ws, res[:8], sp)
    torch.cuda.synchronize()
    r = res[:8].tolist()
    print("enc fill", fill, "res", r, flush=True)
    outs.append((r[0], t_out))
for i in range(1, len(outs)):
    same = outs[i][0] == outs[0][0] and torch.equal(outs[i][1][:outs[0][0]], outs[0][1][:outs[0][0]])
    print("enc run", i, "identical to run 0:", same)
r0, t_c = outs[0]
for fill in (0x00, 0xFF, 0x01, None):
    if fill is not None: ws.fill_(fill)
    t_dec = torch.zeros(n + 128, dtype=torch.uint8, device=dev)
    hs.decompress_device_async(name, t_c, r0, t_dec, n, ws, res[8:], sp)
    torch.cuda.synchronize()
    print("dec fill", fill, "res", res[8:].tolist(), "ok", bool(torch.equal(t_dec[:n], t_in)), flush=True)
