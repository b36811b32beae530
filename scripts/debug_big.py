#!/usr/bin/env python
"""Diagnosis at large sizes: GPU vs compiled reference on a device-generated run-mixed stream.  usage: debug_big.py codec log2n"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
import numpy as np, torch
import hsrle_b200 as hs
from common import CODEC_BY_NAME, ref_compress

name = sys.argv[1]; n = 1 << int(sys.argv[2])
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(77)
starts = torch.rand(n, device=dev, generator=g) < (1.0 / 37.0)
seg = torch.cumsum(starts.to(torch.int32), 0)
t_in = (seg.to(torch.int64) * 2654435761 % 251).to(torch.uint8)
del starts, seg
codec = CODEC_BY_NAME[name]
cap = n + n // 256 + 512
t_out = torch.empty(cap, dtype=torch.uint8, device=dev)
r = hs.compress_device(name, t_in, t_out)
print("gpu compress", r, hs.last_error())
want = ref_compress(codec, t_in.cpu().numpy())
print("ref compress", len(want))
got = t_out[:r].cpu().numpy()
m = min(len(got), len(want))
d = np.flatnonzero(got[:m] != want[:m])
print("enc first diff", int(d[0]) if len(d) else None, "len equal", len(got) == len(want))
t_ref = torch.from_numpy(want).to(dev)
t_pad = torch.zeros(len(want) + 256, dtype=torch.uint8, device=dev); t_pad[:len(want)] = t_ref
t_dec = torch.zeros(n + 128, dtype=torch.uint8, device=dev)
rd = hs.decompress_device(name, t_pad, len(want), t_dec, n)
print("gpu decompress of ref stream", rd, hs.last_error())
neq = (t_dec[:n] != t_in)
cnt = int(neq.sum().item())
print("dec mismatching bytes", cnt)
if cnt:
    idx = torch.nonzero(neq)[:, 0]
    print("first", int(idx[0]), "last", int(idx[-1]), "n", n)
    i0 = int(idx[0])
    print("got ", t_dec[i0 - 8: i0 + 24].cpu().numpy())
    print("want", t_in[i0 - 8: i0 + 24].cpu().numpy())
    # run structure of mismatches
    brk = torch.nonzero(idx[1:] - idx[:-1] > 1)[:, 0]
    print("mismatch ranges", len(brk) + 1, "first ranges:", [(int(idx[0]),)] + [(int(idx[b + 1]),) for b in brk[:8]])
