#!/usr/bin/env python
"""Replays tests/test_gpu_parity.py::test_one_gib_frame_properties in one process and reports where a decode differs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
import numpy as np, torch
import hsrle_b200 as hs

names = sys.argv[1].split(","); kinds = sys.argv[2].split(","); n = 1 << int(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = torch.device("cuda:0")
def gen(kind):
    g = torch.Generator(device=dev); g.manual_seed(77)
    if kind == "single_symbol": return torch.full((n,), 0x5A, dtype=torch.uint8, device=dev)
    if kind == "random": return torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev, generator=g)
    if kind == "alternating": return torch.arange(n, dtype=torch.int32, device=dev).bitwise_and_(1).to(torch.uint8)
    starts = torch.rand(n, device=dev, generator=g) < (1.0 / 37.0)
    seg = torch.cumsum(starts.to(torch.int32), 0)
    return (seg.to(torch.int64) * 2654435761 % 251).to(torch.uint8)
for name in names:
    for kind in kinds:
        for rep in range(reps):
            t_in = gen(kind)
            cap = n + n // 256 + 512
            t_out = torch.empty(cap, dtype=torch.uint8, device=dev)
            r = hs.compress_device(name, t_in, t_out)
            t_dec = torch.empty(n + 128, dtype=torch.uint8, device=dev)
            t_dec.fill_(0xEE)
            rd = hs.decompress_device(name, t_out, r, t_dec, n)
            neq = (t_dec[:n] != t_in)
            cnt = int(neq.sum().item())
            print(name, kind, rep, "r", r, "rd", rd, "mismatching bytes", cnt, flush=True)
            if cnt:
                idx = torch.nonzero(neq)[:, 0]
                brk = torch.nonzero(idx[1:] - idx[:-1] > 1)[:, 0]
                print("  first", int(idx[0]), "last", int(idx[-1]), "ranges", len(brk) + 1)
                i0 = int(idx[0])
                print("  got ", t_dec[max(i0 - 8, 0): i0 + 24].cpu().numpy())
                print("  want", t_in[max(i0 - 8, 0): i0 + 24].cpu().numpy())
                starts_ = [int(idx[0])] + [int(idx[b + 1]) for b in brk[:10]]
                ends_ = [int(idx[b]) for b in brk[:10]] + [int(idx[-1])]
                print("  ranges:", list(zip(starts_, ends_))[:10])
            del t_in, t_out, t_dec
