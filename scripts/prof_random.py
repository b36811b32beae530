#!/usr/bin/env python
"""Per-kernel times of one decode of a 2^k-byte random (incompressible) frame.  usage: prof_random.py [codec] [log2n]"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
import torch
import hsrle_b200 as hs
name = sys.argv[1] if len(sys.argv) > 1 else "rle8_multi"
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 30)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(77)
t_in = torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev, generator=g)
cap = n + n // 256 + 512
ws = torch.empty(max(hs.compress_workspace_size(name, n), hs.decompress_workspace_size(name, cap, n)), dtype=torch.uint8, device=dev)
t_out = torch.empty(cap, dtype=torch.uint8, device=dev); t_dec = torch.empty(n + 128, dtype=torch.uint8, device=dev)
res = torch.zeros(16, dtype=torch.int32, device=dev); sp = torch.cuda.current_stream().cuda_stream
hs.compress_device_async(name, t_in, t_out, ws, res[:8], sp); torch.cuda.synchronize(); r = int(res[0].item())
hs.decompress_device_async(name, t_out, r, t_dec, n, ws, res[8:], sp); torch.cuda.synchronize()
buf = ctypes.create_string_buffer(8192)
hs.lib.hsrle_timing_begin()
hs.compress_device_async(name, t_in, t_out, ws, res[:8], sp)
hs.decompress_device_async(name, t_out, r, t_dec, n, ws, res[8:], sp)
hs.lib.hsrle_timing_end(buf, 8192)
print(name, n, r, " ".join(f"{p.split(':')[0][2:]}={1e3*float(p.split(':')[2]):.0f}us" for p in buf.value.decode().split(";") if p))
