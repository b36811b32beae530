#!/usr/bin/env python
"""One 2^30-byte frame of the configs[3] run-mixed stream, device resident: encode + decode time and the
per-kernel breakdown.  usage: prof_frame.py [codec,codec...] [frame] [reps] [what]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
import torch
import hsrle_b200 as hs
from common import gen_run_mixed_pieces, RM_PIECE

names = (sys.argv[1] if len(sys.argv) > 1 else "rle8_multi,rle64_byte,rle64_byte_packed").split(",")
frame = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
what = sys.argv[4] if len(sys.argv) > 4 else "both"
n = 1 << 30
dev = torch.device("cuda:0")
t_in = gen_run_mixed_pieces(frame * (n // RM_PIECE), n // RM_PIECE, dev)
cap = n + n // 256 + 512
sp = torch.cuda.current_stream().cuda_stream
for name in names:
    ws = torch.empty(max(hs.compress_workspace_size(name, n), hs.decompress_workspace_size(name, cap, n)), dtype=torch.uint8, device=dev)
    t_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    t_dec = torch.empty(n + 128, dtype=torch.uint8, device=dev)
    res = torch.zeros(16, dtype=torch.int32, device=dev)
    hs.compress_device_async(name, t_in, t_out, ws, res[:8], sp); torch.cuda.synchronize()
    r = int(res[0].item()) & 0xFFFFFFFF
    hs.decompress_device_async(name, t_out, r, t_dec, n, ws, res[8:], sp); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    if what in ("enc", "both"):
        for _ in range(reps): hs.compress_device_async(name, t_in, t_out, ws, res[:8], sp)
    ev[1].record()
    if what in ("dec", "both"):
        for _ in range(reps): hs.decompress_device_async(name, t_out, r, t_dec, n, ws, res[8:], sp)
    ev[2].record(); torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(8192)
    hs.lib.hsrle_timing_begin()
    if what in ("enc", "both"): hs.compress_device_async(name, t_in, t_out, ws, res[:8], sp)
    if what in ("dec", "both"): hs.decompress_device_async(name, t_out, r, t_dec, n, ws, res[8:], sp)
    hs.lib.hsrle_timing_end(buf, 8192)
    kt = " ".join(f"{p.split(':')[0][2:]}={1e3*float(p.split(':')[2]):.0f}" for p in buf.value.decode().split(";") if p)
    print(name, "frame", frame, "clen", r, "enc ms", round(ev[0].elapsed_time(ev[1]) / reps, 3), "dec ms", round(ev[1].elapsed_time(ev[2]) / reps, 3),
          "| kernel us:", kt, "| res", res.tolist(), "ok", bool(torch.equal(t_dec[:n], t_in)) if what != "enc" else None, flush=True)
    del ws, t_out, t_dec
