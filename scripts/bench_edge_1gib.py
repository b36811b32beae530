#!/usr/bin/env python
"""BASELINE configs[4] at the 1 GiB frame size: encode / decode time per call for the edge inputs (device resident,
CUDA events, async C ABI on the current stream).  usage: bench_edge_1gib.py codec[,codec...] [log2n]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
import torch
import hsrle_b200 as hs

names = (sys.argv[1] if len(sys.argv) > 1 else "rle8_multi,rle64_byte_packed").split(",")
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 30)
dev = torch.device("cuda:0")
def gen(kind):
    g = torch.Generator(device=dev); g.manual_seed(77)
    if kind == "single_symbol": return torch.full((n,), 0x5A, dtype=torch.uint8, device=dev)
    if kind == "random": return torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev, generator=g)
    if kind == "alternating": return torch.arange(n, dtype=torch.int32, device=dev).bitwise_and_(1).to(torch.uint8)
    starts = torch.rand(n, device=dev, generator=g) < (1.0 / 37.0)
    seg = torch.cumsum(starts.to(torch.int32), 0)
    return (seg.to(torch.int64) * 2654435761 % 251).to(torch.uint8)
cap = n + n // 256 + 512
sp = torch.cuda.current_stream().cuda_stream
for kind in ("single_symbol", "random", "alternating", "run_mixed"):
    t_in = gen(kind)
    for name in names:
        ws = torch.empty(max(hs.compress_workspace_size(name, n), hs.decompress_workspace_size(name, cap, n)), dtype=torch.uint8, device=dev)
        t_out = torch.empty(cap, dtype=torch.uint8, device=dev)
        t_dec = torch.empty(n + 128, dtype=torch.uint8, device=dev)
        res = torch.zeros(16, dtype=torch.int32, device=dev)
        hs.compress_device_async(name, t_in, t_out, ws, res[:8], sp); torch.cuda.synchronize()
        r = int(res[0].item())
        hs.decompress_device_async(name, t_out, r, t_dec, n, ws, res[8:], sp); torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps = 3
        ev[0].record()
        for _ in range(reps): hs.compress_device_async(name, t_in, t_out, ws, res[:8], sp)
        ev[1].record()
        for _ in range(reps): hs.decompress_device_async(name, t_out, r, t_dec, n, ws, res[8:], sp)
        ev[2].record(); torch.cuda.synchronize()
        te, td = ev[0].elapsed_time(ev[1]) / reps, ev[1].elapsed_time(ev[2]) / reps
        import ctypes
        buf = ctypes.create_string_buffer(8192)
        hs.lib.hsrle_timing_begin()
        hs.compress_device_async(name, t_in, t_out, ws, res[:8], sp)
        hs.decompress_device_async(name, t_out, r, t_dec, n, ws, res[8:], sp)
        hs.lib.hsrle_timing_end(buf, 8192)
        kt = {p.split(':')[0][2:]: round(1e3 * float(p.split(':')[2])) for p in buf.value.decode().split(";") if p}
        print(json.dumps({"input": kind, "codec": name, "n": n, "stream_bytes": r, "enc_ms": round(te, 3), "dec_ms": round(td, 3),
                          "enc_GBps": round(n / te / 1e6, 1), "dec_GBps": round(n / td / 1e6, 1),
                          "enc_roofline": round((n + r) / te / 1e6 / 6551.4, 3), "dec_roofline": round((n + r) / td / 1e6 / 6551.4, 3),
                          "roundtrip_ok": bool(torch.equal(t_dec[:n], t_in)), "kernel_us": kt}), flush=True)
        del ws, t_out, t_dec
    del t_in
