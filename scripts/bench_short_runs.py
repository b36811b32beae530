#!/usr/bin/env python
"""BASELINE configs[2], the codecs in scope: every exported codec on the short-run-heavy stream of SURVEY App. E.2
(alphabet 4-8 symbols, runs of 2-9 symbols, random gaps of 0-16 bytes; the automaton-chain stress case) at the
88,473,600-byte size.  Device resident, CUDA events, async C ABI on the current stream; the decode is compared with
the input (bit-exactness against the reference on this shape is tests/test_gpu_parity.py::test_structured_streams).
usage: bench_short_runs.py [codec,codec,...|all] [nbytes]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
import torch
import hsrle_b200 as hs
from common import CODECS

sel = sys.argv[1] if len(sys.argv) > 1 else "all"
codecs = [c for c in CODECS if sel == "all" or c.name in sel.split(",")]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 88473600
dev = torch.device("cuda:0")


def gen_short_runs(W, seed=7):
    """App. E.2 on the device: segments run, gap, run, gap, ...; a run repeats one of 4-8 W-byte symbols 2-9 times."""
    g = torch.Generator(device=dev); g.manual_seed(seed)
    m = n // (5 * W + 8) + 64                                      # mean run 5.5 W, mean gap 8: more segments than needed
    nsym = 4 + seed % 5
    alpha = torch.randint(0, 256, (nsym, W), dtype=torch.uint8, device=dev, generator=g)
    lens = torch.empty(2 * m, dtype=torch.int64, device=dev)
    lens[0::2] = torch.randint(2, 10, (m,), device=dev, generator=g) * W
    lens[1::2] = torch.randint(0, 17, (m,), device=dev, generator=g)
    ends = torch.cumsum(lens, 0)
    assert int(ends[-1].item()) >= n
    pos = torch.arange(n, device=dev)
    seg = torch.searchsorted(ends, pos, right=True)
    start = ends[seg] - lens[seg]
    sym = torch.randint(0, nsym, (m,), device=dev, generator=g)
    run = alpha[sym[seg // 2], (pos - start) % W]
    gap = torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev, generator=g)
    return torch.where(seg % 2 == 0, run, gap)


cap = n + n // 256 + 512
sp = torch.cuda.current_stream().cuda_stream
streams = {}
for c in codecs:
    for W in sorted({1, c.W}):
        if W not in streams:
            streams[W] = gen_short_runs(W, seed=7 + W)
        t_in = streams[W]
        name = c.name
        ws = torch.empty(max(hs.compress_workspace_size(name, n), hs.decompress_workspace_size(name, cap, n)), dtype=torch.uint8, device=dev)
        t_out = torch.empty(cap, dtype=torch.uint8, device=dev)
        t_dec = torch.empty(n + 128, dtype=torch.uint8, device=dev)
        res = torch.zeros(16, dtype=torch.int32, device=dev)
        hs.compress_device_async(name, t_in, t_out, ws, res[:8], sp); torch.cuda.synchronize()
        r = int(res[0].item())
        assert r > 0, (name, res.tolist())
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
        kus = {q.split(":")[0][2:]: round(1e3 * float(q.split(":")[2])) for q in buf.value.decode().split(";") if q}
        rl = res.tolist()
        diag = {"records": rl[2], "super_chunks": rl[3], "serial_sc": rl[4], "inner_serial": rl[5], "dirty_r0": rl[7], "dirty_r1": rl[6] & 0xFFFF, "dirty_r2": rl[6] >> 16}
        print(json.dumps({"input": "short_runs_W%d" % W, "codec": name, "n": n, "stream_bytes": r, "enc_ms": round(te, 3), "dec_ms": round(td, 3),
                          "enc_GBps": round(n / te / 1e6, 1), "dec_GBps": round(n / td / 1e6, 1),
                          "enc_roofline": round((n + r) / te / 1e6 / 6551.4, 3), "dec_roofline": round((n + r) / td / 1e6 / 6551.4, 3),
                          "roundtrip_ok": bool(torch.equal(t_dec[:n], t_in)), "kernel_us": kus, "enc_diag": diag}), flush=True)
        del ws, t_out, t_dec
