#!/usr/bin/env python
"""Decoder debugging aid: encodes a few inputs with the oracle, decodes on the GPU, prints the result words and the first
mismatch.  usage: dec_debug.py [codec ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
import numpy as np, torch
import hsrle_b200 as hs
from common import CODEC_BY_NAME, gen_dct, gen_fuzz, gen_run_mixed, gen_short_runs, oracle_compress

names = sys.argv[1:] or ["rle8_multi", "rle8_packed_multi", "rle8_3symlut", "rle16_sym", "rle24_byte_packed", "rle32_7symlut_sym", "rle64_byte"]
dev = torch.device("cuda:0")
rng = np.random.default_rng(5)
inputs = [("mixed4m", gen_run_mixed(4 << 20, seed=24)), ("tiny", np.array([7] * 7 + [1, 2, 3, 3, 3], dtype=np.uint8)), ("fuzz3k", gen_fuzz(rng, 3000)), ("dct100k", gen_dct(100000, seed=2)),
          ("dct3m", gen_dct(3 << 20, seed=21)), ("short2m", gen_short_runs(2 << 20, seed=23, W=1)), ("mixed4m", gen_run_mixed(4 << 20, seed=24)),
          ("zeros1m", np.zeros(1 << 20, dtype=np.uint8)), ("random1m", rng.integers(0, 256, size=1 << 20, dtype=np.uint8))]
bad = 0
for name in names:
    codec = CODEC_BY_NAME[name]
    for label, data in inputs:
        n = len(data)
        s = oracle_compress(codec, data)
        t_s = torch.from_numpy(s).to(dev)
        t_dec = torch.full((n + 256,), 0xEE, dtype=torch.uint8, device=dev)
        ws = torch.empty(hs.decompress_workspace_size(name, len(s), n), dtype=torch.uint8, device=dev)
        res = torch.zeros(8, dtype=torch.int32, device=dev)
        hs.decompress_device_async(name, t_s, len(s), t_dec, n, ws, res, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        r = res.tolist()
        got = t_dec[:n].cpu().numpy()
        ok = r[0] == n and r[1] == 0 and np.array_equal(got, data) and bool((t_dec[n:] == 0xEE).all())
        if not ok:
            bad += 1
            d = np.nonzero(got != data)[0]
            print("FAIL", name, label, "n", n, "clen", len(s), "res", r, "first diff", d[:5], "ndiff", len(d), "tail ok", bool((t_dec[n:] == 0xEE).all()), flush=True)
        else:
            print("ok  ", name, label, r[:4], flush=True)
print("bad =", bad)
