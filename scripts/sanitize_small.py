#!/usr/bin/env python
"""Small encode+decode of a few codecs: the command to run under compute-sanitizer.  usage: sanitize_small.py [nbytes]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
import numpy as np, torch
import hsrle_b200 as hs
from common import gen_dct, gen_fuzz, gen_short_runs, gen_run_mixed

n = int(sys.argv[1]) if len(sys.argv) > 1 else 700000
dev = torch.device("cuda:0")
rng = np.random.default_rng(3)
inputs = [gen_dct(n, seed=4), gen_fuzz(rng, n // 2, long_every=7), gen_short_runs(n // 2, seed=5, W=1),
          gen_run_mixed(2 * n, seed=6, max_run_log2=19, max_lit_log2=17)]      # long literals and runs: grid-wide operations in the decoder
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["rle8_multi", "rle8_packed_multi", "rle8_3symlut", "rle8_7symlut", "rle16_7symlut_byte", "rle24_byte_packed", "rle32_3symlut_sym", "rle48_sym", "rle64_byte_packed"]
for data in inputs:
    for name in names:
        m = len(data)
        cap = m + m // 256 + 512
        t_in = torch.from_numpy(data).to(dev)
        t_out = torch.zeros(cap, dtype=torch.uint8, device=dev)
        ws = torch.empty(max(hs.compress_workspace_size(name, m), hs.decompress_workspace_size(name, cap, m)), dtype=torch.uint8, device=dev)
        res = torch.zeros(16, dtype=torch.int32, device=dev)
        sp = torch.cuda.current_stream().cuda_stream
        hs.compress_device_async(name, t_in, t_out, ws, res[:8], sp)
        torch.cuda.synchronize()
        r = int(res[0].item())
        t_dec = torch.zeros(m + 128, dtype=torch.uint8, device=dev)
        hs.decompress_device_async(name, t_out, r, t_dec, m, ws, res[8:], sp)
        torch.cuda.synchronize()
        print(name, m, r, int(res[8].item()), bool(torch.equal(t_dec[:m], t_in)), flush=True)
