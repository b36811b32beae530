#!/usr/bin/env python
"""Long randomized GPU-vs-oracle parity sweep (diagnosis tool, not a test): multi-super-chunk inputs that stress the LUT
sensitivity scheme, the automaton repair paths and the decoder chain.  usage: gpu_fuzz.py seconds [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
import numpy as np
import hsrle_b200 as hs
from common import CODECS, gen_dct, gen_fuzz, gen_short_runs, gen_run_mixed, oracle_compress, out_capacity

budget = float(sys.argv[1]); seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
t0 = time.time(); nbad = 0; ncase = 0
def lut_stress(n, W):
    """few symbols, runs of 2..6 symbols, short gaps: every candidate is marginal or close to it"""
    alpha = rng.integers(0, 256, size=(int(rng.integers(3, 12)), W), dtype=np.uint8)
    parts = []; total = 0
    while total < n:
        s = alpha[int(rng.integers(0, len(alpha)))]
        r = int(rng.integers(2, 7)); g = int(rng.integers(0, 5))
        parts.append(np.tile(s, r)); total += r * W
        if g: parts.append(rng.integers(0, 256, size=g, dtype=np.uint8)); total += g
    return np.concatenate(parts)[:n].copy()
while time.time() - t0 < budget:
    kind = int(rng.integers(0, 6))
    n = int(rng.integers(200000, 3000000))
    W = int(rng.choice([1, 2, 3, 4, 6, 8]))
    if kind == 0: data = gen_dct(n, seed=int(rng.integers(1, 1 << 30)))
    elif kind == 1: data = gen_short_runs(n, seed=int(rng.integers(1, 1 << 30)), W=W)
    elif kind == 2: data = lut_stress(n, W)
    elif kind == 3: data = gen_fuzz(rng, n, long_every=int(rng.integers(0, 12)))
    elif kind == 4: data = gen_run_mixed(n, seed=int(rng.integers(1, 1 << 30)), max_run_log2=12, max_lit_log2=12)
    else: data = np.concatenate([gen_dct(n // 2, seed=int(rng.integers(1, 1 << 30))), lut_stress(n - n // 2, W)])
    sel = [c for c in CODECS if c.variant >= 2] if rng.random() < 0.6 else list(CODECS)
    for c in rng.choice(len(sel), size=min(6, len(sel)), replace=False):
        c = sel[int(c)]
        want = oracle_compress(c, data)
        got = hs.compress(c.cname, data, out_capacity(len(data)))
        ok = np.array_equal(got, want)
        r, dec = hs.decompress(c.dname, want, len(data))
        okd = r == len(data) and np.array_equal(dec, data)
        ncase += 1
        if not ok or not okd:
            nbad += 1
            print(f"BAD {c.name} kind={kind} n={n} W={W} seed={seed} case={ncase}: enc ok={ok} len {len(got)}/{len(want)} dec ok={okd} err={hs.last_error()!r}", flush=True)
            np.save(os.path.join(ROOT, "gpurun_out", f"fuzz_bad_{seed}_{ncase}.npy"), data)
print(f"FUZZ cases {ncase} bad {nbad} in {time.time() - t0:.0f} s")
