"""Ad-hoc GPU diagnosis (not a test): prints per-codec first-mismatch details and timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
from common import *
import hsrle_b200 as hs

def first_diff(a, b):
    m = min(len(a), len(b))
    d = np.flatnonzero(a[:m] != b[:m])
    return int(d[0]) if len(d) else (m if len(a) != len(b) else -1)

def main():
    names = sys.argv[1:] or [c.name for c in CODECS]
    rng = np.random.default_rng(5)
    inputs = {"tiny": np.array([7]*7+[1,2,3,3,3], dtype=np.uint8), "fuzz3k": gen_fuzz(rng, 3000), "fuzz200k": gen_fuzz(rng, 200000, long_every=9),
              "dct1m": gen_dct(1 << 20, seed=2), "short": gen_short_runs(300000, seed=3, W=1), "rand": rng.integers(0, 256, size=70000, dtype=np.uint8),
              "zeros": np.zeros(100000, dtype=np.uint8)}
    nbad = 0
    for nm in names:
        c = CODEC_BY_NAME[nm]
        for k, data in inputs.items():
            want = oracle_compress(c, data)
            t0 = time.time()
            got = hs.compress(c.cname, data, out_capacity(len(data)))
            t1 = time.time()
            ok = np.array_equal(got, want)
            r, dec = hs.decompress(c.dname, want, len(data))
            t2 = time.time()
            okd = (r == len(data)) and np.array_equal(dec, data)
            if not ok or not okd:
                nbad += 1
                print(f"BAD {nm} {k}: enc ok={ok} len got={len(got)} want={len(want)} firstdiff={first_diff(got, want)} | dec ok={okd} r={r} firstdiff={first_diff(dec, data) if r else None} err={hs.last_error()!r}", flush=True)
        print(f"{nm}: done ({(t1-t0)*1e3:.2f} ms enc / {(t2-t1)*1e3:.2f} ms dec on last input)", flush=True)
    print("TOTAL BAD", nbad)

main()
