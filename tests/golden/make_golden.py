"""Generates tests/golden/golden_small.npz and tests/golden/golden_hashes.json from the compiled,
UNMODIFIED reference (oracle/_ref/libhsrle_ref.so, built from /root/reference/src by
oracle/Makefile).  The reference ships no golden vectors (its tests are round-trip only), so these
fixtures are what pins the oracle and the CUDA path on machines where /root/reference is absent.

Run from the repo root, in the build container:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from common import CODECS, gen_dct, gen_fuzz, gen_run_mixed, gen_short_runs, ref_compress, ref_lib  # noqa: E402


def small_inputs():
    rng = np.random.default_rng(20261017)
    ins = {
        "fuzz_a": gen_fuzz(rng, 1500),
        "fuzz_b": gen_fuzz(rng, 1200, max_sym=3, p_run=0.7),
        "dct": gen_dct(1536, seed=3),
        "short1": gen_short_runs(1024, seed=5, W=1),
        "short2": gen_short_runs(1024, seed=6, W=2),
        "tiny": np.array([7, 7, 7, 7, 7, 7, 7, 1, 2, 3, 3, 3], dtype=np.uint8),
        "one": np.array([42], dtype=np.uint8),
        "zeros": np.zeros(700, dtype=np.uint8),
        "altern": np.tile(np.array([1, 2], dtype=np.uint8), 300),
    }
    return ins


def large_inputs():
    """Seeded inputs regenerated at test time; only the reference stream hashes are stored."""
    rng = np.random.default_rng(99)
    return {
        "dct_1m": gen_dct(1 << 20, seed=0x5EED),
        "short1_512k": gen_short_runs(1 << 19, seed=7, W=1),
        "short3_256k": gen_short_runs(1 << 18, seed=8, W=3),
        "mixed_1m": gen_run_mixed(1 << 20, seed=11),
        "fuzz_300k": gen_fuzz(rng, 300000, long_every=9),
        "random_100k": np.random.default_rng(5).integers(0, 256, size=100000, dtype=np.uint8),
    }


def main():
    assert ref_lib() is not None, "compiled reference missing: make -C oracle ref"
    small = small_inputs()
    blob = {}
    for k, v in small.items():
        blob["in__" + k] = v
        for c in CODECS:
            blob[f"out__{k}__{c.name}"] = ref_compress(c, v)
    # mode-1 (single symbol) streams: only the reference's own single encoders can make them
    # (src/rle8_extreme_cpu.h:346-700; the single encoders themselves are out of scope)
    import ctypes
    lib = ref_lib()
    u8p = ctypes.POINTER(ctypes.c_uint8)
    rng = np.random.default_rng(77)
    for nm in ("rle8_single_compress", "rle8_packed_single_compress"):
        f = getattr(lib, nm)
        f.restype = ctypes.c_uint32
        f.argtypes = [u8p, ctypes.c_uint32, u8p, ctypes.c_uint32]
        for i, n in enumerate((40, 700, 3000, 20000)):
            data = gen_fuzz(rng, n, max_sym=1, p_run=0.6)
            data[rng.random(n) < 0.5] = 0
            buf = np.zeros(n + 64, dtype=np.uint8); buf[:n] = data
            out = np.zeros(n + 1024, dtype=np.uint8)
            r = f(buf.ctypes.data_as(u8p), n, out.ctypes.data_as(u8p), len(out))
            assert r > 0 and out[8] == 1, (nm, n, int(out[8]))
            blob[f"single_in__{nm}__{i}"] = data
            blob[f"single_out__{nm}__{i}"] = out[:r].copy()
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **blob)
    hashes = {}
    for k, v in large_inputs().items():
        hashes[k] = {"input_sha256": hashlib.sha256(v.tobytes()).hexdigest(), "n": int(len(v)), "streams": {}}
        for c in CODECS:
            s = ref_compress(c, v)
            hashes[k]["streams"][c.name] = {"len": int(len(s)), "sha256": hashlib.sha256(s.tobytes()).hexdigest()}
    with open(os.path.join(HERE, "golden_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)
    print("wrote", len(blob), "arrays and", len(hashes), "hash sets")
    write_full_size()


FULL_N = 88473600      # BASELINE configs[0,1]: the video-frame.raw shape (README.md:25)


def write_full_size():
    """Reference stream length + sha256 of every codec on the full 88,473,600-byte DCT stream (the bench input):
    tests/golden/golden_hashes_88m.json, asserted by tests/test_gpu_parity.py and by bench.py before it times."""
    v = gen_dct(FULL_N)
    out = {"input": "gen_dct(88473600) (SURVEY App. E.1)", "input_sha256": hashlib.sha256(v.tobytes()).hexdigest(), "n": FULL_N, "streams": {}}
    for c in CODECS:
        s = ref_compress(c, v)
        out["streams"][c.name] = {"len": int(len(s)), "sha256": hashlib.sha256(s.tobytes()).hexdigest()}
    with open(os.path.join(HERE, "golden_hashes_88m.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote full-size hashes for", len(out["streams"]), "codecs")


if __name__ == "__main__":
    main()
