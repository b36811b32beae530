"""CPU tests of the product's stage logic.  tests/sim/sim_pipeline.cpp drives the SAME per-item stage
functions of the ENCODER kernels (hsrle_core.cuh, hsrle_enc.cuh) and of the round-1 decoder design (tests/sim/hsrle_dec_v1.cuh, kept as a
second, independent decoder model -- the product decoder is covered by the GPU parity tests) from host loops and
must reproduce the oracle bit for bit.  The simulator is a test tool; it is not part of the product."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from common import CODECS, ROOT, gen_dct, gen_fuzz, gen_run_mixed, gen_short_runs, oracle_compress, out_capacity

SIM_DIR = os.path.join(ROOT, "tests", "sim")
_u8p = ctypes.POINTER(ctypes.c_uint8)


@pytest.fixture(scope="module")
def sim():
    so = os.path.join(SIM_DIR, "libsim.so")
    src = os.path.join(SIM_DIR, "sim_pipeline.cpp")
    hdrs = [os.path.join(ROOT, "hypersonic-rle-kit_b200", "csrc", h) for h in ("hsrle_core.cuh", "hsrle_enc.cuh")] + \
           [os.path.join(SIM_DIR, "hsrle_dec_v1.cuh")]
    newest = max(os.path.getmtime(p) for p in [src] + hdrs)
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.run(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    lib.sim_compress.restype = ctypes.c_uint32
    lib.sim_compress.argtypes = [ctypes.c_int] * 3 + [_u8p, ctypes.c_uint32, _u8p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
    lib.sim_decompress.restype = ctypes.c_uint32
    lib.sim_decompress.argtypes = [ctypes.c_int] * 3 + [_u8p, ctypes.c_uint32, _u8p, ctypes.c_uint32]
    return lib


def _enc(sim, c, data, rounds=3, maxit=6):
    """rounds: grid-level verify rounds before the sequential repair; maxit: in-CTA fixed-point rounds
    before the in-CTA sequential pass (0 forces the sequential paths)."""
    out = np.zeros(out_capacity(len(data)), dtype=np.uint8)
    r = sim.sim_compress(c.W, c.align, c.variant, data.ctypes.data_as(_u8p), len(data), out.ctypes.data_as(_u8p), len(out), rounds, maxit)
    return out[:r]


def _dec(sim, c, stream, n):
    stream = np.ascontiguousarray(stream)
    out = np.zeros(n + 16, dtype=np.uint8)
    r = sim.sim_decompress(c.W, c.align, c.variant, stream.ctypes.data_as(_u8p), len(stream), out.ctypes.data_as(_u8p), n)
    return r, out[:n]


def _inputs():
    rng = np.random.default_rng(4242)
    ins = []
    for n in (1, 2, 3, 7, 15, 16, 17, 31, 32, 33, 47, 48, 49, 64, 65, 100, 255, 256, 257, 1000, 4095, 4096, 4097, 9000):
        ins.append(gen_fuzz(rng, n))
    ins.append(gen_fuzz(rng, 150000, long_every=5))
    ins.append(gen_dct(200000, seed=9))
    ins.append(gen_short_runs(120000, seed=3, W=1))
    ins.append(gen_short_runs(120000, seed=4, W=4))
    ins.append(gen_run_mixed(300000, seed=2, max_run_log2=14, max_lit_log2=13))
    ins.append(np.zeros(70000, dtype=np.uint8))
    ins.append(np.tile(np.array([1, 2], dtype=np.uint8), 5000))
    ins.append(np.tile(np.array([1, 1, 2, 2], dtype=np.uint8), 5000))
    ins.append(rng.integers(0, 256, size=50000, dtype=np.uint8))
    return ins


@pytest.mark.parametrize("codec", CODECS, ids=lambda c: c.name)
def test_stage_pipeline_matches_oracle(sim, codec):
    for data in _inputs():
        want = oracle_compress(codec, data)
        for rounds, maxit in ((3, 6), (0, 6), (3, 0), (0, 1)):   # 0: force the sequential repair / in-CTA sequential pass
            got = _enc(sim, codec, data, rounds, maxit)
            assert np.array_equal(got, want), f"{codec.name}: staged encoder differs, n={len(data)} rounds={rounds} maxit={maxit}"
        r, dec = _dec(sim, codec, want, len(data))
        assert r == len(data) and np.array_equal(dec, data), f"{codec.name}: staged decoder differs, n={len(data)}"


def test_stage_decoder_single_mode(sim, golden_small):
    from common import CODEC_BY_NAME
    for k in [k for k in golden_small.files if k.startswith("single_in__")]:
        _, nm, i = k.split("__")
        data = golden_small[k]
        stream = golden_small[f"single_out__{nm}__{i}"]
        codec = CODEC_BY_NAME["rle8_packed_multi" if "packed" in nm else "rle8_multi"]
        r, dec = _dec(sim, codec, stream, len(data))
        assert r == len(data) and np.array_equal(dec, data)
