"""The drop-in harness: the reference's own benchmark / test driver (src/main.c, codecCallbacks[] of
src/codec_funcs.h:262-410) linked against libhsrle_b200.so in place of the reference's objects for the 90 replaced entry
points (oracle/Makefile target `dropin`).  CPU part: the link recipe actually works and every
replaced symbol resolves to the product library; without a CUDA device the product fails loudly (no CPU fallback).  GPU
part: `hsrlekit --extreme --test` drives the sm_100a kernels through the reference's table and validates every round trip."""
import os
import re
import subprocess

import numpy as np
import pytest

from common import ORACLE_DIR, ROOT, gen_dct, gen_fuzz

EXE = os.path.join(ORACLE_DIR, "_ref", "hsrlekit_b200")
LIB = os.path.join(ROOT, "hypersonic-rle-kit_b200", "libhsrle_b200.so")


def _harness():
    if os.path.isdir("/root/reference/src"):
        newest_lib = os.path.getmtime(LIB)
        if not os.path.exists(EXE) or os.path.getmtime(EXE) < newest_lib:
            subprocess.run(["make", "-s", "-C", ORACLE_DIR, "dropin"], check=True)
    if not os.path.exists(EXE):
        pytest.skip("drop-in harness not built (needs /root/reference at build time)")
    return EXE


def test_replaced_symbols_resolve_to_the_product_library():
    exe = _harness()
    und = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True, check=True).stdout
    names = set(re.findall(r"\bU (rle\w+)", und))
    assert len(names) == 90, sorted(names)                      # 44 codec pairs + rle_compress_bounds + rle_decompress_additional_size
    assert {"rle8_multi_compress", "rle8_decompress", "rle24_7symlut_sym_decompress", "rle64_byte_packed_compress", "rle_compress_bounds"} <= names
    exported = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True, check=True).stdout
    have = set(re.findall(r"\bT (rle\w+)", exported))
    assert names <= have
    # the reference's own definitions of those names were renamed away, its out-of-scope codecs are still in the binary
    syms = subprocess.run(["nm", exe], capture_output=True, text=True, check=True).stdout
    assert " T ref_rle8_multi_compress" in syms and re.search(r" T rle8_single_compress\b", syms)
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libhsrle_b200.so" in ldd and "not found" not in ldd.split("libhsrle_b200.so")[1].split("\n")[0]


def test_without_a_gpu_the_harness_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    exe = _harness()
    f = tmp_path / "in.raw"
    gen_dct(50000, seed=2).tofile(f)
    r = subprocess.run([exe, str(f), "--extreme", "--test", "--runs", "0", "--min-time", "0"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "FAILED" in r.stdout          # the very first codec call returns 0: there is no CPU path


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["dct", "fuzz"])
def test_reference_harness_drives_the_gpu_codecs(tmp_path, kind):
    exe = _harness()
    rng = np.random.default_rng(3)
    data = gen_dct(3 << 20, seed=8) if kind == "dct" else gen_fuzz(rng, 400000, long_every=9)
    f = tmp_path / "in.raw"
    data.tofile(f)
    r = subprocess.run([exe, str(f), "--extreme", "--test", "--runs", "0", "--min-time", "0"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "FAILED" not in r.stdout
    for label in ("8 Bit ", "8 Bit Packed", "64 Bit 7LUT (Byte)", "24 Bit Packed (Symbol)"):
        assert label in r.stdout, (label, r.stdout[-2000:])
