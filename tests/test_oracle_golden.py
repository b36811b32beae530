"""CPU tests: the oracle (oracle/rle_oracle.c) against the committed golden vectors that were
generated from the compiled reference (tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from common import CODECS, oracle_compress, oracle_decompress
from golden.make_golden import large_inputs


@pytest.mark.parametrize("codec", CODECS, ids=lambda c: c.name)
def test_oracle_matches_golden_small(codec, golden_small):
    names = sorted(k[4:] for k in golden_small.files if k.startswith("in__"))
    assert len(names) >= 9
    for k in names:
        data = golden_small["in__" + k]
        want = golden_small[f"out__{k}__{codec.name}"]
        got = oracle_compress(codec, data)
        assert np.array_equal(got, want), f"{codec.name} encoder differs from reference on {k}"
        r, dec = oracle_decompress(codec, want, len(data))
        assert r == len(data) and np.array_equal(dec, data), f"{codec.name} decoder fails on {k}"


@pytest.fixture(scope="module")
def large():
    return large_inputs()


@pytest.mark.parametrize("codec", CODECS, ids=lambda c: c.name)
def test_oracle_matches_golden_hashes(codec, golden_hashes, large):
    for k, v in large.items():
        h = golden_hashes[k]
        assert hashlib.sha256(v.tobytes()).hexdigest() == h["input_sha256"], "generator drifted: " + k
        got = oracle_compress(codec, v)
        assert len(got) == h["streams"][codec.name]["len"]
        assert hashlib.sha256(got.tobytes()).hexdigest() == h["streams"][codec.name]["sha256"], (codec.name, k)
        r, dec = oracle_decompress(codec, got, len(v))
        assert r == len(v) and np.array_equal(dec, v)


def test_oracle_decodes_golden_single_mode_streams(golden_small):
    """mode-1 streams (src/rle8_extreme_cpu.h:736-757) made by the reference's single encoders."""
    from common import CODEC_BY_NAME
    keys = [k for k in golden_small.files if k.startswith("single_in__")]
    assert len(keys) == 8
    for k in keys:
        _, nm, i = k.split("__")
        data = golden_small[k]
        stream = golden_small[f"single_out__{nm}__{i}"]
        codec = CODEC_BY_NAME["rle8_packed_multi" if "packed" in nm else "rle8_multi"]
        r, dec = oracle_decompress(codec, stream, len(data))
        assert r == len(data) and np.array_equal(dec, data)
