"""CPU tests: differential testing of the oracle against the compiled, unmodified reference
(oracle/_ref/libhsrle_ref.so).  Skipped when the reference has not been built (no /root/reference
and no prebuilt _ref/)."""
import numpy as np
import pytest

from common import (CODECS, CODEC_BY_NAME, gen_fuzz, oracle_compress, oracle_decompress, ref_compress, ref_decompress,
                    ref_lib)

pytestmark = pytest.mark.skipif(ref_lib() is None, reason="compiled reference not available")

SIZES = [1, 2, 3, 5, 8, 15, 16, 17, 31, 32, 33, 34, 40, 63, 64, 65, 66, 100, 255, 256, 257, 300, 1000, 5000, 20000]


@pytest.mark.parametrize("codec", CODECS, ids=lambda c: c.name)
def test_differential_fuzz(codec):
    rng = np.random.default_rng(hash(codec.name) & 0xFFFF)
    for it in range(120):
        n = SIZES[it % len(SIZES)] if it % 2 else int(rng.integers(1, 3000))
        data = gen_fuzz(rng, n)
        a = ref_compress(codec, data)
        b = oracle_compress(codec, data)
        assert len(a) > 0
        assert np.array_equal(a, b), f"{codec.name}: encoder mismatch n={n} it={it}"
        r, d = oracle_decompress(codec, a, n)
        assert r == n and np.array_equal(d, data)
        r, d = ref_decompress(codec, b, n)
        assert r == n and np.array_equal(d, data)


@pytest.mark.parametrize("codec", CODECS, ids=lambda c: c.name)
def test_differential_field_boundaries(codec):
    """Sections around the 8-bit / 16-bit field switch points (src/rle_fuzz.c:30-44)."""
    rng = np.random.default_rng(5)
    for ln in (250, 254, 255, 256, 257, 260, 65530, 65535, 65536, 65540, 70000):
        for lit in (0, 1, 126, 127, 128, 129, 254, 255, 256, 300, 65534, 65535, 65536, 66000):
            if ln > 60000 and lit > 60000 and lit != 65536:
                continue
            sym = rng.integers(0, 256, size=codec.W, dtype=np.uint8)
            run = np.tile(sym, -(-ln // codec.W))[:ln]
            data = np.concatenate([rng.integers(0, 256, size=7, dtype=np.uint8), run,
                                   rng.integers(0, 256, size=lit, dtype=np.uint8), run[: 40],
                                   rng.integers(0, 256, size=3, dtype=np.uint8)])
            a = ref_compress(codec, data)
            b = oracle_compress(codec, data)
            assert np.array_equal(a, b), f"{codec.name}: ln={ln} lit={lit}"
            r, d = oracle_decompress(codec, a, len(data))
            assert r == len(data) and np.array_equal(d, data)


def test_single_mode_streams_decode():
    """rle8_decompress / rle8_packed_decompress also accept mode-1 (single symbol) streams
    (src/rle8_extreme_cpu.h:736-757); the single *encoders* are out of scope, so use the reference's."""
    import ctypes
    lib = ref_lib()
    u8p = ctypes.POINTER(ctypes.c_uint8)
    rng = np.random.default_rng(3)
    for nm, cn in (("rle8_single_compress", "rle8_multi"), ("rle8_packed_single_compress", "rle8_packed_multi")):
        f = getattr(lib, nm)
        f.restype = ctypes.c_uint32
        f.argtypes = [u8p, ctypes.c_uint32, u8p, ctypes.c_uint32]
        for it in range(40):
            n = int(rng.integers(1, 5000))
            data = gen_fuzz(rng, n, max_sym=1, p_run=0.6)
            data[rng.random(n) < 0.5] = 0
            buf = np.zeros(n + 64, dtype=np.uint8)
            buf[:n] = data
            out = np.zeros(n + 1024, dtype=np.uint8)
            r = f(buf.ctypes.data_as(u8p), n, out.ctypes.data_as(u8p), len(out))
            assert r > 0
            rr, d = oracle_decompress(CODEC_BY_NAME[cn], out[:r], n)
            assert rr == n and np.array_equal(d, data), (nm, it, n, int(out[8]))
