"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the product's C ABI
(libhsrle_b200.so); the checkers are the oracle (oracle/liboracle.so), the committed golden vectors
made from the compiled reference, and -- when the prebuilt oracle/_ref travelled along -- the
compiled reference itself."""
import hashlib

import numpy as np
import pytest

from common import (CODECS, CODEC_BY_NAME, gen_dct, gen_fuzz, gen_run_mixed, gen_short_runs, oracle_compress,
                    oracle_decompress, out_capacity, ref_compress, ref_decompress, ref_lib)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hs():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import hsrle_b200
    assert hsrle_b200.lib.hsrle_device() >= 0, hsrle_b200.last_error()
    return hsrle_b200


def gpu_enc(hs, c, data):
    return hs.compress(c.cname, data, out_capacity(len(data)))


def gpu_dec(hs, c, stream, n):
    return hs.decompress(c.dname, stream, n)


@pytest.mark.parametrize("codec", CODECS, ids=lambda c: c.name)
def test_golden_small(hs, codec, golden_small):
    for k in sorted(k[4:] for k in golden_small.files if k.startswith("in__")):
        data = golden_small["in__" + k]
        want = golden_small[f"out__{k}__{codec.name}"]
        got = gpu_enc(hs, codec, data)
        assert np.array_equal(got, want), f"{codec.name}: GPU stream differs from reference on {k} ({hs.last_error()})"
        r, dec = gpu_dec(hs, codec, want, len(data))
        assert r == len(data) and np.array_equal(dec, data), f"{codec.name}: GPU decode fails on {k}"


@pytest.fixture(scope="module")
def large():
    from golden.make_golden import large_inputs
    return large_inputs()


@pytest.mark.parametrize("codec", CODECS, ids=lambda c: c.name)
def test_golden_hashes(hs, codec, golden_hashes, large):
    for k, v in large.items():
        h = golden_hashes[k]["streams"][codec.name]
        got = gpu_enc(hs, codec, v)
        assert len(got) == h["len"], (codec.name, k, len(got), h["len"])
        assert hashlib.sha256(got.tobytes()).hexdigest() == h["sha256"], (codec.name, k)
        r, dec = gpu_dec(hs, codec, got, len(v))
        assert r == len(v) and np.array_equal(dec, v), (codec.name, k)


SIZES = [1, 2, 3, 5, 8, 15, 16, 17, 31, 32, 33, 34, 40, 47, 48, 49, 63, 64, 65, 66, 100, 255, 256, 257, 300, 1000,
         4095, 4096, 4097, 5000, 20000]


@pytest.mark.parametrize("codec", CODECS, ids=lambda c: c.name)
def test_fuzz_vs_oracle(hs, codec):
    rng = np.random.default_rng(hash(codec.name) & 0xFFFF)
    for it in range(60):
        n = SIZES[it % len(SIZES)] if it % 2 else int(rng.integers(1, 3000))
        data = gen_fuzz(rng, n)
        want = oracle_compress(codec, data)
        got = gpu_enc(hs, codec, data)
        assert np.array_equal(got, want), f"{codec.name}: n={n} it={it}"
        r, dec = gpu_dec(hs, codec, want, n)
        assert r == n and np.array_equal(dec, data), f"{codec.name}: decode n={n} it={it}"


@pytest.mark.parametrize("codec", CODECS, ids=lambda c: c.name)
def test_edge_inputs(hs, codec):
    """configs[4]: all-random, single symbol, alternating 1-byte / 2-byte runs, 1 KB .. 1 MB."""
    rng = np.random.default_rng(17)
    for n in (1024, 65536, 1 << 20):
        for data in (rng.integers(0, 256, size=n, dtype=np.uint8), np.full(n, 0xAB, dtype=np.uint8),
                     np.tile(np.array([1, 2], dtype=np.uint8), n // 2), np.tile(np.array([1, 1, 2, 2], dtype=np.uint8), n // 4)):
            want = oracle_compress(codec, data)
            got = gpu_enc(hs, codec, data)
            assert np.array_equal(got, want), (codec.name, n)
            r, dec = gpu_dec(hs, codec, got, n)
            assert r == n and np.array_equal(dec, data), (codec.name, n)


@pytest.mark.parametrize("name", ["rle8_multi", "rle8_packed_multi", "rle8_3symlut", "rle16_sym", "rle24_byte_packed",
                                  "rle32_7symlut_sym", "rle48_byte", "rle64_byte_packed", "rle64_3symlut_byte"])
def test_structured_streams(hs, name):
    """DCT-shaped, short-run-heavy and run-mixed streams at a few MiB."""
    codec = CODEC_BY_NAME[name]
    for data in (gen_dct(3 << 20, seed=21), gen_short_runs(2 << 20, seed=22, W=codec.W), gen_short_runs(2 << 20, seed=23, W=1),
                 gen_run_mixed(4 << 20, seed=24)):
        want = oracle_compress(codec, data)
        got = gpu_enc(hs, codec, data)
        assert len(got) == len(want) and np.array_equal(got, want), (name, len(got), len(want))
        r, dec = gpu_dec(hs, codec, got, len(data))
        assert r == len(data) and np.array_equal(dec, data)


@pytest.mark.skipif(ref_lib() is None, reason="prebuilt compiled reference not available")
@pytest.mark.parametrize("codec", CODECS, ids=lambda c: c.name)
def test_cross_decode_with_compiled_reference(hs, codec):
    """reference decodes GPU streams; GPU decodes reference streams."""
    rng = np.random.default_rng(99)
    for n in (777, 30000, 250000):
        data = gen_fuzz(rng, n, long_every=11)
        a = ref_compress(codec, data)
        b = gpu_enc(hs, codec, data)
        assert np.array_equal(a, b), (codec.name, n)
        r, d = ref_decompress(codec, b, n)
        assert r == n and np.array_equal(d, data)
        r, d = gpu_dec(hs, codec, a, n)
        assert r == n and np.array_equal(d, data)


def test_single_mode_streams(hs, golden_small):
    for k in [k for k in golden_small.files if k.startswith("single_in__")]:
        _, nm, i = k.split("__")
        data = golden_small[k]
        stream = golden_small[f"single_out__{nm}__{i}"]
        codec = CODEC_BY_NAME["rle8_packed_multi" if "packed" in nm else "rle8_multi"]
        r, dec = gpu_dec(hs, codec, stream, len(data))
        assert r == len(data) and np.array_equal(dec, data), k


def test_error_conventions(hs):
    """return 0 on bad arguments / headers (src/rle8_extreme_cpu.h:88,704-712,759-760)."""
    c = CODEC_BY_NAME["rle8_multi"]
    data = gen_dct(5000, seed=1)
    assert len(hs.compress(c.cname, data, out_size=len(data))) == 0          # outSize < rle_compress_bounds
    s = gpu_enc(hs, c, data)
    assert gpu_dec(hs, c, s, len(data) - 1)[0] == 0                           # uncompressedLength > outSize
    assert gpu_dec(hs, c, s[:-5], len(data))[0] == 0                          # compressedLength > inSize
    bad = s.copy(); bad[8] = 7
    assert gpu_dec(hs, c, bad, len(data))[0] == 0                             # unknown mode
    # the decoder must not depend on bytes after compressedLength (src/rle_fuzz.c:628-633)
    padded = np.concatenate([s, np.full(300, 0xFF, dtype=np.uint8)])
    r, d = gpu_dec(hs, c, padded, len(data))
    assert r == len(data) and np.array_equal(d, data)


def test_device_resident_and_async_paths(hs):
    import torch
    dev = torch.device("cuda:0")
    for name in ("rle8_multi", "rle32_byte_packed", "rle16_7symlut_sym"):
        codec = CODEC_BY_NAME[name]
        data = gen_dct(1 << 20, seed=5)
        want = oracle_compress(codec, data)
        t_in = torch.from_numpy(data).to(dev)
        t_out = torch.empty(out_capacity(len(data)), dtype=torch.uint8, device=dev)
        r = hs.compress_device(name, t_in, t_out)
        assert r == len(want) and np.array_equal(t_out[:r].cpu().numpy(), want)
        # async: caller-owned workspace, current stream, result read back afterwards
        ws = torch.empty(hs.compress_workspace_size(name, len(data)), dtype=torch.uint8, device=dev)
        res = torch.zeros(8, dtype=torch.int32, device=dev)
        t_out.zero_()
        hs.compress_device_async(name, t_in, t_out, ws, res, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        rr = res.cpu().numpy()
        assert rr[0] == len(want) and rr[1] == 0
        assert np.array_equal(t_out[: rr[0]].cpu().numpy(), want)
        t_dec = torch.empty(len(data) + 128, dtype=torch.uint8, device=dev)
        dws = torch.empty(hs.decompress_workspace_size(name, int(rr[0]), len(data)), dtype=torch.uint8, device=dev)
        hs.decompress_device_async(name, t_out, int(rr[0]), t_dec, len(data), dws, res, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        rr = res.cpu().numpy()
        assert rr[0] == len(data) and rr[1] == 0
        assert np.array_equal(t_dec[: len(data)].cpu().numpy(), data)


@pytest.fixture(scope="module")
def full88():
    """The bench input (BASELINE configs[0,1]): 88,473,600-byte DCT stream, resident on the device once per module."""
    import json
    import os
    import torch
    n = 88473600
    data = gen_dct(n)
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_hashes_88m.json")) as f:
        gold = json.load(f)
    assert gold["n"] == n and hashlib.sha256(data.tobytes()).hexdigest() == gold["input_sha256"]
    return data, torch.from_numpy(data).to(torch.device("cuda:0")), gold["streams"]


@pytest.mark.parametrize("codec", CODECS, ids=lambda c: c.name)
def test_full_size_88mb_all_codecs(hs, codec, full88):
    """BASELINE configs[0,1] at full size, every codec: the GPU stream is byte-identical to the compiled reference's
    (length + sha256 from tests/golden/golden_hashes_88m.json, made by make_golden.py from oracle/_ref) and decodes back."""
    import torch
    data, t_in, gold = full88
    n = len(data)
    dev = t_in.device
    t_out = torch.empty(out_capacity(n), dtype=torch.uint8, device=dev)
    r = hs.compress_device(codec.name, t_in, t_out)
    assert r == gold[codec.name]["len"], (codec.name, r, gold[codec.name]["len"], hs.last_error())
    assert hashlib.sha256(t_out[:r].cpu().numpy().tobytes()).hexdigest() == gold[codec.name]["sha256"], codec.name
    t_dec = torch.empty(n + 128, dtype=torch.uint8, device=dev)
    rd = hs.decompress_device(codec.name, t_out, r, t_dec, n)
    assert rd == n, (codec.name, hs.last_error())
    assert torch.equal(t_dec[:n], t_in)


@pytest.mark.parametrize("name", ["rle8_multi", "rle8_packed_multi", "rle64_byte_packed", "rle24_3symlut_sym"])
def test_full_size_88mb_vs_oracle(hs, name, full88):
    """Same size against the oracle port (the second, independent checker)."""
    import torch
    codec = CODEC_BY_NAME[name]
    data, t_in, _ = full88
    n = len(data)
    want = oracle_compress(codec, data)
    t_out = torch.empty(out_capacity(n), dtype=torch.uint8, device=t_in.device)
    r = hs.compress_device(name, t_in, t_out)
    assert r == len(want), (r, len(want), hs.last_error())
    assert torch.equal(t_out[:r].cpu(), torch.from_numpy(want))


def _frame_input(kind, n, dev):
    import torch
    g = torch.Generator(device=dev); g.manual_seed(77)
    if kind == "single_symbol":
        return torch.full((n,), 0x5A, dtype=torch.uint8, device=dev)
    if kind == "random":
        return torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev, generator=g)
    if kind == "alternating":
        return torch.arange(n, dtype=torch.int32, device=dev).bitwise_and_(1).to(torch.uint8)     # 1-byte runs: no candidates for W = 1
    # runs of 1 .. 4096 equal bytes: segment id = cumulative sum of "a new run starts here" flags
    starts = torch.rand(n, device=dev, generator=g) < (1.0 / 37.0)
    seg = torch.cumsum(starts.to(torch.int32), 0)
    del starts
    t_in = (seg.to(torch.int64) * 2654435761 % 251).to(torch.uint8)
    del seg
    return t_in


def _check_frame(hs, name, t_in, kind=None):
    """Encode one frame on the device, compare the stream with the compiled reference's (oracle/_ref, ~1 s per GiB on the
    host), check the header, decode back.  Returns the stream length."""
    import torch
    codec = CODEC_BY_NAME[name]
    n = t_in.numel()
    dev = t_in.device
    cap = n + n // 256 + 512
    t_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    r = hs.compress_device(name, t_in, t_out)
    assert r > 0, hs.last_error()
    got = t_out[:r].cpu().numpy()
    assert int.from_bytes(got[0:4].tobytes(), "little") == n and int.from_bytes(got[4:8].tobytes(), "little") == r
    if codec.W == 1 and codec.variant in (0, 1):
        assert got[8] == 0                                   # multi mode
    if kind == "single_symbol":
        assert r < 64                                        # header + one token + terminator
    if kind == "random":
        assert n <= r <= n + n // 256 + 64                   # (almost) one literal; a few accidental short runs at most
    if ref_lib() is not None:
        want = ref_compress(codec, t_in.cpu().numpy())
        assert len(want) == r and np.array_equal(got, want), (name, kind, r, len(want))
        del want
    del got
    t_dec = torch.empty(n + 128, dtype=torch.uint8, device=dev)
    rd = hs.decompress_device(name, t_out, r, t_dec, n)
    assert rd == n, hs.last_error()
    assert torch.equal(t_dec[:n], t_in)
    return r


@pytest.mark.parametrize("name", ["rle8_multi", "rle64_byte_packed", "rle32_3symlut_byte"])
@pytest.mark.parametrize("kind", ["single_symbol", "random", "alternating", "run_mixed"])
def test_one_gib_frame_vs_reference(hs, name, kind):
    """BASELINE configs[4] at the frame size (2^30 bytes, the largest input rle_compress_bounds accepts,
    src/rle8_extreme_cpu.c:24-25): the stream must equal the compiled reference's byte for byte (oracle/_ref travels
    to the GPU box), plus header fields, the size bounds the format implies and the round trip."""
    import torch
    dev = torch.device("cuda:0")
    _check_frame(hs, name, _frame_input(kind, 1 << 30, dev), kind)


@pytest.mark.parametrize("name", ["rle8_multi", "rle64_byte"])
def test_four_gib_single_symbol_as_frames(hs, name):
    """BASELINE configs[4] "single-symbol 4 GiB (one run)": above 2^30 bytes rle_compress_bounds returns 0
    (src/rle8_extreme_cpu.c:22-28) and the header fields are u32, so a caller of the u32 API cuts the input into frames
    (hsrle_b200.sliced.frame_bounds); every frame is a complete reference-identical stream."""
    import torch
    from hsrle_b200.sliced import frame_bounds
    dev = torch.device("cuda:0")
    total = 4 << 30
    t_in = torch.full((total,), 0x5A, dtype=torch.uint8, device=dev)
    frames = frame_bounds(total)
    assert len(frames) == 4 and all(b - a == 1 << 30 for a, b in frames)
    sizes = [_check_frame(hs, name, t_in[a:b], "single_symbol") for a, b in frames]
    assert len(set(sizes)) == 1


@pytest.mark.parametrize("name", ["rle8_packed_multi", "rle48_byte", "rle16_7symlut_sym"])
@pytest.mark.parametrize("n", [(1 << 30) - 1, 1 << 30])
def test_frame_size_limits_vs_reference(hs, name, n):
    """The API ceiling itself: frames of exactly 2^30 and 2^30 - 1 bytes (run-mixed content, the last byte inside a run)."""
    import torch
    dev = torch.device("cuda:0")
    t_in = _frame_input("run_mixed", 1 << 30, dev)[:n].clone()
    _check_frame(hs, name, t_in, "run_mixed")


def test_two_devices_in_one_process(hs):
    """Kernel attributes (dynamic shared memory) are per device: encode + decode on cuda:0, then on cuda:1."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    data = gen_dct(1 << 20, seed=9)
    for d in (0, 1):
        with torch.cuda.device(d):
            dev = torch.device("cuda", d)
            for name in ("rle8_multi", "rle32_3symlut_byte"):
                want = oracle_compress(CODEC_BY_NAME[name], data)
                t_in = torch.from_numpy(data).to(dev)
                t_out = torch.empty(out_capacity(len(data)), dtype=torch.uint8, device=dev)
                r = hs.compress_device(name, t_in, t_out)
                assert r == len(want) and np.array_equal(t_out[:r].cpu().numpy(), want), (d, name, hs.last_error())
                t_dec = torch.empty(len(data) + 128, dtype=torch.uint8, device=dev)
                assert hs.decompress_device(name, t_out, r, t_dec, len(data)) == len(data), (d, name, hs.last_error())
                assert torch.equal(t_dec[: len(data)], t_in)


@pytest.mark.parametrize("name", ["rle8_multi", "rle8_packed_multi", "rle16_7symlut_byte", "rle48_3symlut_sym", "rle64_byte"])
def test_workspace_contents_do_not_matter(hs, name):
    """The caller-owned workspace of the async entry points may hold anything (another call's tables, another codec's
    state): only the region the library clears itself is assumed zero."""
    import torch
    dev = torch.device("cuda:0")
    codec = CODEC_BY_NAME[name]
    data = gen_dct(3 << 20, seed=11)
    want = oracle_compress(codec, data)
    n = len(data)
    cap = out_capacity(n)
    t_in = torch.from_numpy(data).to(dev)
    ws = torch.empty(max(hs.compress_workspace_size(name, n), hs.decompress_workspace_size(name, cap, n)), dtype=torch.uint8, device=dev)
    res = torch.zeros(16, dtype=torch.int32, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    for fill in (0xFF, 0x01, 0x80):
        ws.fill_(fill)
        t_out = torch.zeros(cap, dtype=torch.uint8, device=dev)
        hs.compress_device_async(name, t_in, t_out, ws, res[:8], sp)
        torch.cuda.synchronize()
        r = res[:8].tolist()
        assert r[0] == len(want) and r[1] == 0, (fill, r)
        assert np.array_equal(t_out[: r[0]].cpu().numpy(), want)
        ws.fill_(fill)
        t_dec = torch.zeros(n + 128, dtype=torch.uint8, device=dev)
        hs.decompress_device_async(name, t_out, r[0], t_dec, n, ws, res[8:], sp)
        torch.cuda.synchronize()
        rd = res[8:].tolist()
        assert rd[0] == n and rd[1] == 0, (fill, rd)
        assert torch.equal(t_dec[:n], t_in)


@pytest.mark.parametrize("name", ["rle8_multi", "rle8_packed_multi", "rle16_byte_packed", "rle32_3symlut_byte", "rle64_7symlut_sym"])
def test_long_literals_at_stream_start(hs, name):
    """D1's scout follows the true chain from the stream start while its tokens jump over whole super-chunks and gives the
    super-chunks in between constant tables: streams that begin with 0..9 such tokens (literals of 64 KiB and more),
    followed by ordinary tokens or by nothing, must decode exactly and the encoder must still match the oracle."""
    codec = CODEC_BY_NAME[name]
    rng = np.random.default_rng(2024)
    W = codec.W
    for nlong, tail in ((0, "dct"), (1, "dct"), (3, "dct"), (9, "dct"), (2, "none"), (1, "run"), (10, "short")):
        parts = []
        for k in range(nlong):
            parts.append(rng.integers(0, 256, size=int(rng.integers(65536 + 20, 200000)), dtype=np.uint8))
            parts.append(np.tile(rng.integers(0, 256, size=W, dtype=np.uint8), 40 + k))          # a run every codec emits
        if tail == "dct":
            parts.append(gen_dct(300000, seed=int(rng.integers(1, 1000))))
        elif tail == "run":
            parts.append(np.tile(rng.integers(0, 256, size=W, dtype=np.uint8), 50000))
        elif tail == "short":
            parts.append(gen_short_runs(100000, seed=3, W=W))
        else:
            parts.append(rng.integers(0, 256, size=70000, dtype=np.uint8))                          # final token with a long literal
        data = np.concatenate(parts)
        want = oracle_compress(codec, data)
        got = gpu_enc(hs, codec, data)
        assert np.array_equal(got, want), (name, nlong, tail)
        r, dec = gpu_dec(hs, codec, want, len(data))
        assert r == len(data) and np.array_equal(dec, data), (name, nlong, tail)


@pytest.mark.parametrize("name", ["rle8_multi", "rle24_byte_packed", "rle32_sym", "rle48_byte", "rle64_3symlut_sym"])
def test_many_huge_literals_in_mid_stream(hs, name):
    """A stream that is mostly literal bytes with a token every 2 KiB .. 1 MiB (what a wide-symbol codec makes of a
    byte-run stream): the encoder's grid-wide literal copy deals the 16-KiB pieces of all huge literals round-robin over
    the CTAs (hundreds of literals: the rotation wraps around the grid), and the decoder's segment composition (D2) keeps
    taking batched hops while many speculative chains enter super-chunk after super-chunk beyond the window."""
    codec = CODEC_BY_NAME[name]
    rng = np.random.default_rng(77)
    W = codec.W
    parts = []
    total = 0
    while total < 40 << 20:
        ll = int(2 ** rng.uniform(11, 20))
        parts.append(rng.integers(0, 256, size=ll, dtype=np.uint8))
        rl = int(rng.integers(12, 60))
        parts.append(np.tile(rng.integers(0, 256, size=W, dtype=np.uint8), rl))
        total += ll + rl * W
    parts.append(gen_short_runs(200000, seed=5, W=1))
    data = np.concatenate(parts)
    want = oracle_compress(codec, data)
    got = gpu_enc(hs, codec, data)
    assert np.array_equal(got, want), name
    r, dec = gpu_dec(hs, codec, want, len(data))
    assert r == len(data) and np.array_equal(dec, data), name


@pytest.mark.timeout(600)
@pytest.mark.parametrize("name", ["rle8_multi", "rle8_packed_multi", "rle8_3symlut", "rle16_sym_packed", "rle24_7symlut_byte", "rle32_byte",
                                  "rle48_3symlut_sym", "rle64_byte_packed"])
def test_corrupt_token_streams_fail_safely(hs, name):
    """The reference trusts the stream (corrupt input = undefined behaviour); the library must not: bit-flipped
    and truncated token bodies behind a VALID header either decode to exactly `uncompressedLength` bytes or return 0 -- within
    the test timeout (no spin-wait may hang on a broken chain) and without writing a byte past the declared output size."""
    import torch
    dev = torch.device("cuda:0")
    codec = CODEC_BY_NAME[name]
    rng = np.random.default_rng(4242)
    for data in (gen_dct(300000, seed=31), gen_fuzz(rng, 70000, long_every=13), gen_short_runs(120000, seed=33, W=codec.W)):
        n = len(data)
        good = gpu_enc(hs, codec, data)
        variants = []
        for k in range(10):
            bad = good.copy()
            nflip = int(rng.integers(1, 6))
            pos = rng.integers(codec.hdr + 1, len(bad), size=nflip)
            bad[pos] ^= (1 << rng.integers(0, 8, size=nflip)).astype(np.uint8)
            variants.append(bad)
        cut = good[: max(codec.hdr + 2, len(good) - int(rng.integers(1, 400)))].copy()
        cut[4:8] = np.frombuffer(np.uint32(len(cut)).tobytes(), dtype=np.uint8)       # truncated, header made consistent
        variants.append(cut)
        junk = good.copy()
        junk[codec.hdr + 1:] = rng.integers(0, 256, size=len(junk) - codec.hdr - 1, dtype=np.uint8)
        variants.append(junk)
        for bad in variants:
            t_s = torch.from_numpy(bad).to(dev)
            t_dec = torch.full((n + 256,), 0xEE, dtype=torch.uint8, device=dev)
            r = hs.decompress_device(name, t_s, len(bad), t_dec, n)
            assert r in (0, n), (name, r)
            assert bool((t_dec[n:] == 0xEE).all()), name + ": wrote past the declared output size"


@pytest.mark.timeout(600)
@pytest.mark.parametrize("name", ["rle8_multi", "rle8_7symlut", "rle16_3symlut_byte", "rle32_byte_packed", "rle48_7symlut_sym", "rle64_byte",
                                  "rle64_3symlut_byte"])
def test_segment_table_mode_of_the_decoder(hs, name):
    """A 48 MiB stream of literals of 256 B .. 8 KiB between short runs: the mean token is between 256 B and a chunk and the stream
    has more segments than the GPU has SMs, so k_dec_map's composers build the all-position segment tables and the resolver
    takes one look-up per segment (hsrle_dec_kernels.cuh: S.mode == 1) -- with and without a LUT in the aggregate."""
    codec = CODEC_BY_NAME[name]
    rng = np.random.default_rng(123)
    W = codec.W
    parts, total = [], 0
    while total < 48 << 20:
        ll = int(2 ** rng.uniform(8, 13))
        parts.append(rng.integers(0, 256, size=ll, dtype=np.uint8))
        rl = int(rng.integers(4, 40))
        parts.append(np.tile(rng.integers(0, 256, size=W, dtype=np.uint8), rl))
        total += ll + rl * W
    data = np.concatenate(parts)
    want = oracle_compress(codec, data)
    assert len(want) > 40 << 20                                         # >= 148 segments of 256 KiB
    got = gpu_enc(hs, codec, data)
    assert np.array_equal(got, want), name
    r, dec = gpu_dec(hs, codec, want, len(data))
    assert r == len(data) and np.array_equal(dec, data), name


@pytest.mark.parametrize("name", ["rle8_3symlut", "rle8_7symlut"])
def test_lut_walk_fallbacks(hs, name):
    """The 8-bit LUT encoders take the table at every super-chunk start from the stretch walk (hsrle_enc_lutwalk.cuh) -- unless the
    input has more than LW_CAP symbol changes between candidate runs, or a stretch of more than LW_BACK records none of which
    is certain.  Then the plain state guess and the verify / repair rounds alone must produce the same stream."""
    codec = CODEC_BY_NAME[name]
    rng = np.random.default_rng(31)
    many_symbols = gen_fuzz(rng, 6 << 20, max_sym=40, p_run=0.3)         # ~80 K stretch boundaries: over the cap
    noise = (np.arange(300, dtype=np.uint32) * 7 % 251 + 1).astype(np.uint8)   # no two equal neighbours
    piece = np.concatenate([noise, np.full(3, 0x61, dtype=np.uint8)])           # a 3-byte run every 303 bytes: never emitted
    long_uncertain = np.concatenate([np.tile(piece, 4000), gen_dct(1 << 20, seed=8), np.tile(piece, 300)])
    sparse = gen_dct(5 << 20, seed=9)                                     # the walk's own territory
    for data in (many_symbols, long_uncertain, sparse):
        want = oracle_compress(codec, data)
        got = gpu_enc(hs, codec, data)
        assert len(got) == len(want) and np.array_equal(got, want), (name, len(got), len(want))
        r, dec = gpu_dec(hs, codec, got, len(data))
        assert r == len(data) and np.array_equal(dec, data)


@pytest.mark.parametrize("codec", CODECS, ids=lambda c: c.name)
def test_decoder_exit_composition_either_way(hs, codec, monkeypatch):
    """k_dec_map composes the exits of a segment as window rows or as all-position segment tables (chosen per call from the stream);
    HSRLE_DEC_MODE forces one: every codec must decode the same streams to the same bytes either way."""
    rng = np.random.default_rng(2024 + codec.W)
    datas = [gen_fuzz(rng, 300000, long_every=5), gen_run_mixed(1 << 20, seed=3, max_run_log2=14, max_lit_log2=12), gen_dct(600000, seed=11)]
    for data in datas:
        stream = oracle_compress(codec, data)
        for mode in ("rows", "segtab"):
            monkeypatch.setenv("HSRLE_DEC_MODE", mode)
            r, dec = gpu_dec(hs, codec, stream, len(data))
            assert r == len(data) and np.array_equal(dec, data), (codec.name, mode, len(data))
    monkeypatch.delenv("HSRLE_DEC_MODE", raising=False)
