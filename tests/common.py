"""Shared test plumbing: codec table, ctypes bindings for the checkers (oracle/ and oracle/_ref) and
for the product C-ABI library, and the synthetic input generators.

The checkers are TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may load anything under oracle/.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libhsrle_ref.so")
PKG_DIR = os.path.join(ROOT, "hypersonic-rle-kit_b200")
if PKG_DIR not in sys.path:
    sys.path.insert(0, PKG_DIR)

PLAIN, PACKED, LUT3, LUT7 = 0, 1, 2, 3
SYM, BYTE = 0, 1


class Codec:
    """One reference codec pair (src/rle.h:100-394): `<name>_compress` / `<dname>_decompress`."""

    def __init__(self, name, W, align, variant, cname=None, dname=None):
        self.name, self.W, self.align, self.variant = name, W, align, variant
        self.cname = cname or name + "_compress"
        self.dname = dname or name + "_decompress"
        self.hdr = 9 if (W == 1 and variant in (PLAIN, PACKED)) else 8

    def __repr__(self):
        return self.name


def _codecs():
    out = [
        Codec("rle8_multi", 1, BYTE, PLAIN, "rle8_multi_compress", "rle8_decompress"),
        Codec("rle8_packed_multi", 1, BYTE, PACKED, "rle8_packed_multi_compress", "rle8_packed_decompress"),
        Codec("rle8_3symlut", 1, BYTE, LUT3),
        Codec("rle8_7symlut", 1, BYTE, LUT7),
    ]
    for bits in (16, 24, 32, 48, 64):
        W = bits // 8
        for an, a in (("sym", SYM), ("byte", BYTE)):
            out.append(Codec(f"rle{bits}_{an}", W, a, PLAIN))
            out.append(Codec(f"rle{bits}_{an}_packed", W, a, PACKED))
            out.append(Codec(f"rle{bits}_3symlut_{an}", W, a, LUT3))
            out.append(Codec(f"rle{bits}_7symlut_{an}", W, a, LUT7))
    return out


CODECS = _codecs()
CODEC_BY_NAME = {c.name: c for c in CODECS}
assert len(CODECS) == 44

_u8p = ctypes.POINTER(ctypes.c_uint8)


def _ptr(a):
    return a.ctypes.data_as(_u8p)


def build_oracle():
    """(Re)build oracle/liboracle.so (and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"], check=True)
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "rle_oracle.c")):
            build_oracle()
        _oracle = ctypes.CDLL(ORACLE_SO)
        for f in (_oracle.oracle_compress, _oracle.oracle_decompress):
            f.restype = ctypes.c_uint32
            f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _u8p, ctypes.c_uint32, _u8p, ctypes.c_uint32]
        _oracle.oracle_compress_bounds.restype = ctypes.c_uint32
        _oracle.oracle_compress_bounds.argtypes = [ctypes.c_uint32]
    return _oracle


def ref_lib():
    """The compiled, unmodified reference, or None when it has not been built (no /root/reference)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            if os.path.isdir("/root/reference/src"):
                build_oracle()
            else:
                return None
        _ref = ctypes.CDLL(REF_SO)
        for c in CODECS:
            for nm in (c.cname, c.dname):
                f = getattr(_ref, nm)
                f.restype = ctypes.c_uint32
                f.argtypes = [_u8p, ctypes.c_uint32, _u8p, ctypes.c_uint32]
        _ref.rle_compress_bounds.restype = ctypes.c_uint32
        _ref.rle_compress_bounds.argtypes = [ctypes.c_uint32]
    return _ref


def out_capacity(n):
    """Output capacity used on both sides of a parity test.  rle_compress_bounds is not a true
    upper bound for rle8_multi, so give n + n/256 + 512."""
    return n + n // 256 + 512


def ref_compress(codec, data):
    """Reference encoder with the padding convention of SURVEY App. C.1: the input copy is followed
    by pad[0] = ~in[n-W] so that the reference's out-of-bounds word compare always fails."""
    lib = ref_lib()
    n = len(data)
    buf = np.zeros(n + 64, dtype=np.uint8)
    buf[:n] = data
    if n >= codec.W:
        buf[n] = (~int(data[n - codec.W])) & 0xFF
    out = np.zeros(out_capacity(n), dtype=np.uint8)
    r = getattr(lib, codec.cname)(_ptr(buf), n, _ptr(out), len(out))
    return out[:r].copy()


def ref_decompress(codec, stream, n):
    lib = ref_lib()
    buf = np.zeros(len(stream) + 128, dtype=np.uint8)
    buf[: len(stream)] = stream
    out = np.zeros(n + 256, dtype=np.uint8)
    r = getattr(lib, codec.dname)(_ptr(buf), len(stream), _ptr(out), n + 128)
    return r, out[:n]


def oracle_compress(codec, data):
    lib = oracle_lib()
    n = len(data)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    out = np.zeros(out_capacity(n), dtype=np.uint8)
    r = lib.oracle_compress(codec.W, codec.align, codec.variant, _ptr(data), n, _ptr(out), len(out))
    return out[:r].copy()


def oracle_decompress(codec, stream, n):
    lib = oracle_lib()
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    r = lib.oracle_decompress(codec.W, codec.align, codec.variant, _ptr(stream), len(stream), _ptr(out), n)
    return r, out[:n]


# --------------------------------------------------------------------------- generators

def gen_fuzz(rng, n, max_sym=8, p_run=0.5, long_every=0):
    """The reference fuzzer's section model (src/rle_fuzz.c:30-44,360-438): alternating random /
    repeating-symbol sections, symbol length 1..max_sym, arbitrary alignment, three length classes
    including the 8-bit and 16-bit field boundaries."""
    parts = []
    total = 0
    k = 0
    while total < n:
        cls = rng.integers(0, 10)
        if cls < 6:
            ln = int(rng.integers(1, 40))
        elif cls < 9:
            ln = int(rng.integers(1, 300))
        else:
            ln = int(rng.integers(240, 1200))
        if long_every and k % long_every == long_every - 1:
            ln = int(rng.integers(65500, 66100))
        k += 1
        if rng.random() < p_run:
            w = int(rng.integers(1, max_sym + 1))
            sym = rng.integers(0, 256, size=w, dtype=np.uint8)
            if rng.random() < 0.3:
                sym[:] = rng.choice(np.array([0, 0x7F, 0xFF, 1, 0x7E, 0x80, 0xFE], dtype=np.uint8))
            reps = -(-ln // w)
            sec = np.tile(sym, reps)[:ln]
        else:
            if rng.random() < 0.5:
                sec = rng.integers(0, 256, size=ln, dtype=np.uint8)
            else:
                sec = rng.integers(0, 4, size=ln, dtype=np.uint8)
        parts.append(sec)
        total += ln
    return np.concatenate(parts)[:n].copy()


def gen_dct(n, seed=0x5EED):
    """Quantised-DCT byte stream, SURVEY App. E.1 (video-frame.raw shape; 88,473,600 B at full size).
    8x8 blocks in zig-zag order; a 2-state Markov chain (coded / skipped, P(coded)=0.20, stay 0.96)
    selects blocks; in a coded block P(coef k != 0) = 0.995*0.97**k, |v| ~ Geometric(1/12) clipped
    to int8 with random sign; skipped blocks are all zero.  Vectorised numpy, fixed seed."""
    rng = np.random.default_rng(seed)
    nblk = -(-n // 64)
    # Markov chain via geometric sojourn times: mean 25 coded / 100 skipped blocks.
    coded = np.zeros(nblk, dtype=bool)
    pos = 0
    state = bool(rng.random() < 0.2)
    # draw sojourns in bulk
    while pos < nblk:
        m = max(1024, (nblk - pos) // 50)
        lc = rng.geometric(1 / 25.0, size=m)
        ls = rng.geometric(1 / 100.0, size=m)
        for a, b in zip(lc, ls):
            if state:
                coded[pos:pos + a] = True
                pos += a
                state = False
            else:
                pos += b
                state = True
            if pos >= nblk:
                break
    idx = np.flatnonzero(coded)
    out = np.zeros((nblk, 64), dtype=np.uint8)
    if len(idx):
        pk = 0.995 * 0.97 ** np.arange(64)
        nz = rng.random((len(idx), 64)) < pk
        mag = np.minimum(rng.geometric(1 / 12.0, size=(len(idx), 64)), 127).astype(np.int16)
        sign = np.where(rng.random((len(idx), 64)) < 0.5, -1, 1).astype(np.int16)
        v = (np.where(nz, mag * sign, 0)).astype(np.int8).view(np.uint8)
        out[idx] = v
    return out.reshape(-1)[:n].copy()


def gen_short_runs(n, seed=7, W=1):
    """Short-run-heavy stream, SURVEY App. E.2: alphabet of 4-8 values, runs U{2..9} (of W-byte
    symbols) alternating with literal gaps U{0..16} of random bytes."""
    rng = np.random.default_rng(seed)
    alpha = rng.integers(0, 256, size=(int(rng.integers(4, 9)), W), dtype=np.uint8)
    est = max(16, n // (6 * W + 8) + 16)
    parts = []
    total = 0
    while total < n:
        runs = rng.integers(2, 10, size=est)
        gaps = rng.integers(0, 17, size=est)
        syms = rng.integers(0, len(alpha), size=est)
        for r, g, s in zip(runs, gaps, syms):
            parts.append(np.tile(alpha[s], r))
            if g:
                parts.append(rng.integers(0, 256, size=g, dtype=np.uint8))
            total += r * W + g
            if total >= n:
                break
    return np.concatenate(parts)[:n].copy()


def gen_run_mixed(n, seed=11, max_run_log2=20, max_lit_log2=16):
    """Run-mixed stream, SURVEY App. E.3: alternating runs (length log-uniform 2..2**20, symbol
    width 1 or 8) and random literals (log-uniform 1..2**16)."""
    rng = np.random.default_rng(seed)
    parts = []
    total = 0
    while total < n:
        rl = int(2 ** rng.uniform(1, max_run_log2))
        w = 1 if rng.random() < 0.5 else 8
        sym = rng.integers(0, 256, size=w, dtype=np.uint8)
        parts.append(np.tile(sym, -(-rl // w))[:rl])
        ll = int(2 ** rng.uniform(0, max_lit_log2))
        parts.append(rng.integers(0, 256, size=ll, dtype=np.uint8))
        total += rl + ll
    return np.concatenate(parts)[:n].copy()


# --------------------------------------------------------------------------- configs[3]: multi-GiB run-mixed stream
RM_PIECE = 64 << 20          # the stream is defined piece by piece: any rank can materialise any piece on its own


def _rm_boundary_symbol(k, seed):
    """8-byte run symbol that straddles the boundary between pieces k-1 and k (also frame and slice boundaries)."""
    r = np.random.default_rng([seed, 0xB0DE, k])
    return r.integers(0, 256, size=8, dtype=np.uint8)


def run_mixed_piece_table(piece, seed=0xC3, piece_bytes=RM_PIECE):
    """Segment table of one piece: alternating runs (length log-uniform 2..2^20, symbol width 1 or 8)
    and random literals (log-uniform 1..2^16).  The first and the last segment are runs whose symbols are shared with
    the neighbouring pieces, so a run straddles every piece boundary by construction.
    Returns (lengths int64[S], is_run bool[S], sym uint8[S, 8])."""
    r = np.random.default_rng([seed, piece])
    lens, isrun, syms = [], [], []
    first = int(r.integers(16, 4096))
    lens.append(first); isrun.append(True); syms.append(_rm_boundary_symbol(piece, seed))
    total = first
    tail = int(r.integers(16, 4096))
    while True:
        ll = int(2 ** r.uniform(0, 16))
        rl = int(2 ** r.uniform(1, 20))
        if total + ll + rl + tail >= piece_bytes:
            break
        lens.append(ll); isrun.append(False); syms.append(np.zeros(8, dtype=np.uint8))
        s = r.integers(0, 256, size=8, dtype=np.uint8)
        if r.random() < 0.5:
            s[:] = s[0]                                   # symbol width 1
        lens.append(rl); isrun.append(True); syms.append(s)
        total += ll + rl
    rest = piece_bytes - total
    ll = max(1, rest - tail)
    lens.append(ll); isrun.append(False); syms.append(np.zeros(8, dtype=np.uint8))
    lens.append(rest - ll); isrun.append(True); syms.append(_rm_boundary_symbol(piece + 1, seed))
    if lens[-1] == 0:
        lens.pop(); isrun.pop(); syms.pop()
    assert sum(lens) == piece_bytes
    return np.array(lens, dtype=np.int64), np.array(isrun, dtype=bool), np.stack(syms)


def gen_run_mixed_pieces(first_piece, n_pieces, device, seed=0xC3, piece_bytes=RM_PIECE):
    """Bytes [first_piece * piece_bytes, (first_piece + n_pieces) * piece_bytes) of the configs[3] stream as a torch
    uint8 tensor on `device` (segment tables from numpy, expansion with torch ops; literal bytes are a counter-based
    hash of the absolute position, run bytes the segment's 8-byte symbol at phase position % 8)."""
    import torch
    out = torch.empty(n_pieces * piece_bytes, dtype=torch.uint8, device=device)
    for i in range(n_pieces):
        piece = first_piece + i
        lens, isrun, syms = run_mixed_piece_table(piece, seed, piece_bytes)
        t_len = torch.from_numpy(lens).to(device)
        seg = torch.repeat_interleave(torch.arange(len(lens), dtype=torch.int16, device=device), t_len)
        pos = torch.arange(piece_bytes, dtype=torch.int64, device=device) + piece * piece_bytes
        x = pos * -7046029254386353131 + (seed * 1000003 + 12345)      # 0x9E3779B97F4A7C15 as int64; wraps
        x ^= x >> 29
        x *= -4658895280553007687                                      # 0xBF58476D1CE4E5B9 as int64
        x ^= x >> 32
        lit = (x & 0xFF).to(torch.uint8)
        del x
        seg_l = seg.to(torch.int64)
        del seg
        t_sym = torch.from_numpy(syms.reshape(-1)).to(device)
        run = t_sym[seg_l * 8 + (pos & 7)]
        t_isrun = torch.from_numpy(isrun).to(device)[seg_l]
        del seg_l, pos
        out[i * piece_bytes:(i + 1) * piece_bytes] = torch.where(t_isrun, run, lit)
        del run, lit, t_isrun
    return out
