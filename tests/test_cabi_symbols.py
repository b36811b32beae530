"""CPU test: the product C-ABI library loads and exports every symbol include/hsrle_b200.h declares.
No compute calls (no GPU here)."""
import ctypes
import os
import re

from common import CODECS, ROOT

LIB = os.path.join(ROOT, "hypersonic-rle-kit_b200", "libhsrle_b200.so")


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "hsrle_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    body = " ".join(l for l in txt.splitlines() if not l.startswith("#"))
    return sorted(set(re.findall(r"\b(\w+)\s*\(", body)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build the product first: make -C hypersonic-rle-kit_b200"
    lib = ctypes.CDLL(LIB)
    names = declared_symbols()
    assert len(names) >= 44 * 2 + 10
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/hsrle_b200.h but not exported"


def test_pure_host_helpers():
    lib = ctypes.CDLL(LIB)
    lib.rle_compress_bounds.restype = ctypes.c_uint32
    assert lib.rle_compress_bounds(1000) == 1193           # src/rle8_extreme_cpu.c:22-28
    assert lib.rle_compress_bounds((1 << 30) + 1) == 0
    assert lib.rle_decompress_additional_size() == 128      # src/rle8_extreme_cpu.c:17-20
    lib.hsrle_codec_id_from_name.restype = ctypes.c_int
    ids = set()
    for c in CODECS:
        i = lib.hsrle_codec_id_from_name(c.name.encode())
        assert i >= 0, c.name
        ids.add(i)
    assert len(ids) == 44
    lib.hsrle_compress_workspace_size.restype = ctypes.c_size_t
    assert lib.hsrle_compress_workspace_size(lib.hsrle_codec_id_from_name(b"rle8_multi"), 1 << 20) > 0
