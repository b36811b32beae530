"""hsrle_b200 -- thin ctypes binding over the product C-ABI library (libhsrle_b200.so).

The library is the product; this module only loads it and marshals numpy arrays / torch tensors.
There is no CPU fallback anywhere: if the CUDA library is missing, import-time loading raises, and
if no CUDA device is usable every codec call returns 0 bytes (the reference's error convention,
src/rle.h:100-394) and `last_error()` says why.

Codec names are the reference's function names without the `_compress` / `_decompress` suffix, e.g.
"rle8_multi", "rle8_packed_multi", "rle24_3symlut_byte", "rle64_sym_packed".
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# HSRLE_LIB picks another build of the same library (the diagnosis variant libhsrle_b200_dbg.so, see the Makefile)
LIB_PATH = os.path.join(os.path.dirname(_HERE), os.environ.get("HSRLE_LIB", "libhsrle_b200.so"))

_u8p = ctypes.POINTER(ctypes.c_uint8)
_u32p = ctypes.POINTER(ctypes.c_uint32)


class HsrleError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise HsrleError(f"{LIB_PATH} is missing: build it with `make -C hypersonic-rle-kit_b200` "
                         "(the product has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    sig_codec = [_u8p, ctypes.c_uint32, _u8p, ctypes.c_uint32]
    lib.rle_compress_bounds.restype = ctypes.c_uint32
    lib.rle_compress_bounds.argtypes = [ctypes.c_uint32]
    lib.rle_decompress_additional_size.restype = ctypes.c_uint32
    lib.hsrle_codec_id_from_name.restype = ctypes.c_int
    lib.hsrle_codec_id_from_name.argtypes = [ctypes.c_char_p]
    lib.hsrle_codec_id.restype = ctypes.c_int
    lib.hsrle_codec_id.argtypes = [ctypes.c_int] * 3
    lib.hsrle_compress_workspace_size.restype = ctypes.c_size_t
    lib.hsrle_compress_workspace_size.argtypes = [ctypes.c_int, ctypes.c_uint32]
    lib.hsrle_decompress_workspace_size.restype = ctypes.c_size_t
    lib.hsrle_decompress_workspace_size.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32]
    for f in (lib.hsrle_compress_device_async, lib.hsrle_decompress_device_async):
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32,
                      ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    for f in (lib.hsrle_compress_device, lib.hsrle_decompress_device):
        f.restype = ctypes.c_uint32
        f.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32]
    for f in (lib.hsrle_compress_host, lib.hsrle_decompress_host):
        f.restype = ctypes.c_uint32
        f.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32]
    lib.hsrle_last_error.restype = ctypes.c_char_p
    lib.hsrle_device.restype = ctypes.c_int
    lib.hsrle_kernel_launches.restype = ctypes.c_uint64
    lib._codec_sig = sig_codec
    return lib


lib = _load()


def last_error():
    return lib.hsrle_last_error().decode()


def kernel_launches():
    return int(lib.hsrle_kernel_launches())


def codec_id(name):
    cid = lib.hsrle_codec_id_from_name(name.encode())
    if cid < 0:
        raise HsrleError(f"unknown codec {name!r}")
    return cid


def compress_bounds(n):
    return int(lib.rle_compress_bounds(n))


def _named(fn_name):
    f = getattr(lib, fn_name)
    f.restype = ctypes.c_uint32
    f.argtypes = lib._codec_sig
    return f


def compress(fn_name, data, out_size=None):
    """Call a reference-named entry point (e.g. "rle8_multi_compress") with host buffers."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    n = len(data)
    if out_size is None:
        out_size = n + n // 256 + 512
    out = np.empty(out_size, dtype=np.uint8)
    r = _named(fn_name)(data.ctypes.data_as(_u8p), n, out.ctypes.data_as(_u8p), out_size)
    return out[:r].copy()


def decompress(fn_name, stream, out_size):
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    out = np.empty(max(out_size, 1), dtype=np.uint8)
    r = _named(fn_name)(stream.ctypes.data_as(_u8p), len(stream), out.ctypes.data_as(_u8p), out_size)
    return int(r), out[:r]


# ------------------------------------------------------------------ torch (device-resident) helpers

def compress_workspace_size(name, n):
    return int(lib.hsrle_compress_workspace_size(codec_id(name), n))


def decompress_workspace_size(name, in_size, out_size):
    return int(lib.hsrle_decompress_workspace_size(codec_id(name), in_size, out_size))


def compress_device_async(name, t_in, t_out, t_ws, t_result, stream_ptr, n=None):
    """Enqueue an encode on `stream_ptr`; all tensors are CUDA uint8/int32 tensors.  No host sync."""
    n = t_in.numel() if n is None else n
    rc = lib.hsrle_compress_device_async(codec_id(name), t_in.data_ptr(), n, t_out.data_ptr(), t_out.numel(),
                                         t_ws.data_ptr(), t_ws.numel(), t_result.data_ptr(), stream_ptr)
    if rc:
        raise HsrleError(f"compress enqueue failed ({rc}): {last_error()}")


def decompress_device_async(name, t_in, in_size, t_out, out_size, t_ws, t_result, stream_ptr):
    rc = lib.hsrle_decompress_device_async(codec_id(name), t_in.data_ptr(), in_size, t_out.data_ptr(), out_size,
                                           t_ws.data_ptr(), t_ws.numel(), t_result.data_ptr(), stream_ptr)
    if rc:
        raise HsrleError(f"decompress enqueue failed ({rc}): {last_error()}")


def _settle(t):
    """The synchronous device-pointer entry points run on a library-owned stream (include/hsrle_b200.h): whatever the
    caller still has in flight on the buffers -- here: torch's current stream -- must have finished before the call."""
    import torch
    torch.cuda.current_stream(t.device).synchronize()


def compress_device(name, t_in, t_out, n=None):
    n = t_in.numel() if n is None else n
    _settle(t_in)
    return int(lib.hsrle_compress_device(codec_id(name), t_in.data_ptr(), n, t_out.data_ptr(), t_out.numel()))


def decompress_device(name, t_in, in_size, t_out, out_size):
    _settle(t_in)
    return int(lib.hsrle_decompress_device(codec_id(name), t_in.data_ptr(), in_size, t_out.data_ptr(), out_size))
