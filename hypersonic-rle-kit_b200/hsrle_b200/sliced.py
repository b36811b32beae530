"""hsrle_b200.sliced -- ONE reference-identical stream encoded by several GPUs (one process per GPU).

The input is cut into contiguous slices, one per rank; every rank runs the encoder phases of the C ABI
(`hsrle_slice_compress_phase`, include/hsrle_b200.h) on its slice and the ranks exchange a 256-byte message with
an all-gather between the phases (NCCL over NVLink when the process group is NCCL): run-boundary records that
span a cut, the emit automaton's state, and the per-rank byte counts that give every rank its offset in the
single stream.  The payload is never moved: a rank's share of the stream stays on the GPU that produced it
(`SlicedResult.part`), `gather_stream` concatenates the shares when one buffer is wanted.

Inputs above the stream format's u32 ceiling are cut into frames first (`frame_bounds`): each frame is a
complete reference-identical stream, exactly what a caller of the reference's u32 API has to do.

`torch` / `torch.distributed` are plumbing here (device buffers, streams, the collective); all codec work happens
in the CUDA library.  The engine argument exists so that the CPU tests can drive the same orchestration with the
host-side stage simulator over gloo; the product default is the CUDA library and there is no CPU fallback.
"""
import ctypes

import torch
import torch.distributed as dist

SLICE_ALIGN = 16 * 8192          # slice starts are multiples of the scan macro-tile (128 KiB)
FRONT = 32                       # halo bytes before / after the slice in a rank's input buffer
TAIL = 32
MSG_WORDS = 64
FRAME_BYTES = 1 << 30            # largest input rle_compress_bounds accepts (src/rle8_extreme_cpu.c:22-28)


class hsrle_slice_job(ctypes.Structure):
    _fields_ = [("codec", ctypes.c_int), ("rank", ctypes.c_int), ("world", ctypes.c_int),
                ("n", ctypes.c_uint32), ("lo", ctypes.c_uint32), ("hi", ctypes.c_uint32),
                ("dIn", ctypes.c_void_p), ("dOut", ctypes.c_void_p), ("outCap", ctypes.c_uint32),
                ("dWorkspace", ctypes.c_void_p), ("workspaceSize", ctypes.c_size_t),
                ("dMsg", ctypes.c_void_p), ("dAll", ctypes.c_void_p), ("dResult", ctypes.c_void_p)]


def slice_bounds(n, world):
    """Contiguous slices of ceil(n / world) bytes rounded up to 128 KiB.  Returns (bounds, active): rank r owns
    [bounds[r], bounds[r + 1]); only the first `active` ranks get a non-empty slice (small inputs)."""
    per = -(-n // world)
    per = -(-per // SLICE_ALIGN) * SLICE_ALIGN
    active = max(1, min(world, -(-n // per)))
    bounds = [min(r * per, n) for r in range(active)] + [n] * (world - active + 1)
    return bounds, active


def frame_bounds(total, frame_bytes=FRAME_BYTES):
    """Frames of a long input: [k * frame_bytes, min((k + 1) * frame_bytes, total))."""
    return [(o, min(o + frame_bytes, total)) for o in range(0, total, frame_bytes)]


def _all_gather(out, inp, group=None):
    """all_gather_into_tensor; a gloo group cannot gather CUDA tensors, so those are staged through the host (used
    by the single-GPU tests, where two ranks share one device and NCCL refuses to run)."""
    if inp.is_cuda and dist.get_backend(group) == "gloo":
        h_out = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather_into_tensor(h_out, inp.cpu(), group=group)
        out.copy_(h_out)
    else:
        dist.all_gather_into_tensor(out, inp, group=group)


class CudaEngine:
    """The product engine: libhsrle_b200.so on the current CUDA device."""

    def __init__(self, device=None):
        from . import lib, last_error
        if not torch.cuda.is_available():
            raise RuntimeError("hsrle_b200.sliced needs a CUDA device (there is no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self._lib, self._err = lib, last_error
        lib.hsrle_slice_compress_phase.restype = ctypes.c_int
        lib.hsrle_slice_compress_phase.argtypes = [ctypes.POINTER(hsrle_slice_job), ctypes.c_int, ctypes.c_void_p]
        lib.hsrle_slice_workspace_size.restype = ctypes.c_size_t
        lib.hsrle_slice_workspace_size.argtypes = [ctypes.c_int, ctypes.c_uint32]

    def codec_id(self, name):
        from . import codec_id
        return codec_id(name)

    def workspace_size(self, codec, nbytes):
        return int(self._lib.hsrle_slice_workspace_size(codec, nbytes))

    def phase(self, job, k):
        rc = self._lib.hsrle_slice_compress_phase(ctypes.byref(job), k, torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            raise RuntimeError(f"slice phase {k} failed ({rc}): {self._err()}")


class SlicedEncoder:
    """Per-rank state of a sliced encode: buffers are allocated once and re-used across calls."""

    def __init__(self, codec_name, n, group=None, engine=None, rank=None, world=None):
        self.engine = engine if engine is not None else CudaEngine()
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.codec = self.engine.codec_id(codec_name)
        self.n = int(n)
        if not 0 < self.n < 0xFFFFFFF0 - 64:
            raise ValueError("one stream holds at most 2^32 - 81 bytes: cut the input into frames (frame_bounds)")
        self.bounds, self.active = slice_bounds(self.n, self.world)
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        dev = self.engine.device
        ln = self.hi - self.lo
        self.out_cap = ln + ln // 256 + 1024
        self.t_out = torch.empty(self.out_cap + 64, dtype=torch.uint8, device=dev)
        self.t_ws = torch.empty(max(self.engine.workspace_size(self.codec, max(ln, 1)), 256), dtype=torch.uint8, device=dev)
        self.t_msg = torch.zeros(MSG_WORDS, dtype=torch.int32, device=dev)
        self.t_all = torch.zeros(self.world * MSG_WORDS, dtype=torch.int32, device=dev)
        self.t_res = torch.zeros(8, dtype=torch.int32, device=dev)
        self.state_rounds = 0

    # ---- input layout: [FRONT halo][slice][TAIL halo]
    def make_input(self, slice_bytes, left_halo=None, right_halo=None):
        """Device buffer for this rank's slice with the 32-byte halos of its neighbours (zeros at the stream ends)."""
        dev = self.engine.device
        ln = self.hi - self.lo
        buf = torch.zeros(FRONT + ln + TAIL + 16, dtype=torch.uint8, device=dev)
        buf[FRONT:FRONT + ln] = slice_bytes
        if left_halo is not None and len(left_halo):
            buf[FRONT - len(left_halo):FRONT] = left_halo
        if right_halo is not None and len(right_halo):
            buf[FRONT + ln:FRONT + ln + len(right_halo)] = right_halo
        return buf

    def exchange_halos(self, buf):
        """Fill the halos of `buf` from the neighbouring ranks (one all-gather of 64 bytes per rank).  Part of
        laying the input out on the GPUs, not of the encode."""
        dev = self.engine.device
        ln = self.hi - self.lo
        edge = torch.zeros(FRONT + TAIL, dtype=torch.uint8, device=dev)
        if ln > 0:
            k = min(ln, TAIL)
            edge[:k] = buf[FRONT:FRONT + k]                    # my first bytes  -> right halo of the rank before me
            k = min(ln, FRONT)
            edge[FRONT + TAIL - k:] = buf[FRONT + ln - k:FRONT + ln]      # my last bytes -> left halo of the rank after me
        allv = torch.zeros(self.world * (FRONT + TAIL), dtype=torch.uint8, device=dev)
        _all_gather(allv, edge, self.group)
        allv = allv.view(self.world, FRONT + TAIL)
        if 0 < self.rank < self.active and ln > 0:
            buf[:FRONT] = allv[self.rank - 1, TAIL:]
        if self.rank + 1 < self.active and ln > 0:
            buf[FRONT + ln:FRONT + ln + TAIL] = allv[self.rank + 1, :TAIL]
        return buf

    def _gather(self):
        _all_gather(self.t_all, self.t_msg, self.group)

    def encode(self, t_in):
        """Encode; `t_in` is this rank's input buffer (see make_input).  Returns (part, offset, total): this rank's
        share of the stream (a view into the encoder's output buffer), its byte offset in the stream and the
        stream's total length.  Raises on error (output too small cannot happen with the buffers sized here)."""
        eng = self.engine
        on = self.rank < self.active
        job = hsrle_slice_job(self.codec, self.rank, self.active, self.n, self.lo, self.hi, t_in.data_ptr(), self.t_out.data_ptr(),
                              self.out_cap, self.t_ws.data_ptr(), self.t_ws.numel(), self.t_msg.data_ptr(), self.t_all.data_ptr(),
                              self.t_res.data_ptr())
        if not on:
            self.t_msg.zero_()
        if on:
            eng.phase(job, 0)
        self._gather()
        if on:
            eng.phase(job, 1)
        self._gather()
        self.state_rounds = 0
        while True:
            if on:
                eng.phase(job, 2)
            self._gather()
            self.state_rounds += 1
            changed = self.t_all.view(self.world, MSG_WORDS)[:self.active, 6]
            if not bool(changed.any().item()):
                break
            if self.state_rounds > self.world + 1:
                # a state travels at least one slice per round: world rounds always suffice.  More means the exchange is broken.
                raise RuntimeError(f"sliced encode: incoming states did not settle after {self.state_rounds} rounds (world {self.world})")
        if on:
            eng.phase(job, 3)
        self._gather()
        if not on:
            return self.t_out[:0], 0, 0
        eng.phase(job, 4)
        res = self.t_res.cpu().numpy().astype("uint32")
        if res[1] != 0:
            raise RuntimeError(f"sliced encode failed on rank {self.rank}: status {int(res[1])}")
        part_len, start, off, total = int(res[0]), int(res[2]), int(res[3]), int(res[4])
        return self.t_out[start:start + part_len], off, total


def gather_stream(part, total, group=None):
    """Concatenate every rank's share into one buffer on every rank (all-gather of the padded shares)."""
    world = dist.get_world_size(group)
    dev = part.device
    ln = torch.tensor([part.numel()], dtype=torch.int64, device=dev)
    lens = torch.zeros(world, dtype=torch.int64, device=dev)
    _all_gather(lens, ln, group)
    lens = [int(x) for x in lens.cpu()]
    mx = max(max(lens), 1)
    pad = torch.zeros(mx, dtype=torch.uint8, device=dev)
    pad[:part.numel()] = part
    allp = torch.zeros(world * mx, dtype=torch.uint8, device=dev)
    _all_gather(allp, pad, group)
    allp = allp.view(world, mx)
    out = torch.cat([allp[r, :lens[r]] for r in range(world)])
    assert total == 0 or out.numel() == total or max(lens) == 0
    return out
