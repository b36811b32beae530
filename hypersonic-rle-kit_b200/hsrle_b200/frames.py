"""hsrle_b200.frames -- inputs above the stream format's u32 ceiling as a FRAME SEQUENCE, on one or several GPUs.

The reference's entry points take `uint32_t` sizes, its headers store u32 lengths and `rle_compress_bounds()` is 0 above
2^30 bytes (src/rle8_extreme_cpu.c:22-28, header src/rle8_extreme_cpu.c:5-15): a caller with more data has to cut it into
frames of at most 2^30 bytes and call the codec once per frame.  This module is that caller for device-resident data:

  * `frame_bounds(total)`              the cuts (multiples of 2^30, the largest size rle_compress_bounds accepts);
  * `deal_frames(nframes, rank, world)` which frames a rank owns when the sequence is sharded over `world` GPUs (one process
                                        per GPU, round-robin; independent streams, so there is no data-path collective);
  * `FrameCodec`                        per-rank encoder/decoder of its frames: every frame is a complete,
                                        reference-identical stream (byte-identical to what the reference's `*_compress`
                                        returns for that frame); the calls of up to `streams` frames are in flight at once on
                                        separate CUDA streams so that one frame's latency-bound kernels overlap another's
                                        bandwidth-bound ones.
  * `concat_layout` / `split_concat`    the container: frames concatenated back to back; each frame is self-delimiting through
                                        the `compressedLength` field of its own header (bytes 4..7).

`torch` is plumbing (device buffers, streams); all codec work happens in libhsrle_b200.so.  No CPU fallback.
"""
import torch

from .sliced import FRAME_BYTES, frame_bounds  # noqa: F401  (re-exported)


def deal_frames(nframes, rank, world):
    """Frames owned by `rank`: rank, rank + world, ...  (frames are equally long except the last: round-robin balances)."""
    return list(range(rank, nframes, world))


def concat_layout(sizes):
    """Byte offsets of the frames in the concatenated container and its total length."""
    offs, o = [], 0
    for s in sizes:
        offs.append(o)
        o += int(s)
    return offs, o


def split_concat(buf):
    """Cut a concatenated container (numpy uint8 / bytes) into its frames by reading each header's compressedLength."""
    import numpy as np
    buf = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    out, o = [], 0
    while o < len(buf):
        if o + 8 > len(buf):
            raise ValueError("truncated frame header")
        clen = int.from_bytes(buf[o + 4:o + 8].tobytes(), "little")
        if clen < 8 or o + clen > len(buf):
            raise ValueError("bad frame length")
        out.append(buf[o:o + clen])
        o += clen
    return out


class StreamPool:
    """CUDA streams with one codec workspace each, shared by every FrameCodec of a rank (the workspace of a 2^30-byte frame is
    a few GiB: one per stream in flight, not one per codec)."""

    def __init__(self, codec_names, max_frame_bytes, device=None, streams=4):
        from . import compress_workspace_size, decompress_workspace_size
        if not torch.cuda.is_available():
            raise RuntimeError("hsrle_b200.frames needs a CUDA device (there is no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        mx = int(max_frame_bytes)
        ws = max(max(compress_workspace_size(c, mx), decompress_workspace_size(c, mx + mx // 256 + 512, mx)) for c in codec_names)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, streams))]
        self.ws = [torch.empty(ws, dtype=torch.uint8, device=self.device) for _ in self.streams]
        self.ws_bytes = ws

    def fork(self):
        """Every pool stream waits for what the current stream has enqueued so far."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        for s in self.streams:
            s.wait_event(ev)

    def join(self):
        """The current stream waits for everything enqueued on the pool streams."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            ev = torch.cuda.Event()
            ev.record(s)
            cur.wait_event(ev)


class FrameCodec:
    """Encoder/decoder of one rank's frames.  Buffers (compressed frames, result words) are allocated once and re-used across
    calls; as many calls as the pool has streams are in flight at once."""

    def __init__(self, codec_name, frame_sizes, device=None, streams=4, pool=None):
        from . import codec_id
        self.name = codec_name
        self.codec = codec_id(codec_name)
        self.sizes = [int(s) for s in frame_sizes]
        if any(s <= 0 or s > FRAME_BYTES for s in self.sizes):
            raise ValueError("a frame holds 1 .. 2^30 bytes (rle_compress_bounds, src/rle8_extreme_cpu.c:22-28)")
        self.caps = [s + s // 256 + 512 for s in self.sizes]
        if pool is None:
            pool = StreamPool([codec_name], max(self.sizes, default=1), device, max(1, min(streams, len(self.sizes) or 1)))
        self.pool = pool
        self.device = pool.device
        self.streams, self.ws = pool.streams, pool.ws
        self.nstreams = len(self.streams)
        self.comp = [torch.empty(c, dtype=torch.uint8, device=self.device) for c in self.caps]
        self.res = torch.zeros(len(self.sizes), 16, dtype=torch.int32, device=self.device)
        self.clen = [None] * len(self.sizes)

    def _fork(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        for s in self.streams:
            s.wait_event(ev)

    def _join(self):
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            ev = torch.cuda.Event()
            ev.record(s)
            cur.wait_event(ev)

    def encode_async(self, frames):
        """Enqueue the encode of every frame (`frames[i]`: device uint8 tensor of sizes[i] bytes).  Stream-ordered behind the
        current stream; no host synchronisation.  `finish_encode()` reads the sizes."""
        from . import compress_device_async
        self._fork()
        for i, t in enumerate(frames):
            j = i % self.nstreams
            compress_device_async(self.name, t, self.comp[i], self.ws[j], self.res[i, :8], self.streams[j].cuda_stream, n=self.sizes[i])
        self._join()

    def finish_encode(self):
        """Synchronise and return the compressed size of every frame (raises on a codec error)."""
        torch.cuda.current_stream(self.device).synchronize()
        r = self.res.cpu().numpy()
        for i in range(len(self.sizes)):
            if r[i, 1] != 0 or r[i, 0] == 0:
                raise RuntimeError(f"{self.name}: frame {i} failed to encode (status {int(r[i, 1])})")
            self.clen[i] = int(r[i, 0]) & 0xFFFFFFFF
        return list(self.clen)

    def decode_async(self, outs, indices=None):
        """Enqueue the decode of the frames `indices` (default: all) into `outs[k]` (device tensors of at least sizes[i]
        bytes; `outs` may be shorter than the frame list when the caller only wants throughput: it is cycled through)."""
        from . import decompress_device_async
        indices = list(range(len(self.sizes))) if indices is None else list(indices)
        self._fork()
        for k, i in enumerate(indices):
            j = k % self.nstreams
            clen = self.clen[i] if self.clen[i] is not None else self.caps[i]
            decompress_device_async(self.name, self.comp[i], clen, outs[k % len(outs)], self.sizes[i], self.ws[j], self.res[i, 8:],
                                    self.streams[j].cuda_stream)
        self._join()

    def roundtrip_async(self, frames, outs, offset=0):
        """Enqueue encode + decode of every frame, frame i on pool stream (i + offset) mod streams: the two calls of a frame are
        stream-ordered, different frames -- and other codecs that share the pool with another `offset` -- overlap, so the
        latency-bound phases of one call run under the bandwidth-bound kernels of another.  No fork / join here: the caller
        brackets a batch of such calls with pool.fork() / pool.join().  `outs[j]`: one output buffer per pool stream."""
        from . import compress_device_async, decompress_device_async
        for i, t in enumerate(frames):
            j = (i + offset) % self.nstreams
            q = self.streams[j].cuda_stream
            compress_device_async(self.name, t, self.comp[i], self.ws[j], self.res[i, :8], q, n=self.sizes[i])
            clen = self.clen[i] if self.clen[i] is not None else self.caps[i]
            decompress_device_async(self.name, self.comp[i], clen, outs[j], self.sizes[i], self.ws[j], self.res[i, 8:], q)

    def finish_decode(self, indices=None):
        torch.cuda.current_stream(self.device).synchronize()
        r = self.res.cpu().numpy()
        for i in (range(len(self.sizes)) if indices is None else indices):
            if r[i, 9] != 0 or (int(r[i, 8]) & 0xFFFFFFFF) != self.sizes[i]:
                raise RuntimeError(f"{self.name}: frame {i} failed to decode (status {int(r[i, 9])})")

    def stream(self, i):
        """Compressed frame i (device view)."""
        return self.comp[i][: self.clen[i]]


def gather_sizes(my_sizes, nframes, group=None):
    """All ranks learn the compressed size of every frame (frame f lives on rank f mod world): one all-gather of u64 counts
    (NCCL over NVLink in the product, gloo in the CPU tests).  Returns the sizes in frame order; with `concat_layout` that is
    every frame's byte offset in the concatenated container -- the only exchange the frame path needs."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-nframes // world)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.zeros(per, dtype=torch.int64, device=dev)
    own = deal_frames(nframes, rank, world)
    assert len(my_sizes) == len(own)
    for k, s in enumerate(my_sizes):
        mine[k] = int(s)
    allv = torch.zeros(world * per, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allv, mine, group=group)
    allv = allv.view(world, per).cpu()
    return [int(allv[f % world, f // world]) for f in range(nframes)]
