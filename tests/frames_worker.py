"""Worker of tests/test_frames.py: the frame-sequence host logic over gloo (one process per rank, no GPU).  The codec calls
are made by the CPU checker (oracle port) standing in for the CUDA library -- what is under test is the dealing of frames to
ranks, the size exchange (hsrle_b200.frames.gather_sizes) and the self-delimiting container."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from common import CODEC_BY_NAME, ROOT, gen_run_mixed_pieces, oracle_compress, oracle_decompress  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))


def main():
    from hsrle_b200 import frames as fr
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    FB, F = 256 * 1024, 5                      # small frames: the same logic as 2^30-byte ones
    total = F * FB - 1000                      # the last frame is shorter
    bounds = fr.frame_bounds(total, FB)
    assert len(bounds) == F
    whole = gen_run_mixed_pieces(0, F, "cpu", piece_bytes=FB).numpy()[:total]
    bad = 0
    for name in ("rle8_multi", "rle64_byte_packed"):
        codec = CODEC_BY_NAME[name]
        mine = fr.deal_frames(F, rank, world)
        streams = [oracle_compress(codec, whole[bounds[f][0]:bounds[f][1]]) for f in mine]
        sizes = fr.gather_sizes([len(s) for s in streams], F)
        offs, tot = fr.concat_layout(sizes)
        # every rank writes its frames at their offsets; rank 0 receives the container
        buf = torch.zeros(tot, dtype=torch.uint8)
        for f, s in zip(mine, streams):
            buf[offs[f]:offs[f] + len(s)] = torch.from_numpy(s)
        dist.reduce(buf, 0, op=dist.ReduceOp.SUM)          # disjoint ranges: the sum is the concatenation
        if rank == 0:
            parts = fr.split_concat(buf.numpy())
            ok = len(parts) == F
            for f, p in enumerate(parts if ok else []):
                a, b = bounds[f]
                want = oracle_compress(codec, whole[a:b])
                r, dec = oracle_decompress(codec, p, b - a)
                ok = ok and np.array_equal(p, want) and r == b - a and np.array_equal(dec, whole[a:b])
            bad += 0 if ok else 1
    t = torch.tensor([bad], dtype=torch.int64)
    dist.broadcast(t, 0)
    dist.destroy_process_group()
    if rank == 0:
        print(f"frames worker: world={world} mismatches={int(t.item())}", flush=True)
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
