#!/usr/bin/env python
"""One 2^30-byte frame of the configs[3] stream encoded as ONE stream by `world` ranks that share cuda:0 (messages over gloo), compared
with the single-GPU encode of the same frame.  usage: torchrun --nproc-per-node W scripts/slice_frame_repro.py codec[,codec] [frame]"""
import os, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
import torch, torch.distributed as dist
from common import gen_run_mixed_pieces, RM_PIECE
import hsrle_b200 as hs
from hsrle_b200 import sliced

codecs = sys.argv[1].split(",")
frame = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(0)
dist.init_process_group("gloo")
n = 1 << 30
dev = torch.device("cuda", 0)
for name in codecs:
    enc = sliced.SlicedEncoder(name, n, engine=None)
    assert enc.lo % RM_PIECE == 0 and enc.hi % RM_PIECE == 0, (enc.lo, enc.hi)
    sl = gen_run_mixed_pieces(frame * (n // RM_PIECE) + enc.lo // RM_PIECE, (enc.hi - enc.lo) // RM_PIECE, dev)
    buf = enc.exchange_halos(enc.make_input(sl))
    try:
        part, off, total = enc.encode(buf)
        print(f"rank {rank} {name}: part {len(part)} off {off} total {total} res {enc.t_res.tolist()[:8]}", flush=True)
    except Exception as e:
        print(f"rank {rank} {name}: FAILED {e}; res {enc.t_res.tolist()[:8]} msgs {enc.t_all.view(world, -1)[:, :12].tolist()}", flush=True)
dist.destroy_process_group()
