#!/usr/bin/env python
"""One reference-identical stream encoded by N GPUs (hsrle_b200.sliced, SURVEY 8e / BASELINE configs[3] shape): time per
1 GiB frame, aggregate GB/s, and a decode round trip of the gathered stream.  Launch with torchrun, one rank per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_sliced.py [codec] [log2n] [reps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))
os.environ.setdefault("NCCL_DEBUG", "WARN")
import torch
import torch.distributed as dist
import hsrle_b200 as hs
from hsrle_b200.sliced import SlicedEncoder, gather_stream, FRONT

name = sys.argv[1] if len(sys.argv) > 1 else "rle8_multi"
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 30)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
# the same run-mixed stream on every rank (runs of 1 .. ~700 equal bytes, mean 37), each rank keeps its slice
g = torch.Generator(device=dev); g.manual_seed(77)
starts = torch.rand(n, device=dev, generator=g) < (1.0 / 37.0)
seg = torch.cumsum(starts.to(torch.int32), 0)
full = (seg.to(torch.int64) * 2654435761 % 251).to(torch.uint8)
del starts, seg
enc = SlicedEncoder(name, n)
t_in = enc.make_input(full[enc.lo:enc.hi], full[max(enc.lo - FRONT, 0):enc.lo], full[enc.hi:min(enc.hi + 32, n)])
part, off, total = enc.encode(t_in)            # warm-up
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    part, off, total = enc.encode(t_in)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
stream = gather_stream(part, total)
ok = None
if rank == 0:
    t_dec = torch.empty(n + 128, dtype=torch.uint8, device=dev)
    pad = torch.zeros(total + 256, dtype=torch.uint8, device=dev); pad[:total] = stream
    rd = hs.decompress_device(name, pad, total, t_dec, n)
    ok = bool(rd == n and torch.equal(t_dec[:n], full))
    # single-call encode of the same input on one GPU for comparison (and stream identity)
    t_out = torch.empty(n + n // 256 + 512, dtype=torch.uint8, device=dev)
    r1 = hs.compress_device(name, full, t_out)
    same = bool(r1 == total and torch.equal(t_out[:r1], stream))
    print(json.dumps({"what": "sliced encode of one stream", "codec": name, "n": n, "n_gpus": world, "ms_per_frame": round(float(ms.item()), 3),
                      "GBps": round(n / float(ms.item()) / 1e6, 1), "stream_bytes": int(total), "roundtrip_ok": ok, "identical_to_single_gpu_stream": same,
                      "state_rounds": enc.state_rounds}), flush=True)
dist.barrier()
dist.destroy_process_group()
