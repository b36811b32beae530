"""Worker of tests/test_sliced_gloo.py and tests/test_gpu_sliced.py: one process per rank.

mode "sim": the orchestration of hsrle_b200.sliced over gloo with the host-side stage simulator as the engine
            (tests/sim, test tool) -- covers the N>1 host logic and the slice bookkeeping without a GPU.
mode "gpu": the product path, NCCL, one GPU per rank.
mode "gpu1": the product path with every rank on cuda:0 and the messages over gloo (for a single-GPU box).
Every rank takes its slice of the same seeded input, the stream is gathered and rank 0 compares it with the oracle.
"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from common import CODEC_BY_NAME, ROOT, gen_dct, gen_fuzz, gen_run_mixed, gen_short_runs, oracle_compress  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))


class SimEngine:
    """hsrle_slice_compress_phase's host twin (tests/sim/sim_pipeline.cpp); CPU tensors."""

    def __init__(self):
        import subprocess
        so = os.path.join(HERE, "sim", "libsim.so")
        src = os.path.join(HERE, "sim", "sim_pipeline.cpp")
        csrc = os.path.join(ROOT, "hypersonic-rle-kit_b200", "csrc")
        newest = max(os.path.getmtime(p) for p in [src] + [os.path.join(csrc, h) for h in os.listdir(csrc) if h.endswith(".cuh")])
        if not os.path.exists(so) or os.path.getmtime(so) < newest:
            if int(os.environ.get("RANK", "0")) == 0:
                subprocess.run(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-o", so + ".tmp", src], check=True)
                os.replace(so + ".tmp", so)
            dist.barrier()
        self.lib = ctypes.CDLL(so)
        self.device = torch.device("cpu")

    def codec_id(self, name):
        c = CODEC_BY_NAME[name]
        wi = {1: 0, 2: 1, 3: 2, 4: 3, 6: 4, 8: 5}[c.W]
        return wi * 8 + (4 if (c.align or c.W == 1) else 0) + c.variant

    def workspace_size(self, codec, nbytes):
        return 256

    def phase(self, job, k):
        rc = self.lib.sim_slice_compress_phase(ctypes.byref(job), k, None)
        assert rc == 0, (k, rc)


def inputs(which):
    """(label, bytes) test inputs; cuts fall at multiples of 128 KiB."""
    A = 128 * 1024
    rng = np.random.default_rng(99)
    out = []
    out.append(("dct", gen_dct(3 * A + 12345, seed=5)))
    out.append(("mixed", gen_run_mixed(4 * A + 777, seed=8, max_run_log2=15, max_lit_log2=13)))
    x = gen_fuzz(rng, 2 * A + 5000, long_every=7)
    out.append(("fuzz", x))
    # a run that spans a cut (and one that spans a whole slice), cuts inside literals, a slice without any run
    y = rng.integers(0, 256, 4 * A + 100, dtype=np.uint8)
    y[A - 1000:A + 3000] = 7
    y[2 * A - 3:2 * A + 2] = 9
    out.append(("span", y))
    z = rng.integers(0, 256, 4 * A, dtype=np.uint8)
    z[A // 2:3 * A + 17] = 0
    out.append(("wide", z))
    out.append(("zeros", np.zeros(3 * A + 1, dtype=np.uint8)))
    out.append(("random", rng.integers(0, 256, 3 * A + 9, dtype=np.uint8)))
    for W in (2, 3, 8):
        p = rng.integers(0, 256, 4 * A + 31, dtype=np.uint8)
        pat = rng.integers(0, 256, W, dtype=np.uint8)
        for c in (A, 2 * A, 3 * A):                     # period-W patterns across every cut, at every phase
            s0 = c - 5 * W - int(rng.integers(0, W))
            ln = 11 * W + int(rng.integers(0, 2 * W))
            p[s0:s0 + ln] = np.resize(pat, ln)
        out.append((f"pat{W}", p))
    out.append(("short", gen_short_runs(3 * A + 50, seed=11, W=1)))
    if which == "quick":
        return out[:5]
    return out


def main():
    mode, codecs, which = sys.argv[1], sys.argv[2].split(","), sys.argv[3]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if mode == "gpu":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        engine = None
    elif mode == "gpu1":       # all ranks share cuda:0; the messages travel over gloo (NCCL needs one GPU per rank)
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
        engine = None
    else:
        dist.init_process_group("gloo")
        engine = SimEngine()
    from hsrle_b200 import sliced
    bad = 0
    ins = inputs(which)
    if mode != "sim" and which == "all":
        # slices of tens of MB: the scan picks its tile size per call from the slice length (hsrle_api.cu: enc_scan_steps) -- sizes
        # that do not divide the slice made the last tile reach over the cut (found by the 4-GPU run of the 1-GiB frame)
        ins.append(("big", gen_run_mixed(world * (24 << 20) + 4321, seed=9)))
    for label, data in ins:
        n = len(data)
        for name in codecs:
            enc = sliced.SlicedEncoder(name, n, engine=engine)
            dev = enc.engine.device
            sl = torch.from_numpy(data[enc.lo:enc.hi].copy()).to(dev)
            buf = enc.exchange_halos(enc.make_input(sl))
            part, off, total = enc.encode(buf)
            stream = sliced.gather_stream(part, total).cpu().numpy()
            if rank == 0:
                want = oracle_compress(CODEC_BY_NAME[name], data)
                ok = len(stream) == len(want) == total and np.array_equal(stream, want)
                if not ok:
                    bad += 1
                    d = np.nonzero(stream[:min(len(stream), len(want))] != want[:min(len(stream), len(want))])[0]
                    print(f"MISMATCH {name} {label} n={n} world={world}: got {len(stream)} want {len(want)} total {total} first diff {d[:3]}", flush=True)
    t = torch.tensor([bad], dtype=torch.int64, device="cuda" if mode == "gpu" else "cpu")
    if mode != "sim" and rank == 0:
        import hsrle_b200
        print("kernel launches:", hsrle_b200.kernel_launches(), flush=True)
    dist.broadcast(t, 0)
    dist.destroy_process_group()
    if rank == 0:
        print(f"sliced worker: mode={mode} world={world} codecs={len(codecs)} mismatches={bad}", flush=True)
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
