#!/usr/bin/env python
"""bench.py -- extreme-RLE encode+decode throughput on B200 (see the task contract in DESIGN.md §Measurement).

A "step" = one pass of the hot path over the BASELINE configs[1] workload: the 88,473,600-byte synthetic
quantised-DCT stream encoded AND decoded with every codec of the configs[1] matrix
(16/24/32/48/64-bit symbols x {packed, 3LUT, 7LUT}, byte-aligned) plus the two 8-bit headline codecs
of configs[0].  `value` = uncompressed bytes through encode + decode per second, device resident.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("NCCL_DEBUG", "WARN")     # NCCL's version banner goes to stdout; the contract is ONE JSON line
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))

import numpy as np  # noqa: E402

N_BYTES = 88473600
WORKLOAD = "configs[1]: 88,473,600-B synthetic quantised-DCT stream (SURVEY App. E.1), encode+decode"
CODEC_SET = ["rle8_multi", "rle8_packed_multi"] + [f"rle{b}_{v}" for b in (16, 24, 32, 48, 64)
                                                    for v in ("byte_packed", "3symlut_byte", "7symlut_byte")]
METRIC = "encode+decode GB/s of uncompressed data, device-resident"
UNIT = "GB/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
        rows = [r for (ts, r) in self.rows if self.t0 is not None and self.t0 - 0.02 <= ts <= (self.t1 or ts) + 0.04]
        window = "timed region"
        if len(rows) < 2:     # a very short timed region: fall back to everything since the sampler started (warm-up included)
            rows, window = [r for (_, r) in self.rows], "warm-up + timed region"
        sm = sorted(int(float(r[1])) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------ CPU arm

def cpu_reference_lib():
    """The reference's own CPU implementation (oracle/_ref, built from /root/reference in the build
    container and shipped prebuilt), else the oracle port."""
    from common import CODEC_BY_NAME, oracle_lib, ref_lib
    lib = ref_lib()
    u8p = ctypes.POINTER(ctypes.c_uint8)
    if lib is not None:
        def enc(name, src, n, dst, cap):
            return getattr(lib, CODEC_BY_NAME[name].cname)(src.ctypes.data_as(u8p), n, dst.ctypes.data_as(u8p), cap)

        def dec(name, src, sz, dst, cap):
            return getattr(lib, CODEC_BY_NAME[name].dname)(src.ctypes.data_as(u8p), sz, dst.ctypes.data_as(u8p), cap)
        return "reference", enc, dec
    ol = oracle_lib()

    def enc(name, src, n, dst, cap):
        c = CODEC_BY_NAME[name]
        return ol.oracle_compress(c.W, c.align, c.variant, src.ctypes.data_as(u8p), n, dst.ctypes.data_as(u8p), cap)

    def dec(name, src, sz, dst, cap):
        c = CODEC_BY_NAME[name]
        return ol.oracle_decompress(c.W, c.align, c.variant, src.ctypes.data_as(u8p), sz, dst.ctypes.data_as(u8p), cap)
    return "port", enc, dec


def cpu_step(enc, dec, data, codecs, threads):
    """One CPU pass: every codec encodes + decodes `data`; one codec per thread at a time."""
    n = len(data)
    cap = n + n // 256 + 1024
    lock = threading.Lock()
    todo = list(codecs)
    ok = [True]

    def work():
        src = np.empty(n + 64, dtype=np.uint8)
        src[:n] = data
        src[n:] = 0
        comp = np.empty(cap + 256, dtype=np.uint8)
        out = np.empty(n + 256, dtype=np.uint8)
        while True:
            with lock:
                if not todo:
                    return
                name = todo.pop()
            r = enc(name, src, n, comp, cap)
            d = dec(name, comp, r, out, n + 128)
            if r == 0 or d != n:
                ok[0] = False

    ts = [threading.Thread(target=work) for _ in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    assert ok[0], "CPU reference round trip failed"
    return dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from common import gen_dct
    kind, enc, dec = cpu_reference_lib()
    cores = os.cpu_count() or 1
    threads = min(cores, len(CODEC_SET))
    sample_n = N_BYTES // 4          # bounded sample: first quarter of the stream, every codec
    data = gen_dct(N_BYTES)[:sample_n].copy()
    for _ in range(args.warmup):
        cpu_step(enc, dec, data, CODEC_SET, threads)
    times = [cpu_step(enc, dec, data, CODEC_SET, threads) for _ in range(args.steps)]
    total = sum(times)
    value = 2.0 * sample_n * len(CODEC_SET) * args.steps / total / 1e9
    sample = f"first {sample_n} B of the stream, all {len(CODEC_SET)} codecs encode+decode per step, {threads} threads (one codec call per thread)"
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(1e3 * total / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "codecs": CODEC_SET, "bytes_per_codec": sample_n},
            "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm

def parse_timing(buf):
    out = {}
    for part in buf.split(";"):
        if part:
            nm, cnt, ms = part.split(":")
            out[nm] = (int(cnt), float(ms))
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import hsrle_b200 as hs
    from common import gen_dct

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = N_BYTES
    data = gen_dct(n)
    cap = n + n // 256 + 512
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    # inputs rotated over 4 distinct copies (352 MB > 126 MB L2) so no call finds its input in L2
    NCOPY = 4
    NSTREAM = args.streams
    t_in = [torch.from_numpy(data).to(dev) for _ in range(NCOPY)]
    t_comp = {c: torch.empty(cap, dtype=torch.uint8, device=dev) for c in CODEC_SET}
    ws_size = max(max(hs.compress_workspace_size(c, n) for c in CODEC_SET), max(hs.decompress_workspace_size(c, cap, n) for c in CODEC_SET))
    # the codec calls of a step are independent: they are dealt round-robin to NSTREAM CUDA streams (own workspace and
    # decode buffer each), so one call's latency-bound resolve/scan kernels overlap another call's bandwidth kernels
    side = [torch.cuda.Stream(device=dev) for _ in range(NSTREAM)]
    t_ws = [torch.empty(ws_size, dtype=torch.uint8, device=dev) for _ in range(NSTREAM)]
    t_dec = [torch.empty(n + 128, dtype=torch.uint8, device=dev) for _ in range(NSTREAM)]
    t_res = {c: torch.zeros(16, dtype=torch.int32, device=dev) for c in CODEC_SET}
    csize = {}

    def enqueue_step(k, on=None):
        for i, c in enumerate(CODEC_SET):
            j = i % NSTREAM
            q = (on if on is not None else side[j]).cuda_stream
            hs.compress_device_async(c, t_in[(k + i) % NCOPY], t_comp[c], t_ws[j], t_res[c][:8], q)
            hs.decompress_device_async(c, t_comp[c], csize.get(c, cap), t_dec[j], n, t_ws[j], t_res[c][8:], q)

    def fork():
        ev = torch.cuda.Event()
        ev.record(stream)
        for s_ in side:
            s_.wait_event(ev)

    def join():
        for s_ in side:
            ev = torch.cuda.Event()
            ev.record(s_)
            stream.wait_event(ev)

    # first pass: learn the (deterministic) compressed sizes, check correctness
    fork(); enqueue_step(0); join()
    torch.cuda.synchronize()
    for c in CODEC_SET:
        r = t_res[c].cpu().numpy()
        assert r[1] == 0 and r[0] > 0 and r[8] == n and r[9] == 0, (c, r)
        csize[c] = int(r[0])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    fork()
    for w in range(max(args.warmup, 3)):
        enqueue_step(w)
    join()
    torch.cuda.synchronize()
    for j in range(NSTREAM):
        assert torch.equal(t_dec[j][:n], t_in[0]), "decode(encode(x)) != x"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = hs.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record(stream)
    fork()
    for k in range(args.steps):
        enqueue_step(k)
    join()
    e1.record(stream)
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = hs.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    bytes_per_step = 2.0 * n * len(CODEC_SET)
    value = world * bytes_per_step * args.steps / (ms * 1e-3) / 1e9

    # ---- e2e: the reference-named host entry points with pinned host buffers (H2D + kernels + D2H timed).
    # The entry points are re-entrant like the reference's; NTHREAD host threads each call them on their own pinned
    # buffers (one codec per call), so one call's H2D overlaps another's kernels and D2H (PCIe is full duplex).
    NTHREAD = args.threads
    h_in = [torch.from_numpy(data).pin_memory() for _ in range(NTHREAD)]
    h_comp = [torch.empty(cap, dtype=torch.uint8).pin_memory() for _ in range(NTHREAD)]
    h_out = [torch.empty(n + 128, dtype=torch.uint8).pin_memory() for _ in range(NTHREAD)]
    u8p = ctypes.POINTER(ctypes.c_uint8)
    from common import CODEC_BY_NAME
    fns = {}
    for c in CODEC_SET:
        cd = CODEC_BY_NAME[c]
        f = getattr(hs.lib, cd.cname); f.restype = ctypes.c_uint32; f.argtypes = [u8p, ctypes.c_uint32, u8p, ctypes.c_uint32]
        g = getattr(hs.lib, cd.dname); g.restype = ctypes.c_uint32; g.argtypes = [u8p, ctypes.c_uint32, u8p, ctypes.c_uint32]
        fns[c] = (f, g)

    def e2e_step():
        todo = list(CODEC_SET)
        lock = threading.Lock()
        moved = [0, 0]
        errs = []

        def work(j):
            torch.cuda.set_device(local)
            while True:
                with lock:
                    if not todo:
                        return
                    c = todo.pop()
                f, g = fns[c]
                r = f(ctypes.cast(h_in[j].data_ptr(), u8p), n, ctypes.cast(h_comp[j].data_ptr(), u8p), cap)
                d = g(ctypes.cast(h_comp[j].data_ptr(), u8p), r, ctypes.cast(h_out[j].data_ptr(), u8p), n + 128)
                if r != csize[c] or d != n:
                    errs.append((c, r, d, hs.last_error()))
                with lock:
                    moved[0] += n + r
                    moved[1] += r + n
        ts = [threading.Thread(target=work, args=(j,)) for j in range(NTHREAD)]
        for t_ in ts:
            t_.start()
        for t_ in ts:
            t_.join()
        assert not errs, errs
        return moved[0], moved[1]

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h2d, d2h = e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_value = world * bytes_per_step * e2e_steps / dt / 1e9
    for j in range(NTHREAD):
        assert np.array_equal(h_out[j][:n].numpy(), data)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel CUDA-event timing on the launching stream (one extra pass, not the timed region)
    buf = ctypes.create_string_buffer(8192)
    hs.lib.hsrle_timing_begin()
    enqueue_step(0, on=stream)
    hs.lib.hsrle_timing_end(buf, 8192)
    kt = parse_timing(buf.value.decode())
    tot_ms = sum(v[1] for v in kt.values())
    top = max(kt.items(), key=lambda kv: kv[1][1])
    peak, peak_src = peaks()
    csum = sum(csize.values())
    # algorithmic bytes of the dominant kernel per launch (DESIGN.md, section "Kernels"): the codec bytes that kernel has
    # to move once -- N = uncompressed bytes, C = mean compressed bytes over the codec set
    cavg = csum / len(CODEC_SET)
    alg = {"k_enc_scan": n, "k_enc_auto": None, "k_enc_emit": 2 * cavg, "k_enc_copy_big": None,
           "k_dec_map": cavg, "k_dec_chain": None, "k_dec_emit": n + cavg, "k_dec_big": None}
    name = top[0]
    alg_bytes = float(alg.get(name) or (n + cavg))     # state-only kernels are charged the whole call (N + C)
    avg_ms = top[1][1] / top[1][0]
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/), if one exists
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(name, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roofline = {"bound": "hbm", "kernel": name, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": peak_src, "kernel_share_of_step": round(top[1][1] / tot_ms, 4),
                "whole_pipeline": {"algorithmic_bytes_per_step": 2 * (n * len(CODEC_SET) + csum),
                                   "achieved": round(2 * (n * len(CODEC_SET) + csum) / (ms / args.steps * 1e-3) / 1e9, 1),
                                   "frac": round(2 * (n * len(CODEC_SET) + csum) / (ms / args.steps * 1e-3) / 1e9 / peak, 4)},
                "kernel_ms": {k: round(v[1], 4) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][1])}}

    # ---- per-codec detail (device-resident, CUDA events, 3 reps each)
    detail = {}
    for i, c in enumerate([] if args.quick else CODEC_SET):
        a, b, d = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        reps = 3
        torch.cuda.synchronize()
        a.record(stream)
        for k in range(reps):
            hs.compress_device_async(c, t_in[k % NCOPY], t_comp[c], t_ws[0], t_res[c][:8], sp)
        b.record(stream)
        for k in range(reps):
            hs.decompress_device_async(c, t_comp[c], csize[c], t_dec[k % NSTREAM], n, t_ws[0], t_res[c][8:], sp)
        d.record(stream)
        torch.cuda.synchronize()
        te, td = a.elapsed_time(b) / reps, b.elapsed_time(d) / reps
        detail[c] = {"ratio": round(csize[c] / n, 4), "enc_GBps": round(n / te / 1e6, 1), "dec_GBps": round(n / td / 1e6, 1),
                     "enc_roofline_frac": round((n + csize[c]) / te / 1e6 / peak, 4), "dec_roofline_frac": round((n + csize[c]) / td / 1e6 / peak, 4)}

    # ---- CPU baseline: the reference's single-threaded CPU path on a bounded sample (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.quick:
        kind, enc, dec = cpu_reference_lib()
        cpu_step(enc, dec, data, CODEC_SET[:2], 1)     # warm the host caches / page in the library
        reps, dt = 0, 0.0
        while dt < 10.0 and reps < 64:                 # bounded sample: about 10 s of single-thread CPU work
            dt += cpu_step(enc, dec, data, CODEC_SET, 1)
            reps += 1
        cpu = {"value": round(2.0 * n * len(CODEC_SET) * reps / dt / 1e9, 4), "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"the full {n}-B stream, all {len(CODEC_SET)} codecs encode+decode, {reps} passes, 1 thread ({dt:.1f} s)"}

    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "codecs": CODEC_SET, "bytes_per_codec": n, "per_gpu": "every rank runs the full workload on its own copy", "streams": NSTREAM,
                       "l2": "inputs rotated over 4 distinct 88 MB copies (352 MB > 126 MB L2); 17 distinct compressed buffers"},
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "reference-named host entry points (rleNN_*_compress/_decompress), pinned host buffers", "steps": e2e_steps,
                    "host_threads": NTHREAD},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "per_codec": detail}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--quick", action="store_true", help="skip the CPU baseline and the per-codec detail (sweeps)")
    ap.add_argument("--streams", type=int, default=8, help="CUDA streams the independent codec calls of a step are dealt to")
    ap.add_argument("--threads", type=int, default=4, help="host threads calling the host-pointer entry points in the e2e leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
