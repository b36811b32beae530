#!/usr/bin/env python
"""bench.py -- extreme-RLE encode+decode throughput on B200 (contract: DESIGN.md section 5).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload auto|configs1|configs3]

Workloads (BASELINE.json `configs`):
  configs1  (default at N = 1)  the 88,473,600-byte synthetic quantised-DCT stream encoded AND decoded with
            every codec of configs[0] (the four 8-bit codecs) and of the configs[1] matrix (16/24/32/48/64-bit symbols x
            {packed, 3LUT, 7LUT}, byte-aligned): 19 codecs.  With N > 1 (only on request) every rank runs it on its own copy:
            replicas, weak scaling, no data-path collective.
  configs3  (default at N > 1)  the north_star's multi-GPU split: a 16 GiB run-mixed stream cut into 16
            frames of 2^30 bytes (the largest input rle_compress_bounds accepts, src/rle8_extreme_cpu.c:22-28), the frames
            dealt round-robin to the ranks (strong scaling: N ranks do 1/N of ONE input), rle8_multi / rle64_byte /
            rle64_byte_packed encode + decode; every frame's stream is compared (length + sha256) with the compiled
            reference's stream for that frame inside the run.  Reported beside it: ONE 1 GiB frame encoded as contiguous
            slices by all N ranks (boundary-run fix-up + all-gathers, hsrle_b200.sliced), gathered and compared with the
            same reference stream; and the same frame sequence run by rank 0 alone (`n1_same_workload`), which is the
            single-GPU figure the N-rank value has to be divided by.

`value` = uncompressed bytes through encode + decode per second, device resident (CUDA events, max over ranks);
`e2e` = the same work through the reference-named host entry points with pinned host buffers, ONE host thread.
`--impl reference` times the compiled reference (oracle/_ref) on the host cores on the same workload / codec set:
`value` = one thread (the reference's design point, README.md:19), `all_cores` beside it.
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's banner / logs must not reach stdout: the contract is ONE JSON line
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "hypersonic-rle-kit_b200"))

import numpy as np  # noqa: E402

N_BYTES = 88473600
WORKLOAD1 = "configs[1]: 88,473,600-B synthetic quantised-DCT stream (SURVEY App. E.1), encode+decode"
CODEC_SET = ["rle8_multi", "rle8_packed_multi", "rle8_3symlut", "rle8_7symlut"] + \
            [f"rle{b}_{v}" for b in (16, 24, 32, 48, 64) for v in ("byte_packed", "3symlut_byte", "7symlut_byte")]
WORKLOAD3 = "configs[3]: 16 GiB synthetic run-mixed stream (SURVEY App. E.3) as 16 frames of 2^30 B, encode+decode"
CODEC_SET3 = ["rle8_multi", "rle64_byte", "rle64_byte_packed"]
SLICE_CODECS = ["rle8_multi", "rle64_byte_packed"]
METRIC = "encode+decode GB/s of uncompressed data, device-resident"
UNIT = "GB/s"
u8p = ctypes.POINTER(ctypes.c_uint8)


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


# ------------------------------------------------------------------------------------------------ CPU side

def cpu_info():
    """CPU model, core count and the ISA path the reference's dispatch takes on this host (src/simd_platform.c:73-157)."""
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    isa = "unknown"
    try:
        from common import ref_lib
        lib = ref_lib()
        if lib is not None:
            a2, a5 = lib.hsrle_ref_has_avx2(), lib.hsrle_ref_has_avx512f()
            isa = ("AVX2" if a2 else "SSE2") + " encode / " + ("AVX-512F" if a5 else "AVX2" if a2 else "SSE4.1") + " decode"
    except Exception:      # noqa: BLE001
        pass
    return {"cpu_model": model, "host_cores": os.cpu_count() or 1, "isa_path": isa,
            "build": "gcc -O3 -flto -mxsave -DNDEBUG (oracle/Makefile; mirrors project.lua Release)"}


def cpu_reference_lib():
    """The reference's own CPU implementation (oracle/_ref, built from /root/reference in the build container and shipped
    prebuilt), else the oracle port."""
    from common import CODEC_BY_NAME, oracle_lib, ref_lib
    lib = ref_lib()
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


def _codec_widths():
    from common import CODECS
    return {c.name: c.W for c in CODECS}


CODEC_W = _codec_widths()


def cpu_pass(enc, dec, inputs, codecs, threads):
    """One CPU pass: every (input, codec) pair is encoded + decoded once; `threads` workers, one call per worker at a time
    (ctypes releases the GIL).  Returns the wall time."""
    lock = threading.Lock()
    todo = [(i, c) for i in range(len(inputs)) for c in codecs]
    ok = [True]

    def work():
        bufs = {}
        while True:
            with lock:
                if not todo:
                    return
                i, name = todo.pop()
            data = inputs[i]
            n = len(data)
            if n not in bufs:
                bufs.clear()
                bufs[n] = (np.empty(n + 64, dtype=np.uint8), np.empty(n + n // 256 + 1024 + 256, dtype=np.uint8), np.empty(n + 256, dtype=np.uint8))
            src, comp, out = bufs[n]
            src[:n] = data
            src[n:] = 0
            W = CODEC_W[name]
            src[n] = (~int(data[n - W])) & 0xFF if n >= W else 0         # padding convention of SURVEY App. C.1
            r = enc(name, src, n, comp, n + n // 256 + 1024)
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
    """The reference's own CPU implementation on the GPU arm's workload, codec set, metric and unit.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = pick_workload(args)
    kind, enc, dec = cpu_reference_lib()
    info = cpu_info()
    cores = info["host_cores"]
    if workload == "configs1":
        from common import gen_dct
        inputs, codecs, wl = [gen_dct(N_BYTES)], CODEC_SET, WORKLOAD1
        sample = f"the full {N_BYTES}-B stream, all {len(codecs)} codecs encode+decode per step"
        cfg = {"workload": wl, "codecs": codecs, "bytes_per_codec": N_BYTES}
    else:
        nfr = max(1, min(args.ref_frames, args.frames))
        inputs = [host_frame(f, args) for f in range(nfr)]
        codecs, wl = CODEC_SET3, WORKLOAD3
        sample = (f"{nfr} of the {args.frames} frames ({args.frame_bytes} B each), all {len(codecs)} codecs encode+decode per step "
                  "(bounded sample of the workload)")
        cfg = {"workload": wl, "codecs": codecs, "frames": args.frames, "frame_bytes": args.frame_bytes, "frames_in_sample": nfr}
    work_bytes = 2.0 * sum(len(x) for x in inputs) * len(codecs)
    for _ in range(args.warmup):
        cpu_pass(enc, dec, inputs, codecs, 1)
    total = sum(cpu_pass(enc, dec, inputs, codecs, 1) for _ in range(args.steps))
    value = work_bytes * args.steps / total / 1e9
    # all host cores, one codec call per thread (context only: the reference is single-threaded by design)
    nth = max(1, min(cores, len(inputs) * len(codecs)))
    reps = max(1, min(args.steps, 3))
    cpu_pass(enc, dec, inputs, codecs, nth)
    tall = sum(cpu_pass(enc, dec, inputs, codecs, nth) for _ in range(reps))
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(1e3 * total / args.steps, 3), "higher_is_better": True,
            "scaling": "weak" if workload == "configs1" else "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
            "cpu_baseline": dict({"value": round(value, 4), "unit": UNIT, "cores": 1, "kind": kind, "sample": sample + ", 1 thread"}, **info),
            "all_cores": {"value": round(work_bytes * reps / tall / 1e9, 4), "unit": UNIT, "cores": nth,
                          "sample": sample + f", {nth} threads (one codec call per thread)"},
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            # same workload, codec set, metric and unit as the GPU arm's line for this --gpus / --workload (configs[3]: a bounded sample of its frames)
            "same_config": True}
    print(json.dumps(line), flush=True)


def host_frame(f, args):
    """Frame f of the configs[3] stream as a numpy array (generated on the GPU when one is there -- data generation is not
    the measured path -- else with torch on the CPU, which takes about a minute per GiB)."""
    import torch
    from common import RM_PIECE, gen_run_mixed_pieces
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if torch.cuda.is_available() else torch.device("cpu")
    pb = min(RM_PIECE, args.frame_bytes)
    ppf = args.frame_bytes // pb
    t = gen_run_mixed_pieces(f * ppf, ppf, dev, piece_bytes=pb)
    out = t.cpu().numpy()
    del t
    return out


# ------------------------------------------------------------------------------------------------ GPU arm, configs[1]

def parse_timing(buf):
    out = {}
    for part in buf.split(";"):
        if part:
            nm, cnt, ms = part.split(":")
            out[nm] = (int(cnt), float(ms))
    return out


def dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local, dev


def run_configs1(args):
    import torch
    import torch.distributed as dist
    import hsrle_b200 as hs
    from common import CODEC_BY_NAME, gen_dct

    world, rank, local, dev = dist_setup()
    n = N_BYTES
    data = gen_dct(n)
    with open(os.path.join(ROOT, "tests", "golden", "golden_hashes_88m.json")) as f:
        gold = json.load(f)
    assert hashlib.sha256(data.tobytes()).hexdigest() == gold["input_sha256"], "synthetic input differs from the golden fixture's"
    cap = n + n // 256 + 512
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    # inputs rotated over 4 distinct copies (352 MB > 126 MB L2) so no call finds its input in L2
    NCOPY = 4
    NSTREAM = args.streams if args.streams > 0 else 12
    t_in = [torch.from_numpy(data).to(dev) for _ in range(NCOPY)]
    t_comp = {c: torch.empty(cap, dtype=torch.uint8, device=dev) for c in CODEC_SET}
    ws_size = max(max(hs.compress_workspace_size(c, n) for c in CODEC_SET), max(hs.decompress_workspace_size(c, cap, n) for c in CODEC_SET))
    # the codec calls of a step are independent: they are dealt round-robin to NSTREAM CUDA streams (own workspace and
    # decode buffer each), so one call's latency-bound kernels overlap another call's bandwidth kernels
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

    # first pass: compressed sizes; every stream must be the compiled reference's, byte for byte (golden fixture)
    fork(); enqueue_step(0); join()
    torch.cuda.synchronize()
    for c in CODEC_SET:
        r = t_res[c].cpu().numpy()
        assert r[1] == 0 and r[0] > 0 and r[8] == n and r[9] == 0, (c, r)
        csize[c] = int(r[0])
        g = gold["streams"][c]
        assert csize[c] == g["len"], (c, csize[c], g["len"])
        assert hashlib.sha256(t_comp[c][:csize[c]].cpu().numpy().tobytes()).hexdigest() == g["sha256"], c + ": stream differs from the reference's"
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

    # ---- e2e: the reference-named host entry points with pinned host buffers (H2D + kernels + D2H timed).  Headline: ONE host
    # thread, like the reference's only caller (src/main.c:835,970); one call is H2D -> kernels -> D2H in sequence by data
    # dependence, so it is PCIe-bound.  Extra: NTHREAD host threads (the entry points are re-entrant), whose transfers overlap.
    NTHREAD = args.threads
    h_in = [torch.from_numpy(data).pin_memory() for _ in range(NTHREAD)]
    h_comp = [torch.empty(cap, dtype=torch.uint8).pin_memory() for _ in range(NTHREAD)]
    h_out = [torch.empty(n + 128, dtype=torch.uint8).pin_memory() for _ in range(NTHREAD)]
    fns = {}
    for c in CODEC_SET:
        cd = CODEC_BY_NAME[c]
        f = getattr(hs.lib, cd.cname); f.restype = ctypes.c_uint32; f.argtypes = [u8p, ctypes.c_uint32, u8p, ctypes.c_uint32]
        g = getattr(hs.lib, cd.dname); g.restype = ctypes.c_uint32; g.argtypes = [u8p, ctypes.c_uint32, u8p, ctypes.c_uint32]
        fns[c] = (f, g)

    def e2e_step(nthread):
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
        if nthread == 1:
            work(0)
        else:
            ts = [threading.Thread(target=work, args=(j,)) for j in range(nthread)]
            for t_ in ts:
                t_.start()
            for t_ in ts:
                t_.join()
        assert not errs, errs
        return moved[0], moved[1]

    def e2e_measure(nthread):
        steps = max(1, min(args.steps, 3))
        e2e_step(nthread)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            h2d, d2h = e2e_step(nthread)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        return world * bytes_per_step * steps / dt / 1e9, h2d, d2h, steps

    e2e_value, h2d, d2h, e2e_steps = e2e_measure(1)
    for j in range(1):
        assert np.array_equal(h_out[j][:n].numpy(), data)
    e2e_multi = e2e_measure(NTHREAD)[0] if NTHREAD > 1 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel CUDA-event timing on the launching stream (one extra pass, not the timed region)
    buf = ctypes.create_string_buffer(16384)
    hs.lib.hsrle_timing_begin()
    enqueue_step(0, on=stream)
    hs.lib.hsrle_timing_end(buf, 16384)
    kt = parse_timing(buf.value.decode())
    tot_ms = sum(v[1] for v in kt.values())
    top = max(kt.items(), key=lambda kv: kv[1][1])
    peak, peak_src = peaks()
    csum = sum(csize.values())
    cavg = csum / len(CODEC_SET)
    # algorithmic bytes of the dominant kernel per launch: the codec bytes that kernel has to move
    # once -- N = uncompressed bytes, C = mean compressed bytes over the codec set; state-only kernels are charged N + C
    alg = json.load(open(os.path.join(ROOT, "profiles", "kernel_alg_bytes.json"))) if os.path.exists(os.path.join(ROOT, "profiles", "kernel_alg_bytes.json")) else {}
    name = top[0]
    form = alg.get(name, "N+C")
    alg_bytes = float({"N": n, "C": cavg, "2C": 2 * cavg, "N+C": n + cavg}.get(form, n + cavg))
    avg_ms = top[1][1] / top[1][0]
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(name, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    whole = 2 * (n * len(CODEC_SET) + csum)
    roofline = {"bound": "hbm", "kernel": name, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes), "algorithmic_bytes_form": form,
                "kernel_share_of_step": round(top[1][1] / tot_ms, 4),
                "whole_pipeline": {"algorithmic_bytes_per_step": whole, "achieved": round(whole / (ms / args.steps * 1e-3) / 1e9, 1),
                                   "frac": round(whole / (ms / args.steps * 1e-3) / 1e9 / peak, 4), "note": f"{NSTREAM} CUDA streams overlapped"},
                "kernel_ms": {k: round(v[1], 4) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][1])},
                "kernel_launches": {k: v[0] for k, v in sorted(kt.items(), key=lambda kv: -kv[1][1])}}

    # ---- per-codec detail: ONE call at a time on one stream (device-resident, CUDA events, 3 reps each)
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
        detail[c] = {"ratio": round(csize[c] / n, 4), "enc_us": round(te * 1e3, 1), "dec_us": round(td * 1e3, 1),
                     "enc_GBps": round(n / te / 1e6, 1), "dec_GBps": round(n / td / 1e6, 1),
                     "enc_roofline_frac": round((n + csize[c]) / te / 1e6 / peak, 4), "dec_roofline_frac": round((n + csize[c]) / td / 1e6 / peak, 4)}
    if detail:
        te = sum(v["enc_us"] for v in detail.values())
        td = sum(v["dec_us"] for v in detail.values())
        roofline["single_call"] = {"note": "one call at a time on one stream, summed over the codec set",
                                   "achieved": round(whole / ((te + td) * 1e-6) / 1e9, 1), "frac": round(whole / ((te + td) * 1e-6) / 1e9 / peak, 4),
                                   "enc_frac": round((n * len(CODEC_SET) + csum) / (te * 1e-6) / 1e9 / peak, 4),
                                   "dec_frac": round((n * len(CODEC_SET) + csum) / (td * 1e-6) / 1e9 / peak, 4)}

    # ---- CPU baseline: the reference's single-threaded CPU path, full stream, bounded number of passes (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.quick:
        kind, enc, dec = cpu_reference_lib()
        cpu_pass(enc, dec, [data], CODEC_SET[:2], 1)     # warm the host caches / page in the library
        reps, dt = 0, 0.0
        while dt < 10.0 and reps < 64:                   # bounded sample: about 10 s of single-thread CPU work
            dt += cpu_pass(enc, dec, [data], CODEC_SET, 1)
            reps += 1
        cpu = dict({"value": round(2.0 * n * len(CODEC_SET) * reps / dt / 1e9, 4), "unit": UNIT, "cores": 1, "kind": kind,
                    "sample": f"the full {n}-B stream, all {len(CODEC_SET)} codecs encode+decode, {reps} passes, 1 thread ({dt:.1f} s)"}, **cpu_info())

    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD1, "codecs": CODEC_SET, "bytes_per_codec": n, "streams": NSTREAM,
                       "sharding": "none (one GPU)" if world == 1 else "replicas: every rank runs the full workload on its own copy (labelled extra; the north_star split is --workload configs3)",
                       "parity": "every compressed stream equals the compiled reference's (length + sha256, tests/golden/golden_hashes_88m.json); decode(encode(x)) == x",
                       "l2": "inputs rotated over 4 distinct 88 MB copies (352 MB > 126 MB L2); 19 distinct compressed buffers"},
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "reference-named host entry points (rleNN_*_compress/_decompress), pinned host buffers", "steps": e2e_steps,
                    "host_threads": 1, "with_host_threads": {str(NTHREAD): round(e2e_multi, 3)} if e2e_multi else None},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "per_codec": detail}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ GPU arm, configs[3]

def reference_streams(frames_host, codecs, threads):
    """(len, sha256) of the compiled reference's stream for every (frame, codec): the in-run parity target."""
    from common import CODEC_BY_NAME, ref_compress, ref_lib, oracle_compress
    use_ref = ref_lib() is not None
    out = {}
    todo = [(f, c) for f in frames_host for c in codecs]
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                if not todo:
                    return
                f, c = todo.pop()
            s = (ref_compress if use_ref else oracle_compress)(CODEC_BY_NAME[c], frames_host[f])
            out[(f, c)] = (int(len(s)), hashlib.sha256(s.tobytes()).hexdigest())
    ts = [threading.Thread(target=work) for _ in range(max(1, threads))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return out, ("reference" if use_ref else "port")


def run_configs3(args):
    import torch
    import torch.distributed as dist
    import hsrle_b200 as hs
    from hsrle_b200 import frames as fr, sliced
    from common import RM_PIECE, gen_run_mixed_pieces

    world, rank, local, dev = dist_setup()
    F, FB = args.frames, args.frame_bytes
    pb = min(RM_PIECE, FB)
    assert FB % pb == 0 and FB <= (1 << 30)
    ppf = FB // pb
    total_bytes = F * FB
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_frames(ids):
        return {f: gen_run_mixed_pieces(f * ppf, ppf, dev, piece_bytes=pb) for f in ids}

    # ---- the frame sequence, dealt to the ranks; rank 0 additionally holds every frame for the single-GPU figure
    mine = fr.deal_frames(F, rank, world)
    t_frames = make_frames(range(F) if (rank == 0 and world > 1 and not args.no_n1) else mine)
    host_threads = max(1, (os.cpu_count() or 8) // world)
    refs, ref_kind = {}, "reference"
    t0 = time.time()
    for f in mine:        # reference streams of MY frames (one frame on the host at a time)
        r, ref_kind = reference_streams({f: t_frames[f].cpu().numpy()}, CODEC_SET3, min(host_threads, len(CODEC_SET3)))
        refs.update(r)
    ref_s = time.time() - t0

    # streams in flight per rank: every one holds a workspace of ~10 GB for a 2^30-byte frame (and rank 0 of a one-GPU run already holds
    # 16 input frames and 48 compressed ones)
    n_streams = args.streams if args.streams > 0 else (4 if len(mine) > 8 else 6)
    pool = fr.StreamPool(CODEC_SET3, FB, device=dev, streams=n_streams)

    def run_sequence(ids, steps, check):
        """encode + decode of the frames `ids` with every codec, `steps` times; returns ms (CUDA events on the launching
        stream).  check: compare every stream with the reference's and every decoded frame with its input."""
        codecs = {c: fr.FrameCodec(c, [FB] * len(ids), pool=pool) for c in CODEC_SET3}
        outs_s = [torch.empty(FB + 128, dtype=torch.uint8, device=dev) for _ in pool.streams]     # one output buffer per pool stream
        outs = outs_s[:max(1, min(len(ids), len(pool.streams)))]
        ins = [t_frames[f] for f in ids]

        def step(verify=False):
            for c, fc in codecs.items():
                if not verify:
                    break
                fc.encode_async(ins)
                if verify:
                    sizes = fc.finish_encode()
                    for i, f in enumerate(ids):
                        if (f, c) in refs:
                            want = refs[(f, c)]
                            assert sizes[i] == want[0], (c, f, sizes[i], want[0])
                            got = hashlib.sha256(fc.stream(i).cpu().numpy().tobytes()).hexdigest()
                            assert got == want[1], f"{c}: frame {f} differs from the {ref_kind} stream"
                    for i0 in range(0, len(ids), len(outs)):       # decode in groups of len(outs) and compare
                        grp = list(range(i0, min(i0 + len(outs), len(ids))))
                        fc.decode_async(outs, grp); fc.finish_decode(grp)
                        for k, i in enumerate(grp):
                            assert torch.equal(outs[k][:FB], ins[i]), f"{c}: frame {ids[i]} does not decode back"
            if not verify:
                # the timed form: every (codec, frame) pair is its own chain -- encode, then decode, on one pool stream -- and all chains of
                # the step overlap (one fork / join per step, not per codec and direction)
                pool.fork()
                for ci, fc in enumerate(codecs.values()):
                    fc.roundtrip_async(ins, outs_s, offset=ci * max(1, len(pool.streams) // len(codecs)))
                pool.join()
        step(verify=check)
        for _ in range(max(args.warmup, 3) - 1):
            step()
        return codecs, step

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = hs.kernel_launches()
    codecs, step = run_sequence(mine, 0, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches1 = hs.kernel_launches()
    sampler.mark_begin()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = hs.kernel_launches() - launches1
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    work = 2.0 * total_bytes * len(CODEC_SET3)
    value = work * args.steps / (ms * 1e-3) / 1e9
    csum_mine = sum(refs[(f, c)][0] for f in mine for c in CODEC_SET3)
    csum_t = torch.tensor([csum_mine], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(csum_t)
    csum = float(csum_t.item())
    comp_mine = {c: sum(refs[(f, c)][0] for f in mine) for c in CODEC_SET3}
    del codecs, step

    # ---- e2e: the reference-named host entry points, pinned host buffers, one host thread per rank, MY frames
    from common import CODEC_BY_NAME
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty(FB, dtype=torch.uint8).pin_memory()
        h_comp = torch.empty(FB + FB // 256 + 1024, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(FB + 128, dtype=torch.uint8).pin_memory()
        h2d = d2h = 0
        dt = 0.0
        for f in mine:
            h_in.copy_(t_frames[f])
            torch.cuda.synchronize()
            for c in CODEC_SET3:
                cd = CODEC_BY_NAME[c]
                fe = getattr(hs.lib, cd.cname); fe.restype = ctypes.c_uint32; fe.argtypes = [u8p, ctypes.c_uint32, u8p, ctypes.c_uint32]
                fd = getattr(hs.lib, cd.dname); fd.restype = ctypes.c_uint32; fd.argtypes = [u8p, ctypes.c_uint32, u8p, ctypes.c_uint32]
                t1 = time.perf_counter()
                r = fe(ctypes.cast(h_in.data_ptr(), u8p), FB, ctypes.cast(h_comp.data_ptr(), u8p), h_comp.numel())
                d = fd(ctypes.cast(h_comp.data_ptr(), u8p), r, ctypes.cast(h_out.data_ptr(), u8p), FB + 128)
                dt += time.perf_counter() - t1
                assert r == refs[(f, c)][0] and d == FB, (c, f, r, d, hs.last_error())
                h2d += FB + r
                d2h += r + FB
            assert torch.equal(h_out[:FB], h_in)
        tt = torch.tensor([dt, float(h2d), float(d2h)], dtype=torch.float64, device=dev)
        if world > 1:
            mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt)
            dt = float(mx[0].item())
        e2e = {"value": round(work / dt / 1e9, 3) if dt > 0 else None, "unit": UNIT, "h2d_bytes_per_step": int(tt[1].item()),
               "d2h_bytes_per_step": int(tt[2].item()), "api": "reference-named host entry points, pinned host buffers, one frame per call",
               "steps": 1, "host_threads": 1}
        del h_in, h_comp, h_out

    # ---- one 1-frame stream encoded as contiguous slices by ALL ranks (boundary-run fix-up + all-gathers)
    sliced_out = None
    if world > 1 and not args.no_slices:
        sliced_out = {}
        for c in SLICE_CODECS:
            enc = sliced.SlicedEncoder(c, FB)
            lo, hi = enc.lo, enc.hi
            p0, p1 = lo // pb, -(-hi // pb)
            if hi > lo:
                blk = t_frames[0][lo:hi] if 0 in t_frames else gen_run_mixed_pieces(p0, p1 - p0, dev, piece_bytes=pb)[lo - p0 * pb:hi - p0 * pb]
            else:
                blk = torch.empty(0, dtype=torch.uint8, device=dev)
            buf = enc.exchange_halos(enc.make_input(blk))
            for _ in range(3):
                part, off, tot = enc.encode(buf)
            reps = max(1, min(args.steps, 5))
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                part, off, tot = enc.encode(buf)
            b.record(stream)
            barrier()
            tms = a.elapsed_time(b) / reps
            tm = torch.tensor([tms], dtype=torch.float64, device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            whole = sliced.gather_stream(part, tot)
            ok = None
            if rank == 0:
                want = refs[(0, c)]
                ok = (whole.numel() == want[0] and hashlib.sha256(whole.cpu().numpy().tobytes()).hexdigest() == want[1])
                assert ok, f"{c}: sliced stream differs from the {ref_kind} stream of frame 0"
            sliced_out[c] = {"encode_ms": round(float(tm.item()), 4), "encode_GBps": round(FB / float(tm.item()) / 1e6, 1), "state_rounds": enc.state_rounds,
                             "stream_bytes": int(tot), "equals_reference_stream": ok}
            del enc, buf, whole

    # ---- the same frame sequence on rank 0 alone: the single-GPU figure of this run (other ranks wait)
    n1 = None
    if world > 1 and not args.no_n1:
        if rank == 0:
            reps = max(1, min(args.steps, 3))
            torch.cuda.empty_cache()               # (the timed run's compressed frames were released above: room for all 16 frames' on this rank)
            _, step1 = run_sequence(list(range(F)), 0, False)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                step1()
            b.record(stream)
            torch.cuda.synchronize()
            m1 = a.elapsed_time(b) / reps
            n1 = {"value": round(work / (m1 * 1e-3) / 1e9, 2), "unit": UNIT, "ms_per_step": round(m1, 3), "steps": reps,
                  "note": "rank 0 runs all frames alone (same box, same run); divide `value` by N x this for the strong-scaling efficiency"}
        barrier()

    if rank == 0:
        peak, peak_src = peaks()
        alg = work + 2 * csum                    # encode reads N writes C, decode reads C writes N: 2 (N + C) per frame and codec
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": {"workload": WORKLOAD3 if (F, FB) == (16, 1 << 30) else f"configs[3] shape, reduced: {F} frames of {FB} B", "codecs": CODEC_SET3,
                           "frames": F, "frame_bytes": FB, "streams": len(pool.streams),
                           "sharding": f"frame sequence: frame f -> rank f mod {world} (independent reference-identical streams, no data-path collective); "
                                       "beside it `one_stream_slices`: one frame cut into contiguous slices over all ranks",
                           "parity": f"every frame's stream == the compiled {ref_kind}'s stream for that frame (length + sha256, computed on this box's host "
                                     f"cores in {ref_s:.0f} s on rank 0); every frame decodes back to its input",
                           "l2": f"every rank cycles through {len(mine)} x {FB} B of input per codec (>> 126 MB L2)"},
                "e2e": e2e, "gpu_launches": int(launches), "launches_first_pass": int(launches1 - launches0), "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "whole step (all kernels)", "achieved": round(alg * args.steps / (ms * 1e-3) / 1e9 / world, 1),
                             "peak": peak, "unit": "GB/s", "frac": round(alg * args.steps / (ms * 1e-3) / 1e9 / world / peak, 4), "traffic": None,
                             "peak_source": peak_src, "note": "per GPU: 2 (N + C) algorithmic bytes per frame and codec / step time"},
                "compressed_bytes_rank0": comp_mine, "one_stream_slices": sliced_out, "n1_same_workload": n1, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def pick_workload(args):
    if args.workload != "auto":
        return args.workload
    return "configs1" if int(os.environ.get("WORLD_SIZE", str(args.gpus))) <= 1 and args.gpus <= 1 else "configs3"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "configs1", "configs3"])
    ap.add_argument("--quick", action="store_true", help="configs1: skip the CPU baseline and the per-codec detail (sweeps)")
    ap.add_argument("--streams", type=int, default=0, help="CUDA streams the independent codec calls of a step are dealt to (default: 12 for configs[1], 8 for configs[3])")
    ap.add_argument("--threads", type=int, default=4, help="configs1: host threads of the extra (multi-threaded) e2e figure")
    ap.add_argument("--frames", type=int, default=16, help="configs3: frames in the sequence")
    ap.add_argument("--frame-bytes", type=int, default=1 << 30, help="configs3: bytes per frame (<= 2^30)")
    ap.add_argument("--ref-frames", type=int, default=1, help="configs3 reference arm: frames in the bounded sample")
    ap.add_argument("--no-n1", action="store_true", help="configs3: skip the single-GPU run of the same workload on rank 0")
    ap.add_argument("--no-slices", action="store_true", help="configs3: skip the one-stream slicing leg")
    ap.add_argument("--no-e2e", action="store_true", help="configs3: skip the host-pointer leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif pick_workload(args) == "configs1":
        run_configs1(args)
    else:
        run_configs3(args)


if __name__ == "__main__":
    main()
