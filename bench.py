#!/usr/bin/env python3
"""bench.py — headline benchmark of the two hot paths (BASELINE.json: "LZ4 compress GB/s & JPEG encode MPix/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by the driver as ``python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N``:
one process per GPU, weak scaling (every rank compresses its own shard of the same shape), no data-path
collective, ONE all-gather of the per-rank compressed byte totals per step.

A "step" is one pass of the hot path over one batch of synthetic input:
  LZ4  (headline `value`): BASELINE.json configs[2] — 4 GiB random_extract-style text, 64 KiB blocks, per GPU
  JPEG (`jpeg` object)   : BASELINE.json configs[3] — one 16384 x 16384 random_image-style RGBA image, per GPU
  JFIF (`jfif` object)   : the same image through the true baseline-JPEG encoder (byte-identical to the stb_image_write.h
                           the reference vendors) at quality 75: 4:4:4 as BASELINE.json words it, and stb's own 4:2:0
`value` is measured with inputs resident in HBM (CUDA events, max over ranks); `e2e` is the same work
through the C ABI's host-buffer entry point with pinned host buffers, H2D and D2H inside the timed region.
Inputs are far larger than the 126 MB L2, so no explicit L2 flush is needed between iterations.

``--impl reference`` times the reference's own CPU code (oracle/_ref, compiled from /root/reference; falls
back to this repo's C port of it) on all host threads over a bounded sample of the same workload.

Every object of the GPU line also carries ``parity_sample``: after the timed region, randomly chosen units of the
benchmark's own output (64 LZ4 blocks, 4096 JPEG groups; for JFIF a 64-row band re-encoded alone) are compared with the
CPU oracle — the one use of oracle/ in the GPU arm besides ``cpu_baseline``, and never inside a timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = 1 << 30
BLOCK_LEN = 65536
LZ4_BYTES = int(os.environ.get("LJB_BENCH_LZ4_BYTES", 4 * GIB))
JPEG_DIM = int(os.environ.get("LJB_BENCH_JPEG_DIM", 16384))
# dram__bytes_read.sum + dram__bytes_write.sum per launch, from the ncu pass over this same command committed as
# profiles/r2/launches_r2aa_bench_steps2_warmup1.csv (the full-size launches of every kernel); bytes at the default workload sizes,
# None for other sizes.
NCU_TRAFFIC = {"lz4": 4.374e9 + 5.915e9 if LZ4_BYTES == 4 * GIB else None, "lz4_decode": 45.6e9 + 4.49e9 if LZ4_BYTES == 4 * GIB else None,
               "jpeg": 1.086e9 + 0.504e9 if JPEG_DIM == 16384 else None,
               "jfif444": 1.074e9 + 0.318e9 if JPEG_DIM == 16384 else None, "jfif420": 1.074e9 + 0.144e9 if JPEG_DIM == 16384 else None}
JFIF_QUALITY = 75
# frames of 1920x1080 per GPU in the batch object (BASELINE configs[4]: 8192 frames across 8 GPUs); 0 skips it
BATCH_IMAGES = int(os.environ.get("LJB_BENCH_BATCH_IMAGES", 1024))


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        smax = max([int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()] or [0])
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline
# ---------------------------------------------------------------------------------------------------------
def _cpu_lz4(sample_blocks: int, threads: int):
    """Reference block_encode over `sample_blocks` 64 KiB blocks of the benchmark text on `threads` host threads."""
    import numpy as np

    from oracle.pyoracle import Oracle, Ref

    orc = Oracle()
    corpus = np.fromfile(os.path.join(ROOT, "tests", "golden", "Metamorphosis.txt"), dtype=np.uint8)
    data = orc.synth_text(corpus, 42, 30000, sample_blocks * BLOCK_LEN)
    if Ref.available("lz4"):
        sec, _ = Ref("lz4").lz4_time_blocks(data, BLOCK_LEN, threads)
        kind = "reference"
    else:  # the C port of the same exhaustive algorithm (mode 0), one block per thread
        from concurrent.futures import ThreadPoolExecutor

        blocks = [data[i * BLOCK_LEN:(i + 1) * BLOCK_LEN] for i in range(sample_blocks)]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda b: orc.lz4_compress(b, BLOCK_LEN, 0), blocks))
        sec = time.perf_counter() - t0
        kind = "port"
    return data.size / sec / 1e9, kind, sec


def _cpu_lz4_detail(threads: int):
    """The three CPU figures BASELINE.json's north_star asks for: the sequential build on one core, the sequential per-block
    function on all threads (the `value` the reference arm reports), and the parallel build's thread body with its global
    locks (Algorithms/parallel/LZ4/LZ4.c:518-628) on all threads.  Bounded samples of the benchmark text."""
    import numpy as np

    from oracle.pyoracle import Oracle, Ref

    out = {}
    if not Ref.available("lz4"):
        return out
    orc = Oracle()
    corpus = np.fromfile(os.path.join(ROOT, "tests", "golden", "Metamorphosis.txt"), dtype=np.uint8)
    d1 = orc.synth_text(corpus, 42, 30000, 2 * BLOCK_LEN)
    sec, _ = Ref("lz4").lz4_time_blocks(d1, BLOCK_LEN, 1)
    out["sequential_1core"] = {"value": d1.size / sec / 1e9, "unit": "GB/s", "cores": 1, "kind": "reference",
                               "sample": f"2 blocks of 64 KiB of the same text, block_encode (Algorithms/sequential/LZ4/LZ4.c:506) on 1 thread, {sec:.1f} s"}
    if Ref.available("lz4_par"):
        sb = max(threads, 16)
        dp = orc.synth_text(corpus, 42, 30000, sb * BLOCK_LEN)
        sec, _ = Ref("lz4_par").lz4par_time_blocks(dp, BLOCK_LEN, threads)
        out["parallel_build"] = {"value": dp.size / sec / 1e9, "unit": "GB/s", "cores": threads, "kind": "reference",
                                 "sample": f"{sb} blocks of 64 KiB of the same text, parallel_block_encode with its global add_seq lock "
                                           f"(Algorithms/parallel/LZ4/LZ4.c:518-628, windows.h shim over pthreads) from a pool of {threads} "
                                           f"threads instead of one thread per block, {sec:.1f} s"}
    return out


def _cpu_jpeg_detail(threads: int):
    from oracle.pyoracle import Oracle, Ref

    out = {}
    if not Ref.available("jpeg"):
        return out
    orc = Oracle()
    img1 = orc.synth_image(42, 2048, 1024)
    sec, _ = Ref("jpeg").jpeg_time_groups(img1, 1)
    out["sequential_1core"] = {"value": 2048 * 1024 / sec / 1e6, "unit": "MPix/s", "cores": 1, "kind": "reference",
                               "sample": f"2048x1024 seed-42 noise image, the sequential build's per-group encode stages on 1 thread, {sec:.1f} s"}
    if Ref.available("jpeg_par"):
        imgp = orc.synth_image(42, 4096, 2048)
        sec = Ref("jpeg_par").jpegpar_time_groups(imgp, threads)
        out["parallel_build"] = {"value": 4096 * 2048 / sec / 1e6, "unit": "MPix/s", "cores": threads, "kind": "reference",
                                 "sample": f"4096x2048 seed-42 noise image, process() (Algorithms/parallel/JPEG/JPEG.c:1103: forward AND inverse chain, "
                                           f"results discarded as in the reference) from a pool of {threads} threads instead of one thread per group, {sec:.1f} s"}
    return out


def _cpu_jpeg(w: int, h: int, threads: int):
    import numpy as np

    from oracle.pyoracle import Oracle, Ref

    orc = Oracle()
    img = orc.synth_image(42, w, h)
    if Ref.available("jpeg"):
        sec, _ = Ref("jpeg").jpeg_time_groups(img, threads)
        kind = "reference"
    else:
        from concurrent.futures import ThreadPoolExecutor

        total = orc.jpeg_group_count(w, h)
        cuts = [total * i // threads for i in range(threads + 1)]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda i: orc.jpeg_encode(img, cuts[i], cuts[i + 1], want_coefs=False), range(threads)))
        sec = time.perf_counter() - t0
        kind = "port"
    return w * h / sec / 1e6, kind, sec


def _cpu_jfif(w: int, h: int, threads: int, subsample: int):
    """The reference's vendored stbi_write_jpg (oracle/_ref) on `threads` host threads, one whole image each; falls back
    to this repo's C restatement of it."""
    from oracle.pyoracle import Oracle, Ref

    orc = Oracle()
    img = orc.synth_image(42, w, h)
    if Ref.available("jfif"):
        reps = max(1, (64 << 20) // (w * h))  # ~5-10 s of stb on every thread
        sec, _ = Ref("jfif").jfif_time_mt(img, JFIF_QUALITY, subsample, threads, reps)
        return reps * threads * w * h / sec / 1e6, "reference", sec
    from concurrent.futures import ThreadPoolExecutor

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda i: orc.jfif_encode(img, JFIF_QUALITY, subsample), range(threads)))
    sec = time.perf_counter() - t0
    return threads * w * h / sec / 1e6, "port", sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample_blocks = max(threads, 16)
    jw, jh = 4096, 2048  # ~25 CPU-seconds of the reference's per-group stages
    lz, jp, jf4, jf2 = [], [], [], []
    kind = "reference"
    for i in range(args.warmup + args.steps):
        v, kind, _ = _cpu_lz4(sample_blocks, threads)
        vj, kindj, _ = _cpu_jpeg(jw, jh, threads)
        v4, kindf, _ = _cpu_jfif(2048, 2048, threads, 0)
        v2, _, _ = _cpu_jfif(2048, 2048, threads, -1)
        if i >= args.warmup:
            lz.append(v)
            jp.append(vj)
            jf4.append(v4)
            jf2.append(v2)
    value = sum(lz) / len(lz)
    jvalue = sum(jp) / len(jp)
    f4value, f2value = sum(jf4) / len(jf4), sum(jf2) / len(jf2)
    sample = f"{sample_blocks} blocks of 64 KiB of the seed-42 random_extract text, block_encode on {threads} threads"
    lz_detail, jp_detail = _cpu_lz4_detail(threads), _cpu_jpeg_detail(threads)  # once, outside the step loop
    line = {
        "impl": "reference", "metric": "LZ4 compress GB/s (headline) & JPEG encode MPix/s (jpeg)", "value": value, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sample_blocks * BLOCK_LEN / (value * 1e9), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "LZ4 block compression of 4 GiB random_extract-style text, 64 KiB blocks (BASELINE configs[2])",
                   "block_len": BLOCK_LEN, "sampled": sample},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": kind, "sample": sample, **lz_detail},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "jpeg": {"metric": "JPEG encode MPix/s", "value": jvalue, "unit": "MPix/s",
                 "cpu_baseline": {"value": jvalue, "unit": "MPix/s", "cores": threads, "kind": kindj,
                                  "sample": f"{jw}x{jh} seed-42 noise image, per-group encode stages on {threads} threads", **jp_detail},
                 "e2e": {"value": jvalue, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
        "jfif": {"metric": "baseline JPEG (JFIF) encode MPix/s, quality 75, 4:4:4", "value": f4value, "unit": "MPix/s",
                 "cpu_baseline": {"value": f4value, "unit": "MPix/s", "cores": threads, "kind": kindf,
                                  "sample": f"2048x2048 seed-42 noise image, stbi_write_jpg quality {JFIF_QUALITY}, 16 images per thread on {threads} threads"},
                 "e2e": {"value": f4value, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "stb_rule_420": {"value": f2value, "unit": "MPix/s"}},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import lz4jpeg_b200 as ljb
    from lz4jpeg_b200 import _native as N

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # Pin this rank to the CPU cores next to its GPU before any host buffer is allocated and touched: the pinned buffers of
        # the end-to-end measurement then live on the GPU's own NUMA node (first touch) instead of all on node 0.
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
        except Exception:
            pass
        dist.init_process_group("nccl", device_id=dev)
    ctx = ljb.Context(local_rank)
    ext_stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    lib = N.lib()
    peak, peak_src = measured_peak_gbs()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ------------------------------- LZ4 -----------------------------------------------------------------
    n = LZ4_BYTES
    nblocks = (n + BLOCK_LEN - 1) // BLOCK_LEN
    cap = n + n // 8 + 16 * nblocks + 4096  # text compresses to ~0.77 n; LJB_E_CAPACITY is reported if ever exceeded
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    ljb.synth.random_extract(n, seed=42 + rank, out=h_in.numpy())
    h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    h_offs = torch.empty(nblocks + 1, dtype=torch.int64, pin_memory=True)
    d_in = h_in.to(dev)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_offs = torch.empty(nblocks + 1, dtype=torch.int64, device=dev)
    d_res = torch.zeros(3, dtype=torch.int64, device=dev)
    # this rank's shard of the global frame: rank r owns blocks [r*nblocks, (r+1)*nblocks)
    first_block, frame_blocks = rank * nblocks, world * nblocks

    def lz4_step_device():
        ljb.lz4.compress_device(d_in, BLOCK_LEN, d_out, d_offs, d_res, ctx, first_block=first_block, frame_blocks=frame_blocks)
        if world > 1:  # the single collective: all-gather of per-rank byte totals -> global base offsets
            with torch.cuda.stream(ext_stream):
                ljb.sharding.gather_totals_device(d_res[0:1])

    out_len = C.c_size_t(0)
    ph = C.c_uint64(0)

    def lz4_step_e2e():
        rc = lib.ljb_lz4_compress(ctx.handle, h_in.data_ptr(), n, BLOCK_LEN, h_out.data_ptr(), cap, h_offs.data_ptr(),
                                  C.byref(out_len), C.byref(ph))
        N.check(rc, "ljb_lz4_compress")
        if world > 1:
            ljb.sharding.gather_totals(int(out_len.value), device=dev)

    def timed(step_fn, steps, warmup, use_events=True):
        """W untimed warm-up steps, then exactly `steps` steps bracketed by barrier + synchronize on both sides.
        Device time = CUDA events recorded on the stream the kernels are launched on; max over ranks."""
        torch.cuda.synchronize()
        for _ in range(warmup):
            step_fn()
        barrier()
        launches0 = ctx.launch_count
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(ext_stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            step_fn()
        e1.record(ext_stream)
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = e0.elapsed_time(e1) if use_events else wall * 1e3
        return max_over_ranks(dev_ms), wall, ctx.launch_count - launches0

    lz_ms, _, lz_launches = timed(lz4_step_device, args.steps, args.warmup)
    # average duration of the kernel alone (CUDA events around the launch, on the launching stream)
    ks = []
    for _ in range(min(3, args.steps)):
        ljb.lz4.compress_device(d_in, BLOCK_LEN, d_out, d_offs, d_res, ctx, first_block=first_block, frame_blocks=frame_blocks)
        ks.append(ctx.last_kernel_ms())
    lz_kernel_ms = sum(ks) / len(ks)
    torch.cuda.synchronize()
    lz_out_bytes = int(d_res[0].item())
    lz_err = int(d_res[2].item())
    if lz_err:
        raise SystemExit(f"LZ4 kernel reported error flags {lz_err}")
    e2e_steps, e2e_warm = args.steps, max(1, args.warmup)  # the end-to-end path is timed over the same K steps (W >= 1 warm-up)
    e2e_ms, e2e_wall, _ = timed(lz4_step_e2e, e2e_steps, e2e_warm, use_events=False)
    lz_e2e_ms = max_over_ranks(e2e_wall * 1e3 / e2e_steps)
    # the host-buffer call encodes a whole frame (leading frame byte); the device call on rank > 0 encodes a shard without it
    assert int(out_len.value) == lz_out_bytes + (1 if first_block else 0), "host-buffer path and device path disagree on the stream length"

    total_in = sum_over_ranks(float(n))
    lz_value = total_in / (lz_ms / args.steps * 1e-3) / 1e9
    lz_e2e_value = total_in / (lz_e2e_ms * 1e-3) / 1e9
    lz_achieved = (n + lz_out_bytes) / (lz_kernel_ms * 1e-3) / 1e9

    # ---- what the host link can do at most: the same bytes, plain pinned copies both ways at once, all ranks together
    def host_ceiling(h_src, d_dst, d_src, h_dst, nbytes_out):
        s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        best = None
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s_up):
                d_dst.copy_(h_src, non_blocking=True)
            with torch.cuda.stream(s_dn):
                h_dst[:nbytes_out].copy_(d_src[:nbytes_out], non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            best = dt if best is None else min(best, dt)
        return best

    lz_ceiling_s = host_ceiling(h_in, d_in, d_out, h_out, lz_out_bytes)
    lz_ceiling = total_in / lz_ceiling_s / 1e9

    # ---- parity sample: 64 random blocks of the benchmark's own output against the CPU oracle (outside every timed region)
    def lz4_parity_sample(k=64):
        from oracle.pyoracle import Oracle

        orc = Oracle()
        ljb.lz4.compress_device(d_in, BLOCK_LEN, d_out, d_offs, d_res, ctx, first_block=first_block, frame_blocks=frame_blocks)
        torch.cuda.synchronize()
        offs = d_offs.cpu().numpy().astype(np.int64)
        rng = np.random.default_rng(1234 + rank)
        bad = 0
        picks = sorted(int(x) for x in rng.choice(nblocks, size=min(k, nblocks), replace=False))
        hin = h_in.numpy()
        for bidx in picks:
            blk = hin[bidx * BLOCK_LEN: min(n, (bidx + 1) * BLOCK_LEN)]
            ref_stream, _, _ = orc.lz4_compress(blk, BLOCK_LEN, 1)  # frame byte + the block
            got = d_out[int(offs[bidx]):int(offs[bidx + 1])].cpu().numpy()
            if not np.array_equal(got, ref_stream[1:]):
                bad += 1
        return {"checked": len(picks), "mismatch": int(sum_over_ranks(float(bad))), "unit": "64 KiB blocks of the timed output vs oracle/lz4_oracle.c"}

    lz_parity = lz4_parity_sample() if not args.no_parity_sample else None

    # ---- decoder: the stream just produced, device-resident, and through the host-buffer call
    d_dec = torch.empty(n, dtype=torch.uint8, device=dev)
    d_blen = torch.empty(nblocks, dtype=torch.int32, device=dev)
    d_dres = torch.zeros(3, dtype=torch.int64, device=dev)

    def lz4_decode_step():
        ljb.lz4.decompress_device(d_out, lz_out_bytes, d_offs, nblocks, BLOCK_LEN, d_dec, d_blen, d_dres, ctx)

    dec_ms, _, dec_launches = timed(lz4_decode_step, args.steps, args.warmup)
    dec_kernel_ms = ctx.last_kernel_ms()
    torch.cuda.synchronize()
    ok_blocks = int(((d_dec.view(nblocks, BLOCK_LEN) == d_in.view(nblocks, BLOCK_LEN)).all(dim=1)).sum().item()) if n % BLOCK_LEN == 0 else None
    phantom = int(d_res[1].item())
    dec_value = total_in / (dec_ms / args.steps * 1e-3) / 1e9
    decode_obj = {"metric": "LZ4 decompress GB/s (of decoded bytes)", "value": dec_value, "unit": "GB/s", "ms_per_step": dec_ms / args.steps,
                  "config": {"workload": "decode of the stream the compress step produced (device-resident, ljb_lz4_decompress_dev)",
                             "undecodable_sequences_in_stream": phantom,
                             "note": "a 257..259-byte match is a sequence the reference's format cannot represent (SURVEY.md A.3-b): the block that "
                                     "holds one is reported as a format error by any decoder; every other block must round-trip"},
                  "roundtrip": {"blocks": nblocks, "blocks_equal_to_input": ok_blocks, "blocks_with_undecodable_sequence_at_most": phantom},
                  "roofline": {"bound": "hbm", "achieved": (n + lz_out_bytes) / (dec_kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": (n + lz_out_bytes) / (dec_kernel_ms * 1e-3) / 1e9 / peak, "traffic": NCU_TRAFFIC["lz4_decode"], "peak_source": peak_src,
                               "kernel": "lz4d::lz4_decode_kernel", "kernel_ms": dec_kernel_ms, "algorithmic_bytes": n + lz_out_bytes}}
    if ok_blocks is not None and ok_blocks + phantom < nblocks:
        raise SystemExit(f"LZ4 round trip: only {ok_blocks} of {nblocks} blocks decode to their input ({phantom} undecodable sequences)")
    del d_dec, d_blen

    # ---- strong scaling (BASELINE configs[2]: 4 GiB in total across 1/2/4/8 GPUs): this rank's 1/N of one 4 GiB stream
    strong = {}
    if world > 1:
        sb = nblocks // world
        d_in_s = d_in[: sb * BLOCK_LEN]

        def lz4_step_strong():
            ljb.lz4.compress_device(d_in_s, BLOCK_LEN, d_out, d_offs, d_res, ctx, first_block=rank * sb, frame_blocks=world * sb)
            with torch.cuda.stream(ext_stream):
                ljb.sharding.gather_totals_device(d_res[0:1])

        s_ms, _, _ = timed(lz4_step_strong, args.steps, args.warmup)
        strong["lz4"] = {"value": sb * BLOCK_LEN * world / (s_ms / args.steps * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": s_ms / args.steps,
                         "config": f"{sb * BLOCK_LEN * world / GIB:g} GiB in total, {sb} blocks per GPU"}
    else:
        strong["lz4"] = {"value": lz_value, "unit": "GB/s", "ms_per_step": lz_ms / args.steps, "config": f"{n / GIB:g} GiB in total on 1 GPU"}
    del h_out, d_out, d_in, h_in
    torch.cuda.empty_cache()

    # ------------------------------- JPEG ----------------------------------------------------------------
    W = H = JPEG_DIM
    ng = ljb.jpeg.group_count(W, H)
    jcap = ng * 96 + 4096  # noise averages ~63 B per group; LJB_E_CAPACITY is reported if ever exceeded
    hj_in = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True)
    ljb.synth.random_image(W, H, seed=42 + rank, out=hj_in.numpy())
    hj_out = torch.empty(jcap, dtype=torch.uint8, pin_memory=True)
    hj_offs = torch.empty(ng + 1, dtype=torch.int64, pin_memory=True)
    dj_in = hj_in.to(dev)
    # the same pixels as r g b, three bytes each: what BASELINE.json's "RGB image" and stbi_load(..., 3) hold; a quarter less to upload
    hj_rgb = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True)
    hj_rgb.copy_(hj_in[:, :, :3])
    dj_rgb = hj_rgb.to(dev)
    dj_out = torch.empty(jcap, dtype=torch.uint8, device=dev)
    dj_offs = torch.empty(ng + 1, dtype=torch.int64, device=dev)
    dj_bits = torch.empty(ng * 3, dtype=torch.int16, device=dev)
    dj_res = torch.zeros(3, dtype=torch.int64, device=dev)

    def jpeg_step_device():
        ljb.jpeg.encode_device(dj_in, W, H, dj_out, dj_offs, dj_bits, dj_res, ctx)
        if world > 1:
            with torch.cuda.stream(ext_stream):
                ljb.sharding.gather_totals_device(dj_res[0:1])

    jout_len = C.c_size_t(0)

    def jpeg_step_e2e():
        rc = lib.ljb_jpeg_encode_rgba(ctx.handle, hj_in.data_ptr(), W, H, 4 * W, 0, ng, hj_out.data_ptr(), jcap, hj_offs.data_ptr(),
                                      None, None, C.byref(jout_len))
        N.check(rc, "ljb_jpeg_encode_rgba")
        if world > 1:
            ljb.sharding.gather_totals(int(jout_len.value), device=dev)

    def jpeg_step_e2e_rgb():
        rc = lib.ljb_jpeg_encode_rgb(ctx.handle, hj_rgb.data_ptr(), W, H, 3 * W, 0, ng, hj_out.data_ptr(), jcap, hj_offs.data_ptr(),
                                     None, None, C.byref(jout_len))
        N.check(rc, "ljb_jpeg_encode_rgb")
        if world > 1:
            ljb.sharding.gather_totals(int(jout_len.value), device=dev)

    jp_ms, _, jp_launches = timed(jpeg_step_device, args.steps, args.warmup)
    ks = []
    for _ in range(min(3, args.steps)):
        ljb.jpeg.encode_device(dj_in, W, H, dj_out, dj_offs, dj_bits, dj_res, ctx)
        ks.append(ctx.last_kernel_ms())
    jp_kernel_ms = sum(ks) / len(ks)
    torch.cuda.synchronize()
    jp_out_bytes = int(dj_res[0].item())
    if int(dj_res[2].item()):
        raise SystemExit(f"JPEG kernel reported error flags {int(dj_res[2].item())}")
    _, je2e_wall, _ = timed(jpeg_step_e2e, e2e_steps, e2e_warm, use_events=False)
    jp_e2e_ms = max_over_ranks(je2e_wall * 1e3 / e2e_steps)
    _, je2e3_wall, _ = timed(jpeg_step_e2e_rgb, e2e_steps, e2e_warm, use_events=False)
    jp_e2e3_ms = max_over_ranks(je2e3_wall * 1e3 / e2e_steps)
    # the stream the three-byte host path left in hj_out is the device-resident path's stream, byte for byte
    jp_e2e3_same = int(jout_len.value) == jp_out_bytes and bool(torch.equal(hj_out[:jp_out_bytes].to(dev), dj_out[:jp_out_bytes]))
    total_px = sum_over_ranks(float(W) * H)
    jp_value = total_px / (jp_ms / args.steps * 1e-3) / 1e6
    jp_e2e_value = total_px / (jp_e2e_ms * 1e-3) / 1e6
    jp_e2e3_value = total_px / (jp_e2e3_ms * 1e-3) / 1e6
    jp_achieved = (4.0 * W * H + jp_out_bytes) / (jp_kernel_ms * 1e-3) / 1e9
    jp_ceiling_s = host_ceiling(hj_in, dj_in, dj_out, hj_out, jp_out_bytes)
    jp_ceiling = total_px / jp_ceiling_s / 1e6
    jp_ceiling3 = total_px / host_ceiling(hj_rgb, dj_rgb, dj_out, hj_out, jp_out_bytes) / 1e6

    def jpeg_parity_sample(runs=16, run_len=256):
        """runs x run_len consecutive groups of the timed output (records, bit lengths) against the CPU oracle."""
        from oracle.pyoracle import Oracle

        orc = Oracle()
        ljb.jpeg.encode_device(dj_in, W, H, dj_out, dj_offs, dj_bits, dj_res, ctx)
        torch.cuda.synchronize()
        offs = dj_offs.cpu().numpy().astype(np.int64)
        bits = dj_bits.cpu().numpy().astype(np.uint16).reshape(-1, 3)
        rng = np.random.default_rng(4321 + rank)
        img = hj_in.numpy()
        bad = checked = 0
        for g0 in sorted(int(x) for x in rng.integers(0, ng - run_len, size=runs)):
            ref = orc.jpeg_encode(img, g0, g0 + run_len, want_coefs=False)
            got = dj_out[int(offs[g0]):int(offs[g0 + run_len])].cpu().numpy()
            same = np.array_equal(got, ref["stream"]) and np.array_equal(offs[g0:g0 + run_len + 1] - offs[g0], ref["offsets"].astype(np.int64)) \
                and np.array_equal(bits[g0:g0 + run_len], np.asarray(ref["bits"]).reshape(-1, 3).astype(np.uint16))
            checked += run_len
            bad += 0 if same else run_len
        return {"checked": checked, "mismatch": int(sum_over_ranks(float(bad))), "unit": "8x8 groups of the timed output vs oracle/jpeg_oracle.c"}

    jp_parity = jpeg_parity_sample() if not args.no_parity_sample else None
    if world > 1:  # strong scaling (BASELINE configs[3]): ONE image, this rank's 1/N of its group rows
        gs = ng // world

        def jpeg_step_strong():
            ljb.jpeg.encode_device(dj_in, W, H, dj_out, dj_offs, dj_bits, dj_res, ctx, first_group=rank * gs, ngroups=gs)
            with torch.cuda.stream(ext_stream):
                ljb.sharding.gather_totals_device(dj_res[0:1])

        s_ms, _, _ = timed(jpeg_step_strong, args.steps, args.warmup)
        strong["jpeg"] = {"value": float(W) * H / (s_ms / args.steps * 1e-3) / 1e6, "unit": "MPix/s", "ms_per_step": s_ms / args.steps,
                          "config": f"one {W}x{H} image in total, {gs} groups per GPU"}
    else:
        strong["jpeg"] = {"value": jp_value, "unit": "MPix/s", "ms_per_step": jp_ms / args.steps, "config": f"one {W}x{H} image on 1 GPU"}

    # ------------------------------- JFIF (true baseline JPEG) ------------------------------------------
    del hj_out, dj_out, dj_offs, dj_bits
    torch.cuda.empty_cache()
    fcap = 607 + 2 + 2 * W * H + 4096  # noise at quality 75 needs ~1.26 B/px (4:4:4); LJB_E_CAPACITY is reported if exceeded
    df_out = torch.empty(fcap, dtype=torch.uint8, device=dev)
    df_res = torch.zeros(3, dtype=torch.int64, device=dev)
    hf_out = torch.empty(fcap, dtype=torch.uint8, pin_memory=True)
    fout_len = C.c_size_t(0)
    jfif = {}
    for name, sub in (("444", 0), ("420", -1)):
        def jfif_step_device():
            ljb.jfif.encode_device(dj_in, W, H, 4, JFIF_QUALITY, sub, df_out, df_res, ctx)
            if world > 1:
                with torch.cuda.stream(ext_stream):
                    ljb.sharding.gather_totals_device(df_res[0:1])

        def jfif_step_e2e():
            rc = lib.ljb_jfif_encode(ctx.handle, hj_in.data_ptr(), W, H, 4, 4 * W, JFIF_QUALITY, sub, hf_out.data_ptr(), fcap,
                                     C.byref(fout_len))
            N.check(rc, "ljb_jfif_encode")
            if world > 1:
                ljb.sharding.gather_totals(int(fout_len.value), device=dev)

        f_ms, _, f_launches = timed(jfif_step_device, args.steps, args.warmup)
        ks = []
        for _ in range(min(3, args.steps)):
            ljb.jfif.encode_device(dj_in, W, H, 4, JFIF_QUALITY, sub, df_out, df_res, ctx)
            ks.append(ctx.last_kernel_ms())
        f_kernel_ms = sum(ks) / len(ks)
        torch.cuda.synchronize()
        f_out_bytes = int(df_res[0].item())
        if int(df_res[2].item()):
            raise SystemExit(f"JFIF kernels reported error flags {int(df_res[2].item())}")
        _, fe2e_wall, _ = timed(jfif_step_e2e, e2e_steps, e2e_warm, use_events=False)
        f_e2e_ms = max_over_ranks(fe2e_wall * 1e3 / e2e_steps)
        assert int(fout_len.value) == f_out_bytes, "host-buffer path and device path disagree on the file length"

        def jfif_step_e2e_rgb():
            rc = lib.ljb_jfif_encode(ctx.handle, hj_rgb.data_ptr(), W, H, 3, 3 * W, JFIF_QUALITY, sub, hf_out.data_ptr(), fcap,
                                     C.byref(fout_len))
            N.check(rc, "ljb_jfif_encode")
            if world > 1:
                ljb.sharding.gather_totals(int(fout_len.value), device=dev)

        _, fe2e3_wall, _ = timed(jfif_step_e2e_rgb, e2e_steps, e2e_warm, use_events=False)
        f_e2e3_ms = max_over_ranks(fe2e3_wall * 1e3 / e2e_steps)
        f_e2e3_same = int(fout_len.value) == f_out_bytes and bool(torch.equal(hf_out[:f_out_bytes].to(dev), df_out[:f_out_bytes]))
        jfif[name] = {"value": total_px / (f_ms / args.steps * 1e-3) / 1e6, "ms_per_step": f_ms / args.steps, "kernel_ms": f_kernel_ms,
                      "out_bytes": f_out_bytes, "e2e_value": total_px / (f_e2e_ms * 1e-3) / 1e6, "e2e_ms": f_e2e_ms,
                      "achieved": (4.0 * W * H + f_out_bytes) / (f_kernel_ms * 1e-3) / 1e9, "launches": f_launches,
                      "e2e3_value": total_px / (f_e2e3_ms * 1e-3) / 1e6, "e2e3_ms": f_e2e3_ms, "e2e3_same": f_e2e3_same}

    def jfif_parity_sample(sub, rows=64):
        """A band of the benchmark image encoded alone by the GPU path, byte for byte against the CPU oracle's file (the
        benchmark file itself is one bit stream whose every unit depends on all units before it)."""
        from oracle.pyoracle import Oracle

        orc = Oracle()
        r0 = int(np.random.default_rng(99 + rank).integers(0, H - rows))
        band = np.ascontiguousarray(hj_in.numpy()[r0:r0 + rows])
        got = ljb.jfif.write_jpg(band, JFIF_QUALITY, sub, ctx=ctx)
        ref = orc.jfif_encode(band, JFIF_QUALITY, sub)
        ref = ref[0] if isinstance(ref, tuple) else ref
        return {"checked": 1, "mismatch": int(sum_over_ranks(0.0 if np.array_equal(got, ref) else 1.0)),
                "unit": f"{rows}-row band ({W}x{rows}) of the benchmark image, whole .jpg file vs oracle/jfif_oracle.c"}

    if not args.no_parity_sample:
        jfif["444"]["parity"] = jfif_parity_sample(0)
        jfif["420"]["parity"] = jfif_parity_sample(-1)
    # ------------------------------- batch (BASELINE configs[4]) -----------------------------------------
    # 8192 frames of 1920x1080 across 8 GPUs = 1024 per GPU: batch JPEG encode -> LZ4 compress of the concatenated bit streams ->
    # LZ4 decompress (the round trip), all device-resident in `value`; `e2e` through the three host-buffer calls.
    batch_obj = None
    if BATCH_IMAGES > 0:
        del df_out, hf_out
        torch.cuda.empty_cache()
        bw_, bh_ = 1920, 1080
        nimg = BATCH_IMAGES
        G = ljb.jpeg.group_count(bw_, bh_)
        distinct = min(32, nimg)
        hb = torch.empty((nimg, bh_, bw_, 4), dtype=torch.uint8, pin_memory=True)
        hbn = hb.numpy()
        for i in range(distinct):
            ljb.synth.random_image(bw_, bh_, seed=42 + 1000 * rank + i, out=hbn[i])
        for i in range(distinct, nimg):
            hbn[i] = hbn[i % distinct]
        db = hb.to(dev)
        bcap = nimg * G * 64 + 4096
        dbo = torch.empty(bcap, dtype=torch.uint8, device=dev)
        dboffs = torch.empty(nimg * G + 1, dtype=torch.int64, device=dev)
        dbres = torch.zeros(3, dtype=torch.int64, device=dev)
        ljb.jpeg.encode_batch_device(db, bw_, bh_, nimg, dbo, dboffs, None, dbres, ctx)
        torch.cuda.synchronize()
        if int(dbres[2].item()):
            raise SystemExit(f"batch JPEG kernel reported error flags {int(dbres[2].item())}")
        jlen = int(dbres[0].item())
        lnb = (jlen + BLOCK_LEN - 1) // BLOCK_LEN
        dlz = torch.empty(jlen + jlen // 8 + 16 * lnb + 4096, dtype=torch.uint8, device=dev)
        dlzo = torch.empty(lnb + 1, dtype=torch.int64, device=dev)
        dlzr = torch.zeros(3, dtype=torch.int64, device=dev)
        ljb.lz4.compress_device(dbo[:jlen], BLOCK_LEN, dlz, dlzo, dlzr, ctx)
        torch.cuda.synchronize()
        lzlen, lzph = int(dlzr[0].item()), int(dlzr[1].item())
        dback = torch.empty(lnb * BLOCK_LEN, dtype=torch.uint8, device=dev)
        dblen = torch.empty(lnb, dtype=torch.int32, device=dev)
        ddres = torch.zeros(3, dtype=torch.int64, device=dev)

        def batch_step_device():
            ljb.jpeg.encode_batch_device(db, bw_, bh_, nimg, dbo, dboffs, None, dbres, ctx)
            ljb.lz4.compress_device(dbo[:jlen], BLOCK_LEN, dlz, dlzo, dlzr, ctx)
            ljb.lz4.decompress_device(dlz, lzlen, dlzo, lnb, BLOCK_LEN, dback, dblen, ddres, ctx)

        b_ms, _, b_launches = timed(batch_step_device, args.steps, args.warmup)
        torch.cuda.synchronize()
        rt_ok = bool(torch.equal(dback[:jlen], dbo[:jlen])) and int(ddres[2].item()) == 0 and int(dbres[0].item()) == jlen
        if not rt_ok or lzph:
            raise SystemExit(f"batch: LZ4 round trip of the JPEG bit streams failed (flags {int(ddres[2].item())}, phantom {lzph})")
        # per-kernel times (events of the library's own stream around each launch)
        ljb.jpeg.encode_batch_device(db, bw_, bh_, nimg, dbo, dboffs, None, dbres, ctx)
        bj_ms = ctx.last_kernel_ms()
        ljb.lz4.compress_device(dbo[:jlen], BLOCK_LEN, dlz, dlzo, dlzr, ctx)
        bl_ms = ctx.last_kernel_ms()
        ljb.lz4.decompress_device(dlz, lzlen, dlzo, lnb, BLOCK_LEN, dback, dblen, ddres, ctx)
        bd_ms = ctx.last_kernel_ms()
        # parity sample: 600 groups of two random frames of the batch output against the oracle
        b_par = None
        if not args.no_parity_sample:
            from oracle.pyoracle import Oracle

            orc = Oracle()
            offs_h = dboffs.cpu().numpy().astype(np.int64)
            bad = 0
            for im in (int(x) for x in np.random.default_rng(5 + rank).integers(0, nimg, size=2)):
                ref = orc.jpeg_encode(hbn[im], 12000, 12600, want_coefs=False)
                got = dbo[int(offs_h[im * G + 12000]):int(offs_h[im * G + 12600])].cpu().numpy()
                bad += 0 if np.array_equal(got, ref["stream"]) else 600
            b_par = {"checked": 1200, "mismatch": int(sum_over_ranks(float(bad))), "unit": "8x8 groups of two frames of the batch output vs oracle/jpeg_oracle.c"}
        del dback, dlz, dbo, dboffs
        torch.cuda.empty_cache()
        # end to end: the three host-buffer calls a user of the public API makes
        hjo = torch.empty(bcap, dtype=torch.uint8, pin_memory=True)
        hlz = torch.empty(jlen + jlen // 8 + 16 * lnb + 4096, dtype=torch.uint8, pin_memory=True)
        hlo = torch.empty(lnb + 1, dtype=torch.int64, pin_memory=True)
        hbk = torch.empty(lnb * BLOCK_LEN, dtype=torch.uint8, pin_memory=True)
        o1, o2, o3 = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)

        def batch_step_e2e():
            N.check(lib.ljb_jpeg_encode_batch(ctx.handle, hb.data_ptr(), bw_, bh_, 4 * bw_, 4 * bw_ * bh_, nimg, hjo.data_ptr(), bcap, None, None,
                                              C.byref(o1)), "ljb_jpeg_encode_batch")
            N.check(lib.ljb_lz4_compress(ctx.handle, hjo.data_ptr(), o1.value, BLOCK_LEN, hlz.data_ptr(), hlz.numel(), hlo.data_ptr(),
                                         C.byref(o2), None), "ljb_lz4_compress")
            N.check(lib.ljb_lz4_decompress(ctx.handle, hlz.data_ptr(), o2.value, hlo.data_ptr(), lnb, BLOCK_LEN, hbk.data_ptr(), hbk.numel(),
                                           C.byref(o3)), "ljb_lz4_decompress")

        _, be_wall, _ = timed(batch_step_e2e, e2e_steps, e2e_warm, use_events=False)
        be_ms = max_over_ranks(be_wall * 1e3 / e2e_steps)
        assert o1.value == jlen and o3.value == jlen and bool(torch.equal(hbk[:jlen], hjo[:jlen])), "batch e2e round trip"
        # the same frames as r g b, three bytes per pixel (a quarter less to upload: the upload is most of the step)
        hb3 = torch.empty((nimg, bh_, bw_, 3), dtype=torch.uint8, pin_memory=True)
        hb3.copy_(hb[:, :, :, :3])

        def batch_step_e2e_rgb():
            N.check(lib.ljb_jpeg_encode_batch_rgb(ctx.handle, hb3.data_ptr(), bw_, bh_, 3 * bw_, 3 * bw_ * bh_, nimg, hjo.data_ptr(), bcap, None, None,
                                                  C.byref(o1)), "ljb_jpeg_encode_batch_rgb")
            N.check(lib.ljb_lz4_compress(ctx.handle, hjo.data_ptr(), o1.value, BLOCK_LEN, hlz.data_ptr(), hlz.numel(), hlo.data_ptr(),
                                         C.byref(o2), None), "ljb_lz4_compress")
            N.check(lib.ljb_lz4_decompress(ctx.handle, hlz.data_ptr(), o2.value, hlo.data_ptr(), lnb, BLOCK_LEN, hbk.data_ptr(), hbk.numel(),
                                           C.byref(o3)), "ljb_lz4_decompress")

        _, be3_wall, _ = timed(batch_step_e2e_rgb, e2e_steps, e2e_warm, use_events=False)
        be3_ms = max_over_ranks(be3_wall * 1e3 / e2e_steps)
        assert o1.value == jlen and o3.value == jlen and bool(torch.equal(hbk[:jlen], hjo[:jlen])), "batch e2e round trip (r g b)"
        del hb3
        total_img = sum_over_ranks(float(nimg))
        batch_obj = {
            "metric": "batched 1080p JPEG encode + LZ4 round trip of the bit streams, frames/s", "value": total_img / (b_ms / args.steps * 1e-3),
            "unit": "frames/s", "mpix_per_s": total_img * bw_ * bh_ / (b_ms / args.steps * 1e-3) / 1e6, "ms_per_step": b_ms / args.steps,
            "config": {"workload": f"{nimg} frames of {bw_}x{bh_} random_image-style RGBA per GPU (BASELINE configs[4]: 8192 frames across 8 GPUs = 1024 "
                                   f"per GPU; {distinct} distinct frames, seed 42 + 1000 rank + index, repeated): ljb_jpeg_encode_batch_dev -> "
                                   f"ljb_lz4_compress_dev (64 KiB blocks over the concatenated bit streams) -> ljb_lz4_decompress_dev",
                       "frames_per_gpu": nimg, "jpeg_stream_bytes": jlen, "lz4_stream_bytes": lzlen, "launches_per_step": 3},
            "round_trip": {"lz4_decompress(lz4_compress(jpeg_streams)) == jpeg_streams": rt_ok},
            "kernels_ms": {"jpgk::jpeg_encode_kernel": bj_ms, "lz4k::lz4_encode_kernel": bl_ms, "lz4d::lz4_decode_kernel": bd_ms},
            "roofline": {"bound": "hbm", "achieved": (4.0 * bw_ * bh_ * nimg + jlen) / (bj_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": (4.0 * bw_ * bh_ * nimg + jlen) / (bj_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "jpgk::jpeg_encode_kernel (the batch's dominant launch by bytes)", "kernel_ms": bj_ms,
                         "algorithmic_bytes": 4 * bw_ * bh_ * nimg + jlen},
            "e2e": {"value": total_img / (be3_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 3 * bw_ * bh_ * nimg + jlen + lzlen + 8 * (lnb + 1),
                    "d2h_bytes_per_step": jlen + lzlen + 8 * (lnb + 1) + jlen, "ms_per_step": be3_ms,
                    "api": "ljb_jpeg_encode_batch_rgb + ljb_lz4_compress + ljb_lz4_decompress (host buffers, pinned; frames of three bytes per pixel)",
                    "steps": e2e_steps, "warmup": e2e_warm},
            "e2e_rgba": {"value": total_img / (be_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 4 * bw_ * bh_ * nimg + jlen + lzlen + 8 * (lnb + 1),
                         "d2h_bytes_per_step": jlen + lzlen + 8 * (lnb + 1) + jlen, "ms_per_step": be_ms,
                         "api": "ljb_jpeg_encode_batch + ljb_lz4_compress + ljb_lz4_decompress (host buffers, pinned; four bytes per pixel)",
                         "steps": e2e_steps, "warmup": e2e_warm},
            "parity_sample": b_par, "launches": b_launches}
        del hb, db, hjo, hlz, hbk
        torch.cuda.empty_cache()

    clocks = sampler.stop()

    # ------------------------------- CPU baseline (rank 0, N = 1 only) ----------------------------------
    cpu_lz = cpu_jp = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sb = max(threads, 16)
        v, kind, sec = _cpu_lz4(sb, threads)
        cpu_lz = {"value": v, "unit": "GB/s", "cores": threads, "kind": kind,
                  "sample": f"{sb} blocks of 64 KiB of the same seed-42 text, reference block_encode on {threads} threads, {sec:.1f} s"}
        cpu_lz.update(_cpu_lz4_detail(threads))  # + the sequential build on 1 core, + the parallel build's thread body
        vj, kindj, secj = _cpu_jpeg(4096, 2048, threads)
        cpu_jp = {"value": vj, "unit": "MPix/s", "cores": threads, "kind": kindj,
                  "sample": f"4096x2048 seed-42 noise image, reference per-group encode stages on {threads} threads, {secj:.1f} s"}
        cpu_jp.update(_cpu_jpeg_detail(threads))
        cpu_batch = None
        if batch_obj:
            from oracle.pyoracle import Ref

            if Ref.available("lz4"):
                noise = np.random.default_rng(42).integers(0, 256, size=max(threads, 16) * BLOCK_LEN, dtype=np.uint8)
                secn, _ = Ref("lz4").lz4_time_blocks(noise, BLOCK_LEN, threads)
                lz_rate = noise.size / secn  # bytes/s on near-uniform bytes (what a Huffman bit stream looks like to LZ4)
                frame_px = 1920 * 1080
                per_frame = frame_px / (vj * 1e6) + (batch_obj["config"]["jpeg_stream_bytes"] / batch_obj["config"]["frames_per_gpu"]) / lz_rate
                cpu_batch = {"value": 1.0 / per_frame, "unit": "frames/s", "cores": threads, "kind": "reference",
                             "sample": f"per frame: the reference's per-group JPEG stages at the {vj:.1f} MPix/s measured above + its block_encode on the "
                                       f"frame's bit-stream bytes at the rate measured on {max(threads, 16)} blocks of 64 KiB of uniform random bytes "
                                       f"({lz_rate / 1e6:.3f} MB/s, {secn:.1f} s), {threads} threads; decoding not counted"}
        cpu_jf = {}
        for name, sub in (("444", 0), ("420", -1)):
            vf, kindf, secf = _cpu_jfif(4096, 4096, threads, sub)
            cpu_jf[name] = {"value": vf, "unit": "MPix/s", "cores": threads, "kind": kindf,
                            "sample": f"4096x4096 seed-42 noise image, stbi_write_jpg quality {JFIF_QUALITY}, 4 images per thread on {threads} threads, {secf:.1f} s"}

    if rank == 0:
        line = {
            "metric": "LZ4 compress GB/s (headline) & JPEG encode MPix/s (jpeg)", "value": lz_value, "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": lz_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"LZ4 block compression of {n / GIB:g} GiB random_extract-style text per GPU, 64 KiB blocks (BASELINE configs[2])",
                       "bytes_per_gpu": n, "block_len": BLOCK_LEN, "blocks_per_gpu": nblocks, "seed": "42+rank",
                       "l2": "inputs far exceed the 126 MB L2; no flush needed",
                       "compressed_bytes_rank0": lz_out_bytes, "ratio": lz_out_bytes / n},
            "roofline": {"bound": "hbm", "achieved": lz_achieved, "peak": peak, "unit": "GB/s", "frac": lz_achieved / peak,
                         "traffic": NCU_TRAFFIC["lz4"], "peak_source": peak_src, "kernel": "lz4k::lz4_encode_kernel",
                         "kernel_ms": lz_kernel_ms, "algorithmic_bytes": n + lz_out_bytes},
            "e2e": {"value": lz_e2e_value, "unit": "GB/s", "h2d_bytes_per_step": n,
                    "d2h_bytes_per_step": lz_out_bytes + 8 * (nblocks + 1) + 24, "ms_per_step": lz_e2e_ms,
                    "api": "ljb_lz4_compress (host buffers, pinned)", "steps": e2e_steps, "warmup": e2e_warm,
                    "host_ceiling_gbs": lz_ceiling,
                    "host_ceiling": "the same bytes as plain pinned cudaMemcpyAsync H2D + D2H at once, all ranks together, best of 3"},
            "parity_sample": lz_parity,
            "lz4_decode": decode_obj,
            "strong_scaling": strong,
            "batch": batch_obj,
            "gpu_launches": lz_launches + jp_launches + dec_launches + sum(v["launches"] for v in jfif.values()) + (batch_obj["launches"] if batch_obj else 0),
            "clocks": clocks,
            "jpeg": {
                "metric": "JPEG encode MPix/s", "value": jp_value, "unit": "MPix/s", "ms_per_step": jp_ms / args.steps,
                "config": {"workload": f"JPEG-like encode of one {W}x{H} random_image-style RGBA image per GPU (BASELINE configs[3])",
                           "groups_per_gpu": ng, "compressed_bytes_rank0": jp_out_bytes},
                "roofline": {"bound": "hbm", "achieved": jp_achieved, "peak": peak, "unit": "GB/s", "frac": jp_achieved / peak,
                             "traffic": NCU_TRAFFIC["jpeg"], "peak_source": peak_src, "kernel": "jpgk::jpeg_encode_kernel",
                             "kernel_ms": jp_kernel_ms, "algorithmic_bytes": 4 * W * H + jp_out_bytes},
                "e2e": {"value": jp_e2e3_value, "unit": "MPix/s", "h2d_bytes_per_step": 3 * W * H,
                        "d2h_bytes_per_step": jp_out_bytes + 8 * (ng + 1) + 24, "ms_per_step": jp_e2e3_ms,
                        "api": "ljb_jpeg_encode_rgb (host buffers, pinned; r g b, three bytes per pixel: BASELINE configs[3] names an RGB image)",
                        "steps": e2e_steps, "warmup": e2e_warm, "host_ceiling_mpix": jp_ceiling3,
                        "stream_equals_device_resident_stream": jp_e2e3_same},
                "e2e_rgba": {"value": jp_e2e_value, "unit": "MPix/s", "h2d_bytes_per_step": 4 * W * H,
                             "d2h_bytes_per_step": jp_out_bytes + 8 * (ng + 1) + 24, "ms_per_step": jp_e2e_ms,
                             "api": "ljb_jpeg_encode_rgba (host buffers, pinned; four bytes per pixel)", "steps": e2e_steps, "warmup": e2e_warm,
                             "host_ceiling_mpix": jp_ceiling},
                "parity_sample": jp_parity,
            },
        }
        def jfif_obj(name, label):
            v = jfif[name]
            return {"metric": f"baseline JPEG (JFIF) encode MPix/s, quality {JFIF_QUALITY}, {label}", "value": v["value"], "unit": "MPix/s",
                    "ms_per_step": v["ms_per_step"],
                    "config": {"workload": f"stbi_write_jpg-identical .jpg of one {W}x{H} random_image-style RGBA image per GPU, quality "
                                           f"{JFIF_QUALITY}, {label}", "file_bytes_rank0": v["out_bytes"],
                               "kernels_per_step": "jfk::jfif_encode_kernel + jfk::jfif_stuff_kernel"},
                    "roofline": {"bound": "hbm", "achieved": v["achieved"], "peak": peak, "unit": "GB/s", "frac": v["achieved"] / peak,
                                 "traffic": NCU_TRAFFIC["jfif" + name], "peak_source": peak_src, "kernel": "jfk::jfif_encode_kernel",
                                 "kernel_ms": v["kernel_ms"], "algorithmic_bytes": 4 * W * H + v["out_bytes"]},
                    "e2e": {"value": v["e2e3_value"], "unit": "MPix/s", "h2d_bytes_per_step": 3 * W * H,
                            "d2h_bytes_per_step": v["out_bytes"] + 24, "ms_per_step": v["e2e3_ms"],
                            "api": "ljb_jfif_encode (host buffers, pinned; comp = 3, r g b)", "steps": e2e_steps, "warmup": e2e_warm,
                            "file_equals_device_resident_file": v["e2e3_same"]},
                    "e2e_rgba": {"value": v["e2e_value"], "unit": "MPix/s", "h2d_bytes_per_step": 4 * W * H,
                                 "d2h_bytes_per_step": v["out_bytes"] + 24, "ms_per_step": v["e2e_ms"],
                                 "api": "ljb_jfif_encode (host buffers, pinned; comp = 4)", "steps": e2e_steps, "warmup": e2e_warm},
                    "parity_sample": v.get("parity")}

        line["jfif"] = jfif_obj("444", "4:4:4 (BASELINE.json's wording)")
        line["jfif"]["stb_rule_420"] = jfif_obj("420", "4:2:0 (what stbi_write_jpg itself does at quality <= 90)")
        if cpu_lz:
            line["cpu_baseline"] = cpu_lz
            line["jpeg"]["cpu_baseline"] = cpu_jp
            line["jfif"]["cpu_baseline"] = cpu_jf["444"]
            line["jfif"]["stb_rule_420"]["cpu_baseline"] = cpu_jf["420"]
            if batch_obj and cpu_batch:
                line["batch"]["cpu_baseline"] = cpu_batch
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-sample", action="store_true", help="skip the post-run oracle check of sampled output units")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
