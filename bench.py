#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark (BASELINE.json: "4K K=256 CIELAB assign Gpixel/s;
SWASA palette evals/s at 1/2/4/8 B200").

A step = one batched SWASA scoring launch: 64 candidate 256-colour palettes over a 3840x2160
synthetic RGB image per GPU (BASELINE configs[2]); at N GPUs the image is 3840 x (2160*N), pixel
rows sharded one block per rank, and the only exchange is an NCCL all-reduce of the integer
result words (weak scaling; at N=8 this is the 64 MP configuration, configs[3]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # the CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's own kernels on the host cores
                                                                  # (oracle/_ref; oracle port if absent), rank 0 only

value  : pixel x candidate assignments per second (Gpixel/s), inputs resident in HBM, device timed
e2e    : the same through the host-buffer C ABI call hq_eval_palettes (H2D palettes, D2H results)
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

METRIC = "cielab_assign_gpixel_per_s_4k_k256"
UNIT = "Gpixel/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--rows-per-gpu", type=int, default=2160)
    ap.add_argument("--colors", type=int, default=256)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--space", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0, help="0 auto, 1 direct, 2 chunked, 3 prefilter (profiling)")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a, world):
    return (f"{a.width}x{a.rows_per_gpu * world} synthetic RGB ({a.width}x{a.rows_per_gpu} rows per GPU), "
            f"{a.colors}-colour palettes, {a.batch} SWASA candidates per launch")


# ------------------------------------------------------------------ clocks sampler (recipe's clocks line)
class ClockSampler:
    """SM clock, power and throttle reasons sampled every 25 ms during the timed region.  NVML in-process (pynvml) when
    it loads — no start-up delay — else a streaming `nvidia-smi -lms 25` child; `wait_ready` blocks until a first sample exists
    so that a short timed region cannot end before the sampler has started."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread, self.source = index, [], None, None, None
        self._stop = threading.Event()

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if ids and self.index < len(ids) and ids[self.index].isdigit():
            return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml as N

            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self._physical_index())
            N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, args=(N, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self._physical_index())], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.source = "nvidia-smi"
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _poll_nvml(self, N, h):
        bits = [("hw_slowdown", getattr(N, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(N, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(N, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(N, "nvmlClocksThrottleReasonSwPowerCap", 0x4))]
        try:
            mx = str(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
        except Exception:
            mx = ""
        while not self._stop.is_set():
            try:
                sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                try:
                    pw = "%.2f" % (N.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pw = ""
                try:
                    mask = int(N.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                except Exception:
                    mask = 0
                self.rows.append((time.perf_counter(), [str(sm), mx, pw] + ["Active" if mask & b else "Not Active" for _, b in bits] + [hex(mask)]))
            except Exception:
                pass
            self._stop.wait(0.025)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def wait_ready(self, seconds: float = 10.0) -> bool:
        t0 = time.perf_counter()
        while not self.rows and self.thread is not None and time.perf_counter() - t0 < seconds:
            time.sleep(0.02)
        return bool(self.rows)

    def stop(self):
        self._stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0: float, t1: float) -> dict:
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 8] or [r for (_, r) in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][1]) if rows[0][1].isdigit() else None,
                "power_w_max": max((float(r[2]) for r in rows if r[2].replace(".", "").isdigit()), default=None),
                "samples": len(rows), "source": self.source, "reasons": reasons}


# ------------------------------------------------------------------ CPU arm (oracle port)
def run_cpu(a, steps: int, warmup: int, candidates_per_step: int, seconds_budget: float | None = None) -> dict:
    """Times the CPU path (the oracle port: oracle/hq_oracle.c, all host cores) on a bounded sample
    of the GPU arm's workload: the full per-GPU image and palette size, `candidates_per_step`
    of the 64 candidates per step."""
    from hybridquantization_b200 import synth
    from oracle import hq_oracle as O

    O.build()
    cores = O.default_threads()
    img = synth.synth_image_rows(a.width, a.rows_per_gpu, synth.SEED_BASE + 3, 0, a.rows_per_gpu)
    pal = synth.synth_palettes(a.batch, a.colors)
    unit, lab = O.image_planes(img, threads=cores)
    n = a.width * a.rows_per_gpu
    times = []
    t_start = time.perf_counter()
    for s in range(warmup + steps):
        sel = [(s * candidates_per_step + i) % a.batch for i in range(candidates_per_step)]
        t0 = time.perf_counter()
        r = O.assign_reduce_planes(unit, lab, pal[sel], a.space, threads=cores)
        dt = time.perf_counter() - t0
        assert (r["counts"].sum(axis=1) == n).all()
        if s >= warmup:
            times.append(dt)
        if seconds_budget is not None and s >= warmup and time.perf_counter() - t_start > seconds_budget:
            break
    total = sum(times)
    gpix = n * candidates_per_step * len(times) / total / 1e9
    return {"value": gpix, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} steps x {candidates_per_step} of {a.batch} candidates, full {a.width}x{a.rows_per_gpu} image, K={a.colors}; "
                      f"oracle/hq_oracle.c -O3 x86-64-v3, {cores} threads",
            "ms_per_step": 1e3 * total / len(times), "steps": len(times),
            "evals_per_s": candidates_per_step * len(times) / total}


def run_reference_kernels(a, steps: int, warmup: int, seconds_budget: float | None = None) -> dict | None:
    """Times the REFERENCE'S OWN kernels (OptimizedConvolution.cl compiled for the CPU into oracle/_ref by
    oracle/ref_build/build_ref.sh) on all host cores.  One step = ONE candidate palette over the full per-GPU image
    (a bounded sample of the GPU arm's 64-candidate step), through the reference's candidate chain for this path:
    quantizeAndConvertToOpp (argmin + used flags) -> Opp2LAB -> CIEDE -> host double mean
    (ImageManipulation.java:644-665,712), i.e. computeQuantizationErrorPopulation with the spatial-filter kernels
    skipped (the identity-filter cost the north star's path scores).  The full chain including
    computeScielabKernelsTemp/End is timed once and reported beside it.  Returns None when oracle/_ref is absent."""
    from hybridquantization_b200 import synth
    from oracle import hq_oracle as O
    from oracle import hq_ref as R

    if not R.build():
        return None
    L = R.load()
    cores = R.default_threads()
    w, h, K = a.width, a.rows_per_gpu, a.colors
    n = w * h
    img = synth.synth_image_rows(w, h, synth.SEED_BASE + 3, 0, h)
    pal = synth.synth_palettes(a.batch, K)
    rgb4 = R.makeinline(R.unit_planes(img))
    # Lab of the original through the reference's kernels (RGB2XYZ -> XYZ2Opp -> Opp2LAB): the comparison image of CIEDE
    xyz = R.rgb_to_xyz(R.unit_planes(img), cores)
    opp0 = np.zeros_like(xyz); lab0 = np.zeros_like(xyz)
    L.refcl_XYZ2Opp(R._ptr(xyz), R._ptr(opp0), n, cores)
    L.refcl_Opp2LAB(R._ptr(opp0), R.D65[0], R.D65[1], R.D65[2], R._ptr(lab0), n, cores)
    del xyz
    opp = np.zeros((n, 4), np.float32); lab = np.zeros((n, 4), np.float32); err = np.zeros(n, np.float32)

    def one(colors, full_chain=False):
        used = np.zeros(K, np.int32)
        L.refcl_quantizeAndConvertToOpp(R._ptr(rgb4), R._ptr(colors), K, R._ptr(used), R._ptr(opp), n, cores)
        src = opp
        if full_chain:
            L.refcl_computeScielabKernelsTemp(R._ptr(opp), R._ptr(pk["filters4"][0]), R._ptr(pk["filters4"][1]), R._ptr(pk["filter3"]), pk["half"], w, h,
                                              R._ptr(t1), R._ptr(t2), R._ptr(t3), n, cores)
            L.refcl_computeScielabKernelsEnd(R._ptr(t1), R._ptr(t2), R._ptr(t3), R._ptr(pk["filters4"][0]), R._ptr(pk["filters4"][1]), R._ptr(pk["absfilter3"]),
                                             pk["half"], h, w, R._ptr(conv), n, cores)
            src = conv
        L.refcl_Opp2LAB(R._ptr(src), R.D65[0], R.D65[1], R.D65[2], R._ptr(lab), n, cores)
        L.refcl_CIEDE(R._ptr(lab0), R._ptr(lab), R._ptr(err), n, cores)
        return float(err.sum(dtype=np.float64)) / n + 2.0 * int((used == 0).sum())   # averageArray + computePenalty (:712)

    times = []
    t_start = time.perf_counter()
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        cost = one(pal[s % a.batch])
        dt = time.perf_counter() - t0
        assert np.isfinite(cost)
        if s >= warmup:
            times.append(dt)
        if seconds_budget is not None and s >= warmup and time.perf_counter() - t_start > seconds_budget:
            break
    f, ab = O.scielab_filters(72, 45.0)
    pk = R.pack_filters(f, ab)
    t1 = np.zeros((n, 4), np.float32); t2 = np.zeros((n, 4), np.float32); t3 = np.zeros(n, np.float32); conv = np.zeros((n, 4), np.float32)
    t0 = time.perf_counter(); one(pal[0], True); full_s = time.perf_counter() - t0
    total = sum(times)
    return {"value": n * len(times) / total / 1e9, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"{len(times)} steps x 1 of {a.batch} candidates, full {w}x{h} image, K={K}; the reference's OptimizedConvolution.cl kernels "
                      f"(quantizeAndConvertToOpp, Opp2LAB, CIEDE) compiled for the CPU (oracle/_ref, g++ -O3 x86-64-v3), {cores} threads, host double mean",
            "ms_per_step": 1e3 * total / len(times), "steps": len(times), "evals_per_s": len(times) / total,
            "full_chain_with_spatial_filters": {"value": n / full_s / 1e9, "unit": UNIT, "ms_per_candidate": 1e3 * full_s}}


def main_reference(a, rank: int, world: int) -> None:
    if rank != 0:
        return
    r = run_reference_kernels(a, a.steps, a.warmup)
    note = ("reference arm: the reference's own OpenCL kernels (quantizeAndConvertToOpp -> Opp2LAB -> CIEDE -> host mean) compiled for the CPU "
            "from /root/reference by oracle/ref_build/build_ref.sh, all host cores, one candidate per step; the Java/JavaCL host cannot run here (no JDK)")
    if r is None:
        r = run_cpu(a, a.steps, a.warmup, candidates_per_step=1)
        note = ("reference arm: oracle/_ref is absent on this machine; this is the C oracle port (oracle/hq_oracle.c) timed on the host cores, "
                "one candidate per step")
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": r["steps"],
            "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a, 1), "note": note},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample") if k in r},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "swasa_evals_per_s": r["evals_per_s"]}
    if "full_chain_with_spatial_filters" in r:
        line["full_chain_with_spatial_filters"] = r["full_chain_with_spatial_filters"]
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
def main_b200(a, rank: int, local_rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist

    from hybridquantization_b200 import EVAL_FORCE_CHUNKED, EVAL_FORCE_DIRECT, EVAL_FORCE_PREFILTER, EVAL_PRUNE, ImageManipulation, build, synth
    from hybridquantization_b200.dist import install_nccl_allreduce, row_shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sampler = ClockSampler(local_rank)   # started first: by the timed region it has long been sampling
    sampler.start()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build_library()
    if world > 1:
        dist.barrier()

    be = ImageManipulation("CIE76", False, True, local_rank)
    info = be.deviceInfo()
    H = a.rows_per_gpu * world
    r0, r1 = row_shard(H, world, rank)
    n_shard, n_total = a.width * (r1 - r0), a.width * H
    K, B = a.colors, a.batch
    flags = {0: 0, 1: EVAL_FORCE_DIRECT, 2: EVAL_FORCE_CHUNKED, 3: EVAL_FORCE_PREFILTER}[a.variant]

    # a dedicated (non-default) stream: the C ABI takes a cudaStream_t and treats NULL as "the
    # context's own stream", so every launch and every timing event below is on this one stream
    bench_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(bench_stream)

    # inputs resident in HBM before the timed region
    img = synth.synth_image_rows(a.width, H, synth.SEED_BASE + 3, r0, r1)
    d_img = torch.from_numpy(img).to(dev)
    stream = bench_stream.cuda_stream
    assert stream != 0
    be.setImageDevice(d_img.data_ptr(), a.width, r1 - r0, stream=stream)
    pal = synth.synth_palettes(B, K)
    d_pal = torch.from_numpy(pal).to(dev)
    words = be.resultWords(K, 0)
    d_res = torch.zeros((B, words), dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    torch.cuda.synchronize()

    peak = be.measureFp32Peak()  # FFMA microbenchmark at this device's current clocks

    def step():
        be.evalPalettesDevice(d_pal.data_ptr(), B, K, d_res.data_ptr(), a.space, flags, stream)
        if world > 1:
            dist.all_reduce(d_res, op=dist.ReduceOp.SUM)

    for _ in range(a.warmup):
        flush.fill_(1)
        step()
    torch.cuda.synchronize()
    counts_ok = bool((d_res[:, 1:1 + K].sum(dim=1) == n_total).all().item())
    if not counts_ok:
        raise SystemExit("bench: counts do not sum to the pixel count — the kernel is not doing the work")

    sampler.wait_ready()
    be.setProfiling(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    kernel_ms = []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for i in range(a.steps):
        flush.fill_(i & 255)          # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        step()
        ev[i][1].record()
        ev[i][1].synchronize()
        kernel_ms.append(be.lastAssignMs())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.perf_counter()
    be.setProfiling(False)
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    value = n_total * B * a.steps / total_s / 1e9
    clocks = sampler.summary(t_wall0, t_wall1)

    # ---- e2e: host buffers through the C ABI call, H2D + D2H inside the timed region
    install_nccl_allreduce(be)
    for _ in range(max(1, a.warmup // 2)):
        be.evalPalettes(pal, a.space, flags=flags)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        r = be.evalPalettes(pal, a.space, flags=flags)
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    assert int(r["counts"][0].sum()) == n_total
    e2e_value = n_total * B * a.steps / float(t_e2e.item()) / 1e9
    sampler.stop()

    # ---- the same step through the EXACT PRUNED kernel (HQ_EVAL_PRUNE, csrc/hq_pruned.cu): identical integers, ~K/S times
    # less arithmetic.  Reported beside the exhaustive kernel, never instead of it: `value`, `e2e` and `roofline` above are
    # the exhaustive sweep the north star's roofline is defined on.
    pruned = None
    if a.space == 0:
        d_res2 = torch.zeros_like(d_res)
        def pstep():
            be.evalPalettesDevice(d_pal.data_ptr(), B, K, d_res2.data_ptr(), a.space, flags | EVAL_PRUNE, stream)
            if world > 1:
                dist.all_reduce(d_res2, op=dist.ReduceOp.SUM)
        for _ in range(a.warmup):
            flush.fill_(1)
            pstep()
        torch.cuda.synchronize()
        if not torch.equal(d_res2, d_res):
            raise SystemExit("bench: the pruned kernel's integers differ from the exhaustive kernel's")
        be.setProfiling(True)
        pev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for i in range(a.steps):
            flush.fill_(i & 255)
            pev[i][0].record()
            pstep()
            pev[i][1].record()
        torch.cuda.synchronize()
        stats = be.pruningStats()
        be.setProfiling(False)
        p_ms = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in pev)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(p_ms, op=dist.ReduceOp.MAX)
        p_s = float(p_ms.item()) * 1e-3
        t0 = time.perf_counter()
        for i in range(a.steps):
            r2 = be.evalPalettes(pal, a.space, flags=flags | EVAL_PRUNE)
        t_p = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_p, op=dist.ReduceOp.MAX)
        assert np.array_equal(r2["err_fx"], r["err_fx"]) and np.array_equal(r2["counts"], r["counts"])
        pruned = {"value": n_total * B * a.steps / p_s / 1e9, "unit": UNIT, "ms_per_step": 1e3 * p_s / a.steps,
                  "swasa_evals_per_s": B * a.steps / p_s, "speedup_over_exhaustive": total_s / p_s,
                  "e2e": {"value": n_total * B * a.steps / float(t_p.item()) / 1e9, "unit": UNIT, "evals_per_s": B * a.steps / float(t_p.item())},
                  "chunks": stats["chunks"], "mean_surviving_colours": stats["mean_survivors"], "of_colours": K,
                  "identical_to_exhaustive": True,
                  "note": "exact geometric pruning (hq_pruned.cu): pixels cell-sorted once per image, per (chunk, candidate) only colours that can be "
                          "nearest to some pixel of the chunk are swept; not the kernel the roofline above describes"}

    # ---- the one-time image conversion kernel (HBM-bound by design: 15 B/pixel), CUDA events around the kernel
    be.setProfiling(True)
    rl_ms = []
    for i in range(5):
        flush.fill_(i)
        be.setImageDevice(d_img.data_ptr(), a.width, r1 - r0, stream=stream)
        rl_ms.append(be.lastRgbToLabMs())
    be.setProfiling(False)
    rl_ms = sorted(rl_ms[1:])[len(rl_ms[1:]) // 2]

    # ---- what a search pays ONCE per image before its first evaluation (not part of `value` / `e2e`, which time the
    # per-iteration call as the reference's loop issues it): host image upload + RGB->Lab, and the cell sort of the pruned path
    setup = {}
    be2 = ImageManipulation("CIE76", False, True, local_rank)
    be2.setImage(img)                       # first call allocates
    t0 = time.perf_counter(); be2.setImage(img); setup["set_image_host_ms"] = 1e3 * (time.perf_counter() - t0)
    setup["h2d_image_bytes"] = int(img.nbytes)
    if a.space == 0:
        be2.evalPalettes(pal[:1], a.space, flags=EVAL_PRUNE)
        be2.setImage(img)
        t0 = time.perf_counter(); be2.evalPalettes(pal[:1], a.space, flags=EVAL_PRUNE); t1 = time.perf_counter()
        be2.evalPalettes(pal[:1], a.space, flags=EVAL_PRUNE); t2 = time.perf_counter()
        setup["pruned_sort_ms"] = 1e3 * max(0.0, (t1 - t0) - (t2 - t1))
    # the same image entering as the plugin's float planes (getDataXYCAsFloat, 12 B/px over PCIe, exact pow per channel on the device)
    planes = np.ascontiguousarray((img.astype(np.float64) / 255.0).astype(np.float32).transpose(2, 0, 1))
    be2.setImageFloat(planes)
    t0 = time.perf_counter(); be2.setImageFloat(planes); setup["set_image_f32_planar_host_ms"] = 1e3 * (time.perf_counter() - t0)
    del planes
    be2.close()

    # ---- roofline of the dominant kernel (assign_reduce_kernel): FP32 CUDA-core bound at K=256
    flops_per_launch = 8.0 * K * n_shard * B
    k_ms = sum(kernel_ms) / len(kernel_ms)
    achieved = flops_per_launch / (k_ms * 1e-3) / 1e12
    peak_tf = max(peak["ffma_tflops"], peak["ffma2_tflops"])
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"assign_reduce_w{a.width}_h{a.rows_per_gpu}_k{K}_b{B}")
    hbm_peak = None
    ppath = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        hbm_peak = json.load(open(ppath)).get("hbm_gbs")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * total_s / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, world), "timing": "CUDA events per step on the launching stream, max over ranks; "
                   "L2 flushed (256 MiB write) between timed iterations", "space": "LAB" if a.space == 0 else "SRGB",
                   "parallelism": f"row-shard x{world} + int64 all-reduce" if world > 1 else "single GPU", "device": info["name"]},
        "swasa_evals_per_s": B * a.steps / total_s,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pal.nbytes), "d2h_bytes_per_step": int(B * words * 8),
                "evals_per_s": B * a.steps / float(t_e2e.item()),
                "call": "hq_eval_palettes (host palettes in, host integers out); the image is uploaded once per search, as in the reference"},
        "gpu_launches": 2 * a.steps,
        "clocks": clocks,
        "roofline": {"bound": "fp32", "kernel": "assign_reduce_kernel", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic,
                     "peak_source": "FFMA/FFMA2 microbenchmark (hq_measure_fp32_peak) in this run; MEASURED_PEAKS.json has no FP32 CUDA-core figure",
                     "peak_ffma_tflops": peak["ffma_tflops"], "peak_ffma2_tflops": peak["ffma2_tflops"],
                     "kernel_ms": k_ms, "flops_per_launch": flops_per_launch, "flop_per_pair": 8,
                     "algorithmic_bytes_per_launch": 12 * n_shard, "hbm_gbs_measured_peak": hbm_peak,
                     "hbm_floor_ms": (12 * n_shard / (hbm_peak * 1e9) * 1e3) if hbm_peak else None},
    }
    line["secondary_rooflines"] = [{"kernel": "rgb_to_lab_kernel", "bound": "hbm", "achieved": 15.0 * n_shard / (rl_ms * 1e-3) / 1e9,
                                    "peak": hbm_peak, "unit": "GB/s", "frac": (15.0 * n_shard / (rl_ms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None,
                                    "kernel_ms": rl_ms, "algorithmic_bytes_per_pixel": 15, "runs": "once per image, not per step",
                                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if hbm_peak else None}]
    if pruned is not None:
        line["pruned"] = pruned
    line["once_per_image"] = setup
    if world == 1 and rank == 0 and not a.no_cpu_baseline:
        cb = run_reference_kernels(a, steps=1000, warmup=1, seconds_budget=a.cpu_baseline_seconds)
        port = run_cpu(a, steps=1000, warmup=1, candidates_per_step=1, seconds_budget=a.cpu_baseline_seconds if cb is None else a.cpu_baseline_seconds / 2)
        if cb is None:
            cb = port
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline"]["oracle_port"] = {k: port[k] for k in ("value", "unit", "cores", "sample")}
    be.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        main_reference(a, rank, world)
    else:
        if world != a.gpus and world == 1 and a.gpus > 1:
            raise SystemExit(f"--gpus {a.gpus} needs torchrun: python -m torch.distributed.run --nnodes=1 --nproc-per-node {a.gpus} "
                             f"--master-addr 127.0.0.1 --master-port 29500 bench.py --gpus {a.gpus} ...")
        main_b200(a, rank, local_rank, world)


if __name__ == "__main__":
    main()
