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
    ap.add_argument("--no-sweeps", action="store_true", help="skip the strong_64mp / k_sweep / reference_faithful sub-records (profiling runs)")
    ap.add_argument("--force-sweeps", action="store_true", help="k_sweep at any N (default: N = 1 and 8, BASELINE configs[4])")
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
K_SWEEP = (8, 16, 32, 64, 128, 256, 512, 1024)


def _file_sha(path: str) -> str:
    import hashlib
    return hashlib.sha256(open(path, "rb").read()).hexdigest()[:12]


def measured_traffic(key: str):
    """DRAM bytes per launch of a kernel from its last `ncu --set full` capture (profiles/traffic.json, written by
    tools/ncu_summary.py).  Only reported while the kernel source is the one that was captured (sha of csrc/hq_kernels.cu)."""
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return None, "no ncu capture"
    t = json.load(open(tpath))
    v = t.get(key)
    if isinstance(v, dict):
        src = os.path.join(REPO, "hybridquantization_b200", "csrc", v.get("source_file", "hq_kernels.cu"))
        if v.get("source_sha") and v["source_sha"] != _file_sha(src):
            return None, f"stale: {v.get('capture')} was taken on another version of {v.get('source_file')}"
        return v.get("bytes"), v.get("capture")
    return v, t.get("source")


def main_b200(a, rank: int, local_rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist

    from hybridquantization_b200 import (EVAL_ALLREDUCE, EVAL_FORCE_CHUNKED, EVAL_FORCE_DIRECT, EVAL_FORCE_PREFILTER, EVAL_PRUNE, PRUNE_OFF, SPACE_SRGB, SWASA,
                                         ImageManipulation, build, synth)
    from hybridquantization_b200.dist import close_peer_exchange, install_native_nccl, row_shard, row_shard_with_halo

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sampler = ClockSampler(local_rank)   # started first: by the timed region it has long been sampling
    sampler.start()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)   # plumbing only: barriers, max over ranks, shipping the NCCL id bytes
    if rank == 0:
        build.build_library()
    if world > 1:
        dist.barrier()

    be = ImageManipulation("CIE76", False, True, local_rank)
    info = be.deviceInfo()
    comm = install_native_nccl(be)   # the exchange step runs inside libhq_b200 (ncclAllReduce on the library's communicator)
    K, B = a.colors, a.batch
    flags = {0: 0, 1: EVAL_FORCE_DIRECT, 2: EVAL_FORCE_CHUNKED, 3: EVAL_FORCE_PREFILTER}[a.variant]

    # a dedicated (non-default) stream: the C ABI takes a cudaStream_t and treats NULL as "the
    # context's own stream", so every launch and every timing event below is on this one stream
    bench_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(bench_stream)
    stream = bench_stream.cuda_stream
    assert stream != 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    peak = be.measureFp32Peak()  # FFMA microbenchmark at this device's current clocks
    peak_tf = max(peak["ffma_tflops"], peak["ffma2_tflops"])
    hbm_peak = None
    ppath = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        hbm_peak = json.load(open(ppath)).get("hbm_gbs")
    hbm_source = "MEASURED_PEAKS.json hbm_gbs (of measured)" if hbm_peak else "fallback 6650 GB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"
    hbm_bw = hbm_peak or 6650.0

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_device_steps(ctx, d_pal, Bn, Kn, d_res, fl, steps, warmup, n_total_px, profile=False):
        """W warm-up + K timed launches of the scoring kernel (+ the library's all-reduce), inputs resident in HBM, one CUDA
        event pair per step on the launching stream, L2 flushed between steps; returns (seconds: max over ranks of the summed
        step times, kernel ms list, wall window)."""
        nwords = d_res.numel()

        def step():
            # scoring + exchange in ONE library call: over NVLink peer memory the exporting CTA of the scoring kernel does the
            # all-reduce itself (small payloads), else ncclAllReduce follows on the same stream
            ctx.evalPalettesDevice(d_pal.data_ptr(), Bn, Kn, d_res.data_ptr(), a.space, fl | (EVAL_ALLREDUCE if world > 1 else 0), stream)

        for _ in range(warmup):
            flush.fill_(1)
            step()
        torch.cuda.synchronize()
        if not bool((d_res[:, 1:1 + Kn].sum(dim=1) == n_total_px).all().item()):
            raise SystemExit("bench: counts do not sum to the pixel count — the kernel is not doing the work")
        if profile:
            ctx.setProfiling(True)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kms = []
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(i & 255)          # L2 flush between timed iterations (outside the event pair)
            ev[i][0].record()
            step()
            ev[i][1].record()
            if profile:
                ev[i][1].synchronize()
                kms.append(ctx.lastAssignMs())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        w1 = time.perf_counter()
        if profile:
            ctx.setProfiling(False)
        return max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in ev)) * 1e-3, kms, (w0, w1)

    # ================================================================== the headline workload (weak: 4K rows per GPU)
    H = a.rows_per_gpu * world
    r0, r1 = row_shard(H, world, rank)
    n_shard, n_total = a.width * (r1 - r0), a.width * H
    img = synth.synth_image_rows(a.width, H, synth.SEED_BASE + 3, r0, r1)
    d_img = torch.from_numpy(img).to(dev)
    be.setImageDevice(d_img.data_ptr(), a.width, r1 - r0, stream=stream)
    pal = synth.synth_palettes(B, K)
    d_pal = torch.from_numpy(pal).to(dev)
    words = be.resultWords(K, 0)
    d_res = torch.zeros((B, words), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    sampler.wait_ready()
    total_s, kernel_ms, (t_wall0, t_wall1) = timed_device_steps(be, d_pal, B, K, d_res, flags, a.steps, a.warmup, n_total, profile=True)
    value = n_total * B * a.steps / total_s / 1e9
    clocks = sampler.summary(t_wall0, t_wall1)

    # ---- e2e: host buffers through the C ABI call, H2D + D2H inside the timed region
    for _ in range(max(1, a.warmup // 2)):
        be.evalPalettes(pal, a.space, flags=flags)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        r = be.evalPalettes(pal, a.space, flags=flags)
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    assert int(r["counts"][0].sum()) == n_total
    e2e_value = n_total * B * a.steps / t_e2e / 1e9

    # ---- the same step through the EXACT PRUNED kernel (HQ_EVAL_PRUNE, csrc/hq_pruned.cu): identical integers, ~K/S times
    # less arithmetic.  Reported beside the exhaustive kernel, never instead of it: `value`, `e2e` and `roofline` above are
    # the exhaustive sweep the north star's roofline is defined on.
    pruned = None
    if a.space == 0:
        d_res2 = torch.zeros_like(d_res)
        p_s, _, _ = timed_device_steps(be, d_pal, B, K, d_res2, flags | EVAL_PRUNE, a.steps, a.warmup, n_total)
        if not torch.equal(d_res2, d_res):
            raise SystemExit("bench: the pruned kernel's integers differ from the exhaustive kernel's")
        be.setProfiling(True)
        be.evalPalettesDevice(d_pal.data_ptr(), B, K, d_res2.data_ptr(), a.space, flags | EVAL_PRUNE, stream)
        torch.cuda.synchronize()
        stats = be.pruningStats()
        be.setProfiling(False)
        t0 = time.perf_counter()
        for i in range(a.steps):
            r2 = be.evalPalettes(pal, a.space, flags=flags | EVAL_PRUNE)
        t_p = max_over_ranks(time.perf_counter() - t0)
        assert np.array_equal(r2["err_fx"], r["err_fx"]) and np.array_equal(r2["counts"], r["counts"])
        pruned = {"value": n_total * B * a.steps / p_s / 1e9, "unit": UNIT, "ms_per_step": 1e3 * p_s / a.steps,
                  "swasa_evals_per_s": B * a.steps / p_s, "speedup_over_exhaustive": total_s / p_s,
                  "e2e": {"value": n_total * B * a.steps / t_p / 1e9, "unit": UNIT, "evals_per_s": B * a.steps / t_p},
                  "chunks": stats["chunks"], "mean_surviving_colours": stats["mean_survivors"], "of_colours": K,
                  "identical_to_exhaustive": True,
                  "note": "exact geometric pruning (hq_pruned.cu): pixels cell-sorted once per image, per (chunk, candidate) only colours that can be "
                          "nearest to some pixel of the chunk are swept; not the kernel the roofline above describes"}

    # ================================================================== parity of the sharded path, checked in this very run (N > 1)
    parity = None
    if world > 1:
        PB, PIT = 8, 30
        sub = pal[:PB]
        got = be.evalPalettes(sub, a.space, sums=True, flags=flags)                 # all-reduced totals, exhaustive kernel
        got_pr = be.evalPalettes(sub, a.space, sums=True, flags=flags | EVAL_PRUNE) if a.space == 0 else got
        sw = dict(population=4, imax=PIT, seed=77760, space=a.space)
        best, err, tr, its = be.findBestQuantization(K, SWASA(**sw), n_total=n_total, trace=True)   # sharded search, native all-reduce per iteration
        blob = torch.from_numpy(np.concatenate([got["err_fx"], got["counts"].astype(np.int64).ravel(), got["sums_fx"].ravel(),
                                                tr.view(np.int64).ravel(), best.view(np.int32).astype(np.int64).ravel()])).to(dev)
        ref = blob.clone()
        dist.broadcast(ref, 0)
        same = torch.tensor([int(torch.equal(blob, ref))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        parity = {"candidates": PB, "search_iterations": PIT, "same_on_all_ranks": bool(same.item()),
                  "pruned_equals_exhaustive": all(np.array_equal(got[k], got_pr[k]) for k in ("err_fx", "counts", "sums_fx")),
                  "exchange": f"payloads <= 4,096 words (the {PIT}-iteration searches): " + ("all-reduce over NVLink peer memory by the exporting CTA "
                              "(hq_comm_open_peers; inside the scoring kernel for K <= 32)" if comm.get("peer_exchange") else "ncclAllReduce") +
                              f"; larger ones (the {PB}-candidate evaluation with sums): ncclAllReduce(int64, sum) on libhq_b200's own communicator "
                              f"(hq_comm_init_rank), NCCL {comm['nccl_version']}, {comm['size']} ranks; torch.distributed only shipped the id / handle bytes",
                  "peer_exchange": bool(comm.get("peer_exchange"))}
        # the plugin's default palette size: the one-launch evaluation whose last CTA exchanges over peer memory
        SK = 8
        spal8 = synth.synth_palettes(4, SK)
        got8 = be.evalPalettes(spal8, a.space, sums=True)
        sw8 = dict(population=4, imax=PIT, seed=77760, space=a.space)
        best8, err8, tr8, its8 = be.findBestQuantization(SK, SWASA(**sw8), n_total=n_total, trace=True)
        if rank == 0:   # the whole image on ONE GPU, no exchange: the integers and the trajectory must be the same
            whole = synth.synth_image_rows(a.width, H, synth.SEED_BASE + 3, 0, H)
            single = ImageManipulation("CIE76", False, True, local_rank)
            single.setImage(whole)
            want = single.evalPalettes(sub, a.space, sums=True, flags=flags)
            single.setPruning(PRUNE_OFF)   # the single-GPU search scores exhaustively; the sharded one above under the default policy
            sbest, serr, str_, sits = single.findBestQuantization(K, SWASA(**sw), trace=True)
            want8 = single.evalPalettes(spal8, a.space, sums=True)
            sbest8, serr8, str8, sits8 = single.findBestQuantization(SK, SWASA(**sw8), trace=True)
            parity["small_k_equal_single_gpu"] = bool(all(np.array_equal(got8[k], want8[k]) for k in ("err_fx", "counts", "sums_fx")) and
                                                      np.array_equal(tr8.view(np.uint64), str8.view(np.uint64)) and err8 == serr8 and its8 == sits8
                                                      and np.array_equal(best8.view(np.uint32), sbest8.view(np.uint32)))
            single.close()
            del whole
            parity["evals_equal_single_gpu"] = all(np.array_equal(got[k], want[k]) for k in ("err_fx", "counts", "sums_fx"))
            parity["search_trace_equal_single_gpu"] = bool(np.array_equal(tr.view(np.uint64), str_.view(np.uint64)) and err == serr and its == sits == PIT
                                                           and np.array_equal(best.view(np.uint32), sbest.view(np.uint32)))
            parity["ok"] = bool(parity["evals_equal_single_gpu"] and parity["search_trace_equal_single_gpu"] and parity["same_on_all_ranks"]
                                and parity["pruned_equals_exhaustive"] and parity["small_k_equal_single_gpu"])
        okt = torch.tensor([int(parity.get("ok", True))], device=dev)
        dist.broadcast(okt, 0)
        if not bool(okt.item()):
            if rank == 0:
                print(json.dumps({"error": "sharded result differs from the single-GPU result", "parity": parity}), file=sys.stderr, flush=True)
            raise SystemExit("bench: the sharded path's integers / trajectory differ from the single-GPU ones")

    # ---- the one-time image conversion kernel (HBM-bound by design: 15 B/pixel), CUDA events around the kernel
    be.setProfiling(True)
    rl_ms = []
    for i in range(5):
        flush.fill_(i)
        be.setImageDevice(d_img.data_ptr(), a.width, r1 - r0, stream=stream)
        rl_ms.append(be.lastRgbToLabMs())
    be.setProfiling(False)
    rl_ms = sorted(rl_ms[1:])[len(rl_ms[1:]) // 2]

    # ================================================================== C5: palette-size sweep on ONE 4K image (strong over the N GPUs)
    k_sweep = None
    if not a.no_sweeps and (world in (1, 8) or a.force_sweeps):
        sw_rows = a.rows_per_gpu
        s0_, s1_ = row_shard(sw_rows, world, rank)
        simg = torch.from_numpy(synth.synth_image_rows(a.width, sw_rows, synth.SEED_BASE + 5, s0_, s1_)).to(dev)
        be.setImageDevice(simg.data_ptr(), a.width, s1_ - s0_, stream=stream)
        n_sw = a.width * sw_rows
        k_sweep = {"image": f"{a.width}x{sw_rows} synthetic RGB, rows sharded over {world} GPU(s)", "steps": 5, "warmup": 3, "rows": [],
                   "roofline": "t_floor = max(12 B x pixels-per-GPU / HBM peak, 8 K flop x pixels-per-GPU x candidates / FP32 peak); frac = t_floor / t_measured "
                               "(at N > 1 the measured time includes the all-reduce, the floor does not)",
                   "hbm_gbs_peak": hbm_bw, "hbm_peak_source": hbm_source, "fp32_tflops_peak": peak_tf}
        for Ks in K_SWEEP:
            row = {"K": Ks}
            for Bs in (1, B):
                spal = torch.from_numpy(synth.synth_palettes(Bs, Ks)).to(dev)
                sres = torch.zeros((Bs, be.resultWords(Ks, 0)), dtype=torch.int64, device=dev)
                t_s, _, _ = timed_device_steps(be, spal, Bs, Ks, sres, 0, 5, 3, n_sw)
                ms = 1e3 * t_s / 5
                t_hbm = 12.0 * (n_sw / world) / (hbm_bw * 1e9) * 1e3
                t_fp = 8.0 * Ks * (n_sw / world) * Bs / (peak_tf * 1e12) * 1e3
                row[f"b{Bs}"] = {"ms_per_step": ms, "gpixel_per_s": n_sw * Bs / (ms * 1e-3) / 1e9, "evals_per_s": Bs / (ms * 1e-3),
                                 "bound": "hbm" if t_hbm > t_fp else "fp32", "roofline_floor_ms": max(t_hbm, t_fp), "frac": max(t_hbm, t_fp) / ms}
            k_sweep["rows"].append(row)
        del simg

    # ================================================================== C4: ONE 64 MP image, strong scaling over the N GPUs
    strong = None
    if not a.no_sweeps:
        SW = SH = 8192
        g0, g1 = row_shard(SH, world, rank)
        gimg = torch.from_numpy(synth.synth_image_rows(SW, SH, synth.SEED_BASE + 4, g0, g1)).to(dev)
        be.setImageDevice(gimg.data_ptr(), SW, g1 - g0, stream=stream)
        s_steps, s_warm = 5, 3
        s_s, s_kms, (sw0, sw1) = timed_device_steps(be, d_pal, B, K, d_res, flags, s_steps, s_warm, SW * SH, profile=True)
        k_ms_s = sum(s_kms) / len(s_kms)
        strong = {"workload": f"{SW}x{SH} (64 MP) synthetic RGB, FIXED image, rows sharded over {world} GPU(s), {K}-colour palettes, {B} candidates per launch",
                  "scaling": "strong", "n_gpus": world, "steps": s_steps, "warmup": s_warm, "ms_per_step": 1e3 * s_s / s_steps,
                  "value": SW * SH * B * s_steps / s_s / 1e9, "unit": UNIT, "swasa_evals_per_s": B * s_steps / s_s,
                  "kernel_ms_rank0": k_ms_s, "roofline_frac_rank0": (8.0 * K * SW * (g1 - g0) * B / (k_ms_s * 1e-3) / 1e12) / peak_tf,
                  "clocks": sampler.summary(sw0, sw1)}
        del gimg

    # ================================================================== the plugin's REAL cost model on the headline image (row f1)
    faithful = None
    if not a.no_sweeps and world == 1:
        be.setImage(img)
        be.scielabConfigure(72, 45.0)
        FB = 4
        fpal = pal[:FB]
        be.scielabImage()                          # S-CIELAB of the original: once per image
        for _ in range(2):
            fr = be.evalPalettesScielab(fpal, SPACE_SRGB)
        torch.cuda.synchronize()
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            fr = be.evalPalettesScielab(fpal, SPACE_SRGB)
        dt = (time.perf_counter() - t0) / reps
        assert int(fr["counts"][0].sum()) == n_shard
        # the filter stage alone (index images -> error sums: the kernel row f1 added), device-timed by the library's events
        be.setProfiling(True)
        stage = []
        for _ in range(3):
            be.evalPalettesScielab(fpal, SPACE_SRGB)
            stage.append(be.lastScielabStageMs())
        be.setProfiling(False)
        st_ms, st_nb = min(stage)
        sc_flop = 650.0   # per pixel and candidate: 2 x 21 taps x 7 filter planes x 2 flop + Opp->Lab + dE
        st_tf = sc_flop * n_shard * st_nb / (st_ms * 1e-3) / 1e12
        faithful = {"workload": f"{a.width}x{a.rows_per_gpu}, {K} colours, {FB} candidates per call, sRGB assignment + 21-tap S-CIELAB filters + CIE76 "
                                "(what the reference plugin computes per candidate, ImageManipulation.java:635-699)",
                    "call": "hq_eval_palettes_scielab (host palettes in, host integers out)",
                    "e2e": {"evals_per_s": FB / dt, "gpixel_per_s": FB * n_shard / dt / 1e9, "ms_per_candidate": 1e3 * dt / FB,
                            "h2d_bytes_per_step": int(fpal.nbytes), "d2h_bytes_per_step": int(FB * words * 8)},
                    "roofline": {"bound": "fp32", "kernel": "sc_candidate_strip21_kernel", "flop_per_pixel_filter_stage": sc_flop,
                                 "filter_stage_ms_per_candidate": st_ms / st_nb, "achieved": st_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": st_tf / peak_tf,
                                 "algorithmic_bytes_per_pixel": 13, "hbm_floor_ms_per_candidate": 13.0 * n_shard / (hbm_bw * 1e9) * 1e3,
                                 "assignment": "indices come from the exact pruned kernel (DESIGN.md 4c), which skips most of the 8 K flop per pixel the "
                                               "exhaustive argmin spends: the whole call is therefore NOT put over a flop roofline; the stage is",
                                 "assignment_ms_per_candidate": 1e3 * dt / FB - st_ms / st_nb}}

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
    sampler.stop()

    # ---- roofline of the dominant kernel (assign_reduce_kernel): FP32 CUDA-core bound at K=256
    flops_per_launch = 8.0 * K * n_shard * B
    k_ms = sum(kernel_ms) / len(kernel_ms)
    achieved = flops_per_launch / (k_ms * 1e-3) / 1e12
    traffic, traffic_src = measured_traffic(f"assign_reduce_w{a.width}_h{a.rows_per_gpu}_k{K}_b{B}")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * total_s / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, world), "timing": "CUDA events per step on the launching stream, max over ranks; "
                   "L2 flushed (256 MiB write) between timed iterations", "space": "LAB" if a.space == 0 else "SRGB",
                   "parallelism": (f"row-shard x{world} + exchange inside libhq_b200: ncclAllReduce(int64) (NCCL {comm['nccl_version']}) for this payload; "
                                   f"payloads <= 4,096 words {'over NVLink peer memory in the exporting CTA' if comm.get('peer_exchange') else 'also on NCCL'}" if world > 1 else "single GPU"),
                   "device": info["name"]},
        "swasa_evals_per_s": B * a.steps / total_s,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pal.nbytes), "d2h_bytes_per_step": int(B * words * 8),
                "evals_per_s": B * a.steps / t_e2e,
                "call": "hq_eval_palettes (host palettes in, host integers out; at N > 1 the all-reduce is the library's own); the image is uploaded once per search, as in the reference"},
        "gpu_launches": 2 * a.steps,
        "clocks": clocks,
        "roofline": {"bound": "fp32", "kernel": "assign_reduce_kernel", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": "FFMA/FFMA2 microbenchmark (hq_measure_fp32_peak) in this run; MEASURED_PEAKS.json has no FP32 CUDA-core figure",
                     "peak_ffma_tflops": peak["ffma_tflops"], "peak_ffma2_tflops": peak["ffma2_tflops"],
                     "kernel_ms": k_ms, "flops_per_launch": flops_per_launch, "flop_per_pair": 8,
                     "frac_kind": "ALGORITHMIC: 8 flop per (pixel, colour) pair as SURVEY 8(d) counts the direct form (3 sub, 3 mul, 2 add) over the measured "
                                  "FFMA peak.  The kernel EXECUTES 3 FFMA (6 flop) per pair in its expanded-form prefilter plus an exact direct-form re-check of "
                                  "one 8-colour chunk per pixel (about 6.4 flop per pair in all); FMA-pipe utilisation from ncu is in profiles/",
                     "executed_flop_per_pair": 6.0 + 8.0 * 8 / K if K > 32 else 8.0,
                     "algorithmic_bytes_per_launch": 12 * n_shard, "hbm_gbs_measured_peak": hbm_peak,
                     "hbm_floor_ms": (12 * n_shard / (hbm_peak * 1e9) * 1e3) if hbm_peak else None},
    }
    line["secondary_rooflines"] = [{"kernel": "rgb_to_lab_kernel", "bound": "hbm", "achieved": 15.0 * n_shard / (rl_ms * 1e-3) / 1e9,
                                    "peak": hbm_bw, "unit": "GB/s", "frac": 15.0 * n_shard / (rl_ms * 1e-3) / 1e9 / hbm_bw,
                                    "kernel_ms": rl_ms, "algorithmic_bytes_per_pixel": 15, "runs": "once per image, not per step",
                                    "peak_source": hbm_source}]
    if pruned is not None:
        line["pruned"] = pruned
    if parity is not None:
        line["parity"] = parity
    if strong is not None:
        line["strong_64mp"] = strong
    if k_sweep is not None:
        line["k_sweep"] = k_sweep
    if faithful is not None:
        line["reference_faithful"] = faithful
    line["once_per_image"] = setup
    if world == 1 and rank == 0 and not a.no_cpu_baseline:
        cb = run_reference_kernels(a, steps=1000, warmup=1, seconds_budget=a.cpu_baseline_seconds)
        port = run_cpu(a, steps=1000, warmup=1, candidates_per_step=1, seconds_budget=a.cpu_baseline_seconds if cb is None else a.cpu_baseline_seconds / 2)
        if cb is None:
            cb = port
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline"]["oracle_port"] = {k: port[k] for k in ("value", "unit", "cores", "sample")}
    if world > 1:
        close_peer_exchange(be)   # every rank unmaps the others' mailboxes before any rank frees its own
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
