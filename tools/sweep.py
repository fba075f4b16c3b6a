#!/usr/bin/env python
"""Measurement sweep for the scope table's rows (BASELINE configs C2..C5), one JSON
object on stdout.  Device-timed with CUDA events on the launching stream, L2 flushed between
timed launches, inputs resident in HBM.

  rgb_to_lab   GB/s of algorithmic traffic (15 B/pixel) vs the measured HBM copy peak
  k_sweep      assign kernel, 4K image, K = 8..1024, single candidate and 64-candidate batch:
               Gpixel/s and fraction of max(12 B*N / BW_hbm, 8*K*N / P_fp32)
  swasa        wall time of full fixed-seed SWASA runs through the C ABI (C1, C2 shortened)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from hybridquantization_b200 import (EVAL_FORCE_CHUNKED, EVAL_FORCE_DIRECT, EVAL_FORCE_PREFILTER, EVAL_PRUNE, PRUNE_AUTO, PRUNE_OFF, SWASA,  # noqa: E402
                                     ImageManipulation, synth)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--skip-swasa", action="store_true")
    ap.add_argument("--only-scielab", action="store_true")
    ap.add_argument("--only-swasa", action="store_true")
    ap.add_argument("--only-ksweep", type=str, default="", help="comma-separated K list: run only the K sweep for these palette sizes")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    be = ImageManipulation("CIE76", False, True, 0)
    info = be.deviceInfo()
    peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    fp32 = be.measureFp32Peak()
    p_fp32 = max(fp32.values())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"device": info, "hbm_gbs_peak": hbm, "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback", "fp32_peak_tflops": p_fp32, "fp32_probe": fp32}

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        ts = []
        for i in range(reps):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)), float(min(ts))

    # ---- RGB -> Lab (15 B/pixel algorithmic: 3 B read + 12 B written)
    out["rgb_to_lab"] = []
    only_k = tuple(int(v) for v in a.only_ksweep.split(",") if v)
    if only_k:
        a.only_scielab = False; a.only_swasa = True; a.skip_swasa = True
    for (w, h) in (() if (a.only_scielab or a.only_swasa) else ((1920, 1080), (3840, 2160), (8192, 8192))):
        img = synth.synth_image_rows(w, h, synth.SEED_BASE + 3, 0, h)
        d_img = torch.from_numpy(img).to(dev)
        be.setProfiling(True)
        kms = []

        def conv():
            be.setImageDevice(d_img.data_ptr(), w, h, stream=st.cuda_stream)

        for i in range(7):
            flush.fill_(i)
            conv()
            kms.append(be.lastRgbToLabMs())  # CUDA events right around rgb_to_lab_kernel
        be.setProfiling(False)
        ms = float(np.median(kms[2:]))
        gbs = 15.0 * w * h / (ms * 1e-3) / 1e9
        out["rgb_to_lab"].append({"w": w, "h": h, "kernel_ms": ms, "algorithmic_gbs": gbs, "frac_of_hbm_peak": gbs / hbm,
                                  "gpixel_per_s": w * h / (ms * 1e-3) / 1e9, "bytes_per_pixel": 15})
        del d_img

    # ---- K sweep on the 4K image
    w, h = 3840, 2160
    n = w * h
    img = synth.synth_image_rows(w, h, synth.SEED_BASE + 5, 0, h)
    d_img = torch.from_numpy(img).to(dev)
    be.setImageDevice(d_img.data_ptr(), w, h, stream=st.cuda_stream)
    out["k_sweep"] = []
    ks = only_k if only_k else (() if (a.only_scielab or a.only_swasa) else (8, 16, 32, 64, 128, 256, 512, 1024))
    for K in ks:
        for B in (1, 64):
            if a.quick and B == 64 and K > 256:
                continue
            pal = synth.synth_palettes(B, K)
            d_pal = torch.from_numpy(pal).to(dev)
            words = be.resultWords(K, 0)
            d_res = torch.zeros((B, words), dtype=torch.int64, device=dev)
            row = {"K": K, "B": B}
            for name, fl in (("auto", 0), ("pruned", EVAL_PRUNE), ("direct", EVAL_FORCE_DIRECT), ("chunked", EVAL_FORCE_CHUNKED), ("prefilter", EVAL_FORCE_PREFILTER)):
                if name not in ("auto", "pruned") and (a.quick or B == 64 and K > 256):
                    continue
                ms, best = timed(lambda: be.evalPalettesDevice(d_pal.data_ptr(), B, K, d_res.data_ptr(), 0, fl, st.cuda_stream), reps=3 if B == 64 else 7)
                assert int(d_res[0, 1:1 + K].sum().item()) == n
                t_floor_ms = max(12.0 * n / (hbm * 1e9), 8.0 * K * n * B / (p_fp32 * 1e12)) * 1e3
                row[name] = {"ms": ms, "gpixel_per_s": n * B / (ms * 1e-3) / 1e9, "tflops_algorithmic": 8.0 * K * n * B / (ms * 1e-3) / 1e12,
                             "gbs_algorithmic": 12.0 * n / (ms * 1e-3) / 1e9, "roofline_floor_ms": t_floor_ms, "frac_of_roofline": t_floor_ms / ms,
                             "bound": "hbm" if 12.0 * n / (hbm * 1e9) > 8.0 * K * n * B / (p_fp32 * 1e12) else "fp32"}
            out["k_sweep"].append(row)

    # ---- S-CIELAB stage (next row 1): full reference cost chain per candidate, through the host-buffer C ABI
    out["scielab"] = []
    for (w, h, K, B) in (() if a.only_swasa else ((3840, 2160, 256, 4),) if a.only_scielab else ((1920, 1080, 256, 4), (3840, 2160, 256, 4), (3840, 2160, 256, 16))):
        img = synth.synth_image_rows(w, h, synth.SEED_BASE + 2, 0, h)
        be.setImage(img)
        be.scielabConfigure(72, 45.0)
        t0 = time.perf_counter(); be.scielabImage(); t_img = time.perf_counter() - t0
        pal = synth.synth_palettes(B, K)
        be.evalPalettesScielab(pal, 1)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            r = be.evalPalettesScielab(pal, 1)
        dt = (time.perf_counter() - t0) / reps
        be.setPruning(PRUNE_OFF)   # the same chain with the exhaustive assignment kernel
        be.evalPalettesScielab(pal, 1)
        t0 = time.perf_counter()
        for _ in range(reps):
            r2 = be.evalPalettesScielab(pal, 1)
        dt_ex = (time.perf_counter() - t0) / reps
        be.setPruning(PRUNE_AUTO)
        assert np.array_equal(r["err_fx"], r2["err_fx"]) and np.array_equal(r["counts"], r2["counts"])
        out["scielab"].append({"w": w, "h": h, "K": K, "B": B, "scielab_of_original_s": t_img, "eval_ms": dt * 1e3, "evals_per_s": B / dt,
                               "gpixel_per_s": B * w * h / dt / 1e9, "eval_ms_exhaustive_assignment": dt_ex * 1e3,
                               "evals_per_s_exhaustive_assignment": B / dt_ex})

    if not a.skip_swasa:
        out["swasa"] = []
        # full fixed-seed searches through the C ABI, exhaustive scoring vs exact pruned scoring (same trajectory, same palette)
        for name, (w, h, K, P, imax, smooth) in {"C1_512x512_k16_p4_i5000": (512, 512, 16, 4, 5000, True),
                                                 "C2_1920x1080_k256_p4_i" + ("100" if a.quick else "1000"): (1920, 1080, 256, 4, 100 if a.quick else 1000, True),
                                                 "C2u_1920x1080_k256_p4_uniform_i" + ("100" if a.quick else "1000"): (1920, 1080, 256, 4, 100 if a.quick else 1000, False),
                                                 "C3_3840x2160_k256_p64_i" + ("10" if a.quick else "100"): (3840, 2160, 256, 64, 10 if a.quick else 100, False)}.items():
            img = synth.synth_image(w, h, synth.SEED_BASE + 2, smooth=smooth)
            be.setImage(img)
            row = {"config": name}
            for mode, label in ((PRUNE_OFF, "exhaustive"), (PRUNE_AUTO, "pruned_auto")):
                be.setPruning(mode)
                sw = SWASA(population=P, imax=imax, seed=77760)
                t0 = time.perf_counter()
                best, err, _, its = be.findBestQuantization(K, sw)
                dt = time.perf_counter() - t0
                row[label] = {"seconds": dt, "iterations": its, "evals_per_s": (its + 1) * P / dt, "best_error": err,
                              "gpixel_per_s": (its + 1) * P * w * h / dt / 1e9, "best_palette_crc": int(np.bitwise_xor.reduce(best.view(np.uint32).ravel()))}
            row["identical"] = row["exhaustive"]["best_error"] == row["pruned_auto"]["best_error"] and \
                row["exhaustive"]["best_palette_crc"] == row["pruned_auto"]["best_palette_crc"]
            row["speedup"] = row["exhaustive"]["seconds"] / row["pruned_auto"]["seconds"]
            out["swasa"].append(row)
        be.setPruning(PRUNE_AUTO)
    be.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
