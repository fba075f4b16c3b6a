#!/usr/bin/env python
"""A/B of the host-buffer evaluation's I/O path (HQ_DIRECT_IO=1: palettes read from pinned host memory, results exported by a
kernel + flag spin; 0: H2D copy, D2H copy, stream wait): wall time of fixed-seed searches, min and median of several runs."""
import json
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

CASES = {"defaults_512x512_k8_p4_i5000": (512, 512, 8, 4, 5000, True), "defaults_1080p_k8_p4_i2000": (1920, 1080, 8, 4, 2000, True),
         "C1_512x512_k16_p4_i5000": (512, 512, 16, 4, 5000, True), "C2_1080p_k256_p4_i1000": (1920, 1080, 256, 4, 1000, True),
         "1080p_k32_p4_i1000": (1920, 1080, 32, 4, 1000, True), "C3_4k_k256_p64_i50": (3840, 2160, 256, 64, 50, False)}


def child():
    import numpy as np
    from hybridquantization_b200 import SWASA, ImageManipulation, synth
    from bench import ClockSampler
    sampler = ClockSampler(0)
    sampler.start()
    be = ImageManipulation("CIE76", False, True, 0)
    out = {}
    sampler.wait_ready()
    w0 = time.perf_counter()
    for name, (w, h, K, P, imax, smooth) in CASES.items():
        be.setImage(synth.synth_image(w, h, synth.SEED_BASE + 2, smooth=smooth))
        ts = []
        for _ in range(7):
            t0 = time.perf_counter()
            best, err, _, its = be.findBestQuantization(K, SWASA(population=P, imax=imax, seed=77760))
            ts.append(time.perf_counter() - t0)
        out[name] = {"min_s": min(ts[1:]), "median_s": float(np.median(ts[1:])), "us_per_iteration_min": min(ts[1:]) / (its + 1) * 1e6, "best_error": err}
    out["clocks"] = sampler.summary(w0, time.perf_counter())   # SM clock / throttle reasons over the whole timed window
    be.close()
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        res = {}
        for rep in range(2):
            # one launch per evaluation (round 2) / two launches with direct host I/O (round 1) / copies + stream wait
            for name, env in (("one_launch", {}), ("two_launches_direct_io", {"HQ_SMALL_EVAL": "0"}), ("copies", {"HQ_SMALL_EVAL": "0", "HQ_DIRECT_IO": "0"})):
                r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True, check=True)
                res[f"{name} run{rep}"] = json.loads(r.stdout.strip().splitlines()[-1])
        print(json.dumps(res, indent=1))
