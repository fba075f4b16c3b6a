#!/usr/bin/env python
"""us per iteration of a REFERENCE-FAITHFUL search (sRGB assignment + 21-tap S-CIELAB filters + CIE76: what the plugin computes) at the
plugin's default parameters (K colours, population P) on small images: python tools/latency_probe_sc.py [K P imax]"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from hybridquantization_b200 import COST_SCIELAB, SPACE_SRGB, SWASA, ImageManipulation, synth  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
P = int(sys.argv[2]) if len(sys.argv) > 2 else 4
imax = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
out = {}
for w, h in ((512, 512), (1024, 1024), (1920, 1080)):
    be = ImageManipulation("CIE76", False, True, 0)
    be.setImage(synth.synth_image(w, h, synth.SEED_BASE + 2, smooth=True))
    be.scielabConfigure(72, 45.0)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        best, err, _, its = be.findBestQuantization(K, SWASA(population=P, imax=imax, seed=77760, space=SPACE_SRGB, costModel=COST_SCIELAB))
        ts.append((time.perf_counter() - t0) / (its + 1) * 1e6)
    out[f"{w}x{h}_k{K}_p{P}"] = {"us_per_iteration_min": min(ts), "best_error": err, "iterations": its}
    be.close()
print(json.dumps(out))
