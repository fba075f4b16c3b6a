#!/usr/bin/env python
"""The launch set of a default-parameter search iteration (512x512, K colours, P candidates): a short fixed-seed search, meant to
be run under `ncu --metrics gpu__time_duration.sum` so that the kernel's own duration can be set against the wall time per
iteration that tools/latency_ab.py measures."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from hybridquantization_b200 import SWASA, ImageManipulation, synth  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
P = int(sys.argv[2]) if len(sys.argv) > 2 else 4
imax = int(sys.argv[3]) if len(sys.argv) > 3 else 40
be = ImageManipulation("CIE76", False, True, 0)
be.setImage(synth.synth_image(512, 512, synth.SEED_BASE + 2, smooth=True))
t0 = time.perf_counter()
best, err, _, its = be.findBestQuantization(K, SWASA(population=P, imax=imax, seed=77760))
print("iterations", its, "us/iteration", (time.perf_counter() - t0) / (its + 1) * 1e6, "error", err)
be.close()
