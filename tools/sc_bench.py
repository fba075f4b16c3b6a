#!/usr/bin/env python
"""S-CIELAB candidate stage A/B: hq_eval_palettes_scielab at 4K, K = 256 with the fused tile kernel (mode 0) and with
round 1's two-kernel path through a 7-plane intermediate (mode 2); wall time of the synchronous host-buffer call, and the
same integers required.  `--once` runs one call per mode (for an ncu launch list / capture)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

from hybridquantization_b200 import PRUNE_AUTO, PRUNE_OFF, SPACE_SRGB, ImageManipulation, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--colors", type=int, default=256)
    ap.add_argument("--modes", default="0,2")
    a = ap.parse_args()
    w, h, K = a.width, a.height, a.colors
    be = ImageManipulation("CIE76", False, True, 0)
    img = synth.synth_image_rows(w, h, synth.SEED_BASE + 3, 0, h)
    be.setImage(img)
    be.scielabConfigure(72, 45.0)
    be.scielabImage()
    out = {"image": f"{w}x{h}", "K": K, "rows": []}
    for B in ((4,) if a.once else (1, 4, 16)):
        pal = synth.synth_palettes(B, K)
        row = {"B": B}
        ref = None
        for mode in [int(m) for m in a.modes.split(",")]:
            be.scielabForceGeneric(mode)
            for prune, pname in ((PRUNE_AUTO, "pruned_assignment"), (PRUNE_OFF, "exhaustive_assignment")):
                if a.once and prune == PRUNE_OFF:
                    continue
                be.setPruning(prune)
                r = be.evalPalettesScielab(pal, SPACE_SRGB)
                if ref is None:
                    ref = r
                assert np.array_equal(r["err_fx"], ref["err_fx"]) and np.array_equal(r["counts"], ref["counts"]), (mode, pname)
                if a.once:
                    continue
                reps = 10
                t0 = time.perf_counter()
                for _ in range(reps):
                    be.evalPalettesScielab(pal, SPACE_SRGB)
                dt = (time.perf_counter() - t0) / reps
                row[f"mode{mode}_{pname}"] = {"ms_per_call": 1e3 * dt, "ms_per_candidate": 1e3 * dt / B, "evals_per_s": B / dt,
                                              "gpixel_per_s": B * w * h / dt / 1e9}
        out["rows"].append(row)
    be.scielabForceGeneric(0)
    be.setPruning(PRUNE_AUTO)
    be.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
