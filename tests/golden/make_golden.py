"""Generates tests/golden/*.json from the CPU oracle (oracle/hq_oracle.c).

The reference ships no golden vectors, so these fixtures pin the ORACLE
against regressions and travel to the GPU box, where the CUDA path is checked against them.
Re-run only when the oracle's definition changes on purpose:
    python tests/golden/make_golden.py
Floats are stored as IEEE-754 bit patterns (hex) so that comparisons are exact.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import hq_oracle as O  # noqa: E402
from hybridquantization_b200 import synth  # noqa: E402


def fbits(a):
    return [f"{int(v):08x}" for v in np.asarray(a, np.float32).reshape(-1).view(np.uint32)]


def main():
    O.build()
    # 1. Lab of fixed sRGB colours (u8 triples and arbitrary floats), both white points
    rng = np.random.default_rng(20261018)
    u8 = np.concatenate([np.array([[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [128, 128, 128],
                                   [10, 10, 10], [11, 11, 11], [1, 2, 3], [254, 1, 127]], np.uint8),
                         rng.integers(0, 256, (54, 3), dtype=np.uint8)])
    unit, lab65 = O.image_planes(u8, O.WHITE_D65, 1)
    _, lab50 = O.image_planes(u8, O.WHITE_D50, 1)
    cols = rng.random((64, 3), dtype=np.float32)
    cols[0] = [0.04045, 0.040450003, 0.0404499]
    lab_f = np.stack([O.srgb_to_lab(c, O.WHITE_D65) for c in cols])
    json.dump({"u8": u8.tolist(), "unit": fbits(unit.T), "lab_d65": fbits(lab65.T), "lab_d50": fbits(lab50.T),
               "float_rgb": fbits(cols), "float_lab_d65": fbits(lab_f),
               "constants": fbits([O.load().hqo_lab_constants(i) for i in range(3)])},
              open(os.path.join(HERE, "lab_vectors.json"), "w"), indent=0)

    # 2. assign + reduce on small synthetic images (uniform and smooth), K=16 and K=37, B=3
    out = {}
    for name, (w, h, K, smooth) in {"uniform_61x47_k16": (61, 47, 16, False), "smooth_64x48_k37": (64, 48, 37, True),
                                    "uniform_33x31_k256": (33, 31, 256, False)}.items():
        img = synth.synth_image(w, h, synth.SEED_BASE + 1, smooth)
        pal = synth.synth_palettes(3, K)
        for space in (O.SPACE_LAB, O.SPACE_SRGB):
            r = O.assign_reduce(img, pal, space, O.WHITE_D65, want_idx=True, threads=1)
            out[f"{name}_space{space}"] = {
                "w": w, "h": h, "K": K, "smooth": smooth, "seed": synth.SEED_BASE + 1, "B": 3, "space": space,
                "err_fx": [int(v) for v in r["err_fx"]], "counts": r["counts"].tolist(), "sums_fx": r["sums_fx"].tolist(),
                "idx_crc": [int(np.bitwise_xor.reduce((r["idx"][b].astype(np.uint64) + 1) * (np.arange(w * h, dtype=np.uint64) * 2654435761 % 4294967291))) for b in range(3)],
            }
    json.dump(out, open(os.path.join(HERE, "assign_vectors.json"), "w"), indent=0)

    # 3. a short fixed-seed SWASA run (trajectory of candidate costs, final palette)
    img = synth.synth_image(48, 40, synth.SEED_BASE + 1, True)
    runs = {}
    for name, kw in {"p4_k8_i60": dict(population=4, imax=60, K=8), "p1_k5_i40": dict(population=1, imax=40, K=5),
                     "p3_k12_i50_noconv_srgb": dict(population=3, imax=50, K=12, convergence=0, space=O.SPACE_SRGB)}.items():
        K = kw.pop("K")
        p = O.swasa_params(seed=77760, iTc=5, **kw)
        best, err, tr = O.find_best_quantization(img, K, p, trace=True, threads=1)
        runs[name] = {"w": 48, "h": 40, "smooth": True, "image_seed": synth.SEED_BASE + 1, "K": K, "population": p.population,
                      "imax": p.imax, "iTc": p.iTc, "convergence": p.convergence, "space": p.space, "seed": 77760,
                      "best_error": float(err).hex(), "best_colors": fbits(best), "trace": [float(v).hex() for v in tr.reshape(-1)]}
    json.dump(runs, open(os.path.join(HERE, "swasa_vectors.json"), "w"), indent=0)
    # 4. S-CIELAB stage (next row 1): filter bank, S-CIELAB of a small image, candidate costs, error image
    f, a = O.scielab_filters(72, 45.0)
    img = synth.synth_image(40, 32, synth.SEED_BASE + 1, True)
    so = O.scielab_image(img, f, a, O.WHITE_D65, 1)
    pal = synth.synth_palettes(3, 9)
    ev = O.scielab_eval(img, f, a, so, pal, O.SPACE_SRGB, O.WHITE_D65, 1)
    quant = O.quantize(img, pal[0], O.SPACE_SRGB)["rgb"].reshape(img.shape)
    ei = O.error_image(img, quant, f, a, O.WHITE_D65, 1)
    json.dump({"dpi": 72, "viewing_distance": 45.0, "filters": fbits(f), "abs3": fbits(a), "taps": int(f.shape[1]),
               "w": 40, "h": 32, "seed": synth.SEED_BASE + 1, "smooth": True, "scielab_image": fbits(so),
               "K": 9, "B": 3, "space": O.SPACE_SRGB, "err_fx": [int(v) for v in ev["err_fx"]], "counts": ev["counts"].tolist(),
               "error_image_mean": float(ei["deltaE"]).hex(), "error_image_u8_sum": int(ei["errorImageU8"].astype(np.int64).sum())},
              open(os.path.join(HERE, "scielab_vectors.json"), "w"), indent=0)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
