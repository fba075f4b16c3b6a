"""Generates tests/golden/ref_vectors.json from the REFERENCE'S OWN SOURCES compiled for the CPU
(oracle/_ref/libhq_ref.so <- oracle/ref_build/build_ref.sh <- /root/reference).  Runs only where
/root/reference exists (the build container); the fixture travels and pins the oracle — and through it the
CUDA path — everywhere else:
    python tests/golden/make_ref_golden.py
No function of oracle/hq_oracle.c is called here except the synthetic-image generator (input data only).
Floats are stored as IEEE-754 bit patterns (hex)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import hq_ref as R  # noqa: E402
from hybridquantization_b200 import synth  # noqa: E402


def fbits(a):
    return [f"{int(v):08x}" for v in np.asarray(a, np.float32).reshape(-1).view(np.uint32)]


def main():
    assert R.build(force=True), "needs /root/reference"
    out = {}
    # 1. Java CPU helpers: OpptoLab(sRGBtoOpp(.)) for u8 and float colours (ScielabProcessor.java:279-311)
    rng = np.random.default_rng(20261018)
    u8 = np.concatenate([np.array([[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [128, 128, 128], [10, 10, 10], [11, 11, 11]], np.uint8),
                         rng.integers(0, 256, (56, 3), dtype=np.uint8)])
    cols = rng.random((32, 3), dtype=np.float32)
    out["java_lab"] = {"u8": u8.tolist(), "lab_d65": fbits(R.srgb_to_lab_java(R.unit_planes(u8)).T), "lab_d50": fbits(R.srgb_to_lab_java(R.unit_planes(u8), True).T),
                       "float_rgb": fbits(cols), "float_lab_d65": fbits(R.srgb_to_lab_java(np.ascontiguousarray(cols.T)).T)}
    # 2. filter bank of the ScielabProcessor constructor at the plugin defaults and one other geometry
    out["filters"] = {}
    for dpi, vd in ((72, 45.0), (150, 30.0)):
        f, a = R.scielab_filters(dpi, vd)
        out["filters"][f"{dpi}_{vd}"] = {"dpi": dpi, "vd": vd, "taps": int(f.shape[1]), "filters": fbits(f), "abs3": fbits(a)}
    # 3. OpenCL kernels: S-CIELAB of an image, quantize, candidate chain
    f, a = R.scielab_filters()
    packed = R.pack_filters(f, a)
    w, h, seed = 36, 28, synth.SEED_BASE + 1
    img = synth.synth_image(w, h, seed, True)
    so4 = R.srgb_to_scielab(img, packed, threads=1)
    pal = synth.synth_palettes(3, 11)
    rgb4 = R.makeinline(R.unit_planes(img))
    sw = R.Swasa(delta=0.5)
    costs, det = R.eval_population(rgb4, so4, w, packed, pal, sw.h, threads=1, depth=0, details=True)
    q, used = R.quantize(rgb4, pal[0], threads=1)
    out["cl"] = {"w": w, "h": h, "seed": seed, "smooth": True, "B": 3, "K": 11, "scielab_image": fbits(so4[:, :3].T),
                 "err_fx": [int(np.rint(d["err"].astype(np.float64) * 2.0 ** 24).astype(np.int64).sum()) for d in det],
                 "used": [[int(v != 0) for v in d["used"]] for d in det], "costs": [float(c).hex() for c in costs],
                 "quantize_rgb_sum": [float(v).hex() for v in q[:, :3].astype(np.float64).sum(axis=0)], "quantize_used": [int(v != 0) for v in used]}
    # 4. SWASA.java: draw order and arithmetic of the neighbour generation, step widths
    sw = R.Swasa(population=3, imax=500, iTc=7, delta=0.25, conv_delay=0.4, conv_spread=0.2, t0=15.0, alpha=0.93, s0=80.0, beta=9.0)
    R.seed(77760)
    c = sw.generateRandomColors(7)
    seq = [c]
    for it in (1, 2, 250, 500):
        seq.append(sw.generateNeighboringColors(seq[-1], it))
    out["swasa"] = {"params": dict(population=3, imax=500, iTc=7, delta=0.25, conv_delay=0.4, conv_spread=0.2, t0=15.0, alpha=0.93, s0=80.0, beta=9.0),
                    "seed": 77760, "K": 7, "iterations": [1, 2, 250, 500], "colors": [fbits(s) for s in seq],
                    "step_width": fbits([sw.maxStepWidth(i) for i in (1, 2, 250, 500)])}
    # 5. the plugin's whole search from reference code only (annealing loop + OpenCL chain + double mean)
    runs = {}
    for name, kw in {"p3_k8_i40": dict(population=3, imax=40, iTc=5, K=8, seed=4242, convergence=True),
                     "p2_k5_i30_noconv": dict(population=2, imax=30, iTc=4, K=5, seed=99, convergence=False)}.items():
        sw = R.Swasa(population=kw["population"], imax=kw["imax"], iTc=kw["iTc"])
        R.seed(kw["seed"])
        best, err, tr = R.reference_plugin_search(img, kw["K"], sw, f, a, convergence=kw["convergence"], trace=True, threads=1, depth=0)
        runs[name] = dict(kw, w=w, h=h, image_seed=seed, best_error=float(err).hex(), best_colors=fbits(best), trace=[float(v).hex() for v in tr.reshape(-1)])
    out["search"] = runs
    json.dump(out, open(os.path.join(HERE, "ref_vectors.json"), "w"), indent=0)
    print("wrote ref_vectors.json", os.path.getsize(os.path.join(HERE, "ref_vectors.json")), "bytes")


if __name__ == "__main__":
    main()
