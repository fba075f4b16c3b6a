"""The oracle against everything that pins it: analytic anchors, java.util.Random
known answers (Appendix B) and the committed golden vectors.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from helpers import bits, from_bits, idx_crc, load_golden


def test_lab_constants_match_java_fp32_evaluation(oracle):
    # ScielabProcessor.java:59-61 evaluated in fp32
    d = np.float32(6.0) / np.float32(29.0)
    d2 = np.float32(d * d)
    d3 = np.float32(d2 * d)
    L = oracle.load()
    assert np.float32(L.hqo_lab_constants(0)) == d3
    assert np.float32(L.hqo_lab_constants(1)) == np.float32(np.float32(3.0) * d2)
    assert np.float32(L.hqo_lab_constants(2)) == np.float32(4.0) / np.float32(29.0)
    assert abs(float(d3) - 0.008856452070) < 1e-11


@pytest.mark.parametrize("rgb,lab", [
    ((0, 0, 0), (0.0, 0.0, 0.0)),
    ((1, 1, 1), (100.0, 0.0, -0.03245)),
    ((1, 0, 0), (53.2408, 80.0925, 67.1947)),
    ((0, 1, 0), (87.7347, -86.1827, 83.1638)),
    ((0, 0, 1), (32.2970, 79.1875, -107.8912)),
    ((128 / 255, 128 / 255, 128 / 255), (53.5850, 0.0, -0.01947)),
])
def test_lab_anchors(oracle, rgb, lab):
    got = oracle.srgb_to_lab(np.array(rgb, np.float32))
    assert np.allclose(got, lab, atol=1e-3), (got, lab)


def test_java_random_known_answers(oracle):
    L = oracle.load()
    r = oracle.Rng()
    L.hqo_rng_seed(C.byref(r), 42)
    assert L.hqo_rng_next(C.byref(r), 32) == -1170105035
    L.hqo_rng_seed(C.byref(r), 0)
    assert L.hqo_rng_next(C.byref(r), 32) == -1155484576
    L.hqo_rng_seed(C.byref(r), 0)
    assert L.hqo_rng_next_double(C.byref(r)) == 0.730967787376657
    L.hqo_rng_seed(C.byref(r), 42)
    got = [L.hqo_rng_next_float(C.byref(r)) for _ in range(3)]
    assert [float(g).hex() for g in got] == ["0x1.74833a0000000p-1", "0x1.bfd1400000000p-5", "0x1.5dcf760000000p-1"]
    L.hqo_rng_seed(C.byref(r), 42)
    assert [L.hqo_rng_next_double(C.byref(r)) for _ in range(2)] == [0.7275636800328681, 0.6832234717598454]


def test_golden_lab_vectors(oracle):
    g = load_golden("lab_vectors.json")
    u8 = np.array(g["u8"], np.uint8)
    unit, lab65 = oracle.image_planes(u8, oracle.WHITE_D65, 2)
    _, lab50 = oracle.image_planes(u8, oracle.WHITE_D50, 1)
    assert np.array_equal(bits(unit.T.copy()).ravel(), bits(from_bits(g["unit"])))
    assert np.array_equal(bits(lab65.T.copy()).ravel(), bits(from_bits(g["lab_d65"])))
    assert np.array_equal(bits(lab50.T.copy()).ravel(), bits(from_bits(g["lab_d50"])))
    cols = from_bits(g["float_rgb"], (-1, 3))
    want = from_bits(g["float_lab_d65"], (-1, 3))
    for c, w in zip(cols, want):
        assert np.array_equal(bits(oracle.srgb_to_lab(c)), bits(w))
    # the pixel path (u8 -> LUT) and the palette path (float) agree on u8-representable colours
    for i in range(len(u8)):
        assert np.array_equal(bits(oracle.srgb_to_lab(unit[:, i].copy())), bits(lab65[:, i].copy()))


def test_golden_assign_vectors(oracle):
    from hybridquantization_b200 import synth

    g = load_golden("assign_vectors.json")
    for name, v in g.items():
        img = synth.synth_image(v["w"], v["h"], v["seed"], v["smooth"])
        pal = synth.synth_palettes(v["B"], v["K"])
        for threads in (1, 3):
            r = oracle.assign_reduce(img, pal, v["space"], oracle.WHITE_D65, want_idx=True, threads=threads)
            assert [int(x) for x in r["err_fx"]] == v["err_fx"], name
            assert r["counts"].tolist() == v["counts"], name
            assert r["sums_fx"].tolist() == v["sums_fx"], name
            assert [idx_crc(r["idx"][b], v["w"] * v["h"]) for b in range(v["B"])] == v["idx_crc"], name


def test_golden_swasa_vectors(oracle):
    from hybridquantization_b200 import synth

    g = load_golden("swasa_vectors.json")
    for name, v in g.items():
        img = synth.synth_image(v["w"], v["h"], v["image_seed"], v["smooth"])
        p = oracle.swasa_params(population=v["population"], imax=v["imax"], iTc=v["iTc"], convergence=v["convergence"],
                                space=v["space"], seed=v["seed"])
        best, err, tr = oracle.find_best_quantization(img, v["K"], p, trace=True, threads=2)
        assert float(err).hex() == v["best_error"], name
        assert np.array_equal(bits(best).ravel(), bits(from_bits(v["best_colors"]))), name
        assert [float(x).hex() for x in tr.reshape(-1)] == v["trace"], name


def test_golden_scielab_vectors(oracle):
    from hybridquantization_b200 import synth

    g = load_golden("scielab_vectors.json")
    f, a = oracle.scielab_filters(g["dpi"], g["viewing_distance"])
    assert f.shape[1] == g["taps"]
    assert np.array_equal(bits(f).ravel(), bits(from_bits(g["filters"]))) and np.array_equal(bits(a), bits(from_bits(g["abs3"])))
    img = synth.synth_image(g["w"], g["h"], g["seed"], g["smooth"])
    so = oracle.scielab_image(img, f, a, oracle.WHITE_D65, 3)
    assert np.array_equal(bits(so).ravel(), bits(from_bits(g["scielab_image"])))
    pal = synth.synth_palettes(g["B"], g["K"])
    ev = oracle.scielab_eval(img, f, a, so, pal, g["space"], oracle.WHITE_D65, 2)
    assert [int(v) for v in ev["err_fx"]] == g["err_fx"] and ev["counts"].tolist() == g["counts"]
    quant = oracle.quantize(img, pal[0], g["space"])["rgb"].reshape(img.shape)
    ei = oracle.error_image(img, quant, f, a)
    assert float(ei["deltaE"]).hex() == g["error_image_mean"]
    assert int(ei["errorImageU8"].astype(np.int64).sum()) == g["error_image_u8_sum"]


# ---------------------------------------------------------------- fixtures generated by the REFERENCE'S OWN code
# tests/golden/ref_vectors.json comes from oracle/_ref (the reference's sources compiled for the CPU,
# tests/golden/make_ref_golden.py); these checks need neither /root/reference nor the _ref library.
def test_reference_golden_java_lab(oracle):
    g = load_golden("ref_vectors.json")["java_lab"]
    u8 = np.array(g["u8"], np.uint8)
    assert np.array_equal(bits(oracle.image_planes(u8, oracle.WHITE_D65, 1)[1].T.copy()).ravel(), bits(from_bits(g["lab_d65"])))
    assert np.array_equal(bits(oracle.image_planes(u8, oracle.WHITE_D50, 1)[1].T.copy()).ravel(), bits(from_bits(g["lab_d50"])))
    for c, w in zip(from_bits(g["float_rgb"], (-1, 3)), from_bits(g["float_lab_d65"], (-1, 3))):
        assert np.array_equal(bits(oracle.srgb_to_lab(c)), bits(w))


def test_reference_golden_filters(oracle):
    for v in load_golden("ref_vectors.json")["filters"].values():
        f, a = oracle.scielab_filters(v["dpi"], v["vd"])
        assert f.shape[1] == v["taps"]
        assert np.array_equal(bits(f).ravel(), bits(from_bits(v["filters"]))) and np.array_equal(bits(a), bits(from_bits(v["abs3"])))


def test_reference_golden_opencl_chain(oracle):
    from hybridquantization_b200 import synth

    g = load_golden("ref_vectors.json")["cl"]
    img = synth.synth_image(g["w"], g["h"], g["seed"], g["smooth"])
    f, a = oracle.scielab_filters()
    so = oracle.scielab_image(img, f, a)
    assert np.array_equal(bits(so).ravel(), bits(from_bits(g["scielab_image"])))
    pal = synth.synth_palettes(g["B"], g["K"])
    ev = oracle.scielab_eval(img, f, a, so, pal, oracle.SPACE_SRGB)
    assert [int(v) for v in ev["err_fx"]] == g["err_fx"]
    assert [[int(c > 0) for c in row] for row in ev["counts"]] == g["used"]
    for i in range(g["B"]):  # reference: double sum of floats / N + penalty; build: 2^-24 fixed point
        assert abs(oracle.cost(ev["err_fx"][i], ev["counts"][i], g["w"] * g["h"], 0.5) - float.fromhex(g["costs"][i])) <= 2.0 ** -25
    q = oracle.quantize(img, pal[0], oracle.SPACE_SRGB)
    assert [float(v).hex() for v in q["f32"][:, :3].astype(np.float64).sum(axis=0)] == g["quantize_rgb_sum"]
    assert [int(c > 0) for c in np.bincount(q["idx"], minlength=g["K"])] == g["quantize_used"]


def test_reference_golden_swasa(oracle):
    g = load_golden("ref_vectors.json")["swasa"]
    L = oracle.load()
    p = oracle.swasa_params(**g["params"])
    r = oracle.Rng(); L.hqo_rng_seed(C.byref(r), g["seed"])
    K = g["K"]
    cur = np.zeros((K, 4), np.float32); L.hqo_generate_random_colors(C.byref(r), K, cur.ctypes.data_as(C.c_void_p))
    assert np.array_equal(bits(cur).ravel(), bits(from_bits(g["colors"][0])))
    for it, want in zip(g["iterations"], g["colors"][1:]):
        nxt = np.zeros_like(cur)
        L.hqo_generate_neighboring_colors(C.byref(p), C.byref(r), cur.ctypes.data_as(C.c_void_p), nxt.ctypes.data_as(C.c_void_p), K, it)
        assert np.array_equal(bits(nxt).ravel(), bits(from_bits(want)))
        cur = nxt
    assert [f"{int(np.float32(L.hqo_max_step_width(C.byref(p), i)).view(np.uint32)):08x}" for i in g["iterations"]] == g["step_width"]


def test_reference_golden_whole_search(oracle):
    """the plugin's search run from reference code only vs the oracle's reference-faithful mode"""
    from hybridquantization_b200 import synth

    for name, v in load_golden("ref_vectors.json")["search"].items():
        img = synth.synth_image(v["w"], v["h"], v["image_seed"], True)
        p = oracle.swasa_params(population=v["population"], imax=v["imax"], iTc=v["iTc"], seed=v["seed"], convergence=int(v["convergence"]),
                                space=oracle.SPACE_SRGB, cost_model=1)
        best, err, tr = oracle.find_best_quantization(img, v["K"], p, trace=True)
        want = np.array([float.fromhex(x) for x in v["trace"]])
        assert np.allclose(tr.reshape(-1), want, rtol=0, atol=2.0 ** -24), name
        assert abs(err - float.fromhex(v["best_error"])) <= 2.0 ** -24, name
        assert np.array_equal(bits(best).ravel(), bits(from_bits(v["best_colors"]))), name
