"""Pins the oracle (oracle/hq_oracle.c) against the REFERENCE'S OWN SOURCES compiled for the CPU
(oracle/_ref/libhq_ref.so, built by oracle/ref_build/build_ref.sh from /root/reference where it lies):
every kernel of OptimizedConvolution.cl, SWASA.java, the Java CPU colour helpers and filter-bank
construction of ScielabProcessor.java, and the annealing loop of ImageManipulation.findBestQuantization.
CPU only.  On a machine without /root/reference the prebuilt .so (git-ignored, ships with the gpurun
snapshot) is used; with neither, these tests skip and tests/test_oracle_golden.py::test_reference_golden_*
still checks the oracle against the fixtures this reference build generated (tests/golden/ref_vectors.json)."""
import ctypes as C

import numpy as np
import pytest

from helpers import bits


@pytest.fixture(scope="module")
def ref():
    from oracle import hq_ref

    if not hq_ref.build():
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    hq_ref.load()
    return hq_ref


def _palettes(B, K, seed):
    rng = np.random.default_rng(seed)
    pal = np.zeros((B, K, 4), np.float32)
    pal[..., :3] = rng.random((B, K, 3), dtype=np.float32)
    return pal


# ---------------------------------------------------------------- shim pins of device-defined builtins
def test_shim_cbrt_and_pow_equal_the_oracles_over_the_whole_domain(oracle, ref):
    """refcl::cbrt = (float)cbrtl, oracle = (float)pow(t, 1.0/3.0); refcl::pow(x,2.4f) vs hqo_pow_2p4 — every float
    the path can produce: t in (LABDELTA3, 1.25], pow base in [0.0031, 1]"""
    lo = int(np.float32(0.0088564).view(np.uint32)); hi = int(np.float32(1.25).view(np.uint32))
    step = 1 << 22
    for first in range(lo, hi, step):
        cnt = min(step, hi - first)
        a = np.empty(cnt, np.float32); ref.load().refcl_builtin_range(0, first, cnt, a.ctypes.data_as(C.c_void_p))
        assert np.array_equal(bits(a), bits(oracle.math_range(0, first, cnt)))
    lo = int(np.float32(0.05).view(np.uint32)); hi = int(np.float32(1.0).view(np.uint32)) + 1
    for first in range(lo, hi, step):
        cnt = min(step, hi - first)
        a = np.empty(cnt, np.float32); ref.load().refcl_builtin_range(1, first, cnt, a.ctypes.data_as(C.c_void_p))
        assert np.array_equal(bits(a), bits(oracle.math_range(1, first, cnt)))


# ---------------------------------------------------------------- Java CPU colour helpers (the graded RGB->Lab)
def test_java_lab_constants(oracle, ref):
    L = ref.load()
    assert np.float32(L.refj_lab_constant(2)) == np.float32(oracle.load().hqo_lab_constants(0))            # LABDELTA3
    assert np.float32(3.0) * np.float32(L.refj_lab_constant(1)) == np.float32(oracle.load().hqo_lab_constants(1))


@pytest.mark.parametrize("d50", [False, True])
def test_java_srgb_to_lab_all_u8_colours(oracle, ref, d50):
    """OpptoLab(sRGBtoOpp(px)) (ScielabProcessor.java:279-311, compiled) == oracle for u8 colours: a 2^21-colour
    lattice + 2^19 random ones per white point (the full 2^24 runs in tools/ref_exhaustive.py)"""
    g = np.arange(0, 256, 2, dtype=np.uint8)
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    rnd = np.random.default_rng(5).integers(0, 256, (1 << 19, 3), dtype=np.uint8)
    u8 = np.concatenate([lat, rnd, np.array([[0, 0, 0], [255, 255, 255], [10, 10, 10], [11, 11, 11]], np.uint8)])
    want = oracle.image_planes(u8, oracle.WHITE_D50 if d50 else oracle.WHITE_D65)[1]
    got = ref.srgb_to_lab_java(ref.unit_planes(u8), d50)
    assert np.array_equal(bits(got), bits(want))


def test_java_srgb_to_lab_float_colours(oracle, ref):
    """palette colours are arbitrary floats in [0,1] (incl. the 0.04045 branch point)"""
    rng = np.random.default_rng(6)
    cols = rng.random((20000, 3), dtype=np.float32)
    cols[:4] = [[0.04045, 0.040450003, 0.0404499], [0, 0, 0], [1, 1, 1], [0.5, 0.0, 1.0]]
    got = ref.srgb_to_lab_java(np.ascontiguousarray(cols.T))
    want = np.stack([oracle.srgb_to_lab(c) for c in cols[:3000]])
    assert np.array_equal(bits(got[:, :3000].T.copy()), bits(want))


def test_java_srgb_to_lab_16bit_image_planes(oracle, ref):
    """a 16-bit Icy image after the rescaling conversion (c/65535, HybridQuantization.java:95): the float-image entry of the
    oracle (hqo_image_planes_f32, what hq_set_image_f32_planar is held to) == the compiled Java helpers, both white points"""
    rng = np.random.default_rng(16)
    c = rng.integers(0, 65536, (3, 1 << 18), dtype=np.uint32)
    c[:, :8] = [0, 1, 2650, 2651, 2652, 32768, 65534, 65535]
    planes = (c.astype(np.float64) / 65535.0).astype(np.float32)
    for d50 in (False, True):
        want = oracle.image_planes_f32(planes, oracle.WHITE_D50 if d50 else oracle.WHITE_D65)[1]
        assert np.array_equal(bits(ref.srgb_to_lab_java(planes, d50)), bits(want))


def test_scielab_of_a_float_image(oracle, ref):
    """XYZtoScielab(RGBtoXYZ(.)) on float planes (what the plugin passes, :374-381) == the oracle's float-image entry"""
    rng = np.random.default_rng(17)
    w, h = 41, 29
    planes = (rng.integers(0, 65536, (3, h, w)).astype(np.float64) / 65535.0).astype(np.float32)
    f, a = oracle.scielab_filters()
    got = ref.xyz_to_scielab(ref.rgb_to_xyz(planes.reshape(3, -1)), ref.pack_filters(f, a), w, ref.D65)
    want = oracle.scielab_image_f32(planes, f, a)
    assert np.array_equal(bits(got[:, :3].T.copy()), bits(want))


def test_cie94_branch_of_the_ciede_kernel(oracle, ref):
    """scope row f4 (not built on the GPU): the oracle's restatement of the CIE94 branch (cl:217-226) against the
    reference kernel compiled with -DCIE94 — same bits, and NaN exactly where the reference's sqrt goes negative"""
    rng = np.random.default_rng(94)
    n = 1 << 20
    a = np.stack([rng.uniform(0, 100, n), rng.uniform(-90, 100, n), rng.uniform(-110, 95, n)], 1).astype(np.float32)
    b = (a + rng.normal(0, 8, a.shape)).astype(np.float32)
    # collinear chroma vectors (same hue, different chroma): da^2 + db^2 - dC^2 cancels and rounding decides its sign
    k = n // 4
    scale = rng.uniform(0.2, 3.0, k).astype(np.float32)
    b[:k, 1] = a[:k, 1] * scale; b[:k, 2] = a[:k, 2] * scale
    b[k:k + 1000] = a[k:k + 1000]                     # identical pixels
    a[k + 1000:k + 2000, 1:] = 0                      # greys
    got, want = ref.ciede94(a, b), oracle.delta_e94(a, b)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w)
    assert np.array_equal(bits(got[~nan_g]), bits(want[~nan_w]))
    # the reference's latent NaN is real: ~40 % of collinear pairs, ~1 in 4 comparisons against a grey, and about 1 in 4,000
    # GENERIC pairs — any real image's mean CIE94 error is NaN, which is why the mode is not built on the GPU
    assert nan_g[:k].sum() > k // 10 and not nan_g[k:k + 1000].any() and nan_g[k + 1000:k + 2000].any()
    assert 0 < nan_g[k + 2000:].sum() < (n - k - 2000) // 1000


# ---------------------------------------------------------------- filter bank (ScielabProcessor ctor)
@pytest.mark.parametrize("dpi,vd", [(72, 45.0), (96, 50.0), (150, 30.0), (300, 60.0), (20, 100.0), (600, 20.0), (224, 57.0)])
def test_filter_bank(oracle, ref, dpi, vd):
    f, a = oracle.scielab_filters(dpi, vd)
    rf, ra = ref.scielab_filters(dpi, vd)
    assert f.shape == rf.shape and np.array_equal(bits(f), bits(rf)) and np.array_equal(bits(a), bits(ra))


# ---------------------------------------------------------------- OpenCL kernels
@pytest.mark.parametrize("w,h,smooth", [(48, 40, True), (37, 53, False), (21, 10, True), (10, 64, False)])
def test_scielab_of_the_original_image(oracle, ref, w, h, smooth):
    """(sizes >= the filter half-width 10: below it the reference's single reflection reads out of bounds, cl:20-27, and the
    product refuses the image)  RGB2XYZ, XYZ2Opp, convolve4Channels x4, convolve1Channel x2, Opp2LAB in the order of XYZtoScielab :285-370"""
    img = oracle.synth_image(w, h, 1000 + w, smooth)
    f, a = oracle.scielab_filters()
    got = ref.srgb_to_scielab(img, ref.pack_filters(f, a))
    want = oracle.scielab_image(img, f, a)
    assert np.array_equal(bits(got[:, :3].T.copy()), bits(want)) and not got[:, 3].any()


@pytest.mark.parametrize("K", [2, 16, 37, 256])
def test_quantize_kernel(oracle, ref, K):
    """quantize (cl:147-170): nearest colour by sRGB float4 distance, strict '<', used flags"""
    img = oracle.synth_image(64, 48, 77 + K, K % 2 == 0)
    pal = _palettes(1, K, K)[0]
    out, used = ref.quantize(ref.makeinline(ref.unit_planes(img)), pal)
    o = oracle.quantize(img, pal, space=oracle.SPACE_SRGB)
    same = (out == o["f32"]).all(axis=1)
    if not same.all():  # only sqrt-merged near ties may differ: gap < 1e-5
        px = ref.unit_planes(img).T[~same]
        d_ref = np.linalg.norm(px - out[~same][:, :3], axis=1); d_or = np.linalg.norm(px - o["f32"][~same][:, :3], axis=1)
        assert np.all(np.abs(d_ref - d_or) < 1e-5) and (~same).sum() < 4
    else:
        assert np.array_equal(used != 0, np.bincount(o["idx"], minlength=K) > 0)


@pytest.mark.parametrize("w,h,K,smooth", [(48, 40, 16, True), (40, 32, 9, False), (33, 31, 64, True)])
def test_candidate_chain(oracle, ref, w, h, K, smooth):
    """quantizeAndConvertToOpp -> computeScielabKernelsTemp -> computeScielabKernelsEnd -> Opp2LAB -> CIEDE, per
    candidate as computeQuantizationErrorPopulation :620-727 enqueues them; host mean of :736-768"""
    img = oracle.synth_image(w, h, 31 * w + h, smooth)
    f, a = oracle.scielab_filters()
    packed = ref.pack_filters(f, a)
    so4 = ref.srgb_to_scielab(img, packed)
    so = oracle.scielab_image(img, f, a)
    pal = _palettes(3, K, w)
    sw = ref.Swasa(delta=0.5)
    costs, det = ref.eval_population(ref.makeinline(ref.unit_planes(img)), so4, w, packed, pal, sw.h, depth=3, details=True)
    ev = oracle.scielab_eval(img, f, a, so, pal, oracle.SPACE_SRGB)
    for i in range(3):
        fx = int(np.rint(det[i]["err"].astype(np.float64) * 2.0 ** 24).astype(np.int64).sum())
        assert fx == int(ev["err_fx"][i])
        assert np.array_equal(det[i]["used"] != 0, ev["counts"][i] > 0)
        want = oracle.cost(ev["err_fx"][i], ev["counts"][i], w * h, 0.5)
        # the reference sums floats in double; the build sums 2^-24 fixed point: |diff| <= 2^-25 per pixel mean
        assert abs(costs[i] - want) <= 2.0 ** -25 + 1e-12 * abs(want)


# ---------------------------------------------------------------- SWASA.java
def test_rng_pin_known_answers(ref):
    L = ref.load()
    ref.seed(42)
    assert [float(L.refj_nextFloat()).hex() for _ in range(3)] == ["0x1.74833a0000000p-1", "0x1.bfd1400000000p-5", "0x1.5dcf760000000p-1"]
    ref.seed(42)
    assert [L.refj_nextDouble() for _ in range(2)] == [0.7275636800328681, 0.6832234717598454]
    ref.seed(0)
    assert L.refj_nextDouble() == 0.730967787376657


def test_swasa_methods(oracle, ref):
    L = oracle.load()
    kw = dict(population=3, imax=700, iTc=7, delta=0.25, conv_delay=0.4, conv_spread=0.2, t0=15.0, alpha=0.93, s0=80.0, beta=9.0)
    sw = ref.Swasa(**kw)
    p = oracle.swasa_params(**kw)
    for it in (1, 2, 10, 350, 699, 700):
        assert np.float32(sw.maxStepWidth(it)) == np.float32(L.hqo_max_step_width(C.byref(p), it))
    r = oracle.Rng()
    ref.seed(77760); L.hqo_rng_seed(C.byref(r), 77760)
    a = sw.generateRandomColors(19)
    b = np.zeros((19, 4), np.float32); L.hqo_generate_random_colors(C.byref(r), 19, b.ctypes.data_as(C.c_void_p))
    assert np.array_equal(bits(a), bits(b))
    for it in (1, 5, 300, 700):
        n1 = sw.generateNeighboringColors(a, it)
        n2 = np.zeros_like(b); L.hqo_generate_neighboring_colors(C.byref(p), C.byref(r), b.ctypes.data_as(C.c_void_p), n2.ctypes.data_as(C.c_void_p), 19, it)
        assert np.array_equal(bits(n1), bits(n2))
        a, b = n1, n2
    assert sw.computePenalty(np.array([0, 1, 0, 3, 0], np.int32)) == 0.75
    assert ref.load().refj_clamp(1.5, 0.0, 1.0) == 1.0 and ref.load().refj_clamp(-0.1, 0.0, 1.0) == 0.0
    assert ref.load().refj_argmin(np.array([3.0, 1.0, 1.0, 2.0]).ctypes.data_as(C.c_void_p), 4) == 1


# ---------------------------------------------------------------- the annealing loop
@pytest.mark.parametrize("kw,K", [(dict(population=4, imax=60, iTc=5), 8), (dict(population=1, imax=40, iTc=5), 5),
                                  (dict(population=3, imax=50, iTc=5, convergence=0), 12), (dict(population=5, imax=80, iTc=3, delta=2.0, t0=0.05), 6)])
def test_annealing_loop_with_the_oracle_cost(oracle, ref, kw, K):
    """findBestQuantization's loop compiled from ImageManipulation.java:490-545 + SWASA.java, scoring candidates with
    the ORACLE's Lab cost: every candidate cost, accept/reject decision, RNG draw and the final palette must equal the
    oracle's own restated loop bit for bit (same cost function on both sides => the loop logic is what is compared)."""
    img = oracle.synth_image(48, 40, 99, True)
    conv = kw.pop("convergence", 1)
    p = oracle.swasa_params(seed=77760, convergence=conv, **kw)
    best, err, tr = oracle.find_best_quantization(img, K, p, trace=True)
    unit, lab = oracle.image_planes(img)

    def evaluate(pal):
        r = oracle.assign_reduce_planes(unit, lab, pal)
        return [oracle.cost(r["err_fx"][i], r["counts"][i], 48 * 40, p.delta) for i in range(pal.shape[0])]

    sw = ref.Swasa(population=p.population, imax=p.imax, iTc=p.iTc, delta=p.delta, conv_delay=p.conv_delay, conv_spread=p.conv_spread,
                   t0=p.t0, alpha=p.alpha, s0=p.s0, beta=p.beta)
    ref.seed(77760)
    rbest, rerr, rtr = ref.find_best_quantization(sw, K, evaluate, bool(conv), trace=True)
    assert np.array_equal(rtr.view(np.uint64), tr.view(np.uint64))
    assert rerr == err and np.array_equal(bits(rbest), bits(best))


def test_whole_plugin_search_reference_vs_oracle(oracle, ref):
    """The plugin's quantization path end to end from reference code only (S-CIELAB of the original, annealing loop,
    OpenCL candidate chain, double mean) vs the oracle in its reference-faithful mode (sRGB assignment, S-CIELAB cost).
    Costs agree to the fixed-point/double-sum difference; decisions and the final palette are identical."""
    img = oracle.synth_image(40, 32, 7, True)
    f, a = oracle.scielab_filters()
    p = oracle.swasa_params(population=3, imax=40, iTc=5, seed=4242, space=oracle.SPACE_SRGB, cost_model=1)
    best, err, tr = oracle.find_best_quantization(img, 8, p, trace=True)
    sw = ref.Swasa(population=3, imax=40, iTc=5, delta=p.delta, conv_delay=p.conv_delay, conv_spread=p.conv_spread, t0=p.t0, alpha=p.alpha, s0=p.s0, beta=p.beta)
    ref.seed(4242)
    rbest, rerr, rtr = ref.reference_plugin_search(img, 8, sw, f, a, trace=True, depth=2)
    assert np.allclose(rtr, tr, rtol=0, atol=2.0 ** -24)
    assert abs(rerr - err) <= 2.0 ** -24 and np.array_equal(bits(rbest), bits(best))
