"""The CUDA path against the REFERENCE'S OWN code: (1) the fixtures that oracle/_ref — the reference's sources
compiled for the CPU — generated (tests/golden/ref_vectors.json), and (2) when the _ref library travelled to this
box, the compiled reference run live on fresh inputs.  No function of oracle/hq_oracle.c sits between the two."""
import numpy as np
import pytest

from helpers import bits, from_bits, load_golden
from hybridquantization_b200 import COST_SCIELAB, SPACE_SRGB, SWASA, WHITEPOINT_D50, synth

pytestmark = pytest.mark.gpu


def test_lab_of_java_helper_fixture(backend):
    g = load_golden("ref_vectors.json")["java_lab"]
    u8 = np.array(g["u8"], np.uint8).reshape(1, -1, 3)
    backend.setImage(u8)
    assert np.array_equal(bits(backend.labImage().T.copy()).ravel(), bits(from_bits(g["lab_d65"])))
    backend.setImage(u8, WHITEPOINT_D50)
    assert np.array_equal(bits(backend.labImage().T.copy()).ravel(), bits(from_bits(g["lab_d50"])))


def test_opencl_chain_fixture(backend):
    g = load_golden("ref_vectors.json")
    cl = g["cl"]
    img = synth.synth_image(cl["w"], cl["h"], cl["seed"], cl["smooth"])
    backend.setImage(img)
    backend.scielabConfigure(72, 45.0)
    f, a = backend.scielabFilters()
    fx = g["filters"]["72_45.0"]
    assert np.array_equal(bits(f).ravel(), bits(from_bits(fx["filters"]))) and np.array_equal(bits(a), bits(from_bits(fx["abs3"])))
    assert np.array_equal(bits(backend.scielabImage()).ravel(), bits(from_bits(cl["scielab_image"])))
    got = backend.evalPalettesScielab(synth.synth_palettes(cl["B"], cl["K"]), SPACE_SRGB)
    assert [int(v) for v in got["err_fx"]] == cl["err_fx"]
    assert [[int(c > 0) for c in row] for row in got["counts"]] == cl["used"]


def test_whole_search_fixture(backend):
    """4K-independent statement of the north star's last clause on a small case: same seed -> the palette the
    reference's own annealing loop + OpenCL chain produce, bit for bit"""
    for name, v in load_golden("ref_vectors.json")["search"].items():
        img = synth.synth_image(v["w"], v["h"], v["image_seed"], True)
        backend.setImage(img)
        backend.scielabConfigure(72, 45.0)
        sw = SWASA(population=v["population"], imax=v["imax"], iTc=v["iTc"], seed=v["seed"], convergence=v["convergence"], space=SPACE_SRGB,
                   costModel=COST_SCIELAB)
        best, err, tr, its = backend.findBestQuantization(v["K"], sw, trace=True)
        want = np.array([float.fromhex(x) for x in v["trace"]])
        assert its == v["imax"] and np.allclose(tr.reshape(-1), want, rtol=0, atol=2.0 ** -24), name
        assert np.array_equal(bits(best).ravel(), bits(from_bits(v["best_colors"]))), name


@pytest.fixture(scope="module")
def ref():
    from oracle import hq_ref

    if not hq_ref.build():
        pytest.skip("oracle/_ref/libhq_ref.so did not travel to this box")
    return hq_ref


@pytest.mark.parametrize("w,h,K,smooth", [(160, 96, 16, True), (97, 64, 256, False)])
def test_live_reference_kernels(backend, ref, w, h, K, smooth):
    img = synth.synth_image(w, h, 4000 + K, smooth)
    pal = synth.synth_palettes(2, K)
    backend.setImage(img)
    backend.scielabConfigure(72, 45.0)
    f, a = ref.scielab_filters(72, 45.0)
    packed = ref.pack_filters(f, a)
    so4 = ref.srgb_to_scielab(img, packed)
    assert np.array_equal(bits(backend.scielabImage()), bits(so4[:, :3].T.copy()))
    assert np.array_equal(bits(backend.labImage()), bits(ref.srgb_to_lab_java(ref.unit_planes(img))))
    costs, det = ref.eval_population(ref.makeinline(ref.unit_planes(img)), so4, w, packed, pal, details=True)
    got = backend.evalPalettesScielab(pal, SPACE_SRGB)
    for i in range(2):
        assert int(np.rint(det[i]["err"].astype(np.float64) * 2.0 ** 24).astype(np.int64).sum()) == int(got["err_fx"][i])
        assert np.array_equal(det[i]["used"] != 0, got["counts"][i] > 0)


def test_live_reference_on_a_16bit_float_image(backend, ref):
    """hq_set_image_f32_planar with the planes of a 16-bit image (c/65535, the plugin's getDataXYCAsFloat layout) against the
    compiled reference fed with the same floats: Java helpers, S-CIELAB of the original, candidate chain, quantize kernel"""
    w, h, K = 131, 77, 24
    rng = np.random.default_rng(1016)
    planes = (rng.integers(0, 65536, (3, h, w)).astype(np.float64) / 65535.0).astype(np.float32)
    flat = planes.reshape(3, -1)
    pal = synth.synth_palettes(2, K)
    backend.setImageFloat(planes)
    backend.scielabConfigure(72, 45.0)
    f, a = ref.scielab_filters(72, 45.0)
    packed = ref.pack_filters(f, a)
    so4 = ref.xyz_to_scielab(ref.rgb_to_xyz(flat), packed, w, ref.D65)
    assert np.array_equal(bits(backend.labImage()), bits(ref.srgb_to_lab_java(flat)))
    assert np.array_equal(bits(backend.scielabImage()), bits(so4[:, :3].T.copy()))
    costs, det = ref.eval_population(ref.makeinline(flat), so4, w, packed, pal, details=True)
    got = backend.evalPalettesScielab(pal, SPACE_SRGB)
    for i in range(2):
        assert int(np.rint(det[i]["err"].astype(np.float64) * 2.0 ** 24).astype(np.int64).sum()) == int(got["err_fx"][i])
        assert np.array_equal(det[i]["used"] != 0, got["counts"][i] > 0)
    rq, _ = ref.quantize(ref.makeinline(flat), pal[0])
    assert np.array_equal(bits(backend.quantize(pal[0], SPACE_SRGB, want_f32=True)["f32"]), bits(rq))


def test_4k_k256_search_and_image_match_the_compiled_reference(backend, ref):
    """north_star's last clause at the NAMED size: a fixed-seed 3840x2160, K=256 SWASA run (shortened to 10 iterations of a
    population of 4 so that the reference's kernels finish on the host cores in about half a minute) through the CUDA path in
    its reference-faithful mode, against the reference's own annealing loop + OpenCL kernel chain + quantize kernel compiled
    for the CPU: same candidate costs (to the double-sum / fixed-point difference), same palette, same quantised image."""
    w, h, K, P, imax, seed = 3840, 2160, 256, 4, 10, 20261018
    img = synth.synth_image(w, h, synth.SEED_BASE + 3, smooth=True)
    f, a = ref.scielab_filters(72, 45.0)
    sw = ref.Swasa(population=P, imax=imax, iTc=3)
    ref.seed(seed)
    rbest, rerr, rtr = ref.reference_plugin_search(img, K, sw, f, a, trace=True, depth=0)
    backend.setImage(img)
    backend.scielabConfigure(72, 45.0)
    best, err, tr, its = backend.findBestQuantization(K, SWASA(population=P, imax=imax, iTc=3, seed=seed, space=SPACE_SRGB, costModel=COST_SCIELAB), trace=True)
    assert its == imax and np.allclose(tr, rtr, rtol=0, atol=2.0 ** -23)
    assert abs(err - rerr) <= 2.0 ** -23 and np.array_equal(bits(best), bits(rbest))
    # the output image: quantize kernel (cl:147-170) vs hq_quantize in sRGB space
    rq, rused = ref.quantize(ref.makeinline(ref.unit_planes(img)), rbest)
    q = backend.quantize(best, SPACE_SRGB, want_f32=True)
    assert np.array_equal(bits(q["f32"]), bits(rq))
