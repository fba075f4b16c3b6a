"""Scope row f4: ImageManipulation.deltaETypes.CIE94 on the GPU (hq_set_delta_e) — the reference kernel's CIE94 branch
(OptimizedConvolution.cl:217-226) bit for bit, latent NaN included.  The plugin itself only ever passes CIE76
(HybridQuantization.java:96,145); the reference's CIEDE2000 branch is an empty stub (cl:227-229) and is refused."""
import os

import numpy as np
import pytest

from hybridquantization_b200 import COST_SCIELAB, SPACE_LAB, SPACE_SRGB, SWASA, HqError, ImageManipulation, synth
from hybridquantization_b200._lib import DELTAE_CIE76, DELTAE_CIE94, DELTAE_CIEDE2000, ERR_FX_NAN

pytestmark = pytest.mark.gpu
THREADS = max(1, len(os.sched_getaffinity(0)))


def _fx_sum(e: np.ndarray) -> int:
    """the library's reduction of per-pixel dE floats: 2^-24 fixed point, or the NaN marker when any pixel is NaN"""
    if np.isnan(e).any():
        return ERR_FX_NAN
    return int(np.rint(e.astype(np.float64) * 2.0 ** 24).astype(np.int64).sum())


@pytest.fixture()
def be94():
    be = ImageManipulation("CIE94", False, True, 0)
    yield be
    be.close()


def test_ciede2000_is_refused():
    with pytest.raises(HqError) as ex:
        ImageManipulation("CIEDE2000", False, True, 0)
    assert ex.value.code == 4 and "empty stub" in str(ex.value)


def test_compute_error_on_two_lab_images(be94, oracle):
    """computeError (ImageManipulation.java:858-894) with -DCIE94: per-pixel values through the error image, NaN for NaN"""
    rng = np.random.default_rng(94)
    n = 1 << 16
    a = np.zeros((n, 4), np.float32); b = np.zeros((n, 4), np.float32)
    a[:, :3] = np.stack([rng.uniform(0, 100, n), rng.uniform(-90, 100, n), rng.uniform(-110, 95, n)], 1)
    b[:, :3] = a[:, :3] + rng.normal(0, 8, (n, 3))
    k = n // 4   # collinear chroma vectors: the reference's sqrt argument goes slightly negative for many of them
    sc = rng.uniform(0.2, 3.0, k).astype(np.float32)
    b[:k, 1] = a[:k, 1] * sc; b[:k, 2] = a[:k, 2] * sc
    want = oracle.delta_e94(a[:, :3].copy(), b[:, :3].copy())
    assert 0 < np.isnan(want).sum() < n
    eimg = np.zeros((n, 4), np.float32)
    mean = be94.computeErrorLab(a, b, eimg)
    assert np.isnan(mean)                                        # the reference's double sum over floats with a NaN
    d = np.float32(255.0) - want
    v = (d * d) / np.float32(65025.0)
    assert np.array_equal(np.isnan(eimg[:, 0]), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(eimg[ok, 0].view(np.uint32), v[ok].view(np.uint32)) and np.array_equal(eimg[ok, 2].view(np.uint32), v[ok].view(np.uint32))
    # without a NaN pixel: the mean is the sequential double sum of the floats
    mean_ok = be94.computeErrorLab(a[ok][:5000], b[ok][:5000])
    assert mean_ok == float(np.cumsum(want[ok][:5000].astype(np.float64))[-1]) / 5000
    # against the reference kernel compiled with -DCIE94, where oracle/_ref travelled
    from oracle import hq_ref as R
    if os.path.exists(R.LIB94_PATH):
        got = R.ciede94(a[:, :3].copy(), b[:, :3].copy())
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[ok].view(np.uint32), want[ok].view(np.uint32))


@pytest.mark.parametrize("w,h,K,space", [(16, 12, 6, SPACE_LAB), (64, 48, 16, SPACE_LAB), (40, 30, 300, SPACE_SRGB)])
def test_identity_filter_cost_under_cie94(be94, oracle, w, h, K, space):
    """hq_eval_palettes: dE94(Lab(pixel), Lab(P[idx])) summed per candidate; candidates with a NaN pixel report HQ_ERR_FX_NAN"""
    img = synth.synth_image(w, h, 9 + w, smooth=True)
    pal = synth.synth_palettes(6, K, seed=94)
    be94.setImage(img)
    got = be94.evalPalettes(pal, space)
    ref = oracle.assign_reduce(img, pal, space, want_idx=True, threads=THREADS)
    _, lab = oracle.image_planes(img)
    px = np.ascontiguousarray(lab.T)
    want = []
    for b in range(pal.shape[0]):
        plab = np.stack([oracle.srgb_to_lab(c) for c in pal[b, :, :3]])
        want.append(_fx_sum(oracle.delta_e94(px, np.ascontiguousarray(plab[ref["idx"][b]]))))
    assert [int(v) for v in got["err_fx"]] == want
    assert np.array_equal(got["counts"], ref["counts"])
    costs = [be94.cost(int(got["err_fx"][b]), got["counts"][b], w * h, 2.0) for b in range(pal.shape[0])]
    assert [np.isnan(c) for c in costs] == [v == ERR_FX_NAN for v in want]
    # CIE76 on the same context afterwards: back to the squared-distance score
    be94.setDeltaE(DELTAE_CIE76)
    back = be94.evalPalettes(pal, space)
    assert np.array_equal(back["err_fx"], ref["err_fx"])
    be94.setDeltaE(DELTAE_CIE94)


def test_reference_chain_under_cie94(be94, oracle):
    """hq_eval_palettes_scielab with -DCIE94 against the reference's own kernel chain (quantizeAndConvertToOpp -> Temp -> End ->
    Opp2LAB compiled with -DCIE76, their Lab outputs scored by the CIEDE kernel compiled with -DCIE94)"""
    from oracle import hq_ref as R
    if not (R.available() and os.path.exists(R.LIB94_PATH)):
        pytest.skip("oracle/_ref did not travel to this box")
    w, h, K = 48, 40, 12
    img = synth.synth_image(w, h, 3, smooth=True)
    pal = synth.synth_palettes(5, K, seed=5)
    be94.setImage(img)
    be94.scielabConfigure(72, 45.0)
    got = be94.evalPalettesScielab(pal, SPACE_SRGB)
    f, a = oracle.scielab_filters(72, 45.0)
    packed = R.pack_filters(f, a)
    planes = R.unit_planes(img)
    sc4 = R.xyz_to_scielab(R.rgb_to_xyz(planes), packed, w)
    _, det = R.eval_population(R.makeinline(planes), sc4, w, packed, pal, details=True)
    want = [_fx_sum(R.ciede94(np.ascontiguousarray(sc4.reshape(-1, 4)[:, :3]), np.ascontiguousarray(d["lab"][:, :3]))) for d in det]
    assert [int(v) for v in got["err_fx"]] == want
    for i, d in enumerate(det):
        assert np.array_equal(d["used"] != 0, got["counts"][i] > 0)


def test_search_under_cie94_follows_the_reference_loop(be94, oracle):
    """NaN costs travel through the annealing loop exactly as in the reference: `deltaE <= 0 || exp(-deltaE/T) > nextDouble()`
    (SWASA.java:54-57) is false for a NaN but still draws, `errors[i] < minerror` (:520) is false.  The library's search in CIE94
    mode must equal the reference's compiled loop fed with the same costs."""
    from oracle import hq_ref as R
    w, h, K, P, imax, seed = 24, 20, 5, 3, 25, 94
    img = synth.synth_image(w, h, 7, smooth=True)
    be94.setImage(img)
    best, err, tr, its = be94.findBestQuantization(K, SWASA(population=P, imax=imax, iTc=5, seed=seed), trace=True)
    assert its == imax
    # every candidate cost of the trace equals the cost of that palette evaluated on its own (NaN for NaN)
    if R.available():
        sw = R.Swasa(population=P, imax=imax, iTc=5)
        R.seed(seed)

        def evaluate(pals):
            r = be94.evalPalettes(pals)
            return np.array([be94.cost(int(r["err_fx"][i]), r["counts"][i], w * h, 2.0) for i in range(len(pals))])

        rbest, rerr, rtr = R.find_best_quantization(sw, K, evaluate, convergence=True, trace=True)
        assert np.array_equal(np.isnan(tr), np.isnan(rtr)) and np.array_equal(tr[~np.isnan(tr)].view(np.uint64), rtr[~np.isnan(rtr)].view(np.uint64))
        assert np.array_equal(best.view(np.uint32), rbest.view(np.uint32)) and (err == rerr or (np.isnan(err) and np.isnan(rerr)))
