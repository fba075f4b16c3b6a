"""S-CIELAB stage on the GPU against the oracle: the S-CIELAB representation of the original image and
the candidate costs through the full kernel chain, bit for bit."""
import os

import numpy as np
import pytest

from helpers import bits, from_bits, load_golden
from hybridquantization_b200 import COST_SCIELAB, SPACE_LAB, SPACE_SRGB, SWASA, WHITEPOINT_D50, HqError, HybridQuantization, synth

pytestmark = pytest.mark.gpu
THREADS = max(1, len(os.sched_getaffinity(0)))


@pytest.mark.parametrize("w,h,smooth,wp", [(64, 48, True, 0), (97, 33, False, 0), (10, 10, False, 0), (300, 11, True, 1), (513, 257, True, 0)])
def test_scielab_of_original_matches_oracle(backend, oracle, w, h, smooth, wp):
    img = synth.synth_image(w, h, 77 + w, smooth)
    backend.setImage(img, wp)
    backend.scielabConfigure(72, 45.0)
    f, a = backend.scielabFilters()
    of, oa = oracle.scielab_filters(72, 45.0)
    assert np.array_equal(f.view(np.uint32), of.view(np.uint32)) and np.array_equal(a.view(np.uint32), oa.view(np.uint32))
    got = backend.scielabImage()
    want = oracle.scielab_image(img, of, oa, wp, THREADS)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("K,space", [(8, SPACE_SRGB), (16, SPACE_LAB), (256, SPACE_SRGB), (300, SPACE_SRGB)])
def test_scielab_candidate_costs_match_oracle(backend, oracle, K, space):
    img = synth.synth_image(211, 97, 5, smooth=True)
    pal = synth.synth_palettes(3, K)
    backend.setImage(img)
    backend.scielabConfigure(72, 45.0)
    of, oa = oracle.scielab_filters(72, 45.0)
    so = oracle.scielab_image(img, of, oa, 0, THREADS)
    got = backend.evalPalettesScielab(pal, space)
    want = oracle.scielab_eval(img, of, oa, so, pal, space, 0, THREADS)
    assert np.array_equal(got["err_fx"], want["err_fx"])
    assert np.array_equal(got["counts"], want["counts"])


def test_generic_and_specialised_kernels_agree(backend, oracle):
    # taps == 21 runs the specialised kernels — for candidates the FUSED tile kernel (mode 0); the generic any-tap kernels
    # (mode 1) and round 1's two-kernel 21-tap path (mode 2) must give the same integers, at sizes that cut the fused
    # kernel's 32 x 128 tiles everywhere (single partial tile, exact multiples, one column / row over)
    for (w, h, K) in ((1037, 53, 24), (64, 300, 24), (2051, 19, 24), (32, 128, 5), (33, 129, 300), (31, 257, 256), (96, 256, 1024)):
        img = synth.synth_image(w, h, 3, smooth=True)
        pal = synth.synth_palettes(2, K)
        backend.setImage(img)
        backend.scielabConfigure(72, 45.0)
        res = []
        for mode in (0, 1, 2):
            backend.scielabForceGeneric(mode)
            r = backend.evalPalettesScielab(pal)
            res.append((backend.scielabImage().view(np.uint32).copy(), r["err_fx"].copy(), r["counts"].copy()))
        backend.scielabForceGeneric(0)
        for other in res[1:]:
            assert all(np.array_equal(a, b) for a, b in zip(res[0], other)), (w, h, K)
        of, oa = oracle.scielab_filters(72, 45.0)
        so = oracle.scielab_image(img, of, oa, 0, THREADS)
        assert np.array_equal(res[0][0], so.view(np.uint32))
        want = oracle.scielab_eval(img, of, oa, so, pal, SPACE_SRGB, 0, THREADS)
        assert np.array_equal(res[0][1], want["err_fx"]) and np.array_equal(res[0][2], want["counts"]), (w, h, K)


def test_other_viewing_conditions_and_custom_filters(backend, oracle):
    img = synth.synth_image(160, 120, 9, smooth=True)
    pal = synth.synth_palettes(2, 12)
    backend.setImage(img, WHITEPOINT_D50)
    for dpi, dist in ((96, 60.0), (150, 40.0)):
        backend.scielabConfigure(dpi, dist)
        of, oa = oracle.scielab_filters(dpi, dist)
        so = oracle.scielab_image(img, of, oa, 1, THREADS)
        assert np.array_equal(backend.scielabImage().view(np.uint32), so.view(np.uint32))
        got = backend.evalPalettesScielab(pal)
        want = oracle.scielab_eval(img, of, oa, so, pal, 1, 1, THREADS)
        assert np.array_equal(got["err_fx"], want["err_fx"])
    # identity filter bank (single tap): the stage degenerates to a per-pixel colour pipeline
    ident = np.zeros((7, 1), np.float32); ident[[0, 3, 5]] = 1.0
    backend.scielabSetFilters(ident, np.zeros(1, np.float32))
    so = oracle.scielab_image(img, ident, np.zeros(1, np.float32), 1, THREADS)
    assert np.array_equal(backend.scielabImage().view(np.uint32), so.view(np.uint32))


def test_too_small_image_is_rejected(backend):
    backend.setImage(synth.synth_image(9, 40, 1))
    backend.scielabConfigure(72, 45.0)
    with pytest.raises(HqError) as e:
        backend.scielabImage()
    assert e.value.code == 4


def test_swasa_with_scielab_cost_matches_oracle(backend, oracle):
    # the reference configuration: assignment by sRGB distance, S-CIELAB cost, reference defaults (P=4)
    img = synth.synth_image(128, 96, 11, smooth=True)
    backend.setImage(img)
    backend.scielabConfigure(72, 45.0)
    backend.convergence = True
    sw = SWASA(population=4, imax=80, seed=555, space=SPACE_SRGB, costModel=COST_SCIELAB)
    best, err, tr, its = backend.findBestQuantization(12, sw, trace=True)
    p = oracle.swasa_params(population=4, imax=80, seed=555, space=oracle.SPACE_SRGB, cost_model=1)
    obest, oerr, otr = oracle.find_best_quantization(img, 12, p, trace=True, threads=THREADS)
    assert its == 80 and np.array_equal(tr.view(np.uint64), otr.view(np.uint64))
    assert err == oerr and np.array_equal(bits(best), bits(obest))


def test_plugin_entry_with_reference_scoring(oracle):
    img = synth.synth_image(80, 64, 21, smooth=True)
    hq = HybridQuantization(nbOfColors=8, populationSize=3, imax=40, seed=9, space=SPACE_SRGB, costModel=COST_SCIELAB, dpi=96, ViewingDistance=60.0)
    res = hq.quantization(img)
    p = oracle.swasa_params(population=3, imax=40, seed=9, space=oracle.SPACE_SRGB, cost_model=1, dpi=96, viewing_distance=60.0)
    obest, oerr, _ = oracle.find_best_quantization(img, 8, p, threads=THREADS)
    assert res["bestError"] == oerr and np.array_equal(bits(res["bestColors"]), bits(obest))
    assert np.array_equal(res["image"].reshape(-1, 3), oracle.quantize(img, obest, oracle.SPACE_SRGB)["rgb"])


def test_error_image_mode_matches_oracle(backend, oracle):
    # HybridQuantization.errorImage (:139-182): original vs an actually quantised image
    img = synth.synth_image(150, 90, 4, smooth=True)
    pal = synth.synth_palettes(1, 10)[0]
    quant = oracle.quantize(img, pal, oracle.SPACE_SRGB)["rgb"].reshape(img.shape)
    backend.setImage(img)
    backend.scielabConfigure(72, 45.0)
    got = backend.computeError(quant)
    of, oa = oracle.scielab_filters(72, 45.0)
    want = oracle.error_image(img, quant, of, oa, 0, THREADS)
    assert got["deltaE"] == want["deltaE"]
    assert np.array_equal(got["errorImage"].reshape(-1).view(np.uint32), want["errorImage"].view(np.uint32))
    assert np.array_equal(got["errorImageU8"].reshape(-1), want["errorImageU8"])
    same = backend.computeError(img)
    assert same["deltaE"] == 0.0 and (same["errorImage"] == 1.0).all()
    res = HybridQuantization(dpi=96, ViewingDistance=60.0).errorImage(img, quant)
    of, oa = oracle.scielab_filters(96, 60.0)
    assert res["deltaE"] == oracle.error_image(img, quant, of, oa, 0, THREADS)["deltaE"]
    with pytest.raises(ValueError):
        HybridQuantization().errorImage(img, quant[:-1])


def test_golden_scielab_vectors_on_gpu(backend):
    g = load_golden("scielab_vectors.json")
    img = synth.synth_image(g["w"], g["h"], g["seed"], g["smooth"])
    backend.setImage(img)
    backend.scielabConfigure(g["dpi"], g["viewing_distance"])
    f, a = backend.scielabFilters()
    assert np.array_equal(bits(f).ravel(), bits(from_bits(g["filters"]))) and np.array_equal(bits(a), bits(from_bits(g["abs3"])))
    assert np.array_equal(bits(backend.scielabImage()).ravel(), bits(from_bits(g["scielab_image"])))
    pal = synth.synth_palettes(g["B"], g["K"])
    ev = backend.evalPalettesScielab(pal, g["space"])
    assert [int(v) for v in ev["err_fx"]] == g["err_fx"] and ev["counts"].tolist() == g["counts"]
    q = backend.quantize(pal[0], g["space"])["rgb"]
    ei = backend.computeError(q)
    assert float(ei["deltaE"]).hex() == g["error_image_mean"]
    assert int(ei["errorImageU8"].astype(np.int64).sum()) == g["error_image_u8_sum"]


@pytest.mark.parametrize("h,world", [(97, 3), (64, 2), (45, 4)])
def test_row_shards_with_halo_add_up(backend, oracle, h, world):
    # what each rank of a multi-GPU run computes, emulated one shard after the other on one GPU:
    # own rows + 10 halo rows; errors / counts of the shards must add up to the whole image exactly
    from hybridquantization_b200.dist import row_shard_with_halo

    w, K = 131, 20
    img = synth.synth_image(w, h, 17, smooth=True)
    pal = synth.synth_palettes(3, K)
    backend.setImage(img)
    backend.scielabConfigure(72, 45.0)
    whole_sc = backend.evalPalettesScielab(pal)
    whole_id = backend.evalPalettes(pal, SPACE_SRGB, sums=True)
    whole_img = backend.scielabImage()
    whole_q = backend.quantize(pal[0], SPACE_SRGB)
    quant = whole_q["rgb"].reshape(img.shape)
    whole_err = backend.computeError(quant)
    err = np.zeros(3, np.int64); cnt = np.zeros((3, K), np.uint64)
    ierr = np.zeros(3, np.int64); icnt = np.zeros((3, K), np.uint64); isum = np.zeros((3, K, 3), np.int64)
    sc_rows, q_rows, e_rows, e_sum = [], [], [], 0.0
    for rank in range(world):
        r0, r1, top, bot = row_shard_with_halo(h, world, rank, 10)
        backend.setImageSharded(img[r0 - top:r1 + bot], top, bot, r0, h)
        assert backend.pixels() == (r1 - r0) * w
        r = backend.evalPalettesScielab(pal)
        err += r["err_fx"]; cnt += r["counts"]
        r = backend.evalPalettes(pal, SPACE_SRGB, sums=True)
        ierr += r["err_fx"]; icnt += r["counts"]; isum += r["sums_fx"]
        sc_rows.append(backend.scielabImage().reshape(3, r1 - r0, w))
        q_rows.append(backend.quantize(pal[0], SPACE_SRGB)["rgb"])
        e = backend.computeError(quant[r0 - top:r1 + bot])
        e_rows.append(e["errorImage"]); e_sum += e["deltaE"] * (r1 - r0) * w
    assert np.array_equal(err, whole_sc["err_fx"]) and np.array_equal(cnt, whole_sc["counts"])
    assert np.array_equal(ierr, whole_id["err_fx"]) and np.array_equal(icnt, whole_id["counts"]) and np.array_equal(isum, whole_id["sums_fx"])
    assert np.array_equal(np.concatenate(sc_rows, axis=1).reshape(3, -1).view(np.uint32), whole_img.view(np.uint32))
    assert np.array_equal(np.concatenate(q_rows, axis=0), whole_q["rgb"])
    assert np.array_equal(np.concatenate(e_rows, axis=0).view(np.uint32), whole_err["errorImage"].view(np.uint32))
    assert abs(e_sum / (h * w) - whole_err["deltaE"]) < 1e-9
    # a shard without the halo rows it needs is rejected for the S-CIELAB stage (but fine for the identity cost)
    r0, r1, _, _ = row_shard_with_halo(h, world, 1, 10)
    backend.setImageSharded(img[r0:r1], 0, 0, r0, h)
    with pytest.raises(HqError) as ex:
        backend.evalPalettesScielab(pal)
    assert ex.value.code == 4
    backend.evalPalettes(pal)
