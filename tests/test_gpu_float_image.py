"""The image as the plugin holds it — planar floats in [0,1] (`im.getDataXYCAsFloat()` after the rescaling conversion,
HybridQuantization.java:95-98) — through hq_set_image_f32_planar: 16-bit and float Icy images, bit for bit against the
oracle, and the same bits as the u8 entry for a u8-derived image."""
import os

import numpy as np
import pytest

from hybridquantization_b200 import (COST_SCIELAB, EVAL_PRUNE, PRUNE_AUTO, PRUNE_ON, SPACE_LAB, SPACE_SRGB, SWASA, WHITEPOINT_D50, HqError,
                                     ImageManipulation, synth)

pytestmark = pytest.mark.gpu
THREADS = max(1, len(os.sched_getaffinity(0)))


def u16_planes(w, h, seed):
    """a 16-bit RGB image after Icy's rescale: (float)(c / 65535.0)"""
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 65536, (3, h, w), dtype=np.uint32)
    c[:, 0, : min(w, 8)] = np.array([0, 1, 2, 2650, 2651, 2652, 65534, 65535], np.uint32)[: min(w, 8)]  # both sides of 0.04045, the ends
    return (c.astype(np.float64) / 65535.0).astype(np.float32)


def float_planes(w, h, seed):
    """arbitrary floats in [0,1]: uniform, denormal-small, the decode threshold and its neighbours, exact 0 and 1"""
    rng = np.random.default_rng(seed)
    p = rng.random((3, h, w), dtype=np.float32)
    t = np.float32(0.04045)
    special = np.array([0.0, 1.0, t, np.nextafter(t, np.float32(0)), np.nextafter(t, np.float32(1)), 1e-30, 1e-42, np.nextafter(np.float32(1), np.float32(0))],
                       np.float32)
    flat = p.reshape(-1)
    flat[: min(flat.size, special.size)] = special[: min(flat.size, special.size)]
    return p


def _same(a, b, sums=True):
    assert np.array_equal(a["err_fx"], b["err_fx"]) and np.array_equal(a["counts"], b["counts"])
    if sums:
        assert np.array_equal(a["sums_fx"], b["sums_fx"])


@pytest.mark.parametrize("maker", [u16_planes, float_planes])
@pytest.mark.parametrize("w,h,wp", [(257, 63, 0), (64, 33, WHITEPOINT_D50), (1, 1, 0), (1025, 5, 0)])
def test_lab_planes_match_oracle(backend, oracle, maker, w, h, wp):
    planes = maker(w, h, 3 * w + h)
    backend.setImageFloat(planes, wp)
    assert backend.pixels() == w * h
    _, want = oracle.image_planes_f32(planes, wp, THREADS)
    assert np.array_equal(backend.labImage().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("space", [SPACE_LAB, SPACE_SRGB])
@pytest.mark.parametrize("K,B", [(7, 3), (64, 2), (300, 2)])
def test_scoring_matches_oracle(backend, oracle, space, K, B):
    planes = u16_planes(320, 211, K)
    pal = synth.synth_palettes(B, K, seed=K)
    backend.setImageFloat(planes)
    unit, lab = oracle.image_planes_f32(planes, 0, THREADS)
    want = oracle.assign_reduce_planes(unit, lab, pal, space, 0, want_idx=True, threads=THREADS)
    _same(backend.evalPalettes(pal, space, sums=True), want)
    if space == SPACE_LAB:
        _same(backend.evalPalettes(pal, space, sums=True, flags=EVAL_PRUNE), want)
    q = backend.quantize(pal[0], space)
    assert np.array_equal(q["idx"], want["idx"][0])
    backend.setPruning(PRUNE_ON)
    try:
        assert np.array_equal(backend.quantize(pal[0], space)["idx"], want["idx"][0])
    finally:
        backend.setPruning(PRUNE_AUTO)


def test_u8_derived_float_image_equals_the_u8_entry(backend):
    img = synth.synth_image(301, 97, 21, smooth=True)
    planes = np.ascontiguousarray((img.astype(np.float64) / 255.0).astype(np.float32).transpose(2, 0, 1))
    pal = synth.synth_palettes(3, 48)
    backend.setImage(img)
    lab8 = backend.labImage().view(np.uint32).copy()
    a = {s: backend.evalPalettes(pal, s, sums=True) for s in (SPACE_LAB, SPACE_SRGB)}
    backend.scielabConfigure(72, 45.0)
    sc8 = backend.scielabImage().view(np.uint32).copy()
    e8 = backend.evalPalettesScielab(pal, SPACE_SRGB)
    backend.setImageFloat(planes)
    assert np.array_equal(backend.labImage().view(np.uint32), lab8)
    for s in (SPACE_LAB, SPACE_SRGB):
        _same(backend.evalPalettes(pal, s, sums=True), a[s])
    assert np.array_equal(backend.scielabImage().view(np.uint32), sc8)
    _same(backend.evalPalettesScielab(pal, SPACE_SRGB), e8, sums=False)
    # and back: the u8 entry after a float image must not see stale planes
    backend.setImage(img)
    _same(backend.evalPalettes(pal, SPACE_SRGB, sums=True), a[SPACE_SRGB])


@pytest.mark.parametrize("space", [SPACE_SRGB, SPACE_LAB])
def test_scielab_stage_on_a_float_image(backend, oracle, space):
    planes = u16_planes(211, 97, 8)
    pal = synth.synth_palettes(3, 40)
    backend.setImageFloat(planes)
    backend.scielabConfigure(72, 45.0)
    of, oa = oracle.scielab_filters(72, 45.0)
    so = oracle.scielab_image_f32(planes, of, oa, 0, THREADS)
    assert np.array_equal(backend.scielabImage().view(np.uint32), so.view(np.uint32))
    want = oracle.scielab_eval_f32(planes, of, oa, so, pal, space, 0, THREADS)
    for mode in (PRUNE_AUTO, PRUNE_ON):
        backend.setPruning(mode)
        try:
            got = backend.evalPalettesScielab(pal, space)
        finally:
            backend.setPruning(PRUNE_AUTO)
        _same(got, want, sums=False)


def test_row_shards_with_halos(backend, oracle):
    planes = u16_planes(96, 80, 4)
    pal = synth.synth_palettes(2, 16)
    of, oa = oracle.scielab_filters(72, 45.0)
    so = oracle.scielab_image_f32(planes, of, oa, 0, THREADS)
    want = oracle.scielab_eval_f32(planes, of, oa, so, pal, SPACE_SRGB, 0, THREADS)
    err = np.zeros(2, np.int64); cnt = np.zeros((2, 16), np.uint64)
    for (lo, hi) in ((0, 40), (40, 80)):
        top, bot = min(10, lo), min(10, 80 - hi)
        backend.setImageFloat(planes[:, lo - top:hi + bot], 0, top, bot, lo, 80)
        backend.scielabConfigure(72, 45.0)
        assert np.array_equal(backend.scielabImage().view(np.uint32), so.reshape(3, 80, 96)[:, lo:hi].reshape(3, -1).view(np.uint32))
        r = backend.evalPalettesScielab(pal, SPACE_SRGB)
        err += r["err_fx"]; cnt += r["counts"]
    assert np.array_equal(err, want["err_fx"]) and np.array_equal(cnt, want["counts"])


def test_search_on_a_float_image_equals_the_u8_search(backend):
    img = synth.synth_image(200, 120, 6, smooth=True)
    planes = np.ascontiguousarray((img.astype(np.float64) / 255.0).astype(np.float32).transpose(2, 0, 1))
    backend.setImage(img)
    a = backend.findBestQuantization(12, SWASA(population=4, imax=40, seed=5))
    backend.setImageFloat(planes)
    b = backend.findBestQuantization(12, SWASA(population=4, imax=40, seed=5))
    assert a[1] == b[1] and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))


def test_plugin_entry_points_take_float_planes(hqlib):
    from hybridquantization_b200 import HybridQuantization
    img = synth.synth_image(96, 64, 5)
    planes = np.ascontiguousarray((img.astype(np.float64) / 255.0).astype(np.float32).transpose(2, 0, 1))
    hq = HybridQuantization(nbOfColors=8, imax=60, seed=4242)
    a = dict(hq.quantization(img)); b = dict(hq.quantization(planes))
    assert a["bestError"] == b["bestError"] and np.array_equal(a["image"], b["image"]) and np.array_equal(a["bestColors"].view(np.uint32), b["bestColors"].view(np.uint32))
    qp = np.ascontiguousarray((a["image"].astype(np.float64) / 255.0).astype(np.float32).transpose(2, 0, 1))
    assert hq.errorImage(img, a["image"])["deltaE"] == hq.errorImage(planes, qp)["deltaE"]


def test_error_image_mode_on_float_images(backend, oracle):
    a = u16_planes(120, 64, 1)
    b = np.clip(a + np.random.default_rng(2).normal(0, 0.02, a.shape).astype(np.float32), 0, 1).astype(np.float32)
    backend.setImageFloat(a)
    backend.scielabConfigure(72, 45.0)
    got = backend.computeErrorFloat(b)
    of, oa = oracle.scielab_filters(72, 45.0)
    want = oracle.error_image_f32(a, b, of, oa, 0, THREADS)
    assert got["deltaE"] == want["deltaE"] and got["deltaE"] > 0
    assert np.array_equal(got["errorImage"].reshape(-1).view(np.uint32), want["errorImage"].view(np.uint32))
    assert np.array_equal(got["errorImageU8"].reshape(-1), want["errorImageU8"])
    # a u8 second image through either entry: same result
    img = synth.synth_image(120, 64, 3)
    p8 = np.ascontiguousarray((img.astype(np.float64) / 255.0).astype(np.float32).transpose(2, 0, 1))
    r8, rf = backend.computeError(img), backend.computeErrorFloat(p8)
    assert r8["deltaE"] == rf["deltaE"] and np.array_equal(r8["errorImageU8"], rf["errorImageU8"])
    bad = b.copy(); bad[2, 5, 5] = 1.5
    with pytest.raises(HqError):
        backend.computeErrorFloat(bad)
    assert backend.computeErrorFloat(b)["deltaE"] == want["deltaE"]   # the context survives the refusal


@pytest.mark.parametrize("bad", [1.0000001, -1e-9, float("nan"), float("inf")])
def test_values_outside_the_unit_interval_are_refused(hqlib, bad):
    be = ImageManipulation("CIE76", False, True, 0)
    try:
        planes = np.full((3, 9, 11), 0.5, np.float32)
        planes[1, 4, 5] = bad
        with pytest.raises(HqError) as e:
            be.setImageFloat(planes)
        assert "[0,1]" in str(e.value)
        with pytest.raises(HqError):   # no image is resident after the refusal
            be.evalPalettes(synth.synth_palettes(1, 4), SPACE_LAB)
    finally:
        be.close()
