"""Parity of the CUDA path (through the C ABI) against the CPU oracle: bit-exact.

Indices, counts, fixed-point error / Lab sums and Lab values must be IDENTICAL to the oracle's;
there is no tolerance anywhere in this file.  Needs a B200: run with `-m gpu`."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import bits, fbits, from_bits, idx_crc, load_golden
from hybridquantization_b200 import (EVAL_FORCE_CHUNKED, EVAL_FORCE_DIRECT, EVAL_FORCE_PREFILTER, SPACE_LAB, SPACE_SRGB, WHITEPOINT_D50,
                                     WHITEPOINT_D65, synth)

pytestmark = pytest.mark.gpu
THREADS = max(1, len(os.sched_getaffinity(0)))


# ------------------------------------------------------------------ arithmetic on the device
@pytest.mark.parametrize("which,lo,hi", [(0, 0.008856452070, 1.25), (1, 0.0625, 1.0), (2, 0.03, 1.0), (3, 0.0, 2.0), (4, 0.0, 2.0), (5, 0.0, 0.01),
                                          (6, 0.0, 2.0), (7, 0.0, 2.0)])
def test_device_math_exhaustive(backend, hqlib, oracle, which, lo, hi):
    a, b = fbits(lo), fbits(hi)
    chunk = 1 << 24
    bad = 0
    for s in range(a, b + 1, chunk):
        c = min(chunk, b + 1 - s)
        x = np.empty(c, np.float32)
        assert hqlib.hq_device_math_range(backend._ctx, which, s, c, x.ctypes.data_as(C.c_void_p)) == 0
        y = oracle.math_range(which, s, c, THREADS)
        bad += int(np.count_nonzero(x.view(np.uint32) != y.view(np.uint32)))
    assert bad == 0


def test_rgb_to_lab_all_16m_colours(backend, oracle):
    # every u8 RGB triple exactly once: 4096 x 4096 pixels
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=1).astype(np.uint8).reshape(4096, 4096, 3)
    for wp in (WHITEPOINT_D65, WHITEPOINT_D50):
        backend.setImage(img, wp)
        got = backend.labImage()
        _, want = oracle.image_planes(img, wp, THREADS)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("w,h", [(1, 1), (3, 1), (5, 7), (1023, 1), (1025, 3), (4097, 2), (640, 480)])
def test_rgb_to_lab_ragged_sizes(backend, oracle, w, h):
    img = synth.synth_image(w, h, synth.SEED_BASE + 9)
    backend.setImage(img)
    _, want = oracle.image_planes(img)
    assert np.array_equal(backend.labImage().view(np.uint32), want.view(np.uint32))


def test_golden_lab_vectors_on_gpu(backend):
    g = load_golden("lab_vectors.json")
    u8 = np.array(g["u8"], np.uint8).reshape(1, -1, 3)
    backend.setImage(u8, WHITEPOINT_D65)
    assert np.array_equal(bits(backend.labImage().T.copy()).ravel(), bits(from_bits(g["lab_d65"])))
    backend.setImage(u8, WHITEPOINT_D50)
    assert np.array_equal(bits(backend.labImage().T.copy()).ravel(), bits(from_bits(g["lab_d50"])))


# ------------------------------------------------------------------ assign + reduce
def _check(backend, oracle, img, pal, space, flags=0, wp=WHITEPOINT_D65):
    backend.setImage(img, wp)
    got = backend.evalPalettes(pal, space, sums=True, flags=flags)
    want = oracle.assign_reduce(img, pal, space, wp, threads=THREADS)
    assert np.array_equal(got["err_fx"], want["err_fx"])
    assert np.array_equal(got["counts"], want["counts"])
    assert np.array_equal(got["sums_fx"], want["sums_fx"])
    light = backend.evalPalettes(pal, space, sums=False, flags=flags)  # the SA scoring variant
    assert np.array_equal(light["err_fx"], want["err_fx"]) and np.array_equal(light["counts"], want["counts"])


@pytest.mark.parametrize("K", [1, 2, 7, 8, 9, 16, 17, 64, 255, 256, 257, 1024])
@pytest.mark.parametrize("space", [SPACE_LAB, SPACE_SRGB])
def test_assign_reduce_palette_sizes(backend, oracle, K, space):
    img = synth.synth_image(211, 97, synth.SEED_BASE + 1, smooth=(K % 2 == 0))
    _check(backend, oracle, img, synth.synth_palettes(3, K), space)


@pytest.mark.parametrize("flags", [EVAL_FORCE_DIRECT, EVAL_FORCE_CHUNKED, EVAL_FORCE_PREFILTER])
@pytest.mark.parametrize("K", [1, 5, 16, 40, 256])
def test_both_kernel_variants(backend, oracle, flags, K):
    img = synth.synth_image(300, 71, synth.SEED_BASE + 2, smooth=True)
    _check(backend, oracle, img, synth.synth_palettes(2, K, seed=3), SPACE_LAB, flags)
    _check(backend, oracle, img, synth.synth_palettes(2, K, seed=4), SPACE_SRGB, flags)


@pytest.mark.parametrize("w,h", [(1, 1), (2, 1), (3, 1), (4, 1), (5, 1), (1023, 1), (1024, 1), (1025, 1), (33, 31), (2049, 3)])
def test_ragged_image_sizes(backend, oracle, w, h):
    img = synth.synth_image(w, h, synth.SEED_BASE + 3)
    _check(backend, oracle, img, synth.synth_palettes(2, 19), SPACE_LAB)
    _check(backend, oracle, img, synth.synth_palettes(2, 19), SPACE_SRGB, EVAL_FORCE_CHUNKED)


def test_batch_of_64_candidates(backend, oracle):
    img = synth.synth_image(256, 128, synth.SEED_BASE + 3, smooth=True)
    _check(backend, oracle, img, synth.synth_palettes(64, 32), SPACE_LAB)


def test_d50_white_point(backend, oracle):
    img = synth.synth_image(97, 53, 77)
    _check(backend, oracle, img, synth.synth_palettes(2, 24), SPACE_LAB, wp=WHITEPOINT_D50)


def test_tie_rule_lowest_index_wins(backend, oracle):
    img = synth.synth_image(128, 64, 7)
    for K, flags in ((6, EVAL_FORCE_DIRECT), (6, EVAL_FORCE_CHUNKED), (40, EVAL_FORCE_CHUNKED), (6, EVAL_FORCE_PREFILTER), (40, EVAL_FORCE_PREFILTER)):
        pal = synth.synth_palettes(1, K)
        pal[0, K - 2] = pal[0, 1]   # duplicates later in the palette must never be chosen
        pal[0, K - 1] = pal[0, 0]
        if K > 20:
            pal[0, 17] = pal[0, 9]  # a duplicate inside another chunk
        backend.setImage(img)
        got = backend.evalPalettes(pal, SPACE_LAB, flags=flags)
        assert got["counts"][0, K - 1] == 0 and got["counts"][0, K - 2] == 0
        want = oracle.assign_reduce(img, pal)
        assert np.array_equal(got["counts"], want["counts"])
        q = backend.quantize(pal[0])
        assert np.array_equal(q["idx"], oracle.quantize(img, pal[0])["idx"])


def test_prefilter_worklist_overflow_and_near_ties(backend, oracle):
    # every pixel identical AND the winning colour duplicated in another chunk: every pixel is
    # ambiguous for the prefilter, the per-CTA worklist (1024) overflows and the in-place exact
    # sweep must give the same integers
    img = np.full((257, 131, 3), 201, np.uint8)
    unit, _ = oracle.image_planes(img[:1, :1])
    pal = synth.synth_palettes(2, 40)
    pal[:, 3, :3] = unit[:, 0]
    pal[:, 29, :3] = unit[:, 0]           # exact duplicate of the winner, chunk 3
    pal[0, 17, :3] = unit[:, 0] + np.float32(1e-6)   # near tie, chunk 2
    _check(backend, oracle, img, pal, SPACE_LAB, EVAL_FORCE_PREFILTER)
    _check(backend, oracle, img, pal, SPACE_SRGB, EVAL_FORCE_PREFILTER)
    # clustered palettes: colours a few ulps apart scattered over different chunks
    rng = np.random.default_rng(3)
    img = synth.synth_image(400, 300, 123, smooth=True)
    pal = synth.synth_palettes(3, 64)
    for b in range(3):
        base = pal[b, :8].copy()
        for k in range(8, 64):
            pal[b, k, :3] = base[k % 8, :3] + (rng.integers(-3, 4, 3) * np.float32(6e-8)).astype(np.float32)
    pal = np.clip(pal, 0, 1).astype(np.float32)
    _check(backend, oracle, img, pal, SPACE_LAB, EVAL_FORCE_PREFILTER)
    backend.setImage(img)
    for b in range(3):
        assert np.array_equal(backend.quantize(pal[b])["idx"], oracle.quantize(img, pal[b])["idx"])


def test_flat_image_worst_case_contention(backend, oracle):
    # every pixel identical: all lanes hit the same colour bin
    img = np.full((300, 200, 3), 77, np.uint8)
    _check(backend, oracle, img, synth.synth_palettes(2, 16), SPACE_LAB)
    _check(backend, oracle, img, synth.synth_palettes(2, 64), SPACE_LAB)


def test_golden_assign_vectors_on_gpu(backend):
    g = load_golden("assign_vectors.json")
    for name, v in g.items():
        img = synth.synth_image(v["w"], v["h"], v["seed"], v["smooth"])
        pal = synth.synth_palettes(v["B"], v["K"])
        backend.setImage(img)
        got = backend.evalPalettes(pal, v["space"], sums=True)
        assert [int(x) for x in got["err_fx"]] == v["err_fx"], name
        assert got["counts"].tolist() == v["counts"], name
        assert got["sums_fx"].tolist() == v["sums_fx"], name
        for b in range(v["B"]):
            q = backend.quantize(pal[b], v["space"])
            assert idx_crc(q["idx"], v["w"] * v["h"]) == v["idx_crc"][b], name


# ------------------------------------------------------------------ final image
@pytest.mark.parametrize("K,space", [(8, SPACE_LAB), (256, SPACE_LAB), (300, SPACE_LAB), (16, SPACE_SRGB)])
def test_quantize_matches_oracle(backend, oracle, K, space):
    img = synth.synth_image(123, 45, synth.SEED_BASE + 4, smooth=True)
    pal = synth.synth_palettes(1, K)[0]
    backend.setImage(img)
    got = backend.quantize(pal, space, want_f32=True)
    want = oracle.quantize(img, pal, space)
    assert np.array_equal(got["idx"], want["idx"])
    assert np.array_equal(got["rgb"].reshape(-1, 3), want["rgb"])
    assert np.array_equal(got["f32"].view(np.uint32), want["f32"].view(np.uint32))


# ------------------------------------------------------------------ error behaviour
def test_errors_are_loud(backend, hqlib):
    from hybridquantization_b200 import HqError, ImageManipulation

    fresh = ImageManipulation("CIE76", False, True, 0)
    with pytest.raises(HqError) as e:
        fresh.evalPalettes(synth.synth_palettes(1, 4))
    assert e.value.code == 3  # HQ_ERR_NO_IMAGE
    fresh.setImage(synth.synth_image(8, 8, 1))
    few = synth.synth_palettes(1, 4)   # K beyond the plugin's own range (HybridQuantization.java:192) is refused before anything is read
    assert fresh._lib.hq_eval_palettes(fresh._ctx, few.ctypes.data, 1, (1 << 24) + 1, 0, 0, None, None, None) == 4  # HQ_ERR_UNSUPPORTED
    bad = synth.synth_palettes(1, 4)
    bad[0, 2, 1] = np.nan
    with pytest.raises(HqError) as e:
        fresh.evalPalettes(bad)                                     # NaN / out-of-range palette colours are refused (SWASA.java:93-106 clamps them)
    assert e.value.code == 1
    bad[0, 2, 1] = 1.5
    with pytest.raises(HqError) as e:
        fresh.quantize(bad[0])
    assert e.value.code == 1
    with pytest.raises(ValueError):
        fresh.setImage(np.zeros((4, 4), np.uint8))  # fewer than 3 channels (HybridQuantization.java:68)
    fresh.close()
    with pytest.raises(HqError):
        ImageManipulation("CIE76", False, True, 99)
