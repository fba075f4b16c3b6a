"""Exact pruning (HQ_EVAL_PRUNE, csrc/hq_pruned.cu) against the oracle and the exhaustive kernel: the integers must be
IDENTICAL — the pruned path only skips colours that provably cannot be nearest (or tie) for any pixel of a chunk."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from hybridquantization_b200 import COST_SCIELAB, EVAL_PRUNE, PRUNE_AUTO, PRUNE_OFF, PRUNE_ON, SPACE_LAB, SPACE_SRGB, SWASA, WHITEPOINT_D50, synth

pytestmark = pytest.mark.gpu
THREADS = max(1, len(os.sched_getaffinity(0)))


def _same(a, b, sums=True):
    assert np.array_equal(a["err_fx"], b["err_fx"])
    assert np.array_equal(a["counts"], b["counts"])
    if sums:
        assert np.array_equal(a["sums_fx"], b["sums_fx"])


@pytest.mark.parametrize("w,h,K,B,smooth,wp", [(640, 480, 256, 4, False, 0), (640, 480, 256, 3, True, 0), (333, 97, 37, 5, False, 1), (1024, 512, 1024, 2, False, 0),
                                               (2048, 1024, 512, 2, True, 0), (64, 64, 8, 2, False, 0), (1, 1, 3, 1, False, 0), (5, 7, 1, 2, True, 0),
                                               (4097, 3, 300, 2, False, 0)])
def test_pruned_equals_oracle(backend, oracle, w, h, K, B, smooth, wp):
    img = synth.synth_image(w, h, 90 + K, smooth)
    pal = synth.synth_palettes(B, K, seed=K)
    backend.setImage(img, wp)
    got = backend.evalPalettes(pal, SPACE_LAB, sums=True, flags=EVAL_PRUNE)
    want = oracle.assign_reduce(img, pal, SPACE_LAB, wp, threads=THREADS)
    _same(got, want)
    _same(backend.evalPalettes(pal, SPACE_LAB, sums=False, flags=EVAL_PRUNE), want, sums=False)
    st_ = backend.pruningStats()
    assert st_["chunks"] >= max(1, (w * h + 2047) // 2048)


def test_pruning_actually_prunes(backend):
    img = synth.synth_image(1920, 1080, 5, False)
    pal = synth.synth_palettes(8, 256)
    backend.setImage(img)
    backend.setProfiling(True)
    a = backend.evalPalettes(pal, SPACE_LAB, flags=EVAL_PRUNE)
    s = backend.pruningStats()
    backend.setProfiling(False)
    b = backend.evalPalettes(pal, SPACE_LAB)
    _same(a, b, sums=False)
    assert 1.0 <= s["mean_survivors"] < 64.0, s   # of 256 colours


def test_search_policy_call(backend):
    """hq_search_eval_flags: what the built-in search, the C++ host and the JNI shim all ask before scoring a population"""
    backend.setImage(synth.synth_image(320, 256, 5))          # 81,920 px
    assert backend.searchEvalFlags(32, SPACE_LAB) == EVAL_PRUNE
    assert backend.searchEvalFlags(31, SPACE_LAB) == 0
    assert backend.searchEvalFlags(256, SPACE_SRGB) == 0        # sRGB search stays exhaustive
    assert backend.searchEvalFlags(256, SPACE_LAB, COST_SCIELAB) == 0   # the S-CIELAB chain has its own policy
    assert backend.searchEvalFlags(1500, SPACE_LAB) == EVAL_PRUNE
    backend.setPruning(PRUNE_OFF)
    assert backend.searchEvalFlags(256, SPACE_LAB) == 0
    assert backend.searchEvalFlags(1500, SPACE_LAB) == EVAL_PRUNE      # above HQ_MAX_COLORS only the pruned kernel exists
    backend.setPruning(PRUNE_ON)
    assert backend.searchEvalFlags(2, SPACE_LAB) == EVAL_PRUNE
    backend.setPruning(PRUNE_AUTO)
    backend.setImage(synth.synth_image(64, 64, 5))
    assert backend.searchEvalFlags(256, SPACE_LAB) == 0


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture], derandomize=True)
@given(w=st.integers(1, 400), h=st.integers(1, 60), K=st.integers(1, 300), B=st.integers(1, 4), wp=st.integers(0, 1), smooth=st.booleans(),
       patho=st.sampled_from(["none", "dup", "clamp", "tiny", "far"]), seed=st.integers(0, 2 ** 31))
def test_pruned_random_configurations(backend, oracle, w, h, K, B, wp, smooth, patho, seed):
    img = synth.synth_image(w, h, seed, smooth)
    pal = synth.synth_palettes(B, K, seed=seed % 100000)
    rng = np.random.default_rng(seed)
    if patho == "dup" and K > 1:       # exact ties: the lowest index must win, both copies must survive or fall together
        src = rng.integers(0, K, K // 2 + 1); dst = rng.integers(0, K, K // 2 + 1)
        pal[:, dst] = pal[:, src]
    elif patho == "clamp":
        pal[..., :3] = np.round(pal[..., :3] * 2) / 2
    elif patho == "tiny":              # colours a few ulps apart: distances differ in the last bits only
        pal[..., :3] = pal[:, :1, :3] + (rng.integers(-4, 5, pal[..., :3].shape) * np.float32(6e-8)).astype(np.float32)
        pal = np.clip(pal, 0, 1).astype(np.float32)
    elif patho == "far":               # every colour in one corner: all pixels far away, U is large
        pal[..., :3] *= np.float32(0.02)
    backend.setImage(img, wp)
    got = backend.evalPalettes(pal, SPACE_LAB, sums=True, flags=EVAL_PRUNE)
    want = oracle.assign_reduce(img, pal, SPACE_LAB, wp, threads=THREADS)
    _same(got, want)


def test_row_shard_own_range(backend, oracle):
    """halo rows of a shard are excluded from the sorted copy exactly as they are from the exhaustive reductions"""
    img = synth.synth_image(320, 90, 8, True)
    pal = synth.synth_palettes(3, 64)
    backend.setImageSharded(img[10:70], 10, 12, 20, 90)   # own rows 20..57, halos 10 above / 12 below
    got = backend.evalPalettes(pal, SPACE_LAB, sums=True, flags=EVAL_PRUNE)
    want = oracle.assign_reduce(img[20:58], pal, SPACE_LAB, 0, threads=THREADS)
    _same(got, want)


def test_search_trajectory_unchanged_by_pruning(backend, oracle):
    img = synth.synth_image(320, 240, 4, True)
    backend.setImage(img)
    p = oracle.swasa_params(population=4, imax=60, iTc=5, seed=123)
    obest, oerr, otr = oracle.find_best_quantization(img, 48, p, trace=True, threads=THREADS)
    res = {}
    for mode in (PRUNE_OFF, PRUNE_ON):
        backend.setPruning(mode)
        best, err, tr, its = backend.findBestQuantization(48, SWASA(population=4, imax=60, iTc=5, seed=123), trace=True)
        res[mode] = (best.copy(), err, tr.copy())
        assert its == 60 and err == oerr and np.array_equal(tr.view(np.uint64), otr.view(np.uint64))
        assert np.array_equal(best.view(np.uint32), obest.view(np.uint32))
    backend.setPruning(1)


# ---------------------------------------------------------------- palettes beyond HQ_MAX_COLORS (the plugin allows up to 2^24 colours)
@pytest.mark.parametrize("K", [1025, 1500, 4096])
def test_large_palettes_are_scored_by_the_pruned_kernel(backend, oracle, K):
    img = synth.synth_image(400, 300, K, K % 2 == 0)
    pal = synth.synth_palettes(2, K, seed=K)
    backend.setImage(img)
    got = backend.evalPalettes(pal, SPACE_LAB, sums=True)           # no flag: K > 1024 selects the pruned kernel by itself
    want = oracle.assign_reduce(img, pal, SPACE_LAB, threads=THREADS)
    _same(got, want)


@pytest.mark.parametrize("K,space", [(2000, SPACE_LAB), (1100, SPACE_SRGB), (300, SPACE_SRGB)])
def test_quantize_with_large_palettes(backend, oracle, K, space):
    img = synth.synth_image(320, 200, 3 + K, True)
    pal = synth.synth_palettes(1, K, seed=K)[0]
    backend.setImage(img)
    got = backend.quantize(pal, space, want_f32=True)
    want = oracle.quantize(img, pal, space, threads=THREADS)
    assert np.array_equal(got["idx"], want["idx"]) and np.array_equal(got["rgb"].reshape(-1, 3), want["rgb"])
    assert np.array_equal(got["f32"].view(np.uint32), want["f32"].view(np.uint32))


# ---------------------------------------------------------------- palettes beyond every staged kernel: the chunked sweep (hq_bigk.cu)
@pytest.mark.parametrize("K,space", [(4097, SPACE_LAB), (6000, SPACE_LAB), (1500, SPACE_SRGB), (9000, SPACE_SRGB)])
def test_palettes_beyond_the_staged_kernels(backend, oracle, K, space):
    """K > 4,096 (or K > 1,024 with sRGB-space scoring, which the pruned kernel cannot do): error, counts and Lab sums equal the oracle's"""
    img = synth.synth_image(160, 120, K, K % 2 == 0)
    pal = synth.synth_palettes(2, K, seed=K)
    backend.setImage(img)
    _same(backend.evalPalettes(pal, space, sums=True), oracle.assign_reduce(img, pal, space, threads=THREADS))


def test_output_image_with_70000_colours(backend, oracle):
    """beyond 16-bit index images: the output image comes from the sweep's own 32-bit assignment; out_idx is refused"""
    from hybridquantization_b200 import HqError
    K = 70000
    img = synth.synth_image(96, 64, 21, True)
    pal = synth.synth_palettes(1, K, seed=3)[0]
    backend.setImage(img)
    with pytest.raises(HqError) as ex:
        backend.quantize(pal, SPACE_LAB)                # asks for the 16-bit index image
    assert ex.value.code == 4
    n = 96 * 64
    import ctypes as C
    rgb = np.empty((n, 3), np.uint8); f32 = np.empty((n, 4), np.float32)
    rc = backend._lib.hq_quantize(backend._ctx, pal.ctypes.data_as(C.c_void_p), K, SPACE_LAB, rgb.ctypes.data_as(C.c_void_p), f32.ctypes.data_as(C.c_void_p), None)
    assert rc == 0
    # the oracle's nearest colour per pixel, by its own sweep over all 70,000 (first wins)
    r = oracle.assign_reduce(img, pal[None], SPACE_LAB, threads=THREADS)
    got = backend.evalPalettes(pal[None], SPACE_LAB)
    assert np.array_equal(got["err_fx"], r["err_fx"]) and np.array_equal(got["counts"], r["counts"])
    used = np.flatnonzero(r["counts"][0])
    assert set(map(bytes, f32.view(np.uint8).reshape(n, 16))) == set(map(bytes, pal[used].view(np.uint8).reshape(-1, 16)))
    assert np.array_equal(rgb, (f32[:, :3] * np.float32(255.0) + np.float32(0.5)).astype(np.int32).astype(np.uint8))


# ---------------------------------------------------------------- index-producing pruning inside the S-CIELAB chain
@pytest.mark.parametrize("w,h,K,space", [(640, 360, 256, SPACE_SRGB), (512, 300, 64, SPACE_LAB), (300, 260, 1500, SPACE_SRGB), (120, 90, 5000, SPACE_SRGB)])
def test_scielab_chain_with_pruned_assignment(backend, oracle, w, h, K, space):
    """hq_eval_palettes_scielab assigns with the pruned kernel (indices scattered through the sort permutation) when it
    pays or K > 1024: errors and counts must equal the oracle's and the exhaustive assignment's"""
    img = synth.synth_image(w, h, 11 + K, True)
    pal = synth.synth_palettes(3, K, seed=K)
    backend.setImage(img)
    backend.scielabConfigure(72, 45.0)
    f, a = oracle.scielab_filters(72, 45.0)
    so = oracle.scielab_image(img, f, a, 0, THREADS)
    want = oracle.scielab_eval(img, f, a, so, pal, space, 0, THREADS)
    backend.setPruning(PRUNE_AUTO)
    got = backend.evalPalettesScielab(pal, space)
    assert np.array_equal(got["err_fx"], want["err_fx"]) and np.array_equal(got["counts"], want["counts"])
    if K <= 1024:
        backend.setPruning(PRUNE_OFF)
        ex = backend.evalPalettesScielab(pal, space)
        backend.setPruning(PRUNE_AUTO)
        assert np.array_equal(ex["err_fx"], got["err_fx"]) and np.array_equal(ex["counts"], got["counts"])


def test_scielab_pruned_assignment_on_row_shards(backend, oracle):
    """halo pixels are assigned (their colours feed the filter) but not counted, exactly as in the exhaustive path"""
    img = synth.synth_image(360, 300, 21, True)
    pal = synth.synth_palettes(2, 128)
    backend.scielabConfigure(72, 45.0)
    f, a = oracle.scielab_filters(72, 45.0)
    so = oracle.scielab_image(img, f, a, 0, THREADS)
    want = oracle.scielab_eval(img, f, a, so, pal, SPACE_SRGB, 0, THREADS)
    err = np.zeros(2, np.int64); cnt = np.zeros((2, 128), np.uint64)
    for (r0, r1) in ((0, 150), (150, 300)):
        top, bot = min(10, r0), min(10, 300 - r1)
        backend.setImageSharded(img[r0 - top:r1 + bot], top, bot, r0, 300)
        backend.scielabConfigure(72, 45.0)
        r = backend.evalPalettesScielab(pal, SPACE_SRGB)
        err += r["err_fx"]; cnt += r["counts"]
    assert np.array_equal(err, want["err_fx"]) and np.array_equal(cnt, want["counts"])


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture], derandomize=True)
@given(w=st.integers(10, 200), h=st.integers(10, 60), K=st.integers(1, 300), space=st.integers(0, 1), smooth=st.booleans(),
       dup=st.booleans(), seed=st.integers(0, 2 ** 31))
def test_index_producing_mode_random_configurations(backend, oracle, w, h, K, space, smooth, dup, seed):
    """HQ_PRUNE_ON forces the index-producing pruned kernel in hq_quantize and the S-CIELAB chain at any size: indices,
    counts and S-CIELAB errors against the oracle"""
    img = synth.synth_image(w, h, seed, smooth)
    pal = synth.synth_palettes(2, K, seed=seed % 100000)
    if dup and K > 1:
        rng = np.random.default_rng(seed)
        pal[:, rng.integers(0, K, K // 2 + 1)] = pal[:, rng.integers(0, K, K // 2 + 1)]
    backend.setPruning(PRUNE_ON)
    try:
        backend.setImage(img)
        q = backend.quantize(pal[0], space)
        want = oracle.quantize(img, pal[0], space, threads=THREADS)
        assert np.array_equal(q["idx"], want["idx"]) and np.array_equal(q["rgb"].reshape(-1, 3), want["rgb"])
        backend.scielabConfigure(72, 45.0)
        f, a = oracle.scielab_filters(72, 45.0)
        so = oracle.scielab_image(img, f, a, 0, THREADS)
        ev = oracle.scielab_eval(img, f, a, so, pal, space, 0, THREADS)
        got = backend.evalPalettesScielab(pal, space)
        assert np.array_equal(got["err_fx"], ev["err_fx"]) and np.array_equal(got["counts"], ev["counts"])
    finally:
        backend.setPruning(PRUNE_AUTO)
