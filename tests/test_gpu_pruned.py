"""Exact pruning (HQ_EVAL_PRUNE, csrc/hq_pruned.cu) against the oracle and the exhaustive kernel: the integers must be
IDENTICAL — the pruned path only skips colours that provably cannot be nearest (or tie) for any pixel of a chunk."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from hybridquantization_b200 import EVAL_PRUNE, PRUNE_OFF, PRUNE_ON, SPACE_LAB, SWASA, WHITEPOINT_D50, synth

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
