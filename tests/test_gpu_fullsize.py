"""BASELINE.json's full sizes: oracle comparison where the oracle finishes in seconds, and
size-independent properties (conservation, shard additivity, idempotence) everywhere else."""
import os

import numpy as np
import pytest

from hybridquantization_b200 import SPACE_LAB, SPACE_SRGB, synth

pytestmark = pytest.mark.gpu
THREADS = max(1, len(os.sched_getaffinity(0)))


def _sum_shards(backend, img, pal, cuts, space=SPACE_LAB):
    err = np.zeros(pal.shape[0], np.int64); cnt = np.zeros(pal.shape[:2], np.uint64); sums = np.zeros(pal.shape[:2] + (3,), np.int64)
    for a, b in zip(cuts[:-1], cuts[1:]):
        backend.setImage(img[a:b])
        r = backend.evalPalettes(pal, space, sums=True)
        err += r["err_fx"]; cnt += r["counts"]; sums += r["sums_fx"]
    return err, cnt, sums


def test_c2_1080p_k256_matches_oracle(backend, oracle):
    img = synth.synth_image(1920, 1080, synth.SEED_BASE + 2)
    pal = synth.synth_palettes(2, 256)
    backend.setImage(img)
    got = backend.evalPalettes(pal, sums=True)
    want = oracle.assign_reduce(img, pal, threads=THREADS)
    assert np.array_equal(got["err_fx"], want["err_fx"]) and np.array_equal(got["counts"], want["counts"]) and np.array_equal(got["sums_fx"], want["sums_fx"])


def test_c3_4k_k256_matches_oracle_and_properties(backend, oracle):
    img = synth.synth_image(3840, 2160, synth.SEED_BASE + 3, smooth=True)
    pal = synth.synth_palettes(64, 256)
    backend.setImage(img)
    got = backend.evalPalettes(pal)                      # 64 candidates per launch
    n = 3840 * 2160
    assert (got["counts"].sum(axis=1) == n).all()        # every pixel assigned exactly once
    want = oracle.assign_reduce(img, pal[:1], threads=THREADS)   # the oracle on one candidate
    assert got["err_fx"][0] == want["err_fx"][0] and np.array_equal(got["counts"][0], want["counts"][0])
    # candidates are independent: evaluating a sub-batch gives the same integers
    sub = backend.evalPalettes(pal[5:9])
    assert np.array_equal(sub["err_fx"], got["err_fx"][5:9]) and np.array_equal(sub["counts"], got["counts"][5:9])
    # quantise, then quantising the quantised image is a fixed point (idempotence)
    q1 = backend.quantize(pal[0])
    backend.setImage(q1["rgb"])
    q2 = backend.quantize(pal[0])
    backend.setImage(q2["rgb"])
    q3 = backend.quantize(pal[0])
    assert np.array_equal(q2["rgb"], q3["rgb"])


def test_c4_64mp_row_shards_add_up(backend):
    # 8192 x 8192 = 64 MP, K=256: the result of 1 / 2 / 4 / 8 row shards is the same integers
    img = synth.synth_image(8192, 8192, synth.SEED_BASE + 4)
    pal = synth.synth_palettes(2, 256)
    whole = _sum_shards(backend, img, pal, [0, 8192])
    assert (whole[1].sum(axis=1) == 8192 * 8192).all()
    for g in (2, 8):
        cuts = [8192 * r // g for r in range(g + 1)]
        part = _sum_shards(backend, img, pal, cuts)
        assert all(np.array_equal(a, b) for a, b in zip(whole, part)), g
    ragged = _sum_shards(backend, img, pal, [0, 1, 4097, 8192])
    assert all(np.array_equal(a, b) for a, b in zip(whole, ragged))


@pytest.mark.parametrize("K", [8, 16, 32, 64, 128, 256, 512, 1024])
def test_c5_palette_sweep_4k_conservation_and_srgb(backend, oracle, K):
    img = synth.synth_image(3840, 2160, synth.SEED_BASE + 5)
    pal = synth.synth_palettes(1, K)
    backend.setImage(img)
    got = backend.evalPalettes(pal, sums=True)
    n = 3840 * 2160
    assert int(got["counts"].sum()) == n
    # total Lab sum is palette independent: it equals the sum over the image
    lab = backend.labImage()
    tot = np.array([np.rint(lab[c].astype(np.float64) * 16777216.0).astype(np.int64).sum() for c in range(3)])
    assert np.array_equal(got["sums_fx"][0].sum(axis=0), tot)
    # the oracle on the same candidate at every K of the sweep (BASELINE configs[4]; < 0.3 s of host time even at K = 1024)
    want = oracle.assign_reduce(img, pal, threads=THREADS)
    assert got["err_fx"][0] == want["err_fx"][0] and np.array_equal(got["counts"], want["counts"]) and np.array_equal(got["sums_fx"], want["sums_fx"])


def test_pruned_scoring_at_full_sizes(backend):
    """BASELINE sizes through the exact pruned kernel: identical integers to the exhaustive kernel at 4K (64 candidates) and
    64 MP; beyond HQ_MAX_COLORS (K = 4096, pruned kernel only) conservation laws hold."""
    from hybridquantization_b200 import EVAL_PRUNE

    img = synth.synth_image(3840, 2160, synth.SEED_BASE + 3)
    pal = synth.synth_palettes(64, 256)
    backend.setImage(img)
    a = backend.evalPalettes(pal, sums=True)
    b = backend.evalPalettes(pal, sums=True, flags=EVAL_PRUNE)
    assert all(np.array_equal(a[k], b[k]) for k in ("err_fx", "counts", "sums_fx"))
    big = synth.synth_palettes(2, 4096)
    r = backend.evalPalettes(big, sums=True)
    assert (r["counts"].sum(axis=1) == 3840 * 2160).all()
    lab = backend.labImage()
    tot = np.array([np.rint(lab[c].astype(np.float64) * 16777216.0).astype(np.int64).sum() for c in range(3)])
    assert np.array_equal(r["sums_fx"][0].sum(axis=0), tot) and np.array_equal(r["sums_fx"][1].sum(axis=0), tot)
    # a superset palette can only lower the error: the first 256 colours of `big` vs all 4096
    sub = backend.evalPalettes(big[:, :256].copy(), flags=EVAL_PRUNE)
    assert (r["err_fx"] <= sub["err_fx"]).all()
    img = synth.synth_image(8192, 8192, synth.SEED_BASE + 4)
    pal = synth.synth_palettes(2, 256)
    backend.setImage(img)
    a = backend.evalPalettes(pal, sums=True)
    b = backend.evalPalettes(pal, sums=True, flags=EVAL_PRUNE)
    assert all(np.array_equal(a[k], b[k]) for k in ("err_fx", "counts", "sums_fx"))
