"""BASELINE.json's named configurations at their FULL schedule lengths ("truncated trajectories"):
every candidate cost of every iteration, the final palette and the output image of the CUDA path against the oracle.

  C1  512 x 512, K = 16, population 4, the plugin's default 5,000 iterations (HybridQuantization.java:197-199)
  C2  1920 x 1080, K = 256, population 4, 1,000 iterations, exact pruning on (the search's default policy)
  4K  3840 x 2160, K = 256, reference-faithful scoring (sRGB assignment + S-CIELAB filters), 50 iterations
The oracle is multithreaded C; the three runs take about two minutes of host time on a 16-core box."""
import os

import numpy as np
import pytest

from helpers import bits
from hybridquantization_b200 import COST_SCIELAB, PRUNE_AUTO, SPACE_LAB, SPACE_SRGB, SWASA, synth

pytestmark = pytest.mark.gpu
THREADS = max(1, len(os.sched_getaffinity(0)))


def _compare(backend, oracle, img, K, space=SPACE_LAB, cost_model=0, **kw):
    backend.convergence = True
    best, err, tr, its = backend.findBestQuantization(K, SWASA(space=space, costModel=cost_model, **kw), trace=True)
    p = oracle.swasa_params(space=space, cost_model=cost_model, **kw)
    obest, oerr, otr = oracle.find_best_quantization(img, K, p, trace=True, threads=THREADS)
    assert its == kw["imax"]
    assert np.array_equal(tr.view(np.uint64), otr.view(np.uint64)), "a candidate cost differs from the oracle's"
    assert err == oerr and np.array_equal(bits(best), bits(obest))
    got, want = backend.quantize(best, space), oracle.quantize(img, obest, space)
    assert np.array_equal(got["rgb"].reshape(-1, 3), want["rgb"]) and np.array_equal(got["idx"], want["idx"])
    return tr


def test_c1_full_5000_iterations(backend, oracle):
    img = synth.synth_image(512, 512, synth.SEED_BASE + 1, smooth=True)
    backend.setImage(img)
    tr = _compare(backend, oracle, img, 16, population=4, imax=5000, seed=77760)
    assert tr.shape == (5001, 4)


def test_c2_1080p_k256_1000_iterations_pruned(backend, oracle):
    img = synth.synth_image(1920, 1080, synth.SEED_BASE + 2, smooth=True)
    backend.setImage(img)
    backend.setPruning(PRUNE_AUTO)
    assert backend.searchEvalFlags(256) != 0   # the search scores with the exact pruned kernel here
    _compare(backend, oracle, img, 256, population=4, imax=1000, seed=77760)


def test_4k_k256_reference_faithful_50_iterations(backend, oracle):
    img = synth.synth_image(3840, 2160, synth.SEED_BASE + 3, smooth=True)
    backend.setImage(img)
    backend.scielabConfigure(72, 45.0)
    _compare(backend, oracle, img, 256, space=SPACE_SRGB, cost_model=COST_SCIELAB, population=2, imax=50, iTc=5, seed=20261018)
