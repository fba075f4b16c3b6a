"""Randomised CPU-vs-GPU property test (hypothesis): arbitrary image shapes, palette sizes, batch sizes,
spaces, white points, kernel variants and palette pathologies (duplicates, clamped colours) — the
integers from the CUDA path must equal the oracle's, and quantize() must reproduce its indices."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from hybridquantization_b200 import EVAL_FORCE_CHUNKED, EVAL_FORCE_DIRECT, EVAL_FORCE_PREFILTER, synth

pytestmark = pytest.mark.gpu
THREADS = max(1, len(os.sched_getaffinity(0)))


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture], derandomize=True)
@given(w=st.integers(1, 300), h=st.integers(1, 40), K=st.integers(1, 130), B=st.integers(1, 5), space=st.integers(0, 1), wp=st.integers(0, 1),
       variant=st.sampled_from([0, EVAL_FORCE_DIRECT, EVAL_FORCE_CHUNKED, EVAL_FORCE_PREFILTER]), smooth=st.booleans(),
       patho=st.sampled_from(["none", "dup", "clamp", "tiny"]), seed=st.integers(0, 2 ** 31))
def test_random_configurations(backend, oracle, w, h, K, B, space, wp, variant, smooth, patho, seed):
    img = synth.synth_image(w, h, seed, smooth)
    pal = synth.synth_palettes(B, K, seed=seed % 100000)
    rng = np.random.default_rng(seed)
    if patho == "dup" and K > 1:      # duplicated colours anywhere in the palette
        src = rng.integers(0, K, K // 2 + 1); dst = rng.integers(0, K, K // 2 + 1)
        pal[:, dst] = pal[:, src]
    elif patho == "clamp":            # what SWASA's clamp() produces late in a run: many channels at exactly 0 or 1
        pal[..., :3] = np.round(pal[..., :3] * 2) / 2
    elif patho == "tiny":             # colours a few ulps apart
        pal[..., :3] = pal[:, :1, :3] + (rng.integers(-4, 5, pal[..., :3].shape) * np.float32(6e-8)).astype(np.float32)
        pal = np.clip(pal, 0, 1).astype(np.float32)
    backend.setImage(img, wp)
    got = backend.evalPalettes(pal, space, sums=True, flags=variant)
    want = oracle.assign_reduce(img, pal, space, wp, want_idx=True, threads=THREADS)
    assert np.array_equal(got["err_fx"], want["err_fx"])
    assert np.array_equal(got["counts"], want["counts"])
    assert np.array_equal(got["sums_fx"], want["sums_fx"])
    assert np.array_equal(backend.quantize(pal[0], space)["idx"], want["idx"][0])
