"""numpy synthetic-input generators (used by bench.py) against the oracle's C versions."""
import ctypes as C

import numpy as np

from hybridquantization_b200 import synth


def test_images_match_oracle(oracle):
    for (w, h) in ((1, 1), (7, 3), (64, 48), (101, 33)):
        for smooth in (False, True):
            a = synth.synth_image(w, h, synth.SEED_BASE + 2, smooth)
            b = oracle.synth_image(w, h, synth.SEED_BASE + 2, smooth)
            assert np.array_equal(a, b), (w, h, smooth)


def test_palettes_follow_generate_random_colors_order(oracle):
    L = oracle.load()
    pal = synth.synth_palettes(3, 5, 1234)
    r = oracle.Rng()
    L.hqo_rng_seed(C.byref(r), 1234)
    want = np.empty((3, 5, 4), np.float32)
    for b in range(3):
        L.hqo_generate_random_colors(C.byref(r), 5, want[b].ctypes.data_as(C.c_void_p))
    assert np.array_equal(pal.view(np.uint32), want.view(np.uint32))


def test_row_generation_matches_whole_image():
    whole = synth.synth_image(37, 29, 99)
    for r0, r1 in ((0, 29), (0, 1), (3, 17), (28, 29), (5, 5)):
        assert np.array_equal(synth.synth_image_rows(37, 29, 99, r0, r1), whole[r0:r1])
