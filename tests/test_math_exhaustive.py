"""The product's single-source arithmetic (csrc/hq_math.h, host instantiation) against the oracle's
libm formulas, EXHAUSTIVELY over the float domains the path can reach.  Bit-exact, CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import fbits

THREADS = max(1, len(os.sched_getaffinity(0)))
CHUNK = 1 << 24


def host_range(lib, which, first, count):
    out = np.empty(count, np.float32)
    assert lib.hq_host_math_range(which, first, count, out.ctypes.data_as(C.c_void_p), THREADS) == 0
    return out


@pytest.mark.parametrize("which,lo,hi,name", [
    (0, 0.008856452070, 1.25, "(float)pow(t, 1.0/3.0) on (LABDELTA3, 1.25]"),
    (1, 0.0625, 1.0, "(float)pow(b, 2.4f) on [1/16, 1]"),
    (2, 0.0, 1.0, "sRGB decode on [0, 1] (every float)"),
    (3, 0.0, 4.0, "x / Xn(D65) on [0, 4]"),
    (4, 0.0, 4.0, "x / Zn(D65) on [0, 4]"),
    (5, 0.0, 4.0, "x / (3*LABDELTA2) on [0, 4]"),
    (6, 0.0, 4.0, "x / Xn(D50) on [0, 4]"),
    (7, 0.0, 4.0, "x / Zn(D50) on [0, 4]"),
])
def test_exhaustive_bit_exact(hqlib, oracle, which, lo, hi, name):
    a, b = fbits(lo), fbits(hi)
    bad = 0
    for s in range(a, b + 1, CHUNK):
        c = min(CHUNK, b + 1 - s)
        x = host_range(hqlib, which, s, c)
        y = oracle.math_range(which, s, c, THREADS)
        bad += int(np.count_nonzero(x.view(np.uint32) != y.view(np.uint32)))
    assert bad == 0, f"{bad} mismatches of {b - a + 1} for {name}"


def test_constant_division_negative_and_special_values(hqlib, oracle):
    # the Markstein path is only taken for 2^-60 <= |x| <= 4; everything else must fall back to '/'
    for which in (3, 4, 5, 6, 7):
        for lo, hi in ((-4.0, -2.0 ** -61), ):
            a, b = fbits(hi), fbits(lo)   # negative floats: larger magnitude = larger bit pattern
            step = 1 << 24
            bad = 0
            for s in range(a, b + 1, step * 8):   # every 8th block of 16M: 12 % sample of the negative range
                c = min(step, b + 1 - s)
                bad += int(np.count_nonzero(host_range(hqlib, which, s, c).view(np.uint32) != oracle.math_range(which, s, c, THREADS).view(np.uint32)))
            assert bad == 0
        special = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-30, 4.0000005, 1e20, 3.4e38, np.inf, -np.inf], np.float32)
        for v in special:
            x = host_range(hqlib, which, int(v.view(np.uint32)), 1)
            y = oracle.math_range(which, int(v.view(np.uint32)), 1, 1)
            assert x.view(np.uint32)[0] == y.view(np.uint32)[0], (which, v)


def test_palette_colour_to_lab_matches_oracle(hqlib, oracle):
    rng = np.random.default_rng(5)
    cols = np.concatenate([rng.random((4000, 3), dtype=np.float32), np.array([[0, 0, 0], [1, 1, 1], [0.04045, 0.04045, 1.0]], np.float32)])
    out = np.empty(3, np.float32)
    for wp in (0, 1):
        for c in cols:
            hqlib.hq_host_srgb_to_lab(c.ctypes.data_as(C.c_void_p), wp, out.ctypes.data_as(C.c_void_p))
            assert np.array_equal(out.view(np.uint32), oracle.srgb_to_lab(c, wp).view(np.uint32)), (c, wp)
