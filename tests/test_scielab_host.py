"""S-CIELAB filter bank: the product's C++ restatement (hq::ScielabProcessor::buildFilters) against the
oracle's C restatement of ScielabProcessor.java:66-181, bit for bit, plus the geometry SURVEY A.5 derives."""
import numpy as np
import pytest

from hybridquantization_b200 import ScielabProcessor


@pytest.mark.parametrize("dpi,dist", [(72, 45.0), (96, 60.0), (300, 45.0), (150, 30.0), (1000, 30.0), (30, 100.0)])
def test_filter_bank_matches_oracle(hqlib, oracle, dpi, dist):
    f, a = ScielabProcessor.buildFilters(dpi, dist)
    of, oa = oracle.scielab_filters(dpi, dist)
    assert f.shape == of.shape
    assert np.array_equal(f.view(np.uint32), of.view(np.uint32))
    assert np.array_equal(a.view(np.uint32), oa.view(np.uint32))
    assert np.array_equal(a, np.abs(f[2]))


def test_default_geometry(hqlib):
    # dpi 72, 45 cm -> 22 samples/degree -> uprate 11 -> 21 taps, half-size 10
    f, a = ScielabProcessor.buildFilters()
    assert f.shape == (7, 21)
    assert np.allclose(f, f[:, ::-1])                       # symmetric
    assert (f[2] < 0).all() and (f[[0, 1, 3, 4, 5, 6]] >= 0).all()  # only the third luminance Gaussian is negative
    # the three luminance filter pairs approximately restore unit DC gain: sum(k1)^2 + sum(k2)^2 + sum|k3|*sum(k3)
    dc = f[0].sum() ** 2 + f[1].sum() ** 2 + a.sum() * f[2].sum()
    assert abs(dc - 1.0) < 0.02
