"""Host side of the product (java.util.Random, SWASA schedule pieces, cost) against the oracle."""
import ctypes as C

import numpy as np

from helpers import bits


def test_java_random_matches_jdk_known_answers(hqlib):
    from hybridquantization_b200 import JavaRandom

    assert JavaRandom(42).nextInt() == -1170105035
    assert JavaRandom(0).nextInt() == -1155484576
    assert JavaRandom(0).nextDouble() == 0.730967787376657
    r = JavaRandom(42)
    assert [float(r.nextFloat()).hex() for _ in range(3)] == ["0x1.74833a0000000p-1", "0x1.bfd1400000000p-5", "0x1.5dcf760000000p-1"]
    r = JavaRandom(42)
    assert [r.nextDouble() for _ in range(2)] == [0.7275636800328681, 0.6832234717598454]


def test_random_stream_matches_oracle(hqlib, oracle):
    from hybridquantization_b200 import JavaRandom

    L = oracle.load()
    for seed in (77760, -5, 1 << 40):
        a = JavaRandom(seed)
        r = oracle.Rng()
        L.hqo_rng_seed(C.byref(r), seed)
        for i in range(200):
            if i % 3 == 0:
                assert a.nextDouble() == L.hqo_rng_next_double(C.byref(r))
            else:
                assert a.nextFloat() == L.hqo_rng_next_float(C.byref(r))


def test_swasa_moves_match_oracle(hqlib, oracle):
    from hybridquantization_b200 import SWASA

    L = oracle.load()
    for imax, beta, s0 in ((5000, 5.3, 100.0), (60, 2.0, 256.0), (7, 0.0, 1.0)):
        sw = SWASA(imax=imax, beta=beta, s0=s0, seed=99)
        p = oracle.swasa_params(imax=imax, beta=beta, s0=s0, seed=99)
        r = oracle.Rng()
        L.hqo_rng_seed(C.byref(r), 99)
        K = 13
        want = np.empty((K, 4), np.float32)
        L.hqo_generate_random_colors(C.byref(r), K, want.ctypes.data_as(C.c_void_p))
        cur = sw.generateRandomColors(K)
        assert np.array_equal(bits(cur), bits(want))
        for ite in (1, 2, imax // 2 + 1, imax):
            assert np.float32(sw.maxStepWidth(ite)) == np.float32(L.hqo_max_step_width(C.byref(p), ite))
            nxt_want = np.empty_like(want)
            L.hqo_generate_neighboring_colors(C.byref(p), C.byref(r), want.ctypes.data_as(C.c_void_p), nxt_want.ctypes.data_as(C.c_void_p), K, ite)
            nxt = sw.generateNeighboringColors(cur, ite)
            assert np.array_equal(bits(nxt), bits(nxt_want))
            assert (nxt[:, :3] >= 0).all() and (nxt[:, :3] <= 1).all() and (nxt[:, 3] == 0).all()
            cur, want = nxt, nxt_want


def test_cost_matches_oracle(hqlib, oracle):
    rng = np.random.default_rng(0)
    for _ in range(50):
        K = int(rng.integers(1, 40))
        counts = rng.integers(0, 3, K).astype(np.uint64) * rng.integers(1, 1000, K).astype(np.uint64)
        err = int(rng.integers(0, 1 << 50))
        n = int(rng.integers(1, 1 << 26))
        got = hqlib.hq_cost(err, counts.ctypes.data_as(C.c_void_p), K, n, 2.0)
        assert got == oracle.cost(err, counts, n, 2.0)


def test_plugin_parameter_defaults(hqlib):
    # HybridQuantization.java:192-233
    from hybridquantization_b200 import HybridQuantization, _lib

    hq = HybridQuantization()
    assert (hq.nbOfColors, hq.populationSize, hq.imax, hq.delta) == (8, 4, 5000, 2.0)
    assert (hq.ConvEnable, hq.ConvDelay, hq.ConvSpread) == (True, 0.75, 0.15)
    assert (hq.T0, hq.iTc, hq.alpha, hq.s0, hq.beta) == (20.0, 20, 0.9, 100.0, 5.3)
    assert (hq.dpi, hq.ViewingDistance, hq.WhitePoint, hq.Verbose) == (72, 45.0, "D65", False)
    p = _lib.SwasaParams()
    hqlib.hq_swasa_default_params(C.byref(p))
    assert (p.population, p.imax, p.iTc, p.convergence) == (4, 5000, 20, 1)
    assert np.float32(p.alpha) == np.float32(0.9) and np.float32(p.beta) == np.float32(5.3)
