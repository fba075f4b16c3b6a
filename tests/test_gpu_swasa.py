"""Fixed-seed SWASA runs: the GPU path must reproduce the oracle's trajectory bit for bit —
every candidate cost of every iteration, the final palette and the quantised image."""
import os

import numpy as np
import pytest

from helpers import bits, from_bits, load_golden
from hybridquantization_b200 import SPACE_LAB, SPACE_SRGB, SWASA, HybridQuantization, synth

pytestmark = pytest.mark.gpu
THREADS = max(1, len(os.sched_getaffinity(0)))


def test_golden_swasa_vectors_on_gpu(backend):
    g = load_golden("swasa_vectors.json")
    for name, v in g.items():
        img = synth.synth_image(v["w"], v["h"], v["image_seed"], v["smooth"])
        backend.setImage(img)
        backend.convergence = True
        sw = SWASA(population=v["population"], imax=v["imax"], iTc=v["iTc"], seed=v["seed"], convergence=bool(v["convergence"]), space=v["space"])
        best, err, tr, its = backend.findBestQuantization(v["K"], sw, trace=True)
        assert its == v["imax"]
        assert [float(x).hex() for x in tr.reshape(-1)] == v["trace"], name
        assert float(err).hex() == v["best_error"], name
        assert np.array_equal(bits(best).ravel(), bits(from_bits(v["best_colors"]))), name


@pytest.mark.parametrize("space", [SPACE_LAB, SPACE_SRGB])
def test_c1_512x512_k16_trajectory_matches_oracle(backend, oracle, space):
    # BASELINE config C1 (reference defaults, population 4) with a shortened schedule
    img = synth.synth_image(512, 512, synth.SEED_BASE + 1, smooth=True)
    imax = 400
    backend.setImage(img)
    backend.convergence = True
    sw = SWASA(population=4, imax=imax, seed=77760, space=space)
    best, err, tr, its = backend.findBestQuantization(16, sw, trace=True)
    p = oracle.swasa_params(population=4, imax=imax, seed=77760, space=space)
    obest, oerr, otr = oracle.find_best_quantization(img, 16, p, trace=True, threads=THREADS)
    assert its == imax
    assert np.array_equal(tr.view(np.uint64), otr.view(np.uint64))
    assert err == oerr
    assert np.array_equal(bits(best), bits(obest))
    got = backend.quantize(best, space)
    want = oracle.quantize(img, obest, space)
    assert np.array_equal(got["rgb"].reshape(-1, 3), want["rgb"]) and np.array_equal(got["idx"], want["idx"])


def test_plugin_entry_point_end_to_end(oracle):
    img = synth.synth_image(96, 64, 5, smooth=True)
    hq = HybridQuantization(nbOfColors=8, populationSize=4, imax=150, seed=4242)
    res = hq.quantization(img)
    p = oracle.swasa_params(population=4, imax=150, seed=4242)
    obest, oerr, _ = oracle.find_best_quantization(img, 8, p, threads=THREADS)
    assert res["bestError"] == oerr and np.array_equal(bits(res["bestColors"]), bits(obest))
    assert np.array_equal(res["image"].reshape(-1, 3), oracle.quantize(img, obest)["rgb"])
    assert res["image"].shape == img.shape and res["image"].dtype == np.uint8
    with pytest.raises(ValueError):
        hq.quantization(np.zeros((0, 0, 3), np.uint8))


def test_stop_request_returns_best_so_far(backend):
    # EzStoppable (HybridQuantization.java:311-319): the flag is polled once per iteration
    import threading
    import time

    img = synth.synth_image(512, 512, 3)
    backend.setImage(img)
    sw = SWASA(population=4, imax=200000, seed=1)
    t = threading.Timer(0.5, backend.requestStop)
    t.start()
    t0 = time.time()
    best, err, _, its = backend.findBestQuantization(64, sw)
    t.join()
    assert 0 < its < 200000 and time.time() - t0 < 30
    assert np.isfinite(err) and best.shape == (64, 4)


def test_progress_hook_every_ten_iterations(backend):
    img = synth.synth_image(64, 64, 2)
    backend.setImage(img)
    seen = []
    backend.setProgress(lambda i, m, e: seen.append((i, m, e)))
    best, err, _, its = backend.findBestQuantization(8, SWASA(population=2, imax=35, seed=3))
    backend.setProgress(None)
    assert [s[0] for s in seen] == [10, 20, 30] and all(s[1] == 35 for s in seen)
    assert seen[-1][2] >= err and all(a[2] >= b[2] for a, b in zip(seen, seen[1:]))   # best error never increases


def test_cuda_graph_replay_gives_the_same_search(backend, oracle):
    """hq_set_graphs: the captured launch set must reproduce the plain launches bit for bit, across image and argument changes"""
    from hybridquantization_b200 import EVAL_PRUNE
    backend.setGraphs(True)
    try:
        for (w, h, K, P, seed) in ((96, 80, 12, 3, 5), (300, 260, 40, 4, 6), (96, 80, 12, 3, 5)):
            img = synth.synth_image(w, h, seed, True)
            backend.setImage(img)
            best, err, tr, its = backend.findBestQuantization(K, SWASA(population=P, imax=30, iTc=5, seed=seed), trace=True)
            obest, oerr, otr = oracle.find_best_quantization(img, K, oracle.swasa_params(population=P, imax=30, iTc=5, seed=seed), trace=True)
            assert its == 30 and err == oerr and np.array_equal(tr.view(np.uint64), otr.view(np.uint64))
            assert np.array_equal(best.view(np.uint32), obest.view(np.uint32))
            pal = synth.synth_palettes(P, K, seed=seed)
            want = oracle.assign_reduce(img, pal)
            for flags in (0, EVAL_PRUNE, 0, 0, EVAL_PRUNE, EVAL_PRUNE, EVAL_PRUNE):   # repeated signatures: plain, captured, replayed
                got = backend.evalPalettes(pal, flags=flags)
                assert np.array_equal(got["err_fx"], want["err_fx"]) and np.array_equal(got["counts"], want["counts"])
    finally:
        backend.setGraphs(False)


def test_allreduce_hook_is_applied_on_every_path(hqlib):
    """One GPU standing in for three identical ranks: the hook multiplies the device result words by 3 on the library's
    stream.  Every entry that reduces (both host I/O paths of hq_eval_palettes, the S-CIELAB chain, the error image) must
    then return exactly three times the hook-less integers."""
    import torch

    from hybridquantization_b200 import EVAL_PRUNE, ImageManipulation
    from hybridquantization_b200.dist import _DevWords

    calls = []

    def triple(d_ptr, n_words, stream):
        t = torch.as_tensor(_DevWords(d_ptr, n_words), device=torch.device("cuda", 0))
        with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
            t.mul_(3)
        calls.append(n_words)
        return 0

    img = synth.synth_image(320, 200, 8, smooth=True)
    be = ImageManipulation("CIE76", False, True, 0)
    try:
        be.setImage(img)
        be.scielabConfigure(72, 45.0)
        small, big = synth.synth_palettes(3, 16), synth.synth_palettes(24, 300, seed=4)   # direct host I/O / DMA copies
        base = {"small": be.evalPalettes(small, sums=True), "big": be.evalPalettes(big), "pruned": be.evalPalettes(big, flags=EVAL_PRUNE),
                "sc": be.evalPalettesScielab(small, SPACE_SRGB), "err": be.computeError(img[::-1].copy())["deltaE"]}
        be.setAllreduce(triple)
        got = {"small": be.evalPalettes(small, sums=True), "big": be.evalPalettes(big), "pruned": be.evalPalettes(big, flags=EVAL_PRUNE),
               "sc": be.evalPalettesScielab(small, SPACE_SRGB), "err": be.computeError(img[::-1].copy())["deltaE"]}
        assert len(calls) == 5
        for k in ("small", "big", "pruned", "sc"):
            assert np.array_equal(got[k]["err_fx"], 3 * base[k]["err_fx"]) and np.array_equal(got[k]["counts"], 3 * base[k]["counts"]), k
        assert np.array_equal(got["small"]["sums_fx"], 3 * base["small"]["sums_fx"])
        # the error image divides the summed error by the GLOBAL pixel count (here the same image): 3x the mean
        assert abs(got["err"] - 3 * base["err"]) <= 1e-12 * base["err"]
    finally:
        be.close()


@pytest.mark.parametrize("space", [SPACE_LAB, SPACE_SRGB])
def test_persistent_evaluator_gives_the_same_search(space, monkeypatch):
    """HQ_PERSIST=1 (off by default): one persistent kernel serves every evaluation of a small LAB-cost search through a pinned mailbox
    (hq_kernels.cu, assign_persist_kernel).  Same trajectory bit for bit — also when the kernel leaves in the middle of the search
    because the host kept it waiting (HQ_PERSIST_IDLE_MS=2 and a progress listener that sleeps): the library then goes back to one
    launch per evaluation."""
    import time

    from hybridquantization_b200 import ImageManipulation

    img = synth.synth_image(300, 211, 99, smooth=True)
    sw = dict(population=4, imax=400, seed=4321, space=space)
    ref = ImageManipulation("CIE76", False, True, 0)
    ref.setImage(img)
    want = ref.findBestQuantization(8, SWASA(**sw), trace=True)
    ref.close()
    for idle_ms, sleeper in (("200", None), ("2", lambda *a: time.sleep(0.02))):
        monkeypatch.setenv("HQ_PERSIST", "1")
        monkeypatch.setenv("HQ_PERSIST_IDLE_MS", idle_ms)
        be = ImageManipulation("CIE76", False, True, 0)
        be.setImage(img)
        if sleeper:
            be.setProgress(sleeper)
        got = be.findBestQuantization(8, SWASA(**sw), trace=True)
        again = be.findBestQuantization(8, SWASA(**sw), trace=True)      # a second search on the same context starts a new kernel
        q = be.quantize(got[0], space)                                    # and the stream is free again afterwards
        be.close()
        for g in (got, again):
            assert g[3] == want[3] == 400 and g[1] == want[1]
            assert np.array_equal(g[2].view(np.uint64), want[2].view(np.uint64)) and np.array_equal(bits(g[0]), bits(want[0]))
        assert q["idx"].size == img.shape[0] * img.shape[1]
