"""The C++ host mirror (include/hq_plugin.hpp) used directly from a C++ program, linked against
libhq_b200.so: HybridQuantization::quantization() and the class-level API, checked against the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from hybridquantization_b200 import synth

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
THREADS = max(1, len(os.sched_getaffinity(0)))


def test_cpp_plugin_end_to_end(oracle, hqlib, tmp_path):
    exe = str(tmp_path / "plugin_cpp_test")
    libdir = os.path.join(REPO, "hybridquantization_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(REPO, "include"), "-o", exe,
                    os.path.join(REPO, "tests", "cpp", "plugin_cpp_test.cpp"), "-L", libdir, "-lhq_b200", f"-Wl,-rpath,{libdir}"], check=True)
    w, h, K, imax = 96, 64, 8, 120
    r = subprocess.run([exe, str(w), str(h), str(K), str(imax)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    img = synth.synth_image(w, h, 5)
    p = oracle.swasa_params(population=4, imax=imax, seed=4242)
    obest, oerr, _ = oracle.find_best_quantization(img, K, p, threads=THREADS)
    assert float.fromhex(got["best_error"]) == oerr
    assert np.array_equal(np.array([float.fromhex(x) for x in got["best_colors"]], np.float32).view(np.uint32), obest.reshape(-1).view(np.uint32))
    q = oracle.quantize(img, obest)["rgb"].reshape(-1)
    hsh = 0
    for b in q.tolist():
        hsh = (hsh * 1099511628211 + b) & 0xFFFFFFFFFFFFFFFF
    assert got["image_hash"] == hsh
    pal = synth.synth_palettes(4, K)
    ores = oracle.assign_reduce(img, pal)
    want = [oracle.cost(int(ores["err_fx"][i]), ores["counts"][i], w * h, 2.0) for i in range(4)]
    assert [float.fromhex(x) for x in got["costs"]] == want
