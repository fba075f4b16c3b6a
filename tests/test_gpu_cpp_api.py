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


@pytest.mark.parametrize("w,h,K,imax", [(96, 64, 8, 120),      # below the pruning policy's thresholds: exhaustive kernel
                                        (320, 256, 32, 30)])    # hq_search_eval_flags picks the exact pruned kernel
def test_cpp_plugin_end_to_end(oracle, hqlib, tmp_path, w, h, K, imax):
    exe = str(tmp_path / "plugin_cpp_test")
    libdir = os.path.join(REPO, "hybridquantization_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(REPO, "include"), "-o", exe,
                    os.path.join(REPO, "tests", "cpp", "plugin_cpp_test.cpp"), "-L", libdir, "-lhq_b200", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([exe, str(w), str(h), str(K), str(imax)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    img = synth.synth_image(w, h, 5)
    p = oracle.swasa_params(population=4, imax=imax, seed=4242)
    obest, oerr, _ = oracle.find_best_quantization(img, K, p, threads=THREADS)
    assert float.fromhex(got["best_error"]) == oerr
    assert got["float_image_same"] == 1   # quantization(float planes) == quantization(u8)
    assert np.array_equal(np.array([float.fromhex(x) for x in got["best_colors"]], np.float32).view(np.uint32), obest.reshape(-1).view(np.uint32))
    q = oracle.quantize(img, obest)["rgb"].reshape(-1)
    hsh = 0
    for b in q.tolist():
        hsh = (hsh * 1099511628211 + b) & 0xFFFFFFFFFFFFFFFF
    assert got["image_hash"] == hsh
    of, oa = oracle.scielab_filters(72, 45.0)
    oe = oracle.error_image(img, q.reshape(img.shape), of, oa, 0, THREADS)
    assert float.fromhex(got["error_image_mean"]) == oe["deltaE"]
    mh = 0
    for b in oe["errorImageU8"].tolist():
        mh = (mh * 1099511628211 + b) & 0xFFFFFFFFFFFFFFFF
    assert got["error_map_hash"] == mh
    pal = synth.synth_palettes(4, K)
    ores = oracle.assign_reduce(img, pal)
    want = [oracle.cost(int(ores["err_fx"][i]), ores["counts"][i], w * h, 2.0) for i in range(4)]
    assert [float.fromhex(x) for x in got["costs"]] == want


def test_jni_shim_through_a_fake_jnienv(oracle, hqlib, tmp_path):
    """java/jni/hq_jni.c executed on the GPU without a JVM: compiled against tests/stubs/jni.h and driven by
    tests/cpp/jni_harness.c, whose JNIEnv pins plain C buffers.  Every native method of CudaImageManipulation runs; the
    integers must equal the oracle's and an unsupported K must raise the Java exception."""
    exe = str(tmp_path / "jni_harness")
    libdir = os.path.join(REPO, "hybridquantization_b200")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(REPO, "tests", "stubs"), "-I", os.path.join(REPO, "include"), "-o", exe,
                    os.path.join(REPO, "tests", "cpp", "jni_harness.c"), os.path.join(REPO, "java", "jni", "hq_jni.c"),
                    "-L", libdir, "-lhq_b200", f"-Wl,-rpath,{libdir}"], check=True)
    w, h, K, B = 320, 240, 48, 3
    r = subprocess.run([exe, str(w), str(h), str(K)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    img = synth.synth_image(w, h, 5)
    pal = np.zeros((B, K, 4), np.float32)
    for b in range(B):
        for k in range(K):
            for c in range(3):
                pal[b, k, c] = np.float32(np.float32((b * 7919 + k * 104729 + c * 1299709) % 1000) / np.float32(999.0))
    want = oracle.assign_reduce(img, pal, threads=THREADS)
    assert got["pixels"] == w * h and got["threw_on_bad_k"] == 1
    assert got["err_fx"] == [int(v) for v in want["err_fx"]]
    assert got["err_fx_float_image"] == got["err_fx"]   # nSetImageFloat with the c/255 planes
    assert got["counts"] == [int(v) for v in want["counts"].reshape(-1)]
    q = oracle.quantize(img, pal[0])["rgb"].reshape(-1)
    hsh = 0
    for v in q.tolist():
        hsh = (hsh * 1099511628211 + v) & 0xFFFFFFFFFFFFFFFF
    assert got["image_hash"] == hsh
