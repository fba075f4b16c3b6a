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


def _fnv(b: bytes) -> int:
    h = 0
    for v in b:
        h = (h * 1099511628211 + v) & 0xFFFFFFFFFFFFFFFF
    return h


def test_jni_shim_through_a_fake_jnienv(oracle, hqlib, tmp_path):
    """java/jni/hq_jni.c executed on the GPU without a JVM: compiled against tests/stubs/jni.h and driven by
    tests/cpp/jni_harness.c, whose JNIEnv hands out plain C buffers with lengths.  EVERY native method of
    CudaImageManipulation runs: image upload (u8 and float planes), both cost models, the filter bank, the reference's
    RGBtoXYZ -> XYZtoScielab route and its installation as the comparison target, quantize, the three error-image forms, the
    library-side search with a progress listener and a stop request, and the error paths (short / null arrays, bad K), which
    must surface as Java exceptions.  The integers and bits must equal the oracle's."""
    exe = str(tmp_path / "jni_harness")
    libdir = os.path.join(REPO, "hybridquantization_b200")
    subprocess.run(["gcc", "-O2", "-Wall", "-I", os.path.join(REPO, "tests", "stubs"), "-I", os.path.join(REPO, "include"), "-o", exe,
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
    assert got["pixels"] == w * h and got["devices"] == 1
    assert got["err_fx"] == [int(v) for v in want["err_fx"]]
    assert got["err_fx_float_image"] == got["err_fx"] == got["err_fx_again"]   # nSetImageFloat with the c/255 planes; intact after refused calls
    assert got["counts"] == [int(v) for v in want["counts"].reshape(-1)]
    q = oracle.quantize(img, pal[0])
    assert got["image_hash"] == _fnv(q["rgb"].reshape(-1).tobytes()) and got["f32_hash"] == _fnv(q["f32"].tobytes())
    # the reference-faithful chain, three ways to the same integers
    of, oa = oracle.scielab_filters(72, 45.0)
    so = oracle.scielab_image(img, of, oa, 0, THREADS)
    wsc = oracle.scielab_eval(img, of, oa, so, pal, 1, 0, THREADS)
    assert got["sc_err_fx"] == [int(v) for v in wsc["err_fx"]] and got["sc_counts"] == [int(v) for v in wsc["counts"].reshape(-1)]
    assert got["sc_err_fx_installed_bank"] == got["sc_err_fx"] == got["sc_err_fx_installed_original"]
    # error image: computeError on two Lab images == the one-call forms == the oracle
    oe = oracle.error_image(img, q["rgb"].reshape(img.shape), of, oa, 0, THREADS)
    assert float.fromhex(got["de_u8"]) == oe["deltaE"] == float.fromhex(got["de_f32"])
    assert abs(float.fromhex(got["de_lab"]) - oe["deltaE"]) <= 2.0 ** -23   # the reference's double sum of floats vs the 2^-24 fixed-point sum
    assert got["map_equal"] == 1 and got["map_equal_f32"] == 1 and got["map_hash"] == _fnv(oe["errorImage"].astype(np.float32).tobytes())
    # the search: trajectory and palette equal the oracle's; progress every 10 iterations; a stop request ends it early
    p = oracle.swasa_params(population=4, imax=45, seed=4242)
    obest, oerr, otr = oracle.find_best_quantization(img, K, p, trace=True, threads=THREADS)
    assert got["iterations"] == 45 and float.fromhex(got["best_error"]) == oerr
    assert got["trace_hash"] == _fnv(otr.tobytes()) and got["best_hash"] == _fnv(obest.tobytes())
    assert got["progress"] == [4, 10, 40, 45] and float.fromhex(got["progress_best"]) >= oerr
    assert 0 < got["iterations_stopped"] <= 11
    assert got["threw_short"] == 1 and got["cls_short"] == "java/lang/IllegalArgumentException"
    assert got["threw_null"] == 1 and got["cls_null"] == "java/lang/NullPointerException"
    assert got["threw_bad_k"] == 1 and got["threw_short_out"] == 1
    assert got["threw_ciede2000"] == 1 and got["de94_is_nan_or_positive"] == 1


def test_reference_one_shot_entries_match_the_compiled_reference(backend, oracle):
    """hq_rgb_to_xyz / hq_xyz_to_scielab / hq_scielab_set_image / hq_delta_e_images (the reference class's RGBtoXYZ :100,
    XYZtoScielab :285, findBestQuantization's inlineScielabOriginal :383, computeError :858) against the oracle, and — where
    oracle/_ref travelled to this box — against the reference's own kernels compiled for the CPU."""
    from oracle import hq_ref as R

    w, h = 211, 97
    img = synth.synth_image(w, h, 23, smooth=True)
    planes = np.ascontiguousarray((img.astype(np.float64) / 255.0).astype(np.float32).transpose(2, 0, 1)).reshape(3, -1)
    backend.setImage(img)
    backend.scielabConfigure(72, 45.0)
    ill = np.array([0.95047, 1.0, 1.0883], np.float32)
    xyz = backend.RGBtoXYZ(planes[0], planes[1], planes[2])
    lab = backend.XYZtoScielab(xyz, w, ill)
    sc = backend.scielabImage()                                  # the library's own S-CIELAB of the resident image, planes [3, n]
    assert np.array_equal(lab[:, :3].T.view(np.uint32), sc.view(np.uint32)) and not lab[:, 3].any()
    of, oa = oracle.scielab_filters(72, 45.0)
    assert np.array_equal(sc.view(np.uint32), oracle.scielab_image(img, of, oa, 0, THREADS).view(np.uint32))
    pal = synth.synth_palettes(2, 24)
    base = backend.evalPalettesScielab(pal)
    # a DIFFERENT target installed by the caller changes the costs; the original one restores them
    backend.scielabSetImage(np.roll(lab, 7, axis=0))
    assert not np.array_equal(backend.evalPalettesScielab(pal)["err_fx"], base["err_fx"])
    backend.scielabSetImage(lab)
    again = backend.evalPalettesScielab(pal)
    assert np.array_equal(again["err_fx"], base["err_fx"]) and np.array_equal(again["counts"], base["counts"])
    # computeError on two Lab images
    other = np.roll(lab, 3, axis=0).copy()
    eimg = np.full(lab.shape, 7.0, np.float32)
    mean = backend.computeErrorLab(lab, other, eimg)
    e = np.sqrt(((lab[:, :3].astype(np.float64) - other[:, :3].astype(np.float64)) ** 2).sum(axis=1))
    assert abs(mean - e.mean()) < 1e-5 and (eimg[:, 3] == 7.0).all() and np.array_equal(eimg[:, 0], eimg[:, 2])
    if R.available() or R.build():
        rx = R.rgb_to_xyz(planes.reshape(3, -1))
        assert np.array_equal(xyz.view(np.uint32), rx.reshape(-1, 4).view(np.uint32))
        rl = R.xyz_to_scielab(rx, R.pack_filters(of, oa), w)
        assert np.array_equal(lab.view(np.uint32), rl.reshape(-1, 4).view(np.uint32))
        L = R.load()
        re = np.zeros(w * h, np.float32)
        L.refcl_CIEDE(R._ptr(np.ascontiguousarray(lab)), R._ptr(other), R._ptr(re), w * h, 1)
        assert mean == float(np.cumsum(re.astype(np.float64))[-1]) / (w * h)   # :886-893: floats summed in a double, in pixel order
        d = np.float32(255.0) - re
        assert np.array_equal(eimg[:, 1].view(np.uint32), ((d * d) / np.float32(65025.0)).view(np.uint32))
