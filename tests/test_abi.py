"""The C-ABI library: loads, exports every symbol include/hq_b200.h declares, and refuses to run
without a GPU (no CPU fallback).  CPU only — no compute calls."""
import ctypes as C
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(REPO, "include", "hq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hq_[a-z0-9_]+)\s*\(", src)) - {"hq_allreduce_fn"})


def test_header_symbols_are_exported(hqlib):
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(hqlib, n), f"{n} declared in include/hq_b200.h but not exported"


def test_binding_covers_header():
    from hybridquantization_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_oracle_or_cpu_fallback_in_product():
    # the product path must never import / link the oracle
    pkg = os.path.join(REPO, "hybridquantization_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(root, f)).read()
                assert "hq_oracle" not in text and "hqo_" not in text, f
    for f in os.listdir(os.path.join(REPO, "include")):
        assert "hqo_" not in open(os.path.join(REPO, "include", f)).read()


def test_create_fails_loudly_without_gpu(hqlib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    rc = hqlib.hq_create(0, C.byref(ctx))
    assert rc == 2 and not ctx  # HQ_ERR_CUDA
    msg = hqlib.hq_last_error(None).decode()
    assert "no CPU fallback" in msg
    from hybridquantization_b200 import HqError, ImageManipulation

    with pytest.raises(HqError):
        ImageManipulation("CIE76", False, True)


def test_result_layout(hqlib):
    assert hqlib.hq_result_words(256, 0) == 257
    assert hqlib.hq_result_words(256, 1) == 1 + 4 * 256


def test_jni_shim_type_checks_against_the_c_abi():
    """No JDK in this image: java/jni/hq_jni.c is compiled with -fsyntax-only against tests/stubs/jni.h (JNI types and the four
    JNIEnv entries it uses) and the real include/hq_b200.h, so a drift between the shim and the C ABI is caught."""
    import subprocess
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = ["gcc", "-fsyntax-only", "-Wall", "-Werror", "-Wno-unused-parameter", "-I" + os.path.join(repo, "tests", "stubs"), "-I" + os.path.join(repo, "include"),
           os.path.join(repo, "java", "jni", "hq_jni.c")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the guard must have seen the stub, i.e. the shim's body was really compiled
    r2 = subprocess.run(cmd[:1] + ["-E", "-dM"] + cmd[3:], capture_output=True, text=True)
    assert "HQ_HAVE_JNI" in r2.stdout
