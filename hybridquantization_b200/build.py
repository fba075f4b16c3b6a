"""In-tree build of the native code: libhq_b200.so (CUDA, sm_100a) and tools/microbench.

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with
the gpurun snapshot.  Nothing here builds or touches oracle/ (see __graft_entry__.build).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(ROOT)
CSRC = os.path.join(ROOT, "csrc")
LIB = os.path.join(ROOT, "libhq_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -ffp-contract=off: the host instantiation of csrc/hq_math.h must round after every fp32
# operation.  x86-64-v3 (AVX2 + FMA, no AVX-512): the library is built here and runs on the
# GPU box's CPU.
HOST_FLAGS = "-fPIC,-ffp-contract=off,-march=x86-64-v3,-fno-math-errno,-Wall"
# -fmad=false: no contraction of C++-level a*b+c; every fused multiply-add of the path is written explicitly (__fmaf_rn /
# fma.rn.f32x2).  It does NOT reach inline PTX: ptxas still fuses mul.rn.f32x2 + add.rn.f32x2 — see add2_of_product in
# csrc/hq_kernels.cu for how the packed pixel path prevents that.
NVCC_FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-fmad=false", "-Xcompiler", HOST_FLAGS]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return nvcc


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def sources() -> list[str]:
    src = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    src += [os.path.join(REPO, "include", f) for f in sorted(os.listdir(os.path.join(REPO, "include")))]
    return src


def build_library(force: bool = False, verbose: bool = False) -> str:
    cu = [os.path.join(CSRC, f) for f in ("hq_kernels.cu", "hq_pruned.cu", "hq_scielab.cu", "hq_api.cu")]
    if not force and not _stale(LIB, sources()):
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + cu
    print("[build]", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return LIB


def build_microbench(force: bool = False) -> str:
    """tools/microbench{,2,3}: FP32-pipe and issue-model probes (binaries are git-ignored, they ship with the gpurun snapshot)"""
    out = ""
    for name in ("microbench", "microbench2", "microbench3", "microbench4"):
        src = os.path.join(REPO, "tools", name + ".cu")
        out = os.path.join(REPO, "tools", name)
        if force or _stale(out, [src]):
            cmd = [_nvcc()] + ARCH + ["-O3", "-lineinfo", "-o", out, src]
            print("[build]", " ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_microbench(force="--force" in sys.argv)
