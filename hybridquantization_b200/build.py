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


CU_FILES = ("hq_kernels.cu", "hq_pruned.cu", "hq_scielab.cu", "hq_bigk.cu", "hq_api.cu", "hq_multi.cu")


def build_library(force: bool = False, verbose: bool = False) -> str:
    """One object per .cu (compiled in parallel, only when stale against its own source and the shared headers), then one link."""
    from concurrent.futures import ThreadPoolExecutor

    if not force and not _stale(LIB, sources()):
        return LIB  # (the objects stay behind on this machine; the .so alone travels to the GPU box)
    headers = [s for s in sources() if not s.endswith(".cu")]
    objdir = os.path.join(ROOT, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for f in CU_FILES:
        src, obj = os.path.join(CSRC, f), os.path.join(objdir, f[:-3] + ".o")
        if force or _stale(obj, [src] + headers):
            jobs.append([_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])
    def run(cmd):
        print("[build]", " ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(objdir, f[:-3] + ".o") for f in CU_FILES]
    # libdl: NCCL is resolved at run time with dlopen (hq_multi.cu), so the library loads on machines without it
    run([_nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"])
    return LIB


def build_microbench(force: bool = False) -> str:
    """tools/microbench{,2,3}: FP32-pipe and issue-model probes (binaries are git-ignored, they ship with the gpurun snapshot)"""
    out = ""
    for name in ("microbench", "microbench2", "microbench3", "microbench4", "microbench5"):
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
