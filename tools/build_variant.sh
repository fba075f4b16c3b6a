#!/bin/bash
# tools/build_variant.sh <name> [-DFLAG=V ...] — an A/B build of the SAME library with extra defines, as
# hybridquantization_b200/libhq_b200_<name>.so (git-ignored; select it with HQ_B200_LIB=<path>; there is still no fallback)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
src=hybridquantization_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-march=x86-64-v3,-fno-math-errno \
     "$@" -shared -o hybridquantization_b200/libhq_b200_$name.so $src/hq_kernels.cu $src/hq_pruned.cu $src/hq_scielab.cu $src/hq_bigk.cu $src/hq_api.cu $src/hq_multi.cu -ldl
echo built hybridquantization_b200/libhq_b200_$name.so
