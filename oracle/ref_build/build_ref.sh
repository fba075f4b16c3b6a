#!/bin/bash
# build_ref.sh — compiles the reference's OWN sources for the CPU, from where they lie under
# /root/reference, into oracle/_ref/libhq_ref.so (git-ignored; travels to the GPU box).
# No reference text is written to disk: the translation unit is piped into g++.
#
#   OptimizedConvolution.cl   whole file, as C++ behind cl_shim.hpp.  One syntactic rewrite: the
#                             OpenCL vector literal `(float4)(a,b,c,d)` -> `float4(a,b,c,d)` (C++ would
#                             parse the former as a cast of a comma expression).  -DCIE76 as
#                             ImageManipulation.java:60-64 builds it.
#   SWASA.java                whole class, behind java_shim.hpp
#   ScielabProcessor.java     :20-23 (white points, minSAMPPERDEG), :44-56 (filter weights/half-widths, fields),
#                             :59-61 (6/29 powers), :185-254 (conv1D, resize1D, extractWithIndices, gauss),
#                             :279-311 (sRGBtoOpp, OpptoLab), and the constructor's filter-bank construction
#                             :78-148 + :154-178 (only the java.util.stream expression :149-153 is restated, in
#                             ref_filters_mid.inc)
#   ImageManipulation.java    :843-856 (argmin), :490-493 + :496-545 (initial population + annealing loop)
# Java rewrites are syntax only and live in java2cpp.py (array types -> JArr<T>, `new T[n]` -> JArr<T>(n),
# `this.` -> `this->`, `class` -> `struct`, access modifiers dropped, System.out.println(...) -> ;).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${HQ_REFERENCE:-/root/reference}/src/plugins/dbrasseur/hybridquantization"
OUT="$HERE/../_ref"
[ -d "$REF" ] || { echo "[build_ref] no reference tree at $REF — keeping any prebuilt $OUT/libhq_ref.so"; exit 0; }
mkdir -p "$OUT"
# the line ranges below are only valid for this exact snapshot
( cd "$REF" && sha256sum -c --quiet - <<'SUMS'
8ab6f1ad8beb2436f9b1289653ca2217a27d2c83838741304bcc780150faea7b  ImageManipulation.java
45b510929def59e2a34db9f7a24bdb1c01f0dc637907d05a41bf168dc297d37c  SWASA.java
0a26697ad3d4093642906972ef5abe16505d629a4b7d89eadce116f648129e58  ScielabProcessor.java
9b2c8a34bf2943a197d2cf733fd3be8e816b6a7566a95d258227b5ae95f2fd32  OptimizedConvolution.cl
SUMS
) || { echo "[build_ref] reference sources differ from the surveyed snapshot" >&2; exit 1; }

java2cpp() { python3 "$HERE/java2cpp.py"; }

{
    echo '#include "cl_shim.hpp"'
    echo '#include "java_shim.hpp"'
    echo 'namespace refcl {'
    sed 's/(float4)(/float4(/g' "$REF/OptimizedConvolution.cl"
    echo '}  // namespace refcl'
    java2cpp < "$REF/SWASA.java"
    echo ';'
    echo 'struct RefScielab {'
    sed -n '20,23p;44,56p;59,61p;185,254p;279,311p' "$REF/ScielabProcessor.java" | java2cpp
    cat "$HERE/ref_filters_head.inc"
    sed -n '78,148p' "$REF/ScielabProcessor.java" | java2cpp
    cat "$HERE/ref_filters_mid.inc"
    sed -n '154,178p' "$REF/ScielabProcessor.java" | java2cpp
    echo '    }'
    echo '};'
    cat "$HERE/ref_search_head.inc"
    sed -n '843,856p' "$REF/ImageManipulation.java" | java2cpp
    cat "$HERE/ref_search_mid.inc"
    sed -n '490,493p;496,545p' "$REF/ImageManipulation.java" | java2cpp
    cat "$HERE/ref_search_tail.inc"
    cat "$HERE/ref_entry.inc"
} | ${CXX:-g++} -x c++ -std=c++17 -O3 -march=x86-64-v3 -ffp-contract=off -fno-math-errno -fPIC -shared -DCIE76 \
        -Wno-unused-variable -Wno-unused-but-set-variable -I"$HERE" -o "$OUT/libhq_ref.so" - -lpthread -lm
echo "[build_ref] built $OUT/libhq_ref.so from $REF"

# The same kernels built as ImageManipulation.java:63 would for deltaETypes.CIE94 ("-DCIE94"): only CIEDE's branch cl:217-226 differs.
# Pins the oracle's restatement of that branch (hqo_delta_e94); the plugin itself never selects it (HybridQuantization.java:96,145).
{
    echo '#include "cl_shim.hpp"'
    echo 'namespace refcl {'
    sed 's/(float4)(/float4(/g' "$REF/OptimizedConvolution.cl"
    echo '}  // namespace refcl'
    cat "$HERE/ref_entry94.inc"
} | ${CXX:-g++} -x c++ -std=c++17 -O3 -march=x86-64-v3 -ffp-contract=off -fno-math-errno -fPIC -shared -DCIE94 \
        -Wno-unused-variable -Wno-unused-but-set-variable -Wno-unused-function -I"$HERE" -o "$OUT/libhq_ref94.so" - -lm
echo "[build_ref] built $OUT/libhq_ref94.so (CIEDE kernel, -DCIE94)"
