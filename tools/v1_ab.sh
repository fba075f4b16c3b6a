#!/bin/bash
# tools/v1_ab.sh — A/B of the small-palette kernel's residency: 3 CTAs/SM (80 registers, default) against 4 (64 registers,
# tools/build_variant.sh v1c4 -DHQ_V1_MIN_CTAS=4): kernel-only throughput (tools/smallk_bench.py) and a default-parameter search
set -u
mkdir -p gpurun_out
for lib in "" hybridquantization_b200/libhq_b200_v1c4.so; do
    echo "== lib [${lib:-default}]"
    if [ -n "$lib" ]; then export HQ_B200_LIB=$PWD/$lib; else unset HQ_B200_LIB; fi
    timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1
    timeout 200 python tools/smallk_bench.py 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for r in d['rows']: print('  K',r['K'],'B',r['B'],round(r['kernel_us'],1),'us',round(r['gpixel_per_s'],1),'Gpx/s frac',round(r['frac_of_roofline'],3),r['bound'])"
    for k in 8 16; do timeout 100 python tools/latency_probe.py $k 4 3000 2>&1 | tail -1; done
done
