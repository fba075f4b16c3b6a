#!/bin/bash
# tools/v3_ab.sh [variant suffixes...] — tools/v3_ab.py for each build hybridquantization_b200/libhq_b200<suffix>.so ("" = the default build)
mkdir -p gpurun_out
for v in "$@"; do
  [ "$v" = "default" ] && v=""
  echo "== build [$v]"
  HQ_B200_LIB=$PWD/hybridquantization_b200/libhq_b200$v.so timeout 300 python tools/v3_ab.py 2> gpurun_out/v3_ab$v.err | tee gpurun_out/v3_ab$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['rows']: print(r['K'], r['B'], round(r['kernel_ms'],3), 'ms', round(r['gpixel_per_s'],2), 'Gpx/s', round(r['frac_fp32'],4), r['digest'], r['equals_direct_form'])"
  tail -2 gpurun_out/v3_ab$v.err
done
