#!/bin/bash
# tools/sc_ab.sh [variant suffixes...] — tools/sc_bench.py (filter stage + assignment, 4K, K=256) for each build
# hybridquantization_b200/libhq_b200<suffix>.so ("default" = the default build)
for v in "$@"; do
  [ "$v" = "default" ] && v=""
  echo "== variant [$v]"
  HQ_B200_LIB=$PWD/hybridquantization_b200/libhq_b200$v.so timeout 300 python tools/sc_bench.py --modes 0 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['rows']: print(r['B'], {k:round(v['ms_per_candidate'],3) for k,v in r.items() if k!='B'})"
done
