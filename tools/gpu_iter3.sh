#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== sweep"; timeout 900 python tools/sweep.py > gpurun_out/sweep.json 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/sweep.json'))
for r in d['rgb_to_lab']: print('rgb2lab', r['w'], r['h'], round(r['kernel_ms'],4), 'ms', round(r['algorithmic_gbs']), 'GB/s', round(r['frac_of_hbm_peak'],3))
for r in d['k_sweep']: print(r['K'], r['B'], {k:(round(v['ms'],3), round(v['gpixel_per_s'],1), round(v['frac_of_roofline'],3), v['bound']) for k,v in r.items() if isinstance(v,dict)})
for r in d['scielab']: print(r)
for r in d.get('swasa',[]): print(r)
PY
