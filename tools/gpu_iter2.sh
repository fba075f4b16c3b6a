#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_swasa.py -x -q -m gpu 2>&1 | tail -4
echo "== sweep quick"; timeout 600 python tools/sweep.py --quick --skip-swasa > gpurun_out/sweep_quick.json 2> gpurun_out/sweep_quick.err; tail -2 gpurun_out/sweep_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/sweep_quick.json'))
for r in d['rgb_to_lab']: print('rgb2lab', r['w'], r['h'], round(r['kernel_ms'],4), 'ms', round(r['algorithmic_gbs']), 'GB/s', round(r['frac_of_hbm_peak'],3))
for r in d['k_sweep']: print(r['K'], r['B'], {k:(round(v['ms'],3), round(v['frac_of_roofline'],3), v['bound']) for k,v in r.items() if isinstance(v,dict)})
PY
