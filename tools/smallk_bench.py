#!/usr/bin/env python
"""Small-palette regime (K <= 32: the plugin's defaults are K = 8, population 4): the scoring kernel alone, device-timed with
CUDA events on the launching stream, L2 flushed between launches, 4K image resident in HBM; one candidate and 64 per launch.
Reports Gpixel/s and the fraction of max(12 B x N / HBM peak, 8 K flop x N x B / FP32 peak)."""
from __future__ import annotations

import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from hybridquantization_b200 import ImageManipulation, synth  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    be = ImageManipulation("CIE76", False, True, 0)
    peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    p_fp32 = max(be.measureFp32Peak().values())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    w, h = 3840, 2160
    n = w * h
    d_img = torch.from_numpy(synth.synth_image_rows(w, h, synth.SEED_BASE + 5, 0, h)).to(dev)
    be.setImageDevice(d_img.data_ptr(), w, h, stream=st.cuda_stream)
    out = {"tma": os.environ.get("HQ_V1_TMA", "1"), "hbm_gbs_peak": hbm, "fp32_tflops_peak": p_fp32, "rows": []}
    for K in (8, 16, 32):
        for B in (1, 4, 64):
            pal = torch.from_numpy(synth.synth_palettes(B, K)).to(dev)
            res = torch.zeros((B, be.resultWords(K, 0)), dtype=torch.int64, device=dev)
            fn = lambda: be.evalPalettesDevice(pal.data_ptr(), B, K, res.data_ptr(), 0, 0, st.cuda_stream)
            for _ in range(3):
                fn()
            be.setProfiling(True)
            ts = []
            for i in range(9):
                flush.fill_(i)
                fn()
                torch.cuda.synchronize()
                ts.append(be.lastAssignMs())
            be.setProfiling(False)
            assert int(res[0, 1:1 + K].sum().item()) == n
            ms = float(np.median(ts))
            floor = max(12.0 * n / (hbm * 1e9), 8.0 * K * n * B / (p_fp32 * 1e12)) * 1e3
            out["rows"].append({"K": K, "B": B, "kernel_us": 1e3 * ms, "gpixel_per_s": n * B / (ms * 1e-3) / 1e9, "frac_of_roofline": floor / ms,
                                "bound": "hbm" if 12.0 * n / (hbm * 1e9) > 8.0 * K * n * B / (p_fp32 * 1e12) else "fp32"})
    be.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
