#!/usr/bin/env python
"""A/B of builds of the prefilter scoring kernel (assign_reduce_kernel variant 3): device-timed kernel alone (CUDA events on
the launching stream, L2 flushed between launches), 4K image resident in HBM, K = 64 ... 1024, and a digest of every result
word so that two builds can be compared for identical integers.  In the same process the variant is also held against the
chunked direct-form kernel (variant 2, EVAL_FORCE_CHUNKED) on the same inputs, with and without per-colour sums.
Select the build with HQ_B200_LIB=<path to libhq_b200_<name>.so>."""
from __future__ import annotations

import hashlib
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from hybridquantization_b200 import ImageManipulation, synth  # noqa: E402
from hybridquantization_b200._lib import EVAL_FORCE_CHUNKED, EVAL_FORCE_PREFILTER, EVAL_SUMS  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    be = ImageManipulation("CIE76", False, True, 0)
    p_fp32 = max(be.measureFp32Peak().values())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    w, h = 3840, 2160
    n = w * h
    d_img = torch.from_numpy(synth.synth_image_rows(w, h, synth.SEED_BASE + 3, 0, h)).to(dev)
    be.setImageDevice(d_img.data_ptr(), w, h, stream=st.cuda_stream)
    out = {"lib": os.environ.get("HQ_B200_LIB", "default"), "fp32_tflops_peak": p_fp32, "rows": []}
    for K, B in ((64, 64), (128, 64), (256, 64), (512, 32), (1024, 16), (1024, 64), (72, 8), (40, 8)):
        pal = torch.from_numpy(synth.synth_palettes(B, K)).to(dev)
        res = torch.zeros((B, be.resultWords(K, 0)), dtype=torch.int64, device=dev)
        fn = lambda: be.evalPalettesDevice(pal.data_ptr(), B, K, res.data_ptr(), 0, EVAL_FORCE_PREFILTER, st.cuda_stream)
        for _ in range(3):
            fn()
        be.setProfiling(True)
        ts = []
        for i in range(7):
            flush.fill_(i)
            fn()
            torch.cuda.synchronize()
            ts.append(be.lastAssignMs())
        be.setProfiling(False)
        words = res.cpu().numpy()
        assert int(words[0, 1:1 + K].sum()) == n
        # parity inside the process: prefilter == chunked direct form, with sums, on a sub-batch
        Bs = min(B, 4)
        r3 = torch.zeros((Bs, be.resultWords(K, EVAL_SUMS)), dtype=torch.int64, device=dev)
        r2 = torch.zeros_like(r3)
        be.evalPalettesDevice(pal.data_ptr(), Bs, K, r3.data_ptr(), 0, EVAL_FORCE_PREFILTER | EVAL_SUMS, st.cuda_stream)
        be.evalPalettesDevice(pal.data_ptr(), Bs, K, r2.data_ptr(), 0, EVAL_FORCE_CHUNKED | EVAL_SUMS, st.cuda_stream)
        torch.cuda.synchronize()
        same = bool(torch.equal(r3, r2)) and bool(torch.equal(r3[:, :1 + K], res[:Bs]))
        ms = float(np.median(ts))
        out["rows"].append({"K": K, "B": B, "kernel_ms": ms, "gpixel_per_s": n * B / (ms * 1e-3) / 1e9,
                            "frac_fp32": 8.0 * K * n * B / (ms * 1e-3) / (p_fp32 * 1e12),
                            "digest": hashlib.sha256(words.tobytes()).hexdigest()[:16], "equals_direct_form": same})
    be.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
