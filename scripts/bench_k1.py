#!/usr/bin/env python
"""Micro-benchmark of the standalone delay transform (sds_delay_transform, K1) against the HBM
roofline: complex64 rows in, complex64 spectra out, flag mask + taper + fftshift + delta_x
fused.  Algorithmic bytes = (8 + 1) in + 8 out per sample."""
import json
import os
import sys

import numpy as np
import torch
from scipy.signal import windows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpleds_b200 import ops  # noqa: E402
from simpleds_b200._lib import SDS_F32, SDS_F64  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    out = []
    for n in (64, 256, 1024, 2048, 203):
        for math, cdt, ebytes in ((SDS_F32, torch.complex64, 8), (SDS_F64, torch.complex128, 16)):
            rows = (1 << 28) // n if math == SDS_F32 else (1 << 27) // n  # 2 GiB of input either way
            x = torch.randn(rows, n, dtype=torch.float32, device=dev).to(cdt) + 1j
            fl = torch.rand(rows, n, device=dev) < 0.05
            plan = ops.get_plan(n, math, windows.blackmanharris)
            for _ in range(3):
                y = ops.delay_transform(plan, x, fl, 97656.25)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                y = ops.delay_transform(plan, x, fl, 97656.25)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            alg = rows * n * (ebytes + 1 + ebytes)
            rec = dict(nfreq=n, dtype=str(cdt).split(".")[-1], rows=rows, ms=ms, gbs=alg / ms / 1e6,
                       frac=alg / ms / 1e6 / peak)
            out.append(rec)
            print(json.dumps(rec), flush=True)
            del x, fl, y
            torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
