#!/usr/bin/env python
"""Micro-benchmarks of the SURVEY 8(f) rows against the HBM roofline (one B200):
f2 ingest gather (copy: 2 x element bytes per element), f3 weighted_average / fold_along_delay /
bootstrap gather at sizes where they are bandwidth-bound, f1 the cosmology evaluation (latency)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpleds_b200 import cosmo, ops  # noqa: E402


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0

    def report(name, ms, alg_bytes, **kw):
        rec = dict(row=name, ms=round(ms, 4), gbs=round(alg_bytes / ms / 1e6, 1), frac=round(alg_bytes / ms / 1e6 / peak, 3), **kw)
        print(json.dumps(rec), flush=True)

    # f2: HERA-350-like slab: 61075 baselines x 4 times, 1024 channels, 4 pols
    for dtype, eb in ((torch.complex64, 8), (torch.float32, 4), (torch.uint8, 1)):
        nblt, F, P = 61075 * 4, 1024, 4
        x = torch.empty((nblt, F, P), dtype=dtype, device=dev)
        x.view(torch.uint8).random_(0, 255)
        perm = torch.randperm(nblt, device=dev)
        ms = timed(lambda: ops.ingest_gather(x, perm))
        report("f2 ingest_gather", ms, 2 * x.numel() * eb, dtype=str(dtype).split(".")[-1], shape=[nblt, F, P])
        del x
        torch.cuda.empty_cache()
    # f3: spectra of 3601 groups x 4 spw x 4 pol x 64 bootstrap samples x 256 delays
    G, S, P, NB, D = 3601, 4, 4, 64, 256
    x = torch.randn((G, S, P, NB, D), dtype=torch.float64, device=dev)
    sig = torch.rand_like(x) + 0.5
    ms = timed(lambda: ops.weighted_average(x, sig, None, 3))
    report("f3 weighted_average (axis of 64, inner 256)", ms, 16 * x.numel(), shape=list(x.shape))
    ms = timed(lambda: ops.weighted_average(x, sig, None, 4))
    report("f3 weighted_average (last axis of 256)", ms, 16 * x.numel(), shape=list(x.shape))
    xc = torch.complex(x[..., :255], x[..., 1:]).contiguous()
    sc = torch.complex(sig[..., :255], sig[..., 1:]).contiguous()
    ms = timed(lambda: ops.fold_along_delay(xc, sc, None, 4, 127, 127, 128, True))
    report("f3 fold_along_delay (complex128, 255 -> 128 bins)", ms, 32 * xc.numel() + 32 * xc.numel() // 255 * 128, shape=list(xc.shape))
    del x, sig, sc
    src = xc[:, :, :, 0].contiguous()                     # (G, S, P, 255) complex128
    idx = ops.bootstrap_indices(G, 16, 1)
    ms = timed(lambda: ops.bootstrap_gather(src, idx, 0))
    report("f3 bootstrap_gather (groups axis, 16 resamples)", ms, 16 * src.numel() * 16 * 2, shape=list(src.shape))
    # f1: 4096 redshifts (E(z) and the distance integral)
    z = np.linspace(4.0, 30.0, 4096)
    t0 = time.perf_counter()
    for _ in range(5):
        cosmo.Planck15._evaluate(z, True, True)
    torch.cuda.synchronize()
    print(json.dumps(dict(row="f1 cosmo_evaluate, 4096 redshifts (incl. H2D/D2H)", ms=round((time.perf_counter() - t0) / 5 * 1e3, 3))))


if __name__ == "__main__":
    main()
