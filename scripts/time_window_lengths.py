"""HERA-350 three legs (complex64) with the 1024 channels cut into windows of 256 / 512 / 1024:
ms per 26 GB batch and fraction of the measured HBM peak per window length."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from simpleds_b200 import synth  # noqa: E402
from simpleds_b200.engine import RedundantPipeline  # noqa: E402

cfg = bench.CONFIGS["hera350"]
dev = torch.device("cuda:0")
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
lattice, basis = synth.hera_350()
red = synth.redundant_groups(lattice, basis)
freq = cfg["f0"] + cfg["df"] * np.arange(cfg["nchan"])
T = 8
data = synth.make_visibilities(red["group_offset"], 4, T, freq, seed=1, device=dev)
for n in (256, 512, 1024):
    wins = [(i, i + n - 1) for i in range(0, cfg["nchan"], n)]
    p = RedundantPipeline(freq, wins, cfg["pols"], red["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                          precision="complex64")
    for _ in range(3):
        p.process(*data, noise="philox", thermal=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        p.process(*data, noise="philox", thermal=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(n, round(ms, 3), "ms", round(p.algorithmic_bytes(T) / (ms * 1e-3) / 1e9 / peak, 3))
