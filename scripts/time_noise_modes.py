"""HERA-350 three legs with the two in-kernel noise draws: the moment-matched discrete law (production)
and Box-Muller (noise="philox_gauss")."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from simpleds_b200 import synth  # noqa: E402
from simpleds_b200.engine import RedundantPipeline  # noqa: E402

cfg = bench.CONFIGS["hera350"]
dev = torch.device("cuda:0")
lattice, basis = synth.hera_350()
red = synth.redundant_groups(lattice, basis)
freq = cfg["f0"] + cfg["df"] * np.arange(cfg["nchan"])
wins = bench.config_windows(cfg)
data = synth.make_visibilities(red["group_offset"], 4, 8, freq, seed=1, device=dev)
p = RedundantPipeline(freq, wins, cfg["pols"], red["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                      precision="complex64")
for noise in ("philox", "philox_gauss"):
    for _ in range(3):
        p.process(*data, noise=noise, thermal=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        p.process(*data, noise=noise, thermal=True)
    e1.record()
    torch.cuda.synchronize()
    print(noise, round(e0.elapsed_time(e1) / 10, 3), "ms")
