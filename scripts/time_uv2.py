import sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
from simpleds_b200 import synth
from simpleds_b200.engine import RedundantPipeline
cfg = bench.CONFIGS["hera350"]; dev = torch.device("cuda:0")
lattice, basis = synth.hera_350(); red = synth.redundant_groups(lattice, basis)
freq = cfg["f0"] + cfg["df"] * np.arange(cfg["nchan"]); wins = bench.config_windows(cfg)
T2 = 4
vis, flags, ns, it = synth.make_visibilities(red["group_offset"], 4, T2, freq, seed=1, device=dev)
two = [torch.stack([x, torch.roll(x, 1, dims=1)]) for x in (vis, flags, ns)]
p2 = RedundantPipeline(freq, wins, cfg["pols"], red["group_offset"], trcvr_k=144.0, beam_area_sr=0.76, precision="complex64", nuv=2)
for kw in (dict(noise=None, thermal=False), dict(noise="philox", thermal=False), dict(noise="philox", thermal=True)):
    for _ in range(2): p2.process(*two, it, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): p2.process(*two, it, **kw)
    e1.record(); torch.cuda.synchronize()
    print(kw, e0.elapsed_time(e1) / 5, "ms")
