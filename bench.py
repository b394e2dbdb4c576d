#!/usr/bin/env python
"""Benchmark of the delay-spectrum hot path (BASELINE.json metric: delay-spectrum
baseline-pairs/sec; achieved HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # reference CPU path (oracle port)

A step = one pass of the fused path over one HBM-resident batch of the synthetic
HERA-350 array (BASELINE.json configs[2]: 350 antennas -> 61 075 baselines in
3 601 redundant groups, 1024 channels in 4 spectral windows of 256, 4 pols,
Blackman-Harris taper): window select + flag mask + taper + delay transform of the
data and of an in-kernel noise realisation, pair sums and thermal sums, for every
group, of `--times` integrations.  The 1000-time job is streamed as such batches;
with N GPUs each rank takes its own times (weak scaling) and one NCCL reduce of
the partial sums closes the job.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "delay-spectrum baseline-pairs/sec"
UNIT = "pair-spectra/s"

CONFIGS = {
    # name: (array, nchan, nwin, npol, f0, df, t_int)
    "hera350": dict(array="hera_350", nchan=1024, nwin=4, npol=4, f0=100e6, df=97656.25, t_int=10.7,
                    pols=[-5, -6, -7, -8], seed=1003, ntimes=1000),
    # BASELINE.json configs[1]: 203 channels.  The fused path stages rows in 16-channel units, so the
    # row pitch is 208 channels and the ONE spectral window is channels 0..202 (nfreq = 203)
    "paper64": dict(array="paper_grid", nchan=208, nwin=1, npol=2, f0=100e6, df=492610.84375, t_int=42.95,
                    pols=[1, 2], seed=1002, windows=[(0, 202)], ntimes=100),
    "ska512": dict(array="ska_512", nchan=2048, nwin=1, npol=2, f0=50e6, df=146484.375, t_int=0.85,
                   pols=[-5, -6], seed=1005, ntimes=64),
}


def config_windows(cfg):
    """Inclusive (start, end) channel ranges of the configuration's spectral windows."""
    if "windows" in cfg:
        return [tuple(w) for w in cfg["windows"]]
    n = cfg["nchan"] // cfg["nwin"]
    return [(w * n, (w + 1) * n - 1) for w in range(cfg["nwin"])]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- CPU arm
def _cpu_sample_inputs(cfg, n, ntime, seed=0):
    rng = np.random.Generator(np.random.PCG64(seed))
    P, F = cfg["npol"], cfg["nchan"]
    shape = (1, 1, P, n, ntime, F)
    common = rng.standard_normal((1, 1, P, 1, ntime, F)) + 1j * rng.standard_normal((1, 1, P, 1, ntime, F))
    vis = (30 * common + rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    flags = rng.random(shape) < 0.06
    ns = rng.integers(32, 61, shape).astype(np.float64)
    freq = (cfg["f0"] + cfg["df"] * np.arange(F))[None]
    wins = config_windows(cfg)
    return dict(vis=vis, flags=flags, nsample=ns, freq_array_hz=freq, trcvr_k=np.full((1, F), 144.0),
                beam_area_sr=np.full((1, P, F), 0.76), integration_time_s=np.full((n, ntime), cfg["t_int"]),
                polarization_array=np.array(cfg["pols"]), spectral_windows=wins)


def _cpu_chunk(args):
    """One time-chunk of the reference path (the reference's tests show time slices
    commute, test_io.py:570-631): windows -> noise -> transform -> cross multiply ->
    thermal -> mean over pairs and time, all via the oracle restatement."""
    from oracle import simpleds_oracle as orc

    pr, sl, seed = args
    q = dict(pr)
    for k in ("vis", "flags", "nsample"):
        q[k] = pr[k][..., sl, :]
    q["integration_time_s"] = pr["integration_time_s"][:, sl]
    res = orc.pipeline_reference(**q, noise_seed=seed)
    out = [orc.average_power(res[k], remove_autos=True) for k in ("power_array", "noise_power")]
    out.append(res["thermal_power"].mean(axis=(2, 3, 4)))
    return [np.asarray(o) for o in out]


def cpu_reference_rate(cfg, n, ntime, procs):
    """pair-spectra/s of the reference CPU path on one group of n baselines."""
    pr = _cpu_sample_inputs(cfg, n, ntime)
    chunks = np.array_split(np.arange(ntime), max(1, min(procs, ntime)))
    jobs = [(pr, slice(int(c[0]), int(c[-1]) + 1), i) for i, c in enumerate(chunks) if len(c)]
    t0 = time.perf_counter()
    if procs > 1:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(len(jobs)) as pool:
            pool.map(_cpu_chunk, jobs)
    else:
        for j in jobs:
            _cpu_chunk(j)
    dt = time.perf_counter() - t0
    pairs = n * n * len(config_windows(cfg)) * cfg["npol"] * ntime
    return pairs / dt, dt


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n, ntime = args.cpu_bl, max(cores, args.cpu_times)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_rate(cfg, n, ntime, cores)
    rates, times = [], []
    for _ in range(max(1, min(args.steps, 8))):
        r, dt = cpu_reference_rate(cfg, n, ntime, cores)
        rates.append(r)
        times.append(dt)
    wins = config_windows(cfg)
    value = float(np.sum([n * n * len(wins) * cfg["npol"] * ntime] * len(times)) / np.sum(times))
    sample = (f"one redundant group of {n} baselines x {cfg['npol']} pols x {len(wins)} windows of "
              f"{wins[0][1] - wins[0][0] + 1} ch x {ntime} times per step, split over {cores} processes by time; "
              "rate per pair-spectrum extrapolated to the whole array (quadratic in the group size); "
              "oracle port of the reference's numpy path incl. noise draw, thermal estimate and the mean")
    line = dict(
        impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=len(times),
        warmup=min(args.warmup, 1), ms_per_step=1e3 * float(np.mean(times)), higher_is_better=True,
        scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
        config=workload_config(args, cfg, None),
        cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
        e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
    )
    print(json.dumps(line), flush=True)


def measured_traffic(args, T):
    """dram__bytes_read + dram__bytes_write of the fused kernel per launch from the committed
    ncu capture of this configuration (profiles/traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            for rec in json.load(f):
                if (rec["config"], rec["precision"], rec["times"], rec["noise"], rec["thermal"]) == \
                        (args.config, args.precision, T, not args.no_noise, not args.no_thermal):
                    return rec["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    return None


def workload_config(args, cfg, pipe):
    c = dict(
        workload=f"{args.config}: synthetic {cfg['array']} redundant array, {cfg['nchan']} channels in "
                 f"{cfg['nwin']} spectral windows, {cfg['npol']} pols, Blackman-Harris taper, "
                 f"{args.times} times per step per GPU (streamed batches of the {cfg['ntimes']}-time job), "
                 "delay transform + noise realisation + redundant pair average + thermal estimate",
        times_per_step_per_gpu=args.times, precision=args.precision,
        l2_policy="inputs larger than L2 (no flush needed)",
    )
    if pipe is not None:
        c.update(nbls=pipe.nbl, ngroups=pipe.ngroup, nfreq=pipe.nfreq, nspw=pipe.nspw)
    return c


# --------------------------------------------------------------------------- GPU arm
class ClockSampler:
    """nvidia-smi polled in the background; samples are attributed to the timed region by
    line count (mark() before and after)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.proc = index, None
        self.path = tempfile.mktemp(prefix="sds_clocks_", suffix=".csv")

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
            return
        t0 = time.time()
        while self.mark() == 0 and time.time() - t0 < 5.0:  # first sample can take a second
            time.sleep(0.05)

    def _lines(self):
        try:
            with open(self.path) as f:
                data = f.read()
        except OSError:
            return []
        lines = data.split("\n")
        return lines[:-1]  # the last element is an incomplete (or empty) line

    def mark(self):
        return len(self._lines())

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()

    def summary(self, lo, hi):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        sm, smax, power, reasons = [], [], [], set()
        for ln in self._lines()[lo:hi]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(smax)), samples=len(sm),
                       power_w_max=float(np.max(power)))
        out["reasons"] = sorted(reasons)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return out


def measure_mode(name, precision, times, steps, dev, peak):
    """One more BASELINE.json configuration measured in the same run (few steps): the fused path
    on a resident batch of `times` integrations, all three legs, device-timed like the headline."""
    import torch

    from simpleds_b200 import synth
    from simpleds_b200.engine import RedundantPipeline

    cfg = CONFIGS[name]
    lattice, basis = getattr(synth, cfg["array"])()
    red = synth.redundant_groups(lattice, basis)
    freq = cfg["f0"] + cfg["df"] * np.arange(cfg["nchan"])
    wins = config_windows(cfg)
    pipe = RedundantPipeline(freq, wins, cfg["pols"], red["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                             precision=precision, seed=0)
    data = synth.make_visibilities(red["group_offset"], cfg["npol"], times, freq, seed=cfg["seed"], device=dev,
                                   t_int=cfg["t_int"])
    out = time_pipeline(pipe, data, times, steps, peak)
    out.update(config=name, precision=precision, nfreq=pipe.nfreq, nbls=pipe.nbl, ngroups=pipe.ngroup,
               fused=bool(pipe.fused), times_per_step=times)
    del data, pipe
    torch.cuda.empty_cache()
    return out


def time_pipeline(pipe, data, times, steps, peak, warmup=3):
    import torch

    vis, flags, nsample, int_time = data
    for _ in range(warmup):
        pipe.process(vis, flags, nsample, int_time, ntime_total=1000)
    torch.cuda.synchronize()
    l0 = pipe.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        pipe.process(vis, flags, nsample, int_time, ntime_total=1000)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    alg = pipe.algorithmic_bytes(times)
    return dict(kernel_ms=ms, bytes_per_launch=alg, achieved=alg / (ms * 1e-3) / 1e9,
                frac=alg / (ms * 1e-3) / 1e9 / peak, value=pipe.pair_spectra(times) / (ms * 1e-3), unit=UNIT,
                steps=steps, launches_per_step=(pipe.kernel_launches - l0) // steps)


def h2d_ceiling(dev, world):
    """Pinned-host -> device copy rate of this rank while every rank copies at once (GB/s): the
    ceiling of the end-to-end number, a property of the box's PCIe / host-memory fabric."""
    import torch
    import torch.distributed as dist

    n = 1 << 30
    src = torch.empty(n, dtype=torch.uint8).pin_memory()
    dst = torch.empty(n, dtype=torch.uint8, device=dev)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    gbs = 3 * n / (time.perf_counter() - t0) / 1e9
    if world > 1:
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        gbs = float(t.item())
    del src, dst
    return gbs


def run_gpu_arm(args, cfg):
    import torch
    import torch.distributed as dist

    from simpleds_b200 import synth
    from simpleds_b200.engine import RedundantPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    lattice, basis = getattr(synth, cfg["array"])()
    red = synth.redundant_groups(lattice, basis)
    F = cfg["nchan"]
    freq = cfg["f0"] + cfg["df"] * np.arange(F)
    wins = config_windows(cfg)
    N = wins[0][1] - wins[0][0] + 1
    pipe = RedundantPipeline(freq, wins, cfg["pols"], red["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                             precision=args.precision, seed=0)
    T = args.times
    ntime_total = max(cfg["ntimes"], T)  # the configuration's whole job (BASELINE.json configs)
    # this rank's resident batch (generation is not timed)
    vis, flags, nsample, int_time = synth.make_visibilities(
        red["group_offset"], cfg["npol"], T, freq, seed=cfg["seed"], device=dev, t_int=cfg["t_int"],
        time_offset=rank * T)
    torch.cuda.synchronize()

    def step(i):
        # step i of this rank covers global times [(i*world + rank) * T, ...)
        toff = ((i * world + rank) * T) % max(ntime_total - T + 1, 1)
        pipe.process(vis, flags, nsample, int_time, time_offset=toff, ntime_total=ntime_total,
                     noise=None if args.no_noise else "philox", thermal=not args.no_thermal)

    def job_reduce():
        # the path's single exchange step: one NCCL reduce of the flat fp64 partial sums
        pipe.reduce(dst=0)

    for i in range(args.warmup):
        step(i)
    job_reduce()
    torch.cuda.synchronize()
    pipe.reset()
    launches0 = pipe.kernel_launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    mark0 = sampler.mark() if rank == 0 else 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    ev[0].record()
    for i in range(args.steps):
        ev[1 + 2 * i].record()
        step(i)
        ev[2 + 2 * i].record()
    job_reduce()
    ev[-1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = pipe.kernel_launches - launches0  # kernels launched inside the timed region
    clocks = None
    if rank == 0:
        mark1 = sampler.mark()
        region = "timed region"
        if mark1 - mark0 < 3 and sampler.proc is not None:
            # a timed region shorter than a few nvidia-smi periods: keep the GPU on the same
            # work (untimed) until three samples under load exist
            keep = {k: v.clone() for k, v in pipe.sums().items()}
            t_probe = time.time()
            while sampler.mark() - mark0 < 3 and time.time() - t_probe < 3.0:
                step(0)
                torch.cuda.synchronize()
            for k, v in pipe.sums().items():
                v.copy_(keep[k])
            pipe.ntimes_done = args.steps * T
            mark1 = sampler.mark()
            region = "timed region + identical untimed steps right after it (region shorter than 3 samples)"
        sampler.stop()
        clocks = sampler.summary(mark0, mark1)
        clocks["sampled_over"] = region
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[1 + 2 * i].elapsed_time(ev[2 + 2 * i]) for i in range(args.steps)]
    if world > 1:
        tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_ms = float(tmax.item())
    pairs_per_step = pipe.pair_spectra(T)
    value = pairs_per_step * args.steps * world / (total_ms * 1e-3)

    # roofline of the dominant kernel (the fused engine launch; the tiny setup launch that
    # precedes it is inside the same event pair)
    peak, peak_src = peaks()
    alg_bytes = pipe.algorithmic_bytes(T)
    kern_ms = float(np.mean(step_ms))
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                    traffic=measured_traffic(args, T), peak_source=peak_src, bytes_per_launch=alg_bytes,
                    kernel_ms=kern_ms, kernel="sds::engine_f32_kernel (+ engine_setup_kernel)"
                    if args.precision == "complex64" else "sds::engine_kernel<double, N, data> + sds::engine_f32_kernel<N, noise | thermal64> (+ setup each)")

    # per-leg breakdown of the fused kernel (explains the roofline fraction): same batch,
    # noise and/or thermal legs switched off
    if world == 1 and not args.no_breakdown and not (args.no_noise or args.no_thermal):
        legs = {}
        for name, kw in (("transform+pair_sums", dict(noise=None, thermal=False)),
                         ("transform+pair_sums+thermal", dict(noise=None, thermal=True)),
                         ("transform+pair_sums+noise", dict(noise="philox", thermal=False))):
            for _ in range(3):
                pipe.process(vis, flags, nsample, int_time, ntime_total=ntime_total, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                pipe.process(vis, flags, nsample, int_time, ntime_total=ntime_total, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            legs[name] = dict(kernel_ms=ms, achieved=alg_bytes / (ms * 1e-3) / 1e9,
                              frac=alg_bytes / (ms * 1e-3) / 1e9 / peak)
        roofline["legs"] = legs
        pipe.reset()
        for i in range(args.steps):  # restore the accumulators of the timed job for the sanity check
            step(i)
        torch.cuda.synchronize()

    # the configuration's fixed job (HERA-350: 1000 times, BASELINE.json configs[2] / [3]): its batches of T times dealt to the
    # ranks, one reduce at the end; at N = 1 this is the sustained number of the whole job, over N
    # it is the STRONG scaling of the path (the headline `value` is weak scaling).  Clocks sampled.
    fixed_job = None
    if not args.no_fixed_job and not (args.no_noise or args.no_thermal):
        keep = {k: v.clone() for k, v in pipe.sums().items()}
        pipe.reset()
        nbatch = max(ntime_total // T, 1)
        mine = list(range(rank, nbatch, world))
        sampler2 = ClockSampler(local)
        if rank == 0:
            sampler2.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        m0 = sampler2.mark() if rank == 0 else 0
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for k in mine:
            pipe.process(vis, flags, nsample, int_time, time_offset=k * T, ntime_total=ntime_total,
                         noise="philox", thermal=True)
        job_reduce()
        f1.record()
        torch.cuda.synchronize()
        job_ms = f0.elapsed_time(f1)
        if world > 1:
            tmax = torch.tensor([job_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            job_ms = float(tmax.item())
        if rank == 0:
            m1 = sampler2.mark()
            sampler2.stop()
            fixed_job = dict(times=nbatch * T, batches=nbatch, batches_on_rank0=len(mine), seconds=job_ms * 1e-3,
                             value=pipe.pair_spectra(nbatch * T) / (job_ms * 1e-3), unit=UNIT,
                             frac_of_hbm_peak_per_gpu=pipe.algorithmic_bytes(T) * len(mine) / (job_ms * 1e-3) / 1e9 / peak,
                             scaling="strong", clocks=sampler2.summary(m0, m1))
        for k, v in pipe.sums().items():
            v.copy_(keep[k])

    # N > 1: the reduced sums of a small sharded job against rank 0 doing the same job alone
    multi_gpu_check = None
    if world > 1 and not args.no_check:
        chk = RedundantPipeline(freq, wins, cfg["pols"], red["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                                precision=args.precision, seed=3)

        def one(r, pl):
            d = synth.make_visibilities(red["group_offset"], cfg["npol"], 1, freq, seed=cfg["seed"] + 17,
                                        device=dev, t_int=cfg["t_int"], time_offset=2000 + r)
            pl.process(*d, time_offset=r, ntime_total=world, noise="philox", thermal=True)
            del d

        one(rank, chk)
        chk.reduce(dst=0)
        torch.cuda.synchronize()
        if rank == 0:
            alone = RedundantPipeline(freq, wins, cfg["pols"], red["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                                      precision=args.precision, seed=3)
            for r in range(world):
                one(r, alone)
            torch.cuda.synchronize()
            worst = 0.0
            for k, v in alone.sums().items():
                ref = v.abs().amax(dim=-1, keepdim=True) if v.ndim == 4 else v.abs()
                err = ((chk.sums()[k] - v).abs() / ref.clamp_min(1e-300)).max()
                worst = max(worst, float(err.item()))
            res_c, res_a = chk.result(), alone.result()
            rel = float(((res_c["power_all"] - res_a["power_all"]).abs().amax(dim=-1)
                         / res_a["power_all"].abs().amax(dim=-1)).max().item())
            multi_gpu_check = dict(max_rel_err=worst, result_power_all_max_rel_err=rel,
                                   what=f"{world} ranks x 1 time each (whole array), NCCL reduce to rank 0, against "
                                        "rank 0 running the same times alone; error relative to each row's maximum")
            del alone
        del chk
        dist.barrier()
        torch.cuda.empty_cache()

    # end to end through the public API with pinned HOST buffers (H2D + D2H timed), every rank
    e2e = None
    if not args.no_e2e:
        res_keep = {k: v.clone() for k, v in pipe.sums().items()}
        e2e = run_e2e(args, cfg, pipe, red, freq, dev, world, rank)
        for k, v in pipe.sums().items():
            v.copy_(res_keep[k])
        pipe.ntimes_done = args.steps * T

    # the other BASELINE.json configurations, same run, few steps each (rank 0 of a 1-GPU run)
    modes = None
    if world == 1 and not args.no_modes and args.config == "hera350" and args.precision == "complex64":
        modes = {}
        p128 = RedundantPipeline(freq, wins, cfg["pols"], red["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                                 precision="complex128", seed=0)
        m = time_pipeline(p128, (vis, flags, nsample, int_time), T, 5, peak)
        m.update(config="hera350", precision="complex128", nfreq=p128.nfreq, nbls=p128.nbl, ngroups=p128.ngroup,
                 fused=bool(p128.fused), times_per_step=T)
        modes["hera350_complex128"] = m
        del p128
        # two data sets (Nuv = 2): the batch cross-multiplied with a copy shifted by one baseline
        T2 = max(T // 2, 1)
        p2 = RedundantPipeline(freq, wins, cfg["pols"], red["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                               precision="complex64", seed=0, nuv=2)
        two = [torch.stack([x[:, :, :T2], torch.roll(x[:, :, :T2], 1, dims=1)]) for x in (vis, flags, nsample)]
        m = time_pipeline(p2, (*two, int_time[:, :T2].contiguous()), T2, 5, peak)
        m.update(config="hera350", precision="complex64", nuv=2, nfreq=p2.nfreq, nbls=p2.nbl, ngroups=p2.ngroup,
                 fused=bool(p2.fused), times_per_step=T2,
                 note="26 B per sample pair algorithmic; the data launch reads 18 B, the noise + thermal launch 10 B")
        modes["hera350_nuv2_complex64"] = m
        del p2, two
        torch.cuda.empty_cache()

    line = None
    if rank == 0:
        res = pipe.result(ntimes=args.steps * T)
        finite = bool(torch.isfinite(res["power_all"].real).all().item())
        cpu = None
        if world == 1 and not args.no_cpu:
            r, dt = cpu_reference_rate(cfg, args.cpu_bl, args.cpu_times, 1)
            cpu = dict(value=r, unit=UNIT, cores=1, kind="port", seconds=dt,
                       sample=f"one redundant group of {args.cpu_bl} baselines x {cfg['npol']} pols x "
                              f"{cfg['nwin']} windows of {N} ch x {args.cpu_times} times, single process "
                              "(the reference's numpy path is single-threaded); oracle port incl. noise "
                              "draw, thermal estimate and the mean over pairs/time")
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="f32" if args.precision == "complex64" else "f64", data="synthetic",
            config=workload_config(args, cfg, pipe), roofline=roofline, cpu_baseline=cpu, e2e=e2e,
            gpu_launches=int(launches), clocks=clocks, outputs_finite=finite,
            input_gb_per_step_per_gpu=alg_bytes / 1e9, fixed_job=fixed_job, multi_gpu_check=multi_gpu_check,
        )
        if modes is not None:
            del vis, flags, nsample, int_time, pipe
            torch.cuda.empty_cache()
            modes["ska512_complex64"] = measure_mode("ska512", "complex64", 2, 5, dev, peak)
            modes["ska512_complex128"] = measure_mode("ska512", "complex128", 2, 3, dev, peak)
            modes["paper64_203ch_complex64"] = measure_mode("paper64", "complex64", 100, 10, dev, peak)
            line["roofline"]["modes"] = modes
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def run_e2e(args, cfg, pipe, red, freq, dev, world=1, rank=0):
    """Same metric through RedundantPipeline.process_host: pinned host arrays in, host
    results out, every step; every rank runs its own batch, time = max over ranks."""
    import torch
    import torch.distributed as dist

    from simpleds_b200 import synth

    Te = max(1, min(args.e2e_times, args.times))
    vis, flags, nsample, int_time = synth.make_visibilities(
        red["group_offset"], cfg["npol"], Te, freq, seed=cfg["seed"] + 1, device=dev, t_int=cfg["t_int"],
        time_offset=rank * Te)
    host = [t.cpu().pin_memory() for t in (vis, flags, nsample, int_time)]
    del vis, flags, nsample, int_time
    torch.cuda.empty_cache()
    h2d = sum(t.numel() * t.element_size() for t in host)
    steps = max(1, min(args.steps, args.e2e_steps))
    out = None
    for _ in range(2):
        pipe.reset()
        out = pipe.process_host(*host, noise=None if args.no_noise else "philox", thermal=not args.no_thermal)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ceiling = h2d_ceiling(dev, world)
    pipe.reset()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # a streamed job: every step uploads its batch from pinned host memory and runs it; the job's
    # result (the per-group means) is read back into pinned host memory ONCE, at its end
    t0 = time.perf_counter()
    for i in range(steps):
        pipe.process_host(*host, time_offset=i * Te, ntime_total=steps * Te,
                          noise=None if args.no_noise else "philox", thermal=not args.no_thermal, fetch=False)
    out = pipe.result_host()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    if world > 1:
        tmax = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dt = float(tmax.item())
    d2h = sum(v.nbytes for v in out.values() if hasattr(v, "nbytes"))
    return dict(value=world * pipe.pair_spectra(Te) / dt, unit=UNIT, h2d_bytes_per_step=int(h2d),
                d2h_bytes_per_step=int(d2h // steps), d2h_bytes_per_job=int(d2h), ms_per_step=1e3 * dt,
                times_per_step_per_gpu=Te, steps=steps,
                h2d_gbs_per_gpu=h2d / dt / 1e9, h2d_ceiling_gbs_per_gpu=ceiling,
                frac_of_h2d_ceiling=h2d / dt / 1e9 / ceiling,
                note="a streamed job: every step uploads its batch from pinned host buffers (chunked H2D on a "
                     "copy stream overlapped with the fused engine); the per-group means are read back into "
                     "pinned buffers once, at the end of the job, inside the timed region; the ceiling is the "
                     "pinned H2D copy rate of a rank measured while all ranks copy at once (PCIe / host fabric)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="hera350", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="complex64", choices=["complex64", "complex128"])
    ap.add_argument("--times", type=int, default=8, help="integrations per step per GPU")
    ap.add_argument("--e2e-times", type=int, default=2)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-bl", type=int, default=64)
    ap.add_argument("--cpu-times", type=int, default=16)
    ap.add_argument("--no-noise", action="store_true")
    ap.add_argument("--no-thermal", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the other BASELINE configurations (roofline.modes)")
    ap.add_argument("--no-fixed-job", action="store_true", help="skip the fixed 1000-time job (sustained / strong scaling)")
    ap.add_argument("--no-check", action="store_true", help="N > 1: skip the check of the reduced sums against one rank")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        if args.warmup < 3:
            args.warmup = 3  # timing hygiene: at least 3 warm-up steps
        run_gpu_arm(args, cfg)


if __name__ == "__main__":
    main()
