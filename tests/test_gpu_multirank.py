"""Two ranks, two GPUs, NCCL: the sharded job on the real kernels.

Each rank runs the fused engine on its shard -- half of the times (noise counters keyed by the
global time index), or ONLY the baselines of the redundant groups the partitioner deals it
(``process(..., groups=mine)``) -- and ONE reduce of the
partial sums closes the job; rank 0 then holds the sums of the one-GPU job: equal up to the
order of the additions (fp32 partial sums per time chunk for time shards; only the fp64 atomics
for group shards, whose results are disjoint).  Skipped on boxes with fewer than two GPUs (the driver's one-GPU tier); the CPU suite
covers the same algebra with `gloo` (test_shard_gloo.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

GROUP_SIZES = [40, 9, 1, 17, 5]
P, T, N = 2, 6, 256
WINS = [(0, N - 1), (N, 2 * N - 1)]
KEYS = ("pow_all", "pow_auto", "noise_all", "noise_auto", "thermal_all", "thermal_auto")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _array():
    rng = np.random.Generator(np.random.PCG64(23))
    B = int(np.sum(GROUP_SIZES))
    F0 = 2 * N
    return dict(
        vis=(rng.standard_normal((P, B, T, F0)) + 1j * rng.standard_normal((P, B, T, F0))).astype(np.complex64),
        flags=rng.random((P, B, T, F0)) < 0.05,
        nsample=rng.integers(0, 61, (P, B, T, F0)).astype(np.float32),
        int_time=(10.7 * (1 + 0.02 * rng.random((B, T)))).astype(np.float32),
        group_offset=np.concatenate([[0], np.cumsum(GROUP_SIZES)]),
        freq=100e6 + 97656.25 * np.arange(F0), pols=np.array([-5, -6]),
    )


def _worker(rank, world, port, mode, out_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from simpleds_b200 import shard
        from simpleds_b200.engine import RedundantPipeline
        a = _array()
        dev = torch.device("cuda", rank)
        pipe = RedundantPipeline(a["freq"], WINS, a["pols"], a["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                                 precision="complex64", seed=99)
        if mode == "time":
            t0, t1 = shard.time_shards(T, world)[rank]
            sl = slice(t0, t1)
            pipe.process(torch.as_tensor(a["vis"][:, :, sl], device=dev), torch.as_tensor(a["flags"][:, :, sl], device=dev),
                         torch.as_tensor(a["nsample"][:, :, sl], device=dev), torch.as_tensor(a["int_time"][:, sl], device=dev),
                         time_offset=t0, ntime_total=T)
        else:
            # this rank holds and processes ONLY the baselines of the groups the partitioner deals it
            mine = shard.group_shards(GROUP_SIZES, world)[rank]
            go = a["group_offset"]
            bl = np.concatenate([np.arange(go[g], go[g + 1]) for g in mine])
            pipe.process(torch.as_tensor(a["vis"][:, bl], device=dev), torch.as_tensor(a["flags"][:, bl], device=dev),
                         torch.as_tensor(a["nsample"][:, bl], device=dev), torch.as_tensor(a["int_time"][bl], device=dev),
                         time_offset=0, ntime_total=T, groups=mine)
        pipe.reduce(dst=0)
        torch.cuda.synchronize()
        if rank == 0:
            out = {k: v.cpu() for k, v in pipe.sums().items()}
            out.update({"result_" + k: v.cpu() for k, v in pipe.result().items() if hasattr(v, "cpu")})
            torch.save(out, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["time", "group"])
def test_two_gpu_nccl_reduce_equals_the_one_gpu_job(cuda_device, tmp_path, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from simpleds_b200.engine import RedundantPipeline
    out_path = str(tmp_path / f"sums_{mode}.pt")
    mp.spawn(_worker, args=(2, _free_port(), mode, out_path), nprocs=2, join=True)
    got = torch.load(out_path)
    a = _array()
    dev = torch.device("cuda", 0)
    pipe = RedundantPipeline(a["freq"], WINS, a["pols"], a["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                             precision="complex64", seed=99)
    pipe.process(*(torch.as_tensor(a[k], device=dev) for k in ("vis", "flags", "nsample", "int_time")),
                 time_offset=0, ntime_total=T)
    torch.cuda.synchronize()
    for k, v in pipe.sums().items():
        want = v.cpu().numpy()
        if mode == "group" or k == "times_done":  # disjoint results; only the order of the fp64 atomic additions differs
            np.testing.assert_allclose(got[k].numpy(), want, rtol=1e-12, atol=0, err_msg=k)
        else:  # the same additions in another order (fp32 partial sums per time chunk, fp64 atomics)
            scale = np.abs(want).max(axis=-1, keepdims=True) if want.ndim == 4 else np.abs(want)
            assert np.all(np.abs(got[k].numpy() - want) <= 2e-6 * scale + 1e-300), k
    # the documented flow `pipe.reduce(dst=0); res = pipe.result()` gives the one-GPU means (the
    # per-group time counts travel with the sums)
    res = {k: v.cpu().numpy() for k, v in pipe.result().items() if hasattr(v, "cpu")}
    sums = {k: v.cpu().numpy() for k, v in pipe.sums().items()}
    n = np.asarray(GROUP_SIZES, dtype=np.float64)
    for base, ka, kb in (("power", "pow_all", "pow_auto"), ("noise_power", "noise_all", "noise_auto"),
                         ("thermal", "thermal_all", "thermal_auto")):
        a_all, a_auto = sums[ka], sums[kb]
        shp = (-1,) + (1,) * (a_all.ndim - 1)
        top = (lambda x: np.abs(x).max(axis=-1, keepdims=True)) if a_all.ndim == 4 else np.abs
        # the sums agree to 2e-6 of their row maximum (above); the means inherit that error
        tol_all = 4e-6 * top(a_all) / (n.reshape(shp) ** 2 * T)
        with np.errstate(divide="ignore", invalid="ignore"):
            tol_no = 4e-6 * (top(a_all) + top(a_auto)) / (n.reshape(shp) * (n.reshape(shp) - 1) * T)
        for key, tol in ((base + "_all", tol_all), (base + "_noautos", tol_no)):
            want = res[key]
            err = np.abs(got["result_" + key].numpy() - want)
            assert np.all((err <= tol + 1e-300) | np.isnan(want)), key
