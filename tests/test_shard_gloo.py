"""Multi-rank host logic on CPU: world_size-2 `gloo` runs of the partitioner and of the
single reduce of partial sums (simpleds_b200/shard.py).  The per-rank partial sums come
from the oracle (engine_ref), so these tests pin the ALGEBRA of the sharded job --
time shards with globally keyed noise counters, group shards as disjoint results --
without a GPU; the CUDA kernels are checked against the same oracle in test_gpu_engine.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import engine_ref
from simpleds_b200 import shard

GROUP_SIZES = [3, 1, 4, 2]
P, T, N = 2, 5, 16
WINS = [(0, N - 1), (N, 2 * N - 1)]
SEED = 0xC0FFEE


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _array():
    rng = np.random.Generator(np.random.PCG64(11))
    B = int(np.sum(GROUP_SIZES))
    go = np.concatenate([[0], np.cumsum(GROUP_SIZES)])
    F0 = 2 * N
    vis = (rng.standard_normal((P, B, T, F0)) + 1j * rng.standard_normal((P, B, T, F0))).astype(np.complex64)
    flags = rng.random((P, B, T, F0)) < 0.1
    ns = rng.integers(32, 61, (P, B, T, F0)).astype(np.float32)
    it = (10.7 * (1 + 0.02 * rng.random((B, T)))).astype(np.float32)
    freq = 100e6 + 97656.25 * np.arange(F0)
    return dict(vis=vis, flags=flags, nsample=ns, int_time=it, group_offset=go, freq=freq,
                pols=np.array([-5, -6]))


def _sigma_coef32(a):
    from simpleds_b200.engine import radiometer_tables
    from scipy.signal import windows

    chans = np.asarray([w[0] for w in WINS])[:, None] + np.arange(N)
    fw = a["freq"][chans]
    df = float(fw[0, 1] - fw[0, 0])
    sc, _, _ = radiometer_tables(fw, np.full((len(WINS), N), 144.0), np.full((len(WINS), P, N), 0.76),
                                 a["pols"], windows.blackmanharris(N), df)
    return sc.astype(np.float32)


def _partial_sums(a, tsl, groups=None):
    """Oracle partial sums over the times `tsl` (slice) for the groups `groups` (all if None)."""
    B = a["vis"].shape[1]
    t0 = tsl.start
    sub = {k: a[k][:, :, tsl] for k in ("vis", "flags", "nsample")}
    it = a["int_time"][:, tsl]
    nt = it.shape[1]
    G, S = len(GROUP_SIZES), len(WINS)
    sums = dict(pow_all=torch.zeros(G, S, P, N, dtype=torch.float64), pow_auto=torch.zeros(G, S, P, N, dtype=torch.float64),
                noise_all=torch.zeros(G, S, P, N, dtype=torch.float64), noise_auto=torch.zeros(G, S, P, N, dtype=torch.float64),
                thermal_all=torch.zeros(G, S, P, dtype=torch.float64), thermal_auto=torch.zeros(G, S, P, dtype=torch.float64))
    if nt == 0:
        return sums
    noise = engine_ref.engine_noise_freq(_sigma_coef32(a), sub["nsample"], it, [w[0] for w in WINS], N, SEED,
                                         time_offset=t0, ntime_total=T, bl_offset=0, nbl_total=B)
    ref = engine_ref.group_means(sub["vis"], sub["flags"], sub["nsample"], it, a["group_offset"], a["freq"],
                                 WINS, a["pols"], 144.0, 0.76, noise_freq=noise)
    for g, n in enumerate(GROUP_SIZES):
        if groups is not None and g not in groups:
            continue
        for key, out in (("power", "pow"), ("noise_power", "noise")):
            s_all = np.real(ref[key + "_all"][g]) * n * n * nt
            s_auto = s_all - (np.real(ref[key + "_noautos"][g]) * n * (n - 1) * nt if n > 1 else 0.0)
            sums[out + "_all"][g] = torch.as_tensor(s_all)
            sums[out + "_auto"][g] = torch.as_tensor(s_auto)
        th_all = ref["thermal_all"][g] * n * n * nt
        th_auto = th_all - (ref["thermal_noautos"][g] * n * (n - 1) * nt if n > 1 else 0.0)
        sums["thermal_all"][g] = torch.as_tensor(th_all)
        sums["thermal_auto"][g] = torch.as_tensor(th_auto)
    return sums


def _worker(rank, world, port, mode, out_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a = _array()
        if mode == "time":
            t0, t1 = shard.time_shards(T, world)[rank]
            sums = _partial_sums(a, slice(t0, t1))
        else:
            mine = shard.group_shards(GROUP_SIZES, world)[rank]
            sums = _partial_sums(a, slice(0, T), groups=mine)
        shard.reduce_partials(sums, dst=0)
        if rank == 0:
            torch.save(sums, out_path)
        # all_ranks variant: everyone ends with the same totals
        again = {k: torch.full_like(v, float(rank + 1)) for k, v in sums.items()}
        shard.reduce_partials(again, all_ranks=True)
        want = float(sum(range(1, world + 1)))
        assert all(bool((v == want).all()) for v in again.values())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["time", "group"])
def test_two_rank_reduce_equals_whole_job(tmp_path, mode):
    out = str(tmp_path / "sums.pt")
    mp.spawn(_worker, args=(2, _free_port(), mode, out), nprocs=2, join=True)
    got = torch.load(out)
    whole = _partial_sums(_array(), slice(0, T))
    for k in shard.SUM_KEYS:
        np.testing.assert_allclose(got[k].numpy(), whole[k].numpy(), rtol=1e-9, atol=1e-9 * float(whole[k].abs().max()),
                                   err_msg=k)


def test_time_shards_cover_and_balance():
    for nt, w in [(1000, 8), (7, 2), (3, 4), (0, 2), (8, 8)]:
        sh = shard.time_shards(nt, w)
        assert len(sh) == w and sh[0][0] == 0 and sh[-1][1] == nt
        assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        sizes = [b - a for a, b in sh]
        assert max(sizes) - min(sizes) <= 1


def test_group_shards_are_a_balanced_partition():
    rng = np.random.default_rng(0)
    sizes = np.concatenate([rng.integers(1, 311, 600), np.ones(3000, dtype=np.int64)])
    for w in (1, 2, 4, 8):
        bins = shard.group_shards(sizes, w)
        flat = sorted(g for b in bins for g in b)
        assert flat == list(range(len(sizes)))  # every group exactly once
        loads = [int(sizes[b].sum()) for b in bins]
        assert max(loads) - min(loads) <= int(sizes.max())  # LPT bound
    assert shard.choose_partition(1000, sizes, 8) == "time"
    assert shard.choose_partition(3, sizes, 8) == "group"


def test_pack_unpack_round_trip():
    sums = dict(pow_all=torch.randn(3, 2, 2, 8, dtype=torch.float64), pow_auto=torch.randn(3, 2, 2, 8, dtype=torch.float64),
                noise_all=torch.randn(3, 2, 2, 8, dtype=torch.complex128), noise_auto=torch.randn(3, 2, 2, 8, dtype=torch.float64),
                thermal_all=torch.randn(3, 2, 2, dtype=torch.float64), thermal_auto=torch.randn(3, 2, 2, dtype=torch.float64))
    flat, layout = shard.pack_sums(sums)
    assert flat.numel() == sum(v.numel() * (2 if v.is_complex() else 1) for v in sums.values())
    back = {k: torch.zeros_like(v) for k, v in sums.items()}
    shard.unpack_sums(flat, layout, back)
    assert all(torch.equal(back[k], sums[k]) for k in sums)
