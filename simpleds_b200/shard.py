"""Partitioning of the delay-spectrum job over the GPUs of one box, and the single
exchange step of the path.

The reference is single-process; what makes the path shardable is pinned by its own
tests: slices along time / polarisation / spectral window commute with the whole
pipeline (simpleDS/tests/test_io.py:570-631, 680-736), and every redundant group is
a separate ``DelaySpectrum`` object (delay_spectrum.py:891-900).  So

* times are split evenly over the ranks (every rank sees every group: perfect
  balance); each rank accumulates the partial pair sums of its own times and ONE
  reduce (sum, fp64) of the flat accumulator buffer closes the job;
* when there are fewer times than ranks, whole groups are dealt to the ranks by
  longest-processing-time bin packing (bytes are proportional to the group size);
  results are then disjoint and the same reduce acts as a gather.

Nothing here touches the data path: no collective runs until the local fused pass is
done.  Works with the ``nccl`` backend on CUDA tensors and with ``gloo`` on CPU
tensors (tests).
"""
import numpy as np
import torch
import torch.distributed as dist

SUM_KEYS = ("pow_all", "pow_auto", "noise_all", "noise_auto", "thermal_all", "thermal_auto")
# per-group count of accumulated times: rides in the same flat buffer when present, so that the
# reduced means are right for time shards (counts add up) and group shards (disjoint) alike
OPTIONAL_KEYS = ("times_done",)


def time_shards(ntimes, world):
    """Contiguous, near-even split of range(ntimes): [(t0, t1)] per rank (possibly empty)."""
    ntimes, world = int(ntimes), int(world)
    if ntimes < 0 or world < 1:
        raise ValueError("ntimes must be >= 0 and world >= 1")
    base, extra = divmod(ntimes, world)
    out, t = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((t, t + n))
        t += n
    return out


def group_shards(group_sizes, world):
    """Longest-processing-time bin packing of redundant groups: returns, per rank, the
    sorted indices of its groups.  Cost of a group = its number of baselines (the
    bytes the fused pass reads for it)."""
    sizes = np.asarray(group_sizes, dtype=np.int64)
    if world < 1:
        raise ValueError("world must be >= 1")
    order = np.argsort(-sizes, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    bins = [[] for _ in range(world)]
    for g in order:
        r = int(np.argmin(load))  # first minimum: deterministic
        bins[r].append(int(g))
        load[r] += sizes[g]
    return [sorted(b) for b in bins]


def choose_partition(ntimes, group_sizes, world):
    """'time' when every rank can get at least one time, else 'group'."""
    return "time" if ntimes >= world else "group"


def pack_sums(sums):
    """Flatten the accumulator dict (SUM_KEYS order) into one fp64 vector + its layout."""
    parts, layout = [], []
    for k in SUM_KEYS + tuple(k for k in OPTIONAL_KEYS if k in sums):
        t = sums[k]
        if t.is_complex():
            t = torch.view_as_real(t)
        if t.dtype != torch.float64:
            raise ValueError(f"{k} must be float64/complex128")
        layout.append((k, tuple(sums[k].shape), sums[k].is_complex(), t.numel()))
        parts.append(t.reshape(-1))
    return torch.cat(parts), layout


def unpack_sums(flat, layout, into):
    """Write the flat vector back into the tensors of ``into`` (in place)."""
    off = 0
    for k, shape, is_c, n in layout:
        seg = flat[off:off + n]
        if is_c:
            seg = torch.view_as_complex(seg.reshape(*shape, 2))
        into[k].copy_(seg.reshape(shape))
        off += n
    return into


def reduce_partials(sums, dst=0, group=None, all_ranks=False):
    """The one exchange step: sum the ranks' partial accumulators (dict of SUM_KEYS
    tensors, modified in place on ``dst``; on every rank with ``all_ranks``).  A single
    collective on one flat fp64 buffer -- (2 x 4 x Ngroups x Nspws x Npols x Ndelays +
    2 x ...) x 8 B, bandwidth-trivial on NVLink."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sums
    flat, layout = pack_sums(sums)
    if all_ranks:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(flat, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if all_ranks or dist.get_rank(group) == dst:
        unpack_sums(flat, layout, sums)
    return sums
