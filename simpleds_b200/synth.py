"""Synthetic redundant arrays and visibilities for the BASELINE.json configs.

Antenna positions live on integer lattices so that redundancy grouping is exact
integer arithmetic (bit-exact pair indexing): a baseline is the integer lattice
vector between its antennas, conjugated so that it points to u > 0 (or u == 0,
v > 0) as pyuvdata's ``get_redundancies(include_conjugates=True)`` does, and a
redundant group is the set of baselines with the same vector.  Inside a group
baselines are ordered by baseline number (ds.py:269-272, 931); groups are ordered
by size, largest first (helps the dynamic item schedule; the reference has no
group order -- one DelaySpectrum object per group).
"""
import numpy as np
import torch


def paper_grid(ncol=8, nrow=8):
    """PAPER-like grid: 30 m E-W x 4 m N-S spacing (BASELINE.json config 2)."""
    c, r = np.meshgrid(np.arange(ncol), np.arange(nrow), indexing="ij")
    lattice = np.stack([c.ravel(), r.ravel()], 1).astype(np.int64)
    basis = np.array([[30.0, 0.0], [0.0, 4.0]])
    return lattice, basis


def hera_350(side=11, spacing=14.6):
    """Hex core of side 11 (331 antennas) + 19 outriggers on the same lattice = 350."""
    pts = []
    n = side - 1
    for a in range(-n, n + 1):
        for b in range(-n, n + 1):
            if abs(a + b) <= n:
                pts.append((a, b))
    core = np.array(sorted(pts), dtype=np.int64)
    assert len(core) == 3 * side * (side - 1) + 1
    ring = []
    k = 0
    while len(ring) < 19:  # deterministic outer positions, pairwise-distinct vectors
        ang = k * 7
        a = 17 + (ang % 13) + 3 * k
        b = -(9 + (5 * k) % 11) - 2 * k
        ring.append((a, b) if k % 2 == 0 else (b - 4, a + 6))
        k += 1
    lattice = np.concatenate([core, np.array(ring, dtype=np.int64)])
    assert len(np.unique(lattice, axis=0)) == 350
    basis = np.array([[spacing, 0.0], [spacing * 0.5, spacing * np.sqrt(3) / 2]])
    return lattice, basis


def ska_512(ncol=32, nrow=16, spacing=20.0):
    """SKA-Low-style 32 x 16 rectangular station lattice (BASELINE.json config 5)."""
    c, r = np.meshgrid(np.arange(ncol), np.arange(nrow), indexing="ij")
    lattice = np.stack([c.ravel(), r.ravel()], 1).astype(np.int64)
    return lattice, np.array([[spacing, 0.0], [0.0, spacing]])


def baseline_number(ant_1, ant_2):
    """pyuvdata antnums_to_baseline, < 2048 antennas (ds.py:269-272)."""
    return 2048 * (np.asarray(ant_1, np.int64) + 1) + (np.asarray(ant_2, np.int64) + 1) + 2**16


def redundant_groups(lattice, basis):
    """All cross baselines of the array grouped by integer lattice vector.

    Returns dict: ant_1, ant_2 (Nbls,) in group order (conjugated so the vector is
    'positive'), baseline_array, group_offset (Ngroups+1,), group_vector (Ngroups,2)
    integer vectors, uvw (Ngroups,2) in metres."""
    n = len(lattice)
    i, j = np.triu_indices(n, k=1)
    vec = lattice[j] - lattice[i]
    # sign of u = vec @ basis[:,0]; decide with exact integer arithmetic when the basis is
    # (1,0),(1/2,h): u ~ 2 a + b ; otherwise u ~ a.  Ties fall back to v ~ b.
    if abs(basis[1, 0]) > 0:
        u_sign, v_sign = 2 * vec[:, 0] + vec[:, 1], vec[:, 1]
    else:
        u_sign, v_sign = vec[:, 0], vec[:, 1]
    flip = (u_sign < 0) | ((u_sign == 0) & (v_sign < 0))
    a1 = np.where(flip, j, i)
    a2 = np.where(flip, i, j)
    vec = np.where(flip[:, None], -vec, vec)
    key, inverse, counts = np.unique(vec, axis=0, return_inverse=True, return_counts=True)
    inverse = inverse.reshape(-1)
    bl = baseline_number(a1, a2)
    gorder = np.lexsort((key[:, 1], key[:, 0], -counts))  # size desc, then vector
    rank = np.empty_like(gorder)
    rank[gorder] = np.arange(len(gorder))
    order = np.lexsort((bl, rank[inverse]))  # by group, then baseline number
    sizes = counts[gorder]
    return dict(
        ant_1=a1[order], ant_2=a2[order], baseline_array=bl[order],
        group_offset=np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64),
        group_vector=key[gorder], uvw=key[gorder] @ basis,
    )


def make_visibilities(group_offset, npol, ntime, freq_hz, seed=1003, device="cuda", t_int=10.7,
                      flag_comb=0.05, flag_random=0.01, time_offset=0, chunk_bl=4096):
    """Device-side synthetic visibilities: per group a common
    smooth foreground + unit-variance 'EoR', plus independent per-baseline noise;
    flags = a fixed RFI comb + random samples; nsample integers in [32, 60].
    Returns vis c64 (P,B,T,F), flags u8, nsample f32, int_time f32 (B,T)."""
    dev = torch.device(device)
    go = np.asarray(group_offset)
    B, F = int(go[-1]), len(freq_hz)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed) + 7919 * int(time_offset))
    ngroup = len(go) - 1
    nu = torch.as_tensor(np.asarray(freq_hz) / 150e6, dtype=torch.float32, device=dev)
    sizes = torch.as_tensor(np.diff(go), device=dev)
    gid = torch.repeat_interleave(torch.arange(ngroup, device=dev), sizes)
    vis = torch.empty((npol, B, ntime, F), dtype=torch.complex64, device=dev)
    flags = torch.empty((npol, B, ntime, F), dtype=torch.uint8, device=dev)
    nsample = torch.empty((npol, B, ntime, F), dtype=torch.float32, device=dev)
    tt = torch.arange(ntime, device=dev, dtype=torch.float32) + float(time_offset)
    amp = 100.0 + 400.0 * torch.rand((npol, ngroup, 1, 1), generator=g, device=dev)
    rate = 0.02 * torch.randn((npol, ngroup, 1, 1), generator=g, device=dev)
    tau = 3.0 * torch.randn((npol, ngroup, 1, 1), generator=g, device=dev)
    comb = torch.zeros(F, dtype=torch.bool, device=dev)
    step = max(int(round(1.0 / flag_comb)), 1) if flag_comb > 0 else 0
    if step:
        comb[7::step] = True
    for p in range(npol):
        for b0 in range(0, B, chunk_bl):
            b1 = min(B, b0 + chunk_bl)
            gi = gid[b0:b1]
            phase = rate[p, gi] * tt.view(1, -1, 1) + tau[p, gi] * (nu.view(1, 1, -1) - 1.0)
            fg = amp[p, gi] * nu.view(1, 1, -1) ** -2.55
            eor_seed = torch.Generator(device=dev)
            eor_seed.manual_seed(int(seed) * 31 + p)
            common = torch.polar(fg.expand(b1 - b0, ntime, F).contiguous(), phase.expand(b1 - b0, ntime, F).contiguous())
            noise = torch.randn((b1 - b0, ntime, F, 2), generator=g, device=dev)
            vis[p, b0:b1] = common + 2.0 * torch.view_as_complex(noise)
            fl = torch.rand((b1 - b0, ntime, F), generator=g, device=dev) < flag_random
            flags[p, b0:b1] = (fl | comb.view(1, 1, -1)).to(torch.uint8)
            nsample[p, b0:b1] = torch.randint(32, 61, (b1 - b0, ntime, F), generator=g, device=dev).to(torch.float32)
    int_time = torch.full((B, ntime), float(t_int), dtype=torch.float32, device=dev)
    return vis, flags, nsample, int_time
