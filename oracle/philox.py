"""Philox4x32-10 + Box-Muller restated in numpy.  TEST INFRASTRUCTURE ONLY.

The reference draws its noise from numpy's global MT19937 (utils.py:501-505); the
device path uses the counter-based Philox4x32-10 of Salmon et al. 2011 ("Parallel
random numbers: as easy as 1, 2, 3") so that noise does not depend on launch
geometry or on how times/groups are sharded.  This file restates that generator
so tests can check the device stream bit for bit on the uniforms and to ~1e-12
on the normals."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_LO = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Philox4x32-R.  R = 10 is the Random123 / cuRAND default; R = 7 is the smallest
    round count Salmon et al. report as Crush-resistant (used by the fused engine)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0 = np.asarray(k0, dtype=np.uint32)
    k1 = np.asarray(k1, dtype=np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(rounds):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = (p1 & _LO).astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = (p0 & _LO).astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return c0, c1, c2, c3


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    return philox4x32(c0, c1, c2, c3, k0, k1, rounds=10)


def normal_pairs(seed, n, offset=0):
    """The (g1, g2) the device draws for flat sample indices offset .. offset+n-1
    (sds_elementwise.cu: philox_normal_pair)."""
    q = np.arange(n, dtype=np.uint64) + np.uint64(offset)
    ctr = q >> np.uint64(1)
    r = philox4x32_10(
        (ctr & _LO).astype(np.uint32), (ctr >> np.uint64(32)).astype(np.uint32),
        np.zeros(n, np.uint32), np.zeros(n, np.uint32),
        np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF),
    )
    odd = (q & np.uint64(1)).astype(bool)
    a = np.where(odd, r[2], r[0]).astype(np.float64)
    b = np.where(odd, r[3], r[1]).astype(np.float64)
    u1 = (a + 0.5) * 2.3283064365386963e-10
    u2 = (b + 0.5) * 2.3283064365386963e-10
    rad = np.sqrt(-2.0 * np.log(u1))
    return rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)
