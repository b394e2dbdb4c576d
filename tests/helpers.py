"""Shared generators / comparators for the parity tests."""
import numpy as np

RTOL_C128 = 1e-9  # north_star: rtol 1e-9 for complex128
RTOL_C64 = 1e-5   # north_star: rtol 1e-5 for complex64 output


def assert_close(got, ref, rtol, axis=-1, what=""):
    """Elementwise rtol plus atol = rtol * max|ref| over the transform row
    (norm-wise criterion of SURVEY.md section 7 'Tolerance semantics')."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} != {ref.shape}"
    if ref.size == 0:
        return
    scale = np.abs(ref).max(axis=axis, keepdims=True)
    err = np.abs(got - ref)
    bad = err > rtol * scale + rtol * np.abs(ref)
    assert not bad.any(), (
        f"{what}: {bad.sum()} of {bad.size} elements off; max err/scale = "
        f"{(err / np.maximum(scale, 1e-300)).max():.3e} (rtol {rtol})"
    )


def make_group(seed=3, S0=1, U=1, P=2, B=5, T=4, F=21, flag_frac=0.1, zero_ns_frac=0.0,
               f0=146.8e6, df=492610.84375, t_int=42.95, trcvr=144.0, beam=0.76):
    """One redundant group in the reference's 6-D layout, complex64-representable
    visibilities (utils.py:38), nsample integers in [32, 60]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    shape = (S0, U, P, B, T, F)
    common = rng.standard_normal((S0, U, P, 1, T, F)) + 1j * rng.standard_normal((S0, U, P, 1, T, F))
    vis = (3.0 * common + rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
    vis = vis.astype(np.complex64)
    flags = rng.random(shape) < flag_frac
    ns = rng.integers(32, 61, shape).astype(np.float64)
    if zero_ns_frac:
        ns[rng.random(shape) < zero_ns_frac] = 0.0
    freq = (f0 + df * np.arange(F))[None].repeat(S0, 0)
    pols = np.array([1, 2, 3, 4, -5, -6])[:P]
    return dict(
        vis=vis, flags=flags, nsample=ns, freq_array_hz=freq,
        trcvr_k=np.full((S0, F), trcvr), beam_area_sr=np.full((S0, P, F), beam),
        integration_time_s=np.full((B, T), t_int) * (1.0 + 0.01 * rng.random((B, T))),
        polarization_array=pols,
    )
