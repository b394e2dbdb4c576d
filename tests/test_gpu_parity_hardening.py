"""Parity where a loose tolerance could hide something ('harden parity'):

* the largest redundant group of the benchmark array (n = 310 baselines): all-pairs AND no-autos means
  against the oracle transform, at the north_star tolerance itself (norm-wise over the delay row;
  measured 7e-7 in complex64, 4e-16 in complex128 -- no factor n);
* a foreground 1e5 times the noise amplitude under the Blackman-Harris taper (12.7 decades of power
  between the peak and the noise-dominated delays): norm-wise AND elementwise errors, the elementwise
  one bounded by what the arithmetic's unit roundoff allows for a bin that many decades below the row
  maximum (eps sqrt(P_max / P_k): half of the first-order bound 2 |X_k| eps |X|_max / P_k); complex128 keeps the science bins to 1e-9, complex64 cannot (the
  reference itself computes in complex128, ds.py:933-939) and the test states by how much.
"""
import numpy as np
import pytest
import torch

from helpers import RTOL_C64, RTOL_C128
from oracle import simpleds_oracle as orc

pytestmark = pytest.mark.gpu
DF = 97656.25


def _run(vis, flags, precision):
    from simpleds_b200.engine import RedundantPipeline

    _, n, T, N = vis.shape
    dev = torch.device("cuda:0")
    freq = 100e6 + DF * np.arange(N)
    pipe = RedundantPipeline(freq, [(0, N - 1)], np.array([-5]), np.array([0, n]), trcvr_k=144.0, beam_area_sr=0.76,
                             precision=precision, seed=1)
    pipe.process(torch.as_tensor(vis, device=dev), torch.as_tensor(flags, device=dev),
                 torch.full((1, n, T, N), 40.0, device=dev), torch.full((n, T), 10.7, device=dev),
                 noise=None, thermal=False)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy()[0, 0, 0] for k, v in pipe.result().items() if hasattr(v, "cpu") and v.ndim == 4}


def _oracle(vis, flags):
    """Reference tensor algebra on the oracle transform (utils.py:509-560, 387-389): the mean over
    ordered pairs (i, j) of X_i conj(X_j) is |sum_i X_i|^2 / n^2; without autos
    (|sum_i X_i|^2 - sum_i |X_i|^2) / (n (n - 1))."""
    X = orc.normalized_fourier_transform(vis.astype(np.complex128) * (~flags), DF)[0]  # (B, T, N)
    n = X.shape[0]
    s1, s2 = X.sum(0), (np.abs(X) ** 2).sum(0)
    return (np.abs(s1) ** 2).mean(0) / n**2, ((np.abs(s1) ** 2 - s2) / (n * (n - 1))).mean(0)


@pytest.mark.parametrize("kind", ["common signal 4x the noise", "noise only"])
@pytest.mark.parametrize("precision,rtol", [("complex64", RTOL_C64), ("complex128", RTOL_C128)])
def test_largest_group_of_the_benchmark_array_at_the_stated_tolerance(cuda_device, precision, rtol, kind):
    n, T, N = 310, 2, 256
    rng = np.random.Generator(np.random.PCG64(1))
    common = rng.standard_normal((1, 1, T, N)) + 1j * rng.standard_normal((1, 1, T, N))
    amp = 4.0 if kind.startswith("common") else 0.0
    vis = (amp * common + rng.standard_normal((1, n, T, N)) + 1j * rng.standard_normal((1, n, T, N))).astype(np.complex64)
    flags = rng.random((1, n, T, N)) < 0.05
    want_all, want_no = _oracle(vis, flags)
    got = _run(vis, flags, precision)
    e_all = np.abs(got["power_all"] - want_all).max() / np.abs(want_all).max()
    # the no-autos mean is a difference of two sums: its error is measured on the scale of the
    # all-pairs mean times n / (n - 1) -- which is its own scale when a common signal dominates
    e_no = np.abs(got["power_noautos"] - want_no).max() / (np.abs(want_all).max() * n / (n - 1))
    print(f"\nn = {n}, {kind}, {precision}: norm-wise error all pairs {e_all:.2e}, no autos {e_no:.2e} (tolerance {rtol:.0e})")
    assert e_all < rtol and e_no < rtol


@pytest.mark.parametrize("precision,eps,rtol", [("complex64", 2.0 ** -24, RTOL_C64), ("complex128", 2.0 ** -53, RTOL_C128)])
def test_foreground_dynamic_range_norm_wise_and_elementwise(cuda_device, precision, eps, rtol):
    n, T, N = 12, 2, 256
    rng = np.random.Generator(np.random.PCG64(2))
    nu = (100e6 + DF * np.arange(N)) / 150e6
    fg = 1e5 * nu ** -2.55 * np.exp(2j * np.pi * 0.7 * (nu - 1.0))  # smooth spectrum, 1e5 x the noise amplitude
    vis = (fg + rng.standard_normal((1, n, T, N)) + 1j * rng.standard_normal((1, n, T, N))).astype(np.complex64)
    flags = np.zeros((1, n, T, N), bool)
    want_all, want_no = _oracle(vis, flags)
    got = _run(vis, flags, precision)
    hi = np.abs(np.arange(N) - N // 2) > 40  # noise-dominated delays
    assert want_all.max() / np.median(want_all[hi]) > 1e12  # the premise: > 12 decades of power
    for key, want in (("power_all", want_all), ("power_noautos", want_no)):
        err = np.abs(got[key] - want)
        normwise = err.max() / np.abs(want).max()
        elem = err / np.abs(want_all)  # relative to the all-pairs power of the same bin
        # an error of eps |X|_max in the transform moves the power of bin k by 2 |X_k| eps |X|_max
        bound = eps * np.sqrt(want_all.max() / want_all) + rtol
        print(f"\n{precision} {key}: norm-wise {normwise:.2e}; elementwise max {elem.max():.2e} "
              f"(noise-dominated delays: max {elem[hi].max():.2e}, median {np.median(elem[hi]):.2e}); "
              f"bound at the noise-dominated delays {np.median(bound[hi]):.2e}")
        assert normwise < rtol
        assert np.all(elem <= bound), f"{key}: elementwise error above the roundoff bound by {(elem / bound).max():.2f}x"
    if precision == "complex128":  # the reference's arithmetic keeps the science bins
        assert (np.abs(got["power_all"] - want_all) / want_all)[hi].max() < 1e-9
