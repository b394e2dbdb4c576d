"""Production-path statistics of the in-kernel noise realisation (fused engine, noise="philox").

The reference draws sigma (g1 + i g2) / sqrt(2) per channel with numpy's global generator
(utils.py:501-505), so parity of a realisation is statistical.  The engine's
production draw is the moment-matched discrete law of csrc/sds_engine.cuh; what must agree with
the reference is the law of the DELAY-domain noise power the engine returns.  Many one-baseline
groups give independent samples y = |N[k]|^2 of one time each; against the expectation computed
from the reference's sigma (oracle) the tests check, per delay bin and pooled:
  * the mean (variance ratio of N[k]),
  * Var(y) / mean(y)^2 = 1 and the whole law of y (Kolmogorov-Smirnov against the exponential law a
    circular complex Gaussian N[k] implies) -- fourth moments,
  * the covariance of neighbouring delay bins, which the taper, the flag comb and the nsample
    profile of the row fix: Cov(y_k, y_k+1) = |E N[k] conj(N[k+1])|^2,
  * zero mean of the cross terms (the no-autos estimate of two-baseline groups).
The Box-Muller draw ("philox_gauss") goes through the same checks.
"""
import numpy as np
import pytest
import torch
from scipy import stats
from scipy.signal import windows

from oracle import simpleds_oracle as orc

pytestmark = pytest.mark.gpu


def _run(N, G, n_per_group, mode, seed, F0=None):
    from simpleds_b200.engine import RedundantPipeline

    F0 = F0 or N
    rng = np.random.Generator(np.random.PCG64(5))
    B = G * n_per_group
    go = np.arange(0, B + 1, n_per_group)
    freq = 100e6 + 97656.25 * np.arange(F0)
    # one flag / nsample profile for every row, so that all rows share one expectation
    fl_row = rng.random(F0) < 0.04
    fl_row[5::37] = True
    ns_row = rng.integers(8, 61, F0).astype(np.float32)
    vis = np.zeros((1, B, 1, F0), np.complex64)
    flags = np.broadcast_to(fl_row, (1, B, 1, F0)).copy()
    ns = np.broadcast_to(ns_row, (1, B, 1, F0)).copy()
    it = np.full((B, 1), 10.7, np.float32)
    dev = torch.device("cuda:0")
    pipe = RedundantPipeline(freq, [(0, N - 1)], np.array([-5]), go, trcvr_k=144.0, beam_area_sr=0.76,
                             precision="complex64", seed=seed)
    pipe.process(torch.as_tensor(vis, device=dev), torch.as_tensor(flags, device=dev), torch.as_tensor(ns, device=dev),
                 torch.as_tensor(it, device=dev), noise=mode, thermal=False)
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in pipe.result().items() if hasattr(v, "cpu")}
    # expectation from the reference's sigma (ds.py:3546-3583): a_c = (1 - flag) taper delta_f sigma_c
    sigma = orc.calculate_noise_power(freq[None, :N], np.full((1, N), 144.0), np.full((1, 1, N), 0.76),
                                      it[:1].astype(np.float64), ns[None, None, :, :1, :, :N].astype(np.float64))
    a2 = ((~fl_row[:N]) * windows.blackmanharris(N) * pipe.delta_f * sigma.reshape(-1)[:N]) ** 2
    mu = a2.sum()
    # covariance of neighbouring bins of a circular Gaussian with independent channels
    rho2 = np.abs((a2 * np.exp(-2j * np.pi * np.arange(N) / N)).sum()) ** 2 / mu**2
    return res, mu, rho2


@pytest.mark.parametrize("N,mode", [(256, "philox"), (1024, "philox"), (64, "philox"), (256, "philox_gauss")])
def test_delay_domain_noise_power_has_the_gaussian_law(cuda_device, N, mode):
    G = 4096
    res, mu, rho2 = _run(N, G, 1, mode, seed=20261020)
    y = res["noise_power_all"][:, 0, 0, :] / mu       # (G, N), one sample |N[k]|^2 / E per group and bin
    assert np.all(np.isfinite(y)) and y.min() >= 0
    # mean per bin: 1 +- 1/sqrt(G) (y is exponential: unit relative scatter); grand mean tighter
    per_bin = y.mean(0)
    assert np.abs(per_bin - 1).max() < 5.5 / np.sqrt(G), f"per-bin variance ratio off: {per_bin.min():.4f}..{per_bin.max():.4f}"
    assert abs(y.mean() - 1) < 0.01, f"pooled variance ratio {y.mean():.5f}"
    # second moment of the power (fourth moment of the noise): Var(y) / mean^2 = 1 for a Gaussian
    r = (y.var(0, ddof=1) / per_bin**2).mean()
    assert abs(r - 1) < 0.03, f"Var(|N|^2) / <|N|^2>^2 = {r:.4f} (Gaussian: 1)"
    # whole law: one bin per group (independent samples) against Exp(1)
    pick = y[np.arange(G), (np.arange(G) * 7) % N]
    ks = stats.kstest(pick, "expon")
    assert ks.statistic < 0.03, f"KS distance to the exponential law {ks.statistic:.4f} (p = {ks.pvalue:.3f})"
    # neighbouring delay bins: Cov(y_k, y_k+1) / (mean_k mean_k+1) = |rho_1|^2, fixed by taper + flags + nsample
    c = ((y[:, :-1] - 1) * (y[:, 1:] - 1)).mean()
    assert abs(c - rho2) < 0.04 * max(rho2, 0.25), f"adjacent-bin covariance {c:.4f}, expectation {rho2:.4f}"


@pytest.mark.parametrize("mode", ["philox", "philox_gauss"])
def test_noise_cross_terms_have_zero_mean(cuda_device, mode):
    """Two-baseline groups: the no-autos noise power is Re n1 conj(n2): mean 0, variance mu^2 / 2."""
    G, N = 4096, 256
    res, mu, _ = _run(N, G, 2, mode, seed=77)
    z = res["noise_power_noautos"][:, 0, 0, :] / mu
    assert abs(z.mean()) < 5 / np.sqrt(2 * G * N / 4), f"mean of the cross terms {z.mean():.5f}"
    assert abs(z.var() - 0.5) < 0.03, f"variance of the cross terms {z.var():.4f} (expect 0.5)"
    # and the all-pairs mean is mu / n = 0.5
    assert abs(res["noise_power_all"][:, 0, 0, :].mean() / mu - 0.5) < 0.01
