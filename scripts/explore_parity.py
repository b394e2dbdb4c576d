"""Exploration (not a test): measured errors of the fused engine for the hardening tests."""
import numpy as np, torch, sys
sys.path.insert(0, ".")
from scipy.signal import windows
from oracle import simpleds_oracle as orc
from simpleds_b200.engine import RedundantPipeline

dev = torch.device("cuda:0")
def run(vis, flags, ns, it, go, freq, wins, precision, noise=None, thermal=False):
    pipe = RedundantPipeline(freq, wins, np.array([-5]), go, trcvr_k=144.0, beam_area_sr=0.76, precision=precision, seed=1)
    pipe.process(torch.as_tensor(vis, device=dev), torch.as_tensor(flags, device=dev), torch.as_tensor(ns, device=dev),
                 torch.as_tensor(it, device=dev), noise=noise, thermal=thermal)
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in pipe.result().items()}

def oracle_means(vis, flags, df):
    x = vis.astype(np.complex128) * (~flags)
    X = orc.normalized_fourier_transform(x, df)     # (P,B,T,N)
    n = X.shape[1]
    s1 = X.sum(1); s2 = (np.abs(X) ** 2).sum(1)
    return (np.abs(s1) ** 2).mean(1) / n**2, ((np.abs(s1) ** 2 - s2) / (n * (n - 1))).mean(1)

rng = np.random.Generator(np.random.PCG64(1))
N, T, df = 256, 2, 97656.25
freq = 100e6 + df * np.arange(N)
for n in (310,):
    for kind in ("signal4", "noise_only"):
        common = rng.standard_normal((1, 1, T, N)) + 1j * rng.standard_normal((1, 1, T, N))
        vis = ((4.0 if kind == "signal4" else 0.0) * common + rng.standard_normal((1, n, T, N)) + 1j * rng.standard_normal((1, n, T, N))).astype(np.complex64)
        flags = rng.random((1, n, T, N)) < 0.05
        ns = np.full((1, n, T, N), 40, np.float32); it = np.full((n, T), 10.7, np.float32)
        go = np.array([0, n])
        want_all, want_no = oracle_means(vis, flags, df)
        for prec in ("complex64", "complex128"):
            r = run(vis, flags, ns, it, go, freq, [(0, N - 1)], prec)
            ea = np.abs(r["power_all"][0, 0] - want_all) ; en = np.abs(r["power_noautos"][0, 0] - want_no)
            print(f"n={n} {kind:10s} {prec}: all  normwise {ea.max()/np.abs(want_all).max():.2e} elementwise {(ea/np.abs(want_all)).max():.2e} | "
                  f"noautos normwise {en.max()/np.abs(want_no).max():.2e} (vs all-scale {en.max()/np.abs(want_all).max():.2e}) elementwise {(en/np.abs(want_no)).max():.2e}")

# dynamic range: smooth foreground 1e5 x the noise amplitude, Blackman-Harris
n = 12
nu = freq / 150e6
fg = 1e5 * nu ** -2.55 * np.exp(2j * np.pi * 0.7 * (nu - 1.0))     # smooth in frequency
vis = (fg[None, None, None, :] + rng.standard_normal((1, n, T, N)) + 1j * rng.standard_normal((1, n, T, N))).astype(np.complex64)
flags = np.zeros((1, n, T, N), bool); ns = np.full((1, n, T, N), 40, np.float32); it = np.full((n, T), 10.7, np.float32)
go = np.array([0, n])
want_all, want_no = oracle_means(vis, flags, df)
print("dynamic range of the oracle row (power): max / median = %.2e" % (want_all.max() / np.median(want_all)))
hi = np.abs(np.arange(N) - N // 2) > 40      # high-delay bins
for prec in ("complex64", "complex128"):
    r = run(vis, flags, ns, it, go, freq, [(0, N - 1)], prec)
    for key, want in (("power_all", want_all), ("power_noautos", want_no)):
        e = np.abs(r[key][0, 0] - want)
        print(f"dynrange {prec} {key}: normwise {e.max()/np.abs(want).max():.2e}; elementwise all bins {(e/np.abs(want)).max():.2e}; high-delay bins {(e[:, hi]/np.abs(want[:, hi])).max():.2e} (median {np.median(e[:, hi]/np.abs(want[:, hi])):.2e})")
