"""Full-size checks of the fused engine (BASELINE.json configs[2]: HERA-350, 61 075 baselines
in 3 601 redundant groups, 1024 channels in 4 windows of 256, 4 pols) through properties that
do not need the O(n^2) oracle:

* Parseval: sum_k |FFT(x)|_k^2 = N sum_n |x_n|^2, so the delay-summed auto power of a group is
  N df^2 sum_rows sum_n |x m w|^2 and the delay-summed all-pairs power is
  N df^2 sum_t sum_n |sum_rows x m w|^2 -- both computed here with plain torch reductions in
  fp64, independent of the kernels;
* the fused path against the chained API kernels (any-length transform + pair sums), two
  independent implementations, on a sample of groups;
* bilinearity (vis * 2 -> power * 4, exactly in binary floating point), time-batch invariance,
  determinism, and the noise leg's expectation (<noise auto power> = thermal-sigma power).
"""
import numpy as np
import pytest
import torch
from scipy.signal import windows

pytestmark = pytest.mark.gpu

NCHAN, NWIN, NPOL, NT = 1024, 4, 4, 2
N = NCHAN // NWIN
DF = 97656.25


@pytest.fixture(scope="module")
def hera(cuda_device):
    from simpleds_b200 import synth
    from simpleds_b200.engine import RedundantPipeline

    lattice, basis = synth.hera_350()
    red = synth.redundant_groups(lattice, basis)
    freq = 100e6 + DF * np.arange(NCHAN)
    wins = [(w * N, (w + 1) * N - 1) for w in range(NWIN)]
    arrs = synth.make_visibilities(red["group_offset"], NPOL, NT, freq, seed=1003, device=cuda_device)
    pipe = RedundantPipeline(freq, wins, [-5, -6, -7, -8], red["group_offset"], trcvr_k=144.0,
                             beam_area_sr=0.76, precision="complex64", seed=0)
    assert pipe.fused and pipe.nbl == 61075 and pipe.ngroup == 3601
    pipe.process(*arrs, noise="philox", thermal=True)
    torch.cuda.synchronize()
    sums = {k: v.clone() for k, v in pipe.sums().items()}
    return dict(pipe=pipe, arrs=arrs, sums=sums, go=red["group_offset"], freq=freq, wins=wins)


def _segment_sum(x, go, dev):
    """Sum x (B, ...) over the baselines of each group -> (G, ...)."""
    sizes = torch.as_tensor(np.diff(go), device=dev)
    gid = torch.repeat_interleave(torch.arange(len(go) - 1, device=dev), sizes)
    out = torch.zeros((len(go) - 1,) + tuple(x.shape[1:]), dtype=x.dtype, device=dev)
    return out.index_add_(0, gid, x)


def test_parseval_at_full_size(hera, cuda_device):
    vis, flags, nsample, int_time = hera["arrs"]
    go, pipe = hera["go"], hera["pipe"]
    taper = torch.as_tensor(windows.blackmanharris(N), dtype=torch.float64, device=cuda_device)
    want_auto = torch.zeros((pipe.ngroup, NWIN, NPOL), dtype=torch.float64, device=cuda_device)
    want_all = torch.zeros_like(want_auto)
    for p in range(NPOL):
        for s in range(NWIN):
            sl = slice(s * N, (s + 1) * N)
            x = vis[p, :, :, sl].to(torch.complex128) * (flags[p, :, :, sl] == 0) * taper  # (B,T,N)
            want_auto[:, s, p] = _segment_sum((x.abs() ** 2).sum(dim=(1, 2)), go, cuda_device)
            want_all[:, s, p] = (_segment_sum(x, go, cuda_device).abs() ** 2).sum(dim=(1, 2))
            del x
    scale = N * DF * DF
    got_auto = hera["sums"]["pow_auto"].sum(dim=-1)
    got_all = hera["sums"]["pow_all"].sum(dim=-1)
    # fp32 transform and row sums, fp64 reference: relative 1e-5 per (group, spw, pol)
    assert torch.allclose(got_auto, want_auto * scale, rtol=1e-5, atol=0)
    assert torch.allclose(got_all, want_all * scale, rtol=1e-5, atol=0)
    assert bool((hera["sums"]["pow_all"] >= 0).all()) and bool(torch.isfinite(hera["sums"]["pow_all"]).all())


def test_fused_equals_chained_kernels_on_sampled_groups(hera, cuda_device):
    """sds_engine_run vs sds_delay_transform + sds_cross_power_avg on the same full-size data."""
    from simpleds_b200 import ops
    from simpleds_b200._lib import SDS_F64

    vis, flags, _, _ = hera["arrs"]
    go = hera["go"]
    plan = ops.get_plan(N, SDS_F64, windows.blackmanharris)
    sizes = np.diff(go)
    picks = [int(np.argmax(sizes)), int(np.argmin(sizes)), 17, 1234, len(sizes) - 1]
    for g in picks:
        b0, b1 = int(go[g]), int(go[g + 1])
        x = ops.delay_transform(plan, vis[:, b0:b1].contiguous(), flags[:, b0:b1].contiguous(), DF,
                                win_start=[w[0] for w in hera["wins"]])          # (S,P,n,T,N) c128
        s1 = x.sum(dim=2)
        want_all = (s1.abs() ** 2).sum(dim=2)                                     # (S,P,N)
        want_auto = (x.abs() ** 2).sum(dim=(2, 3))
        for got, want, what in ((hera["sums"]["pow_all"][g], want_all, "all"),
                                (hera["sums"]["pow_auto"][g], want_auto, "auto")):
            err = (got - want).abs().amax(dim=-1) / want.abs().amax(dim=-1)
            assert float(err.max()) < 4e-5, f"group {g} (n={b1 - b0}) {what}: {float(err.max())}"


def test_bilinearity_batches_and_determinism(hera, cuda_device):
    vis, flags, nsample, int_time = hera["arrs"]
    pipe = hera["pipe"]
    base = hera["sums"]
    # same input, same seed -> bit-identical outputs? (fp64 atomics commute only approximately:
    # allow the last bits) and identical noise realisation
    pipe.reset()
    pipe.process(vis, flags, nsample, int_time, noise="philox", thermal=True)
    for k in ("pow_all", "noise_all", "thermal_all"):
        assert torch.allclose(pipe.sums()[k], base[k], rtol=1e-12, atol=0), k
    # two single-time batches == one two-time batch (noise counters are global coordinates)
    pipe.reset()
    for t in range(NT):
        pipe.process(vis[:, :, t:t + 1].contiguous(), flags[:, :, t:t + 1].contiguous(),
                     nsample[:, :, t:t + 1].contiguous(), int_time[:, t:t + 1].contiguous(),
                     time_offset=t, ntime_total=NT, noise="philox", thermal=True)
    for k in ("pow_all", "pow_auto", "noise_all", "noise_auto", "thermal_all", "thermal_auto"):
        assert torch.allclose(pipe.sums()[k], base[k], rtol=5e-6, atol=0), k
    # the data leg alone is another instantiation of the kernel (no lockstep noise transform;
    # ptxas contracts a few mul+add pairs differently): same sums to fp32 rounding, norm-wise
    pipe.reset()
    pipe.process(vis, flags, nsample, int_time, noise=None, thermal=False)
    one = {k: pipe.sums()[k].clone() for k in ("pow_all", "pow_auto")}
    for k in one:
        err = float(((one[k] - base[k]).abs().amax(dim=-1) / base[k].abs().amax(dim=-1)).max())
        assert err < 2e-6, (k, err)
    # power is bilinear in the visibilities: scaling by 2 is exact in binary floating point
    pipe.reset()
    pipe.process(vis * 2, flags, nsample, int_time, noise=None, thermal=False)
    for k in one:
        err = float(((pipe.sums()[k] - 4 * one[k]).abs() / one[k].abs().clamp_min(1e-300)).max())
        assert err < 1e-12, (k, err)


def test_noise_power_expectation_at_full_size(hera, cuda_device):
    """<sum_k |N_i[k]|^2> = N df^2 sum_n sigma_n^2 m_n w_n^2 (white noise through the taper);
    summed over all rows of the array the relative sampling error is ~1e-4."""
    _, flags, nsample, int_time = hera["arrs"]
    pipe = hera["pipe"]
    taper2 = torch.as_tensor(windows.blackmanharris(N) ** 2, dtype=torch.float64, device=cuda_device)
    want = 0.0
    for p in range(NPOL):
        for s in range(NWIN):
            sl = slice(s * N, (s + 1) * N)
            sig2 = pipe.sigma_coef64[s, p].view(1, 1, N) ** 2 / (int_time.view(-1, NT, 1).double() *
                                                                   nsample[p, :, :, sl].double())
            want += float((sig2 * (flags[p, :, :, sl] == 0) * taper2).sum())
    got = float(hera["sums"]["noise_auto"].sum())
    assert abs(got / (want * N * DF * DF) - 1.0) < 2e-3
    # the all-pairs noise power has the same expectation (independent rows: cross terms average out)
    got_all = float(hera["sums"]["noise_all"].sum())
    assert abs(got_all / (want * N * DF * DF) - 1.0) < 5e-2
