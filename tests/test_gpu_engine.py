"""GPU parity tests of the fused engine (sds_engine_run through RedundantPipeline)
against 'reference tensor, then numpy mean' (oracle/engine_ref.py)."""
import numpy as np
import pytest
import torch
from scipy.signal import windows

from helpers import RTOL_C64, RTOL_C128, assert_close
from oracle import engine_ref

pytestmark = pytest.mark.gpu


def make_array(seed, P, group_sizes, T, F0, flag_frac=0.08, zero_ns_frac=0.0, f0=100e6, df=97656.25):
    rng = np.random.Generator(np.random.PCG64(seed))
    B = int(np.sum(group_sizes))
    go = np.concatenate([[0], np.cumsum(group_sizes)])
    vis = np.zeros((P, B, T, F0), np.complex64)
    for g in range(len(group_sizes)):
        common = rng.standard_normal((P, 1, T, F0)) + 1j * rng.standard_normal((P, 1, T, F0))
        n = group_sizes[g]
        vis[:, go[g]:go[g + 1]] = (4.0 * common + rng.standard_normal((P, n, T, F0))
                                   + 1j * rng.standard_normal((P, n, T, F0))).astype(np.complex64)
    flags = rng.random((P, B, T, F0)) < flag_frac
    flags[..., 5::37] = True  # RFI comb
    ns = rng.integers(32, 61, (P, B, T, F0)).astype(np.float32)
    if zero_ns_frac:
        ns[rng.random(ns.shape) < zero_ns_frac] = 0.0
    it = (10.7 * (1.0 + 0.02 * rng.random((B, T)))).astype(np.float32)
    freq = f0 + df * np.arange(F0)
    pols = np.array([-5, -6, -7, -8, 1, 2])[:P]
    return dict(vis=vis, flags=flags, nsample=ns, int_time=it, group_offset=go, freq=freq, pols=pols)


def run_pipeline(a, wins, precision, noise, noise_in=None, thermal=True, seed=7, batches=None, taper=None):
    from simpleds_b200.engine import RedundantPipeline

    dev = torch.device("cuda:0")
    pipe = RedundantPipeline(a["freq"], wins, a["pols"], a["group_offset"], trcvr_k=144.0,
                             beam_area_sr=0.76, precision=precision, seed=seed, taper=taper)
    T = a["vis"].shape[2]
    batches = batches or [(0, T)]
    for t0, t1 in batches:
        sl = slice(t0, t1)
        pipe.process(
            torch.as_tensor(a["vis"][:, :, sl], device=dev), torch.as_tensor(a["flags"][:, :, sl], device=dev),
            torch.as_tensor(a["nsample"][:, :, sl], device=dev), torch.as_tensor(a["int_time"][:, sl], device=dev),
            time_offset=t0, ntime_total=T, noise=noise, thermal=thermal,
            noise_in=None if noise_in is None else torch.as_tensor(noise_in[:, :, sl], device=dev),
        )
    torch.cuda.synchronize()
    return pipe, {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in pipe.result().items()}


def check_against_oracle(res, ref, group_sizes, rtol, keys):
    """Norm-wise over the delay row at the tolerance itself (measured at n = 310: 7e-7 in complex64,
    tests/test_gpu_parity_hardening.py).  The no-autos mean is a difference of two sums and can be
    arbitrarily small next to them (noise-like data), so its error is measured on the scale of the
    all-pairs mean of the same quantity times n / (n - 1)."""
    for key in keys:
        for g, n in enumerate(group_sizes):
            want = np.asarray(ref[key][g])
            if key.endswith("noautos") and n == 1:
                assert np.all(np.isnan(res[key][g])), f"{key} group {g}: one baseline has no cross pairs"
                continue
            if key.startswith("thermal"):
                np.testing.assert_allclose(res[key][g], want, rtol=4 * rtol, err_msg=f"{key} group {g}")
                continue
            got = np.asarray(res[key][g])
            scale = np.abs(want).max(axis=-1, keepdims=True)
            if key.endswith("noautos"):
                all_ref = np.abs(np.asarray(ref[key.replace("noautos", "all")][g])).max(axis=-1, keepdims=True)
                scale = np.maximum(scale, all_ref * n / (n - 1))
            err = np.abs(got - want)
            assert np.all(err <= rtol * scale), (
                f"{key} group {g} (n={n}): max err/scale = {(err / scale).max():.3e} (rtol {rtol})")


@pytest.mark.parametrize(
    "N,nwin,precision",
    [(64, 2, "complex64"), (128, 1, "complex64"), (256, 4, "complex64"), (512, 1, "complex64"),
     (1024, 1, "complex64"), (2048, 1, "complex64"),
     (64, 1, "complex128"), (256, 2, "complex128"), (1024, 1, "complex128"), (2048, 1, "complex128")],
)
def test_fused_engine_matches_oracle_injected_noise(cuda_device, N, nwin, precision):
    group_sizes = [5, 1, 2, 9] if N <= 256 else [3, 1, 4]
    T = 3 if N <= 512 else 2
    a = make_array(100 + N, 2, group_sizes, T, N * nwin, zero_ns_frac=0.03)
    wins = [(w * N, (w + 1) * N - 1) for w in range(nwin)]
    rng = np.random.default_rng(5)
    noise_in = (rng.standard_normal(a["vis"].shape) + 1j * rng.standard_normal(a["vis"].shape)).astype(np.complex64)
    pipe, res = run_pipeline(a, wins, precision, "inject", noise_in=noise_in)
    assert pipe.fused
    noise_freq = np.stack([noise_in[..., w * N:(w + 1) * N] for w in range(nwin)])
    ref = engine_ref.group_means(a["vis"], a["flags"], a["nsample"], a["int_time"], a["group_offset"],
                                 a["freq"], wins, a["pols"], 144.0, 0.76, noise_freq=noise_freq)
    rtol = RTOL_C64 if precision == "complex64" else RTOL_C128
    check_against_oracle(res, ref, group_sizes, rtol, ["power_all", "power_noautos"])
    # the noise leg runs in fp32 in both precisions (its parity is statistical in production)
    check_against_oracle(res, ref, group_sizes, RTOL_C64, ["noise_power_all", "noise_power_noautos"])
    check_against_oracle(res, ref, group_sizes, RTOL_C64 if precision == "complex64" else 1e-10 / 4,
                         ["thermal_all", "thermal_noautos"])
    assert np.abs(res["power_all"].imag).max() == 0


@pytest.mark.parametrize("N,precision,mode", [(256, "complex64", "philox"), (128, "complex128", "philox"),
                                              (1024, "complex64", "philox"), (64, "complex64", "philox"),
                                              (256, "complex64", "philox_gauss"), (128, "complex128", "philox_gauss")])
def test_fused_engine_philox_noise_matches_restated_stream(cuda_device, N, precision, mode):
    """Both in-kernel draws against their restated streams (oracle/engine_ref.py), value by value
    through the noise pair sums: "philox" = the moment-matched discrete law, "philox_gauss" =
    Box-Muller per channel."""
    group_sizes = [4, 7, 1, 6]
    a = make_array(7, 2, group_sizes, 3, 2 * N, zero_ns_frac=0.02)
    wins = [(0, N - 1), (N, 2 * N - 1)]
    pipe, res = run_pipeline(a, wins, precision, mode, seed=0x1234567890)
    restate = engine_ref.engine_noise_freq if mode == "philox_gauss" else engine_ref.engine_noise_freq_mm
    noise_freq = restate(pipe.sigma_coef32.cpu().numpy(), a["nsample"], a["int_time"], pipe.win_start, N,
                         0x1234567890)
    ref = engine_ref.group_means(a["vis"], a["flags"], a["nsample"], a["int_time"], a["group_offset"],
                                 a["freq"], wins, a["pols"], 144.0, 0.76, noise_freq=noise_freq)
    # fast-math rsqrt (and log/sincos for Box-Muller) in the kernel vs numpy: 1e-4 of the row maximum
    check_against_oracle(res, ref, group_sizes, 2.5e-5, ["noise_power_all", "noise_power_noautos"])
    # sigma of the reference's radiometer equation is what the stream is scaled by
    sig = ref["sigma"][1]  # (S,P,B,T,N) of group 1
    g1 = noise_freq[:, :, 4:11]
    ok = sig > 0
    ratio = np.mean(np.abs(g1[ok]) ** 2) / np.mean(sig[ok] ** 2)
    assert abs(ratio - 1.0) < 0.05  # <|sigma g|^2> = sigma^2  (t_utils.py:243-249 style)


def test_fused_engine_time_batches_and_shards_are_invariant(cuda_device):
    """Time chunks commute (t_io.py:570-631) -- also with the in-kernel noise, whose
    counters are global sample coordinates."""
    group_sizes = [6, 3]
    a = make_array(21, 2, group_sizes, 6, 256)
    wins = [(0, 255)]
    _, whole = run_pipeline(a, wins, "complex64", "philox")
    _, parts = run_pipeline(a, wins, "complex64", "philox", batches=[(0, 2), (2, 3), (3, 6)])
    # fp32 partial sums are grouped differently; the no-autos mean is a difference of
    # sums ~n times larger, so the criterion is norm-wise over the delay row
    for key in ("power_all", "noise_power_all"):
        assert_close(parts[key], whole[key], 2e-6, what=key)
    for key in ("power_noautos", "noise_power_noautos"):
        assert_close(parts[key], whole[key], 2e-5, what=key)
    np.testing.assert_allclose(parts["thermal_all"], whole["thermal_all"], rtol=2e-6)


def test_fused_engine_large_group_and_many_items(cuda_device):
    """A group wider than one item (n > 64 rows) and enough items for several waves."""
    group_sizes = [150, 40, 1, 1, 17]
    a = make_array(33, 1, group_sizes, 5, 256)
    wins = [(0, 127), (128, 255)]
    a2 = dict(a)
    _, res = run_pipeline(a2, wins, "complex64", None, thermal=True)
    # oracle via the factorised identities on the oracle transform (materialising 150^2 is slow)
    from oracle import simpleds_oracle as orc

    for g, n in enumerate(group_sizes):
        sl = slice(a["group_offset"][g], a["group_offset"][g + 1])
        for s, (w0, w1) in enumerate(wins):
            x = a["vis"][:, sl, :, w0:w1 + 1].astype(complex) * (~a["flags"][:, sl, :, w0:w1 + 1])
            X = orc.normalized_fourier_transform(x, 97656.25)
            s1 = X.sum(1)
            want_all = (np.abs(s1) ** 2).mean(1) / n**2
            assert_close(res["power_all"][g, s], want_all, RTOL_C64, what=f"all g{g}")
            if n > 1:
                s2 = (np.abs(X) ** 2).sum(1)
                want = ((np.abs(s1) ** 2 - s2) / (n * (n - 1))).mean(1)
                scale = np.maximum(np.abs(want).max(-1, keepdims=True), np.abs(want_all).max(-1, keepdims=True) * n / (n - 1))
                assert np.all(np.abs(res["power_noautos"][g, s] - want) <= RTOL_C64 * scale), f"noautos g{g}"


@pytest.mark.parametrize("wins,F0", [([(0, 202)], 203), ([(95, 115)], 203), ([(3, 66), (70, 133)], 160)])
def test_unfused_fallback_matches_oracle(cuda_device, wins, F0):
    """Window lengths / origins the fused path does not take (203 = 7 * 29 of the bundled
    PAPER files, unaligned starts) go through the chained API kernels."""
    group_sizes = [4, 1, 6]
    a = make_array(55, 2, group_sizes, 3, F0, zero_ns_frac=0.03, df=492610.84375)
    N = wins[0][1] - wins[0][0] + 1
    rng = np.random.default_rng(9)
    noise_in = (rng.standard_normal(a["vis"].shape) + 1j * rng.standard_normal(a["vis"].shape)).astype(np.complex64)
    for precision, rtol in (("complex128", RTOL_C128), ("complex64", RTOL_C64)):
        pipe, res = run_pipeline(a, wins, precision, "inject", noise_in=noise_in)
        assert not pipe.fused
        noise_freq = np.stack([noise_in[..., w0:w1 + 1] for w0, w1 in wins])
        ref = engine_ref.group_means(a["vis"], a["flags"], a["nsample"], a["int_time"], a["group_offset"],
                                     a["freq"], wins, a["pols"], 144.0, 0.76, noise_freq=noise_freq)
        check_against_oracle(res, ref, group_sizes, rtol,
                             ["power_all", "power_noautos", "noise_power_all", "noise_power_noautos"])
        check_against_oracle(res, ref, group_sizes, 1e-10 / 4, ["thermal_all", "thermal_noautos"])
    _, res = run_pipeline(a, wins, "complex64", "philox")
    assert np.all(np.isfinite(res["noise_power_all"])) and res["noise_power_all"].real.min() > 0


def test_process_host_chunks_match_resident_path(cuda_device):
    """RedundantPipeline.process_host (pinned host arrays, uploaded in chunks of whole groups
    on a copy stream) == process() on device-resident arrays, noise realisation included
    (its counters are keyed by the global baseline index, whatever the chunking)."""
    group_sizes = [9, 1, 4, 30, 2, 6]
    a = make_array(77, 2, group_sizes, 3, 256, zero_ns_frac=0.02)
    wins = [(0, 127), (128, 255)]
    _, whole = run_pipeline(a, wins, "complex64", "philox", seed=11)
    from simpleds_b200.engine import RedundantPipeline

    pipe = RedundantPipeline(a["freq"], wins, a["pols"], a["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                             precision="complex64", seed=11)
    host = [torch.as_tensor(a[k]).pin_memory() for k in ("vis", "flags", "nsample", "int_time")]
    per_bl = 13 * 2 * 3 * 256
    assert len(pipe._host_chunks(3, 12 * per_bl)) >= 4  # several chunks, some holding several groups
    got = pipe.process_host(*host, noise="philox", thermal=True, chunk_bytes=12 * per_bl)
    for key in ("power_all", "power_noautos", "noise_power_all", "noise_power_noautos", "thermal_all",
                "thermal_noautos"):
        np.testing.assert_allclose(got[key], whole[key], rtol=1e-12, atol=0, equal_nan=True, err_msg=key)
    # a second call accumulates (streaming) and reuses the staging buffers
    got2 = pipe.process_host(*host, noise=None, thermal=False, chunk_bytes=12 * per_bl)
    assert pipe.ntimes_done == 6 and np.all(np.isfinite(got2["power_all"]))


@pytest.mark.parametrize("N,F0,starts", [(203, 208, [0]), (21, 64, [0, 32]), (100, 224, [16, 112]), (11, 16, [0]),
                                         (600, 608, [0])])
def test_chirpz_engine_matches_oracle(cuda_device, N, F0, starts):
    """Window lengths that are not powers of two (21 and 203 channels of the reference's bundled
    files) go through the fused chirp-z engine in complex64; complex128 keeps the chained kernels."""
    group_sizes = [5, 1, 2, 9] if N < 300 else [3, 1]
    a = make_array(300 + N, 2, group_sizes, 3, F0, zero_ns_frac=0.03, df=492610.84375)
    wins = [(s, s + N - 1) for s in starts]
    rng = np.random.default_rng(6)
    noise_in = (rng.standard_normal(a["vis"].shape) + 1j * rng.standard_normal(a["vis"].shape)).astype(np.complex64)
    pipe, res = run_pipeline(a, wins, "complex64", "inject", noise_in=noise_in)
    assert pipe.fused
    noise_freq = np.stack([noise_in[..., s:s + N] for s in starts])
    ref = engine_ref.group_means(a["vis"], a["flags"], a["nsample"], a["int_time"], a["group_offset"],
                                 a["freq"], wins, a["pols"], 144.0, 0.76, noise_freq=noise_freq)
    check_against_oracle(res, ref, group_sizes, RTOL_C64,
                         ["power_all", "power_noautos", "noise_power_all", "noise_power_noautos",
                          "thermal_all", "thermal_noautos"])
    # in-kernel noise: the restated stream of the chirp-z engine, value by value
    pipe, res = run_pipeline(a, wins, "complex64", "philox", seed=0xABCDEF0123)
    nz = engine_ref.engine_noise_freq_chirpz(pipe.sigma_coef32.cpu().numpy(), a["nsample"], a["int_time"],
                                             pipe.win_start, N, 0xABCDEF0123)
    ref = engine_ref.group_means(a["vis"], a["flags"], a["nsample"], a["int_time"], a["group_offset"],
                                 a["freq"], wins, a["pols"], 144.0, 0.76, noise_freq=nz)
    check_against_oracle(res, ref, group_sizes, 2.5e-5, ["noise_power_all", "noise_power_noautos"])
    # batches commute here too
    _, parts = run_pipeline(a, wins, "complex64", "philox", seed=0xABCDEF0123, batches=[(0, 1), (1, 3)])
    for key in ("power_all", "noise_power_all"):
        assert_close(parts[key], res[key], 2e-6, what=key)
    pipe128, _ = run_pipeline(a, wins, "complex128", None)
    assert not pipe128.fused


def test_fused_engine_empty_batch_is_a_no_op(cuda_device):
    """A streamed job may hand over a batch with no times; the sums must not move."""
    a = make_array(5, 1, [3, 2], 2, 256)
    pipe, before = run_pipeline(a, [(0, 255)], "complex64", "philox")
    dev = pipe.device
    z = lambda x: torch.as_tensor(x, device=dev)[:, :, :0].contiguous()  # noqa: E731
    pipe.process(z(a["vis"]), z(a["flags"]), z(a["nsample"]),
                 torch.as_tensor(a["int_time"], device=dev)[:, :0].contiguous())
    torch.cuda.synchronize()
    after = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in pipe.result().items()}
    for k, v in before.items():
        if isinstance(v, np.ndarray):
            assert np.array_equal(v, after[k], equal_nan=True), k


def test_fused_engine_on_a_second_device_of_the_same_process(cuda_device):
    """Kernel attributes (dynamic shared memory) are per device: one process driving two GPUs must
    get the same sums on both."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from simpleds_b200.engine import RedundantPipeline
    a = make_array(11, 2, [4, 3], 2, 256)
    out = []
    for d in (0, 1):
        with torch.cuda.device(d):
            dev = torch.device("cuda", d)
            pipe = RedundantPipeline(a["freq"], [(0, 255)], a["pols"], a["group_offset"], trcvr_k=144.0,
                                     beam_area_sr=0.76, precision="complex64", seed=3)
            pipe.process(*(torch.as_tensor(a[k], device=dev) for k in ("vis", "flags", "nsample", "int_time")))
            torch.cuda.synchronize()
            out.append({k: v.cpu().numpy() for k, v in pipe.result().items() if hasattr(v, "cpu")})
    for k in out[0]:
        assert np.array_equal(out[0][k], out[1][k], equal_nan=True), k


def test_fused_engine_group_subsets_equal_the_whole_job(cuda_device):
    """Sharding by redundant group (every group is its own object in the reference,
    ds.py:891-900): two calls that each hold ONLY the baselines of their groups give the sums of
    the whole job -- noise included, its stream being keyed by the global baseline index -- and a
    process that saw only some groups reports NaN means for the others."""
    from simpleds_b200.engine import RedundantPipeline

    group_sizes = [6, 1, 9, 3, 17, 2]
    a = make_array(91, 2, group_sizes, 3, 512, zero_ns_frac=0.02)
    wins = [(0, 255), (256, 511)]
    whole_pipe, whole = run_pipeline(a, wins, "complex64", "philox", seed=5)
    dev = torch.device("cuda:0")
    pipe = RedundantPipeline(a["freq"], wins, a["pols"], a["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                             precision="complex64", seed=5)
    go = a["group_offset"]
    for mine in ([0, 2, 3], [1, 4, 5]):
        bl = np.concatenate([np.arange(go[g], go[g + 1]) for g in mine])
        pipe.process(torch.as_tensor(a["vis"][:, bl], device=dev), torch.as_tensor(a["flags"][:, bl], device=dev),
                     torch.as_tensor(a["nsample"][:, bl], device=dev), torch.as_tensor(a["int_time"][bl], device=dev),
                     noise="philox", groups=mine)
        if mine == [0, 2, 3]:
            part = pipe.result()
            assert bool(torch.isnan(part["power_all"][1]).all()) and bool(torch.isfinite(part["power_all"][0]).all())
    torch.cuda.synchronize()
    for k, v in whole_pipe.sums().items():
        np.testing.assert_allclose(pipe.sums()[k].cpu().numpy(), v.cpu().numpy(), rtol=1e-12, atol=0, err_msg=k)
    with pytest.raises(ValueError):
        pipe.process(torch.as_tensor(a["vis"], device=dev), torch.as_tensor(a["flags"], device=dev),
                     torch.as_tensor(a["nsample"], device=dev), torch.as_tensor(a["int_time"], device=dev),
                     groups=[2, 1])
    with pytest.raises(ValueError):  # a later batch of a streamed job must say how long the job is
        pipe.process(torch.as_tensor(a["vis"], device=dev), torch.as_tensor(a["flags"], device=dev),
                     torch.as_tensor(a["nsample"], device=dev), torch.as_tensor(a["int_time"], device=dev),
                     time_offset=3)


@pytest.mark.parametrize("N,nwin", [(64, 2), (256, 2), (1024, 1)])
def test_fused_engine_two_data_sets_match_oracle(cuda_device, N, nwin):
    """Nuv = 2 (ds.py:3620-3633, thermal 3655-3662): power[i, j] = A2_i conj(A1_j) with A1 / A2 the
    transforms of the two data sets; the fused pair means are complex.  Injected noise for both sets
    (exact), then the in-kernel draw against its restated stream (one realisation per set)."""
    from simpleds_b200.engine import RedundantPipeline

    group_sizes = [5, 1, 2, 9] if N <= 256 else [3, 1, 4]
    T = 3 if N <= 256 else 2
    a, b = (make_array(500 + N + u, 2, group_sizes, T, N * nwin, zero_ns_frac=0.03) for u in (0, 1))
    wins = [(w * N, (w + 1) * N - 1) for w in range(nwin)]
    vis, flags, ns = (np.stack([a[k], b[k]]) for k in ("vis", "flags", "nsample"))
    rng = np.random.default_rng(8)
    noise_in = (rng.standard_normal(vis.shape) + 1j * rng.standard_normal(vis.shape)).astype(np.complex64)
    dev = torch.device("cuda:0")

    def run(noise, seed=7):
        pipe = RedundantPipeline(a["freq"], wins, a["pols"], a["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                                 precision="complex64", seed=seed, nuv=2)
        assert pipe.fused
        pipe.process(torch.as_tensor(vis, device=dev), torch.as_tensor(flags, device=dev), torch.as_tensor(ns, device=dev),
                     torch.as_tensor(a["int_time"], device=dev), noise=noise,
                     noise_in=torch.as_tensor(noise_in, device=dev) if noise == "inject" else None)
        torch.cuda.synchronize()
        return pipe, {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in pipe.result().items()}

    pipe, res = run("inject")
    noise_freq = np.stack([noise_in[..., w * N:(w + 1) * N] for w in range(nwin)])  # (S,2,P,B,T,N)
    ref = engine_ref.group_means(vis, flags, ns, a["int_time"], a["group_offset"], a["freq"], wins, a["pols"],
                                 144.0, 0.76, noise_freq=noise_freq)
    assert np.iscomplexobj(res["power_all"]) and np.abs(res["power_all"].imag).max() > 0
    check_against_oracle(res, ref, group_sizes, RTOL_C64,
                         ["power_all", "power_noautos", "noise_power_all", "noise_power_noautos",
                          "thermal_all", "thermal_noautos"])  # the thermal sums ride in the noise launch (fp32)
    # in-kernel draw: the restated stream of each set
    pipe, res = run("philox", seed=0x51ED5EED)
    nz = np.stack([engine_ref.engine_noise_freq_mm(pipe.sigma_coef32.cpu().numpy(), ns[u], a["int_time"],
                                                   pipe.win_start, N, 0x51ED5EED, nuv=2, uv=u) for u in (0, 1)], axis=1)
    ref = engine_ref.group_means(vis, flags, ns, a["int_time"], a["group_offset"], a["freq"], wins, a["pols"],
                                 144.0, 0.76, noise_freq=nz)
    check_against_oracle(res, ref, group_sizes, 2.5e-5, ["noise_power_all", "noise_power_noautos"])
    check_against_oracle(res, ref, group_sizes, RTOL_C64, ["thermal_all", "thermal_noautos"])
    # without a noise leg the thermal estimate goes through the chained kernels (fp64)
    _, res = run(None)
    check_against_oracle(res, ref, group_sizes, RTOL_C64, ["power_all", "power_noautos"])
    check_against_oracle(res, ref, group_sizes, 1e-10 / 4, ["thermal_all", "thermal_noautos"])
    with pytest.raises(NotImplementedError):
        RedundantPipeline(a["freq"], wins, a["pols"], a["group_offset"], trcvr_k=144.0, beam_area_sr=0.76,
                          precision="complex128", nuv=2).process(
            torch.as_tensor(vis, device=dev), torch.as_tensor(flags, device=dev), torch.as_tensor(ns, device=dev),
            torch.as_tensor(a["int_time"], device=dev))


@pytest.mark.parametrize("cplx", [False, True])
def test_engine_means_match_the_numpy_formulas(cuda_device, cplx):
    """sds_engine_means (one launch) against the mean over ordered pairs and times written out in numpy
    (tutorial cell 41; utils.remove_auto_correlations): groups of 1 baseline (no cross pairs: NaN), a group
    that saw no time (NaN), per-group time counts, the ntimes override, complex sums of two data sets."""
    import ctypes as C

    from simpleds_b200 import _lib
    from simpleds_b200._lib import check, ptr, stream_ptr

    rng = np.random.default_rng(5)
    G, S, P, N = 5, 2, 2, 64
    sizes = np.array([4, 1, 7, 2, 3], np.int32)
    tdone = np.array([3.0, 3.0, 0.0, 5.0, 3.0])
    L = N * (2 if cplx else 1)
    sums = {k: rng.standard_normal((G, S, P, L)) for k in ("pa", "pu", "na", "nu")}
    th = {k: rng.standard_normal((G, S, P)) for k in ("ta", "tu")}
    dev = cuda_device
    d = {k: torch.as_tensor(v, device=dev) for k, v in {**sums, **th}.items()}
    outs = {k: torch.empty_like(d["pa"]) for k in ("pa", "pn", "na", "nn")}
    outs.update({k: torch.empty_like(d["ta"]) for k in ("ta", "tn")})
    gs, td = torch.as_tensor(sizes, device=dev), torch.as_tensor(tdone, device=dev)
    lib = _lib.load()
    for ntimes in (0.0, 4.0):
        check(lib.sds_engine_means(G, S * P, L, ptr(gs), ptr(td), ntimes, ptr(d["pa"]), ptr(d["pu"]), ptr(d["na"]),
                                   ptr(d["nu"]), ptr(d["ta"]), ptr(d["tu"]), ptr(outs["pa"]), ptr(outs["pn"]),
                                   ptr(outs["na"]), ptr(outs["nn"]), ptr(outs["ta"]), ptr(outs["tn"]), stream_ptr()),
              "sds_engine_means")
        torch.cuda.synchronize()
        n = sizes.astype(float)
        T = np.full(G, ntimes) if ntimes > 0 else np.where(tdone > 0, tdone, np.nan)
        with np.errstate(divide="ignore", invalid="ignore"):
            for a, u, oa, on, shape in (("pa", "pu", "pa", "pn", (G, 1, 1, 1)), ("na", "nu", "na", "nn", (G, 1, 1, 1)),
                                        ("ta", "tu", "ta", "tn", (G, 1, 1))):
                A = {**sums, **th}[a]
                U = {**sums, **th}[u]
                nn_, TT = n.reshape(shape), T.reshape(shape)
                want_all = A / (nn_ * nn_ * TT)
                want_no = np.where(nn_ > 1, (A - U) / (nn_ * (nn_ - 1) * TT), np.nan)
                np.testing.assert_allclose(outs[oa].cpu().numpy(), want_all, rtol=1e-14, equal_nan=True)
                np.testing.assert_allclose(outs[on].cpu().numpy(), want_no, rtol=1e-14, equal_nan=True)
    # a sum without its partner is refused
    assert lib.sds_engine_means(G, S * P, L, ptr(gs), ptr(td), 0.0, ptr(d["pa"]), None, None, None, None, None,
                                ptr(outs["pa"]), ptr(outs["pn"]), None, None, None, None, stream_ptr()) != 0
