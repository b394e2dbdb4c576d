"""GPU parity tests of the reference-facing API (utils mirror + DelaySpectrum)
against the oracle and the golden vectors.  Everything here goes through the C
ABI of libsds.so."""
import os
import warnings

import numpy as np
import pytest
import torch
from scipy.signal import windows

from helpers import RTOL_C64, RTOL_C128, assert_close, make_group
from oracle import philox as ophilox
from oracle import simpleds_oracle as orc

pytestmark = pytest.mark.gpu

FT_SIZES = (5, 11, 21, 31, 64, 203, 256, 1024)


def _q(v, u=""):
    from simpleds_b200.units import Quantity

    return Quantity(v, u)


# ----------------------------------------------------------------- transform
@pytest.mark.parametrize("n", FT_SIZES)
def test_ft_matches_reference_golden(cuda_device, golden, n):
    from simpleds_b200 import utils

    x, dx = golden[f"ft_in_{n}"], float(golden[f"ft_dx_{n}"])
    out = utils.normalized_fourier_transform(x, _q(dx, "Hz"))
    assert_close(out.numpy(), golden[f"ft_fwd_bh_{n}"], RTOL_C128, what=f"fwd bh N={n}")
    out = utils.normalized_fourier_transform(x, _q(dx, "Hz"), taper=windows.boxcar)
    assert_close(out.numpy(), golden[f"ft_fwd_box_{n}"], RTOL_C128, what=f"fwd box N={n}")
    out = utils.normalized_fourier_transform(x, _q(1e-9, "s"), inverse=True)
    assert_close(out.numpy(), golden[f"ft_inv_bh_{n}"], RTOL_C128, what=f"inv N={n}")
    # complex64 mode (fp32 math)
    out = utils.normalized_fourier_transform(x.astype(np.complex64), _q(dx, "Hz"))
    assert out.dtype == np.complex64
    ref32 = orc.normalized_fourier_transform(x.astype(np.complex64).astype(complex), dx)
    assert_close(out.numpy(), ref32, RTOL_C64, what=f"fwd c64 N={n}")


def test_ft_kats_and_units(cuda_device, golden):
    from simpleds_b200 import utils
    from simpleds_b200.units import Unit

    fake = np.zeros((1, 13, 21))
    fake[0, 7, 11] += 1
    out = utils.normalized_fourier_transform(fake, _q(1.0), taper=windows.boxcar, axis=2)
    assert np.allclose(out.numpy(), golden["ft_kat_delta"])  # t_utils.py:379-389
    fake = np.zeros((3, 2, 13, 31))
    fake[:, 0, 7, 11] += 1.0
    out = utils.normalized_fourier_transform(fake, _q(1.0), taper=windows.boxcar, axis=3)
    assert out.shape == (3, 2, 13, 31)  # t_utils.py:392-399
    assert np.allclose(out.numpy(), np.fft.fftshift(np.fft.fft(fake, axis=-1), axes=-1))
    out = utils.normalized_fourier_transform(_q(fake, "m"), _q(1.0, "Hz"), taper=windows.boxcar, axis=3)
    assert out.unit == Unit("m Hz")  # t_utils.py:415-423
    with pytest.raises(ValueError):  # reference quirk: last axis only
        utils.normalized_fourier_transform(np.zeros((4, 21, 3)), _q(2.5), axis=1)
    # the reference multiplies by delta_x.si (utils.py:560): 0.5 MHz is 5e5 Hz, 4 ns is 4e-9 s
    hz = utils.normalized_fourier_transform(_q(fake, "Jy"), _q(5e5, "Hz"), axis=3)
    mhz = utils.normalized_fourier_transform(_q(fake, "Jy"), _q(0.5, "MHz"), axis=3)
    assert mhz.unit == Unit("Jy Hz") and np.allclose(mhz.numpy(), hz.numpy(), rtol=1e-14, atol=0)
    nsec = utils.normalized_fourier_transform(_q(fake, "Jy"), _q(4.0, "ns"), axis=3)
    assert nsec.unit == Unit("Jy s") and np.allclose(nsec.numpy(), hz.numpy() * (4e-9 / 5e5), rtol=1e-13, atol=0)


@pytest.mark.parametrize("n,f0", [(21, 21), (10, 21), (64, 203), (256, 1024), (203, 203), (128, 512),
                                  (512, 515), (1024, 1024), (2048, 2048)])
def test_transform_flags_and_windows(cuda_device, n, f0):
    from simpleds_b200 import ops
    from simpleds_b200._lib import SDS_F32, SDS_F64

    rng = np.random.default_rng(n * 7 + f0)
    x = (rng.standard_normal((3, 4, f0)) + 1j * rng.standard_normal((3, 4, f0))).astype(np.complex64)
    fl = rng.random((3, 4, f0)) < 0.25
    nwin = max(1, f0 // n)
    starts = [w * n for w in range(nwin)]
    if f0 - n > 3:
        starts.append(3)  # unaligned, overlapping window (t_io.py:719-722)
    df = 97656.25
    ref = np.stack([
        orc.normalized_fourier_transform(x[..., s:s + n].astype(complex) * (~fl[..., s:s + n]).astype(float), df)
        for s in starts
    ])
    xd, fd = torch.as_tensor(x, device=cuda_device), torch.as_tensor(fl, device=cuda_device)
    out = ops.delay_transform(ops.get_plan(n, SDS_F64, windows.blackmanharris), xd, fd, df, win_start=starts)
    assert_close(out.cpu().numpy(), ref, RTOL_C128, what="windows c128")
    out = ops.delay_transform(ops.get_plan(n, SDS_F32, windows.blackmanharris), xd, fd, df, win_start=starts)
    assert out.dtype == torch.complex64
    assert_close(out.cpu().numpy(), ref, RTOL_C64, what="windows c64")


def test_take_windows_bit_exact(cuda_device):
    from simpleds_b200 import ops

    rng = np.random.default_rng(5)
    for dtype in (np.complex128, np.complex64, np.float64, np.float32, bool):
        a = rng.standard_normal((2, 1, 2, 3, 4, 10))
        a = (a > 0) if dtype is bool else a.astype(dtype)
        for wins in ([(0, 9), (10, 19)], [(2, 6), (4, 8), (13, 17)], [(0, 19)]):
            _, chans = orc.normalize_spectral_windows(wins, 10)
            ref = orc.take_spectral_windows_from_data_like_array(a, chans)
            out = ops.take_windows(torch.as_tensor(a, device=cuda_device), chans)
            assert np.array_equal(out.cpu().numpy(), ref)


# ----------------------------------------------------------------- cross multiply
def test_cross_multiply_matches_reference_golden(cuda_device, golden):
    from simpleds_b200 import utils
    from simpleds_b200.units import Unit

    a1, a2 = golden["xm_a1"], golden["xm_a2"]
    for out, key in (
        (utils.cross_multiply_array(a1, axis=2), "xm_auto_axis2"),
        (utils.cross_multiply_array(a1, a2, axis=2), "xm_cross_axis2"),
        (utils.cross_multiply_array(a1, a2, axis=0), "xm_cross_axis0"),
    ):
        assert out.shape == golden[key].shape
        assert np.allclose(out, golden[key], rtol=1e-14, atol=1e-14)
    # t_utils.py:192-233
    with pytest.raises(ValueError):
        utils.cross_multiply_array(np.zeros((1, 2, 3)), np.zeros((2, 3, 4)), 2)
    assert utils.cross_multiply_array(np.ones((1, 3)), axis=1).shape == (1, 3, 3)
    assert utils.cross_multiply_array(np.ones((1, 3)).tolist(), axis=1).shape == (1, 3, 3)
    assert utils.cross_multiply_array(np.ones((1, 3)), np.ones((1, 3)).tolist(), axis=1).shape == (1, 3, 3)
    q = utils.cross_multiply_array(_q(np.ones((1, 3)), "Hz"), axis=1)
    assert q.shape == (1, 3, 3) and q.unit == Unit("Hz") ** 2


def test_combine_nsamples_and_remove_autos(cuda_device, golden):
    from simpleds_b200 import utils

    n1, n2 = golden["ns_n1"], golden["ns_n2"]
    assert np.allclose(utils.combine_nsamples(n1, axis=2), golden["ns_auto_axis2"], rtol=1e-15)
    assert np.allclose(utils.combine_nsamples(n1, n2, axis=2), golden["ns_cross_axis2"], rtol=1e-15)
    with pytest.raises(ValueError):  # t_utils.py:446-451
        utils.combine_nsamples(np.ones((2, 13, 21)), np.ones((3, 13, 21)))
    assert np.allclose(utils.combine_nsamples(np.ones((2, 13, 21)) * 3, axis=0), 3)
    out = utils.combine_nsamples(np.ones((3, 2, 13, 21)) * 3, np.ones((3, 2, 13, 21)) * 2, axis=1)
    assert out.shape == (3, 2, 2, 13, 21) and np.allclose(out, np.sqrt(6), rtol=1e-15)
    out = utils.remove_auto_correlations(torch.as_tensor(golden["rac_in"], device=cuda_device), axes=(1, 2))
    assert np.array_equal(out.cpu().numpy(), golden["rac_out_axes12"])
    assert utils.remove_auto_correlations(np.ones((3, 3, 11, 21))).shape == (6, 11, 21)
    with pytest.raises(ValueError):
        utils.remove_auto_correlations(np.ones((3, 3, 11, 21)), axes=(0, 2))


# ----------------------------------------------------------------- noise
def test_generate_noise_injected_normals_match_reference(cuda_device, golden):
    from simpleds_b200 import utils

    sig = golden["gn_sigma"]
    np.random.seed(0)  # the reference's stream: all reals, then all imags (utils.py:501-504)
    g_re = np.random.normal(size=sig.shape)
    g_im = np.random.normal(size=sig.shape)
    out = utils.generate_noise(sig, normals=(g_re, g_im))
    assert np.allclose(out, golden["gn_seed0"], rtol=1e-15, atol=0)


def test_generate_noise_philox_stream_and_statistics(cuda_device):
    from simpleds_b200 import utils

    sig = np.full((100, 1000), 3.0)
    out = utils.generate_noise(sig, seed=1234)
    assert out.shape == sig.shape
    g1, g2 = ophilox.normal_pairs(1234, sig.size)
    ref = (3.0 * (g1 + 1j * g2) / np.sqrt(2)).reshape(sig.shape)
    assert np.allclose(out, ref, rtol=1e-10, atol=1e-10)
    # t_utils.py:243-249
    assert np.isclose(out.std(), 3, atol=out.std(1).std())
    assert abs(out.real.mean()) < 0.03 and abs(out.imag.mean()) < 0.03
    assert abs(np.mean(out.real * out.imag)) < 0.05
    again = utils.generate_noise(sig, seed=1234)
    assert np.array_equal(out, again)  # counter-based: reproducible
    other = utils.generate_noise(sig, seed=1235)
    assert not np.array_equal(out, other)


# ----------------------------------------------------------------- DelaySpectrum
def _build_ds(pr, precision="complex128", **kw):
    from simpleds_b200 import DelaySpectrum

    return DelaySpectrum.from_arrays(
        pr["vis"], pr["flags"], pr["nsample"], pr["freq_array_hz"], pr["integration_time_s"],
        pr["polarization_array"], trcvr=pr["trcvr_k"], beam_area=pr["beam_area_sr"],
        precision=precision, **kw,
    )


def _oracle_with_normals(pr, spectral_windows, g_re, g_im, taper=windows.blackmanharris):
    state = dict(
        data_array=pr["vis"].astype(np.complex128), noise_array=np.zeros(pr["vis"].shape, complex),
        nsample_array=pr["nsample"], flag_array=pr["flags"], freq_array=pr["freq_array_hz"],
        trcvr=pr["trcvr_k"], beam_area=pr["beam_area_sr"], beam_sq_area=pr["beam_area_sr"],
    )
    state = orc.select_spectral_windows(state, spectral_windows)
    sigma = orc.calculate_noise_power(state["freq_array"], state["trcvr"], state["beam_area"],
                                      pr["integration_time_s"], state["nsample_array"])
    state["noise_array"] = sigma * (g_re + 1j * g_im) / np.sqrt(2)
    res = orc.calculate_delay_spectrum(
        state["data_array"], state["noise_array"], state["flag_array"], state["nsample_array"],
        state["freq_array"], state["trcvr"], state["beam_area"], pr["integration_time_s"],
        pr["polarization_array"], taper,
    )
    res["sigma"], res["freq_noise"] = sigma, state["noise_array"]
    res["delay_array_ns"] = state["delay_array_ns"]
    return res


@pytest.mark.parametrize(
    "kw,wins,precision",
    [
        (dict(P=1, B=7, T=5, F=21), [(0, 20)], "complex128"),           # C1 shape family
        (dict(P=2, B=6, T=4, F=21), [(0, 9), (10, 19)], "complex128"),  # t_ds.py:37 windows
        (dict(P=2, B=9, T=3, F=203, zero_ns_frac=0.05, flag_frac=0.27), [(95, 115)], "complex128"),
        (dict(P=4, B=12, T=6, F=256), [(0, 63), (64, 127), (128, 191), (192, 255)], "complex128"),
        (dict(U=2, P=2, B=5, T=4, F=64), [(0, 63)], "complex128"),      # two data sets (Nuv=2)
        (dict(P=2, B=6, T=4, F=128), [(0, 127)], "complex64"),
    ],
)
def test_delay_spectrum_pipeline_matches_oracle(cuda_device, kw, wins, precision):
    from simpleds_b200.units import Unit

    pr = make_group(seed=11, **kw)
    rtol = RTOL_C128 if precision == "complex128" else RTOL_C64
    ds = _build_ds(pr, precision=precision)
    ds.select_spectral_windows(wins)
    S, N = len(wins), wins[0][1] - wins[0][0] + 1
    U, P, B, T = pr["vis"].shape[1:5]
    assert ds.data_array.shape == (S, U, P, B, T, N) and ds.Nspws == S and ds.Nfreqs == N
    rng = np.random.default_rng(0)
    g_re, g_im = rng.standard_normal(ds.data_array.shape), rng.standard_normal(ds.data_array.shape)
    ref = _oracle_with_normals(pr, wins, g_re, g_im)
    assert_close(ds.delay_array.value, ref["delay_array_ns"], 1e-12, what="delay_array")
    sigma = ds.calculate_noise_power()
    assert sigma.unit == Unit("Jy")
    assert_close(sigma.numpy(), ref["sigma"], 1e-12, what="sigma")
    ds.generate_noise(normals=(g_re, g_im))
    assert_close(ds.noise_array.numpy(), ref["freq_noise"], rtol if precision == "complex64" else 1e-12, what="noise")
    ds.calculate_delay_spectrum()
    ds.remove_cosmology()  # the raw products, as t_ds.py:1122-1147 compares them
    assert ds.data_type == "delay" and ds.data_array.unit == Unit("Jy Hz")
    assert_close(ds.data_array.numpy(), ref["data_array"], rtol, what="delay data")
    assert_close(ds.noise_array.numpy(), ref["noise_array"], rtol, what="delay noise")
    assert ds.power_array.shape == (S, P, B, B, T, N)  # t_ds.py:889-909
    assert ds.thermal_power.shape == (S, P, B, B, T)
    assert ds.power_array.unit == Unit("Jy Hz") ** 2
    # product of two transforms: tolerance doubles; scale per (pair) row
    assert_close(ds.power_array.numpy(), ref["power_array"], 4 * rtol, what="power")
    assert_close(ds.noise_power.numpy(), ref["noise_power"], 4 * rtol, what="noise power")
    assert_close(ds.thermal_power.numpy(), ref["thermal_power"], 1e-11, what="thermal")
    # averaged path == reference tensor then numpy mean (tutorial cell 41)
    for remove_autos in (False, True):
        for method in ("factorised", "explicit"):
            av = ds.calculate_averaged_delay_spectrum(remove_autos=remove_autos, method=method)
            for key, full in (("power", ref["power_array"]), ("noise_power", ref["noise_power"])):
                want = orc.average_power(full, remove_autos=remove_autos, over_time=True)
                scale_tol = 4 * rtol * (B if remove_autos else 1)
                assert_close(av[key].numpy(), want, scale_tol, what=f"avg {key} autos={remove_autos} {method}")


def test_delay_transform_round_trip_units_and_warning(cuda_device):
    from simpleds_b200.units import Unit

    pr = make_group(seed=2, P=1, B=3, T=2, F=21)
    ds = _build_ds(pr)
    ds.delay_transform()
    assert ds.data_array.unit == Unit("Jy Hz") and ds.data_type == "delay"
    ds.delay_transform()  # t_ds.py:829-839: back to Jy (inverse branch, not a true inverse)
    assert ds.data_array.unit == Unit("Jy") and ds.data_type == "frequency"
    fwd, _ = orc.ds_normalized_fourier_transform(pr["vis"].astype(complex), np.zeros(pr["vis"].shape, complex), pr["flags"], pr["freq_array_hz"], windows.blackmanharris)
    back, _ = orc.ds_normalized_fourier_transform(fwd, np.zeros(fwd.shape, complex), pr["flags"], pr["freq_array_hz"], windows.blackmanharris, inverse=True, delay_array_ns_=orc.delay_array_ns(pr["freq_array_hz"]))
    assert_close(ds.data_array.numpy(), back, RTOL_C128, what="inverse branch")
    from simpleds_b200 import DelaySpectrum

    un = DelaySpectrum.from_arrays(pr["vis"], pr["flags"], pr["nsample"], pr["freq_array_hz"],
                                   pr["integration_time_s"], pr["polarization_array"], vis_units="uncalib")
    with pytest.warns(UserWarning, match="Fourier Transforming uncalibrated data"):  # t_ds.py:854-860
        un.delay_transform()
    assert un.data_array.unit == Unit("Hz")
    un.data_type = "bogus"
    with pytest.raises(ValueError, match="Unknown data type"):
        un.delay_transform()


def test_select_spectral_windows_errors_and_copy(cuda_device):
    pr = make_group(seed=4, P=1, B=3, T=2, F=21)
    ds = _build_ds(pr)
    with pytest.raises(ValueError, match="only have two"):
        ds.select_spectral_windows([(0, 5, 7)])
    with pytest.raises(ValueError, match="same size"):
        ds.select_spectral_windows([(0, 5), (6, 13)])
    new = ds.select_spectral_windows([(1, 10)], inplace=False)
    assert new.Nfreqs == 10 and ds.Nfreqs == 21 and new.data_array.shape[-1] == 10
    # like the reference (ds.py:3306): the cosmological axes exist right after the selection
    assert new.redshift is not None and new.k_parallel is not None
    assert np.asarray(new.redshift).shape == (1,) and tuple(np.asarray(new.k_parallel.value).shape) == (1, 10)
    assert np.array_equal(new.data_array.numpy()[0], pr["vis"][0][..., 1:11].astype(complex))
    byfreq = ds.select_spectral_windows(freqs=pr["freq_array_hz"][:, 3:9], inplace=False)
    assert np.array_equal(byfreq.freq_array.value, pr["freq_array_hz"][:, 3:9])
    with pytest.raises(ValueError, match="not found in frequency array"):
        ds.select_spectral_windows(freqs=np.array([[1.0, 2.0]]))


def test_partition_invariance_time_pol_spw(cuda_device):
    # t_io.py:570-631, 680-736: slices along time / pol / spw commute with the path
    pr = make_group(seed=9, P=2, B=5, T=6, F=32)
    wins = [(0, 15), (16, 31)]
    full = _build_ds(pr)
    full.select_spectral_windows(wins)
    full.calculate_delay_spectrum()
    full.remove_cosmology()
    fp, ft = full.power_array.numpy(), full.thermal_power.numpy()

    def sub(tsl=slice(None), psl=slice(None), w=wins):
        q = dict(pr)
        q["vis"], q["flags"], q["nsample"] = (pr[k][:, :, psl][..., tsl, :] for k in ("vis", "flags", "nsample"))
        q["integration_time_s"] = pr["integration_time_s"][:, tsl]
        q["beam_area_sr"] = pr["beam_area_sr"][:, psl]
        q["polarization_array"] = pr["polarization_array"][psl]
        d = _build_ds(q)
        d.select_spectral_windows(w)
        d.calculate_delay_spectrum()
        d.remove_cosmology()
        return d.power_array.numpy(), d.thermal_power.numpy()

    p, t = sub(tsl=slice(2, 5))
    assert np.array_equal(p, fp[..., 2:5, :]) and np.array_equal(t, ft[..., 2:5])
    p, t = sub(psl=slice(1, 2))
    assert np.array_equal(p, fp[:, 1:2]) and np.array_equal(t, ft[:, 1:2])
    p, t = sub(w=[wins[1]])
    assert np.array_equal(p, fp[1:2]) and np.array_equal(t, ft[1:2])


def test_uvdatalite_ingest_ordering(cuda_device):
    """ds.py:913-973: baselines sorted by baseline number, times by unwrapped LST."""
    from simpleds_b200 import DelaySpectrum, UVDataLite

    rng = np.random.default_rng(8)
    ants = [(1, 4), (0, 3), (44, 47), (2, 5)]
    T, F, P = 5, 21, 1
    lst = (6.2 + 0.05 * np.arange(T)) % (2 * np.pi)  # wraps through 2 pi
    order = rng.permutation(len(ants))
    a1, a2, lsts, data, uvw = [], [], [], [], []
    for t in range(T):
        for b in order:
            a1.append(ants[b][0]); a2.append(ants[b][1]); lsts.append(lst[t])
            data.append((b + 1) * 100 + t + 1j * np.arange(F))
            uvw.append([30.0, 0.0, 0.0])
    data = np.array(data)[:, :, None]
    uv = UVDataLite(data, np.zeros(data.shape, bool), np.full(data.shape, 40.0), a1, a2, lsts,
                    np.full(len(a1), 42.95), 146.8e6 + 492610.84375 * np.arange(F), [1], uvw_array=uvw)
    ds = DelaySpectrum(uv=uv)
    bl = orc.baseline_number(np.array([a[0] for a in ants]), np.array([a[1] for a in ants]))
    assert np.array_equal(ds.baseline_array, np.sort(bl))
    assert 69637 in ds.baseline_array  # (1,4) <-> 69637 (t_ds.py:1532)
    assert ds.data_array.shape == (1, 1, 1, 4, T, F)
    got = ds.data_array.numpy()[0, 0, 0, :, :, 0].real
    srt = np.argsort(bl)
    want = np.array([[(b + 1) * 100 + t for t in range(T)] for b in srt], dtype=float)
    assert np.array_equal(got, want)
    assert np.all(np.diff(ds.lst_array.value) > 0)
    assert np.all(ds.noise_array.numpy() == 0)  # beam/trcvr are inf at load (ds.py:975-987)
    # two groups -> ValueError (ds.py:894-900)
    uvw2 = np.array(uvw); uvw2[::4] = [60.0, 0, 0]
    uv2 = UVDataLite(data, np.zeros(data.shape, bool), np.full(data.shape, 40.0), a1, a2, lsts,
                     np.full(len(a1), 42.95), 146.8e6 + 492610.84375 * np.arange(F), [1], uvw_array=uvw2)
    with pytest.raises(ValueError, match="single baseline vector"):
        DelaySpectrum(uv=uv2)


def test_spectrum_on_no_data_and_twice(cuda_device):
    """t_ds.py:742-745: no data -> ValueError; t_ds.py:1389-1436: calculate_delay_spectrum may be
    called twice in a row (the second call finds delay-domain data and skips the transform,
    ds.py:3617-3618) and gives the same power."""
    from simpleds_b200 import DelaySpectrum

    with pytest.raises(ValueError, match="No data has be loaded"):
        DelaySpectrum().calculate_delay_spectrum()
    with pytest.raises(ValueError):  # t_ds.py:736-739: trcvr alone cannot initialise
        DelaySpectrum(trcvr=9.0)
    pr = make_group(seed=6, P=2, B=4, T=3, F=21)
    ds = _build_ds(pr)
    ds.select_spectral_windows([(1, 3), (4, 6)])
    ds.calculate_delay_spectrum()
    first = ds.power_array.numpy().copy()
    thermal = ds.thermal_power.numpy().copy()
    ds.calculate_delay_spectrum()
    assert ds.data_type == "delay"
    assert np.array_equal(ds.power_array.numpy(), first)
    assert np.array_equal(ds.thermal_power.numpy(), thermal)


@pytest.mark.gpu
@pytest.mark.parametrize("P,F,dtype", [(1, 21, np.complex64), (2, 203, np.complex64), (4, 5, np.float32),
                                        (4, 1024, np.complex64), (2, 64, np.uint8), (3, 17, np.complex128)])
def test_ingest_gather_is_the_per_baseline_transpose(cuda_device, P, F, dtype):
    """utils.py:15-168: data[:, c] = transpose(uv.data_array[inds], [2, 0, 1]) per baseline, with the
    LST sort of ds.py:913-916 -- one device gather, bit-exact."""
    import torch

    from simpleds_b200 import ops
    rng = np.random.default_rng(P * 1000 + F)
    B, T = 7, 5
    nblt = B * T
    if np.issubdtype(dtype, np.complexfloating):
        x = (rng.standard_normal((nblt, F, P)) + 1j * rng.standard_normal((nblt, F, P))).astype(dtype)
    elif dtype == np.uint8:
        x = rng.integers(0, 2, (nblt, F, P)).astype(np.uint8)
    else:
        x = rng.standard_normal((nblt, F, P)).astype(dtype)
    bl = np.tile(rng.permutation(B), T)                       # time-major blt axis, shuffled baselines
    lst_inds = rng.permutation(T)
    want = np.zeros((P, B, T, F), dtype=dtype)
    perm = np.zeros((B, T), dtype=np.int64)
    for c in range(B):
        inds = np.nonzero(bl == c)[0]
        want[:, c] = np.transpose(x[inds], [2, 0, 1])[:, lst_inds]
        perm[c] = inds[lst_inds]
    got = ops.ingest_gather(torch.as_tensor(x, device="cuda"), torch.as_tensor(perm.reshape(-1), device="cuda"))
    assert np.array_equal(got.cpu().numpy().reshape(P, B, T, F), want)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["complex128", "complex64"])
def test_noise_transform_is_the_chain_of_the_three_calls(cuda_device, precision):
    """sds_noise_transform == sds_noise_sigma -> sds_generate_noise ->
    sds_delay_transform, bit for bit, with injected normals and with the Philox stream."""
    import torch
    from scipy.signal import windows

    from simpleds_b200 import ops
    from simpleds_b200._lib import SDS_F32, SDS_F64
    rng = np.random.default_rng(77)
    S, U, P, B, T, F = 2, 1, 2, 3, 4, 32
    cdt = torch.complex128 if precision == "complex128" else torch.complex64
    plan = ops.get_plan(F, SDS_F64 if precision == "complex128" else SDS_F32, windows.blackmanharris)
    coef = torch.as_tensor(rng.uniform(1.0, 2.0, (S, P, F)), device="cuda")
    it = torch.as_tensor(rng.uniform(5.0, 10.0, (B, T)), device="cuda")
    ns = torch.as_tensor(rng.integers(0, 50, (S, U, P, B, T, F)).astype(np.float64), device="cuda")
    fl = torch.as_tensor(rng.random((S, U, P, B, T, F)) < 0.1, device="cuda")
    g = (rng.standard_normal(ns.shape), rng.standard_normal(ns.shape))
    for normals in (g, None):
        nf, nd = ops.noise_transform(plan, coef, it, ns, fl, 97656.25, seed=5, normals=normals, out_dtype=cdt)
        sig = ops.noise_sigma(coef, it, ns)
        want_f = ops.generate_noise(sig, seed=5, normals=normals, out_dtype=cdt)
        want_d = ops.delay_transform(plan, want_f.reshape(-1, F), fl.reshape(-1, F), 97656.25)
        assert torch.equal(nf, want_f) and torch.equal(nd.reshape(-1), want_d.reshape(-1))


def test_delay_spectrum_file_round_trip_and_write_partial(cuda_device, tmp_path):
    """delay_spectrum.py:2423-3194: write -> read gives the object back ((Jy Hz)^2 units: the
    cosmological conversion is taken out for the file and restored afterwards); a file initialised
    from the whole object and filled by write_partial from two time slices equals the whole write."""
    from simpleds_b200 import DelaySpectrum, hdf5lite
    from simpleds_b200.units import Unit

    pr = make_group(seed=12, P=2, B=4, T=6, F=32)
    lst = 0.1 + 1e-3 * np.arange(6)

    def build(sl):
        ds = DelaySpectrum.from_arrays(
            pr["vis"][..., sl, :], pr["flags"][..., sl, :], pr["nsample"][..., sl, :], pr["freq_array_hz"],
            pr["integration_time_s"][:, sl], pr["polarization_array"], trcvr=pr["trcvr_k"], beam_area=pr["beam_area_sr"],
            precision="complex128", lst_array=lst[sl], baseline_array=np.array([67611, 69637, 71714, 73763]),
            uvw=[30.0, 0.0, 0.0], noise_seed=3)
        ds.select_spectral_windows([(0, 15), (16, 31)])
        ds.calculate_delay_spectrum()
        return ds

    whole = build(slice(0, 6))
    assert whole.power_array.unit == Unit("mK^2 Mpc^3")
    path = str(tmp_path / "whole.h5")
    with pytest.warns(UserWarning, match="cosmological units"):
        whole.write(path)
    assert whole.power_array.unit == Unit("mK^2 Mpc^3")  # restored (ds.py:2739-2741)
    with pytest.raises(IOError):
        whole.write(path)
    f = hdf5lite.File(path)
    assert f.keys("/") == ["Data", "Header"] and f.attrs("Data/data_power")["unit"] == "Hz^2 Jy^2"
    assert f.shape("Data/data_power") == (2, 2, 4, 4, 6, 16) and f.read("Data/thermal_power").dtype == np.float64
    back = DelaySpectrum(precision="complex128").read(path)
    assert (back.Nspws, back.Nuv, back.Npols, back.Nbls, back.Ntimes, back.Nfreqs) == (2, 1, 2, 4, 6, 16)
    assert back.data_type == "delay" and back.vis_units == "Jy" and back.data_array.unit == Unit("Jy Hz")
    np.testing.assert_array_equal(back.data_array.numpy(), whole.data_array.numpy())
    np.testing.assert_array_equal(back.flag_array.cpu().numpy(), whole.flag_array.cpu().numpy())
    np.testing.assert_array_equal(back.freq_array.value, whole.freq_array.value)
    np.testing.assert_array_equal(back.baseline_array, whole.baseline_array)
    back.update_cosmology()  # the file's (Jy Hz)^2 power back to mK^2 Mpc^3
    np.testing.assert_allclose(back.power_array.numpy(), whole.power_array.numpy(), rtol=1e-12)
    np.testing.assert_allclose(back.thermal_power.numpy(), whole.thermal_power.numpy(), rtol=1e-12)
    # streamed sink: the file laid out once, filled by two time slices
    part = str(tmp_path / "partial.h5")
    whole.initialize_save_file(part)
    assert not hdf5lite.File(part).read("Data/data_power").any()
    with pytest.raises(AssertionError):
        whole.write_partial(str(tmp_path / "missing.h5"))
    with pytest.warns(UserWarning, match="cosmological units"):
        build(slice(0, 2)).write_partial(part)
        build(slice(2, 6)).write_partial(part)
    a, b = hdf5lite.File(path), hdf5lite.File(part)
    for name in ("visdata", "flags", "data_power", "thermal_power"):
        np.testing.assert_array_equal(a.read("Data/" + name), b.read("Data/" + name), err_msg=name)
    # the noise realisation of a time slice is its own draw: only its shape and unit are compared
    assert b.read("Data/visnoise").shape == a.read("Data/visnoise").shape
    np.testing.assert_allclose(a.read("Data/nsamples"), b.read("Data/nsamples"))


def test_uvdatalite_from_uvh5_feeds_add_uvdata(cuda_device, tmp_path):
    """The real-file front end: a UVH5 file (here written from the fixture cut out of the
    reference's test_redundant_array_odd_file.uvh5, old 4-D axis order) is read with the package's
    own HDF5 reader and goes through add_uvdata (baseline numbers of t_ds.py:271-289).
    tests/test_hdf5lite.py reads the h5py-written original where it exists."""
    from simpleds_b200 import DelaySpectrum, hdf5lite
    from simpleds_b200.delay_spectrum import UVDataLite

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "uvh5_odd.npz"))
    path = str(tmp_path / "odd.uvh5")
    w = hdf5lite.Writer(path)
    w.create_dataset("Data/visdata", g["data_array"][:, None])
    w.create_dataset("Data/flags", g["flag_array"][:, None])
    w.create_dataset("Data/nsamples", g["nsample_array"][:, None])
    for k in ("ant_1_array", "ant_2_array", "lst_array", "integration_time", "uvw_array", "polarization_array"):
        w.create_dataset("Header/" + k, g[k])
    w.create_dataset("Header/freq_array", g["freq_array"][None])
    w.create_dataset("Header/vis_units", str(g["vis_units"]).encode())
    w.create_dataset("Header/Nants_telescope", np.int64(g["Nants_telescope"]))
    w.close()
    uv = UVDataLite.from_uvh5(path)
    assert np.array_equal(uv.data_array, g["data_array"]) and uv.vis_units == "Jy" and uv.Nbls == 51
    ds = DelaySpectrum(uv=uv)
    assert ds.Nbls == 51 and ds.Ntimes == 21 and ds.Nfreqs == 21 and ds.baseline_array[0] == 67611
