"""The oracle reproduces outputs of the reference's own utils.py (golden vectors
made by tests/golden/make_golden.py) and the restatable KATs of the reference's
test suite (simpleDS/tests/test_utils.py, test_delay_spectrum.py, test_io.py)."""
import numpy as np
import pytest
from scipy.signal import windows

from oracle import simpleds_oracle as orc

FT_SIZES = (5, 11, 21, 31, 64, 203, 256, 1024)


@pytest.mark.parametrize("n", FT_SIZES)
def test_ft_matches_reference_bitwise(golden, n):
    x, dx = golden[f"ft_in_{n}"], float(golden[f"ft_dx_{n}"])
    assert np.array_equal(orc.normalized_fourier_transform(x, dx), golden[f"ft_fwd_bh_{n}"])
    assert np.array_equal(
        orc.normalized_fourier_transform(x, dx, taper=windows.boxcar),
        golden[f"ft_fwd_box_{n}"],
    )
    assert np.array_equal(
        orc.normalized_fourier_transform(x, 1e-9, inverse=True), golden[f"ft_inv_bh_{n}"]
    )


def test_ft_kat_delta_boxcar(golden):
    # simpleDS/tests/test_utils.py:379-389
    fake = np.zeros((1, 13, 21))
    fake[0, 7, 11] += 1
    out = orc.normalized_fourier_transform(fake, 1.0, taper=windows.boxcar, axis=2)
    assert np.array_equal(out, golden["ft_kat_delta"])
    assert np.allclose(out, np.fft.fftshift(np.fft.fft(fake, axis=-1), axes=-1))


def test_ft_only_last_axis_quirk(golden):
    assert bool(golden["ft_axis1_raises"])
    with pytest.raises(ValueError):
        orc.normalized_fourier_transform(np.zeros((4, 21, 3)), 2.5, axis=1)


def test_ft_closed_form():
    # Y[k'] = dx * sum_n x[n] w[n] exp(-2 pi i (k'-N//2) n / N)
    rng = np.random.default_rng(1)
    for n in (5, 8, 21):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        w = windows.blackmanharris(n)
        k = np.arange(n)[:, None] - n // 2
        direct = 3.0 * (np.exp(-2j * np.pi * k * np.arange(n)[None] / n) @ (x * w))
        assert np.allclose(orc.normalized_fourier_transform(x[None], 3.0)[0], direct)


def test_cross_multiply_matches_reference(golden):
    a1, a2 = golden["xm_a1"], golden["xm_a2"]
    assert np.array_equal(orc.cross_multiply_array(a1, axis=2), golden["xm_auto_axis2"])
    assert np.array_equal(orc.cross_multiply_array(a1, a2, axis=2), golden["xm_cross_axis2"])
    assert np.array_equal(orc.cross_multiply_array(a1, a2, axis=0), golden["xm_cross_axis0"])
    # index order: P[..., i, j, ...] = a2[i] * conj(a1[j])
    out = golden["xm_cross_axis2"]
    assert np.allclose(out[1, 0, 3, 1], a2[1, 0, 3] * np.conj(a1[1, 0, 1]))


def test_cross_multiply_kats():
    # simpleDS/tests/test_utils.py:192-219
    with pytest.raises(ValueError):
        orc.cross_multiply_array(np.zeros((1, 2, 3)), np.zeros((2, 3, 4)), 2)
    assert orc.cross_multiply_array(np.ones((1, 3)), axis=1).shape == (1, 3, 3)
    assert orc.cross_multiply_array(np.ones((1, 3)).tolist(), axis=1).shape == (1, 3, 3)
    assert orc.cross_multiply_array(np.ones((1, 3)), np.ones((1, 3)).tolist(), axis=1).shape == (1, 3, 3)


def test_combine_nsamples(golden):
    n1, n2 = golden["ns_n1"], golden["ns_n2"]
    assert np.array_equal(orc.combine_nsamples(n1, axis=2), golden["ns_auto_axis2"])
    assert np.array_equal(orc.combine_nsamples(n1, n2, axis=2), golden["ns_cross_axis2"])
    # simpleDS/tests/test_utils.py:446-467
    with pytest.raises(ValueError):
        orc.combine_nsamples(np.ones((2, 13, 21)), np.ones((3, 13, 21)))
    assert np.allclose(orc.combine_nsamples(np.ones((2, 13, 21)) * 3, axis=0), 3)
    out = orc.combine_nsamples(np.ones((3, 2, 13, 21)) * 3, np.ones((3, 2, 13, 21)) * 2, axis=1)
    assert np.all(out == np.ones((3, 2, 2, 13, 21)) * np.sqrt(6))


def test_remove_autos(golden):
    assert np.array_equal(
        orc.remove_auto_correlations(golden["rac_in"], axes=(1, 2)), golden["rac_out_axes12"]
    )
    # simpleDS/tests/test_utils.py:491-509
    assert orc.remove_auto_correlations(np.ones((3, 3, 11, 21))).shape == (6, 11, 21)
    assert orc.remove_auto_correlations(np.ones((4, 3, 3, 11, 21)), axes=(1, 2)).shape == (4, 6, 11, 21)
    with pytest.raises(ValueError):
        orc.remove_auto_correlations(np.ones((3, 3, 11, 21)), axes=(0, 2))


def test_generate_noise_stream_order(golden):
    np.random.seed(0)
    assert np.array_equal(orc.generate_noise(golden["gn_sigma"]), golden["gn_seed0"])
    # simpleDS/tests/test_utils.py:243-249
    np.random.seed(5)
    noise = orc.generate_noise(np.ones((100, 1000)) * 3)
    assert np.isclose(noise.std(), 3, atol=noise.std(1).std())


def test_small_helpers(golden):
    assert orc.noise_equivalent_bandwidth(windows.blackmanharris(203)) == float(golden["neb_bh_203"])
    wa, we = orc.weighted_average(golden["wavg_arr"], golden["wavg_err"], axis=-1)
    assert np.array_equal(wa, golden["wavg_out"]) and np.array_equal(we, golden["wavg_err_out"])
    # simpleDS/tests/test_utils.py:578-584
    a, e = orc.weighted_average(np.array([1.0, 3.0, 10.0]), np.array([1.0, 3.0, 10.0]))
    assert np.isclose(a, 1.2784935579781962) and np.isclose(e, 0.9444428250308379)


def test_delay_array_kats():
    # simpleDS/tests/test_delay_spectrum.py:1619 and test_io.py:294,314
    df = 492610.84375
    f21 = (146.80e6 + df * np.arange(21))[None]
    d = orc.delay_array_ns(f21)
    assert np.isclose(d[5], -483.333327140625, rtol=0, atol=1e-6)
    st = dict(
        data_array=np.zeros((1, 1, 1, 2, 2, 21), complex), noise_array=np.zeros((1, 1, 1, 2, 2, 21), complex),
        nsample_array=np.ones((1, 1, 1, 2, 2, 21)), flag_array=np.zeros((1, 1, 1, 2, 2, 21), bool),
        freq_array=f21, trcvr=np.full((1, 21), 144.0), beam_area=np.ones((1, 1, 21)),
        beam_sq_area=np.ones((1, 1, 21)),
    )
    out = orc.select_spectral_windows(st, [(0, 9), (10, 19)])
    assert out["data_array"].shape == (2, 1, 1, 2, 2, 10)
    assert np.isclose(out["delay_array_ns"][3], -405.999994798125, rtol=0, atol=1e-6)
    assert np.array_equal(out["freq_array"][1], f21[0, 10:20])


def test_jy_to_mk_value():
    # lambda = 1.998616 m at 150 MHz; lambda^2 / (2 k_B) * 1e-23 = 1.446590 mK sr / Jy
    assert np.isclose(orc.jy_to_mk(150e6), 1.446590, rtol=1e-6)


def _small_problem(seed=3, S0=1, U=1, P=2, B=5, T=4, F=21, zeros=False):
    rng = np.random.default_rng(seed)
    shape = (S0, U, P, B, T, F)
    vis = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    flags = rng.random(shape) < 0.1
    ns = rng.integers(32, 61, shape).astype(np.float64)
    if zeros:
        ns[rng.random(shape) < 0.05] = 0
    freq = (146.8e6 + 492610.84375 * np.arange(F))[None]
    return dict(
        vis=vis, flags=flags, nsample=ns, freq_array_hz=freq, trcvr_k=np.full((1, F), 144.0),
        beam_area_sr=np.full((1, P, F), 0.76), integration_time_s=np.full((B, T), 42.95),
        polarization_array=np.array([1, 2])[:P],
    )


def test_pipeline_shapes_and_identities():
    # shapes of simpleDS/tests/test_delay_spectrum.py:889-955
    pr = _small_problem()
    res = orc.pipeline_reference(**pr, spectral_windows=[(0, 9), (10, 19)])
    S, P, B, T, D = 2, 2, 5, 4, 10
    assert res["power_array"].shape == (S, P, B, B, T, D)
    assert res["noise_power"].shape == (S, P, B, B, T, D)
    assert res["thermal_power"].shape == (S, P, B, B, T)
    # Nuv == 1 -> Hermitian in (i, j), real non-negative diagonal
    pw = res["power_array"]
    assert np.allclose(pw, np.conj(np.swapaxes(pw, 2, 3)))
    # factorised average identities used by the fused kernels
    X = res["data_array"][:, 0]
    s1 = X.sum(axis=2)
    s2 = (np.abs(X) ** 2).sum(axis=2)
    allp = orc.average_power(pw, remove_autos=False, over_time=False)
    noauto = orc.average_power(pw, remove_autos=True, over_time=False)
    assert np.allclose(allp, np.abs(s1) ** 2 / B**2)
    assert np.allclose(noauto, (np.abs(s1) ** 2 - s2) / (B * (B - 1)))


def test_partition_invariance_time_pol():
    # simpleDS/tests/test_io.py:570-631, 680-697: slices along time / pol commute
    pr = _small_problem(seed=7)
    full = orc.calculate_delay_spectrum(
        pr["vis"].astype(complex), np.zeros(pr["vis"].shape, complex), pr["flags"], pr["nsample"],
        pr["freq_array_hz"], pr["trcvr_k"], pr["beam_area_sr"], pr["integration_time_s"],
        pr["polarization_array"], windows.blackmanharris,
    )
    for sl in (slice(0, 2), slice(2, 4)):
        part = orc.calculate_delay_spectrum(
            pr["vis"][..., sl, :].astype(complex), np.zeros(pr["vis"][..., sl, :].shape, complex),
            pr["flags"][..., sl, :], pr["nsample"][..., sl, :], pr["freq_array_hz"], pr["trcvr_k"],
            pr["beam_area_sr"], pr["integration_time_s"][:, sl], pr["polarization_array"],
            windows.blackmanharris,
        )
        assert np.array_equal(part["power_array"], full["power_array"][..., sl, :])
        assert np.array_equal(part["thermal_power"], full["thermal_power"][..., sl])


def test_thermal_masks_zero_nsample():
    pr = _small_problem(seed=11, zeros=True)
    th = orc.calculate_thermal_sensitivity(
        pr["freq_array_hz"], pr["trcvr_k"], pr["beam_area_sr"], pr["integration_time_s"],
        pr["nsample"], pr["polarization_array"], windows.blackmanharris,
    )
    assert np.all(np.isfinite(th)) and np.all(th >= 0)
    sig = orc.calculate_noise_power(
        pr["freq_array_hz"], pr["trcvr_k"], pr["beam_area_sr"], pr["integration_time_s"], pr["nsample"]
    )
    assert np.all(sig[pr["nsample"] == 0] == 0) and np.all(sig[pr["nsample"] > 0] > 0)
    # all-inf trcvr/beam (state inside add_uvdata, delay_spectrum.py:975-987) -> zero noise
    sig0 = orc.calculate_noise_power(
        pr["freq_array_hz"], np.full((1, 21), np.inf), np.full((1, 2, 21), np.inf),
        pr["integration_time_s"], pr["nsample"],
    )
    assert np.all(sig0 == 0)


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors (Salmon et al. 2011): pins the counter-based generator that
    the kernels' noise streams are compared against."""
    from oracle.philox import philox4x32_10

    u = np.uint32
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kats:
        got = philox4x32_10(*(np.array([c], dtype=u) for c in ctr), u(key[0]), u(key[1]))
        assert [int(g[0]) for g in got] == want


def test_noise_power_and_thermal_match_the_reference_method_bodies():
    """tests/golden/ds_golden.npz was made by executing the reference's own
    DelaySpectrum.calculate_noise_power / calculate_thermal_sensitivity / utils.jy_to_mk source
    (cut from the files with ast) under a dimension-checked mini units module
    (tests/golden/make_ds_golden.py).  The oracle restatements must reproduce them."""
    import os

    from scipy.signal import windows as _w

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ds_golden.npz"))
    np.testing.assert_allclose(orc.jy_to_mk(z["jy_freqs"]), z["jy_to_mk"], rtol=1e-14)
    for tag in "abc":
        args = (z[f"{tag}_freq_array"], z[f"{tag}_trcvr"], z[f"{tag}_beam_area"], z[f"{tag}_integration_time"],
                z[f"{tag}_nsample"])
        got = orc.calculate_noise_power(*args)
        want = z[f"{tag}_noise_power_jy"]
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)
        if tag == "c":  # infinite receiver temperature in one channel -> sigma masked to 0 (3575-3582)
            assert np.all(want[..., 3] == 0) and np.all(got[..., 3] == 0) and np.all(want[..., 0] > 0)
        else:
            th = orc.calculate_thermal_sensitivity(*args, z[f"{tag}_pols"], _w.blackmanharris)
            np.testing.assert_allclose(th, z[f"{tag}_thermal_power"], rtol=1e-13, atol=0)


def test_select_spectral_windows_matches_the_reference_method_body():
    """ds_golden.npz w1..w3: the reference's own select_spectral_windows and
    _take_spectral_windows_from_data_like_array (delay_spectrum.py:3196-3338) executed on seeded
    arrays -- the gather is integer index work and must be reproduced bit for bit; the delay
    axis to the last ulp."""
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ds_golden.npz"))
    for tag in ("w1", "w2", "w3"):
        state = {k: z[f"{tag}_in_{k}"] for k in ("data_array", "noise_array", "nsample_array", "flag_array",
                                                  "freq_array", "trcvr", "beam_area", "beam_sq_area")}
        out = orc.select_spectral_windows(state, [tuple(w) for w in z[f"{tag}_wins"]])
        for k in state:
            assert np.array_equal(out[k], z[f"{tag}_out_{k}"]), (tag, k)
        np.testing.assert_allclose(out["delay_array_ns"], z[f"{tag}_out_delay_ns"], rtol=1e-15, atol=1e-12)


def test_calculate_delay_spectrum_matches_the_reference_driver():
    """ds_golden.npz d1 (Nuv = 1), d2 (Nuv = 2): the reference's DelaySpectrum.
    normalized_fourier_transform / delay_transform / calculate_delay_spectrum bodies
    (delay_spectrum.py:3487-3542, 3585-3635) driving its real utils functions.  The oracle's
    driver must give the same delay-domain data and noise, power tensors and thermal estimate."""
    import os

    from scipy.signal import windows as _w

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ds_golden.npz"))
    for tag in ("d1", "d2"):
        i = {k: z[f"{tag}_in_{k}"] for k in ("data_array", "noise_array", "flag_array", "nsample_array",
                                              "freq_array", "trcvr", "beam_area", "integration_time")}
        res = orc.calculate_delay_spectrum(i["data_array"], i["noise_array"], i["flag_array"], i["nsample_array"],
                                           i["freq_array"], i["trcvr"], i["beam_area"], i["integration_time"],
                                           z[f"{tag}_pols"], _w.blackmanharris)
        for k in ("data_array", "noise_array", "power_array", "noise_power"):
            assert np.array_equal(res[k], z[f"{tag}_out_{k}"]), (tag, k)  # same numpy calls: bit for bit
        np.testing.assert_allclose(res["thermal_power"], z[f"{tag}_out_thermal_power"], rtol=1e-13, atol=0)
