"""CPU oracle for the simpleDS delay-spectrum hot path.  TEST INFRASTRUCTURE ONLY.

This module is a value-level, astropy-free numpy restatement of the reference
(rasg-affiliates/simpleDS, mounted at /root/reference).  It exists so the CUDA
path can be checked; it is NOT part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``simpleds_b200/`` imports it.

Why a restatement: the reference imports astropy / pyuvdata / h5py /
setuptools_scm at module import time (``simpleDS/__init__.py:6-11``,
``utils.py:8-12``, ``delay_spectrum.py:12-23``) and none of those exist in this
image, so the package cannot be imported here or on the GPU box.  The only
third-party *arithmetic* on the path (``numpy.fft``, ``numpy.random``,
``scipy.signal.windows``, the trapezoid rule) is present and is called
directly, so results are pinned to the same libraries the reference would use.

Pinning status ("Oracle"):
  * utils-level functions (``normalized_fourier_transform``,
    ``cross_multiply_array``, ``combine_nsamples``, ``remove_auto_correlations``,
    ``generate_noise``, ``noise_equivalent_bandwidth``, ``weighted_average``)
    are PINNED: ``tests/golden/make_golden.py`` executes the reference's own
    ``simpleDS/utils.py`` (unmodified, loaded from /root/reference with a stub
    ``astropy`` namespace that only supplies the names the module touches at
    import time) on seeded inputs and the oracle reproduces those outputs
    bit for bit (``tests/test_oracle_golden.py``).
  * ``calculate_noise_power``, ``calculate_thermal_sensitivity`` and ``jy_to_mk`` are PINNED
    the same way: ``tests/golden/make_ds_golden.py`` cuts the reference's method bodies out
    of ``delay_spectrum.py`` / ``utils.py`` with ``ast`` and executes them unmodified under a
    small dimension-checked units module (SI scales, CODATA-2018 c and k_B); the oracle
    reproduces ``tests/golden/ds_golden.npz`` to 1e-13 (``tests/test_oracle_golden.py``),
    including the sigma -> 0 masking of non-finite values and the Nuv = 2 case.
  * ``select_spectral_windows`` / ``_take_spectral_windows_from_data_like_array`` /
    ``delay_array`` (delay_spectrum.py:3196-3338) and the driver
    ``DelaySpectrum.normalized_fourier_transform`` / ``delay_transform`` /
    ``calculate_delay_spectrum`` (3487-3542, 3585-3635; Nuv = 1 and 2) are pinned the same way
    (bit-exact gather, data, noise and power tensors), and additionally by the reference
    tests' literal constants (delay values of ``test_delay_spectrum.py:1619`` /
    ``test_io.py:294``, shapes of ``test_delay_spectrum.py:889-955``) on the bundled files
    (``tests/test_uvh5_golden.py``).  Thermal parity on samples with ``nsample == 0`` is
    UNPINNED (how the mask of ``np.ma.masked_invalid`` survives the product with an astropy
    Quantity at ``delay_spectrum.py:3696-3699`` is version dependent):
    the oracle follows the documented intent (invalid samples drop out of the trapezoid).

Units are not tracked: every function documents the unit its value is in.
All ``file:line`` citations are relative to /root/reference/simpleDS/.
"""
import numpy as np
from scipy.signal import windows

# astropy.constants values the reference reads (utils.py:482, cosmo.py:16)
C_M_PER_S = 299792458.0
K_B_J_PER_K = 1.380649e-23


# --------------------------------------------------------------------------
# utils.py restatements
# --------------------------------------------------------------------------
def normalized_fourier_transform(
    data_array, delta_x, axis=-1, taper=windows.blackmanharris, inverse=False
):
    """utils.py:509-566.  ``delta_x`` is the SI value of the reference's Quantity."""
    data_array = np.asarray(data_array)
    n_axis = data_array.shape[axis]
    data_shape = np.ones_like(data_array.shape)  # utils.py:549
    data_shape[axis] = n_axis
    win = np.broadcast_to(taper(n_axis), data_shape)  # utils.py:552
    if not inverse:
        fourier_array = np.fft.fft(data_array * win, axis=axis)  # utils.py:558
        fourier_array = np.fft.fftshift(fourier_array, axes=axis)  # utils.py:559
        fourier_array = fourier_array * delta_x  # utils.py:560
    else:
        fourier_array = np.fft.ifft(data_array, axis=axis)  # utils.py:562
        fourier_array = np.fft.ifftshift(fourier_array, axes=axis)  # utils.py:563
        fourier_array = fourier_array / win * delta_x  # utils.py:564
    return fourier_array


def cross_multiply_array(array_1, array_2=None, axis=0):
    """utils.py:334-393: out[.., i, j, ..] = conj(a1[.., j, ..]) * a2[.., i, ..]."""
    array_1 = np.asarray(array_1)  # utils.py:361-362
    if array_2 is None:
        array_2 = array_1.copy()  # utils.py:364-365
    array_2 = np.asarray(array_2)  # utils.py:367-368
    if array_2.shape != array_1.shape:  # utils.py:379-385
        raise ValueError(
            "array_1 and array_2 must have the same shapes. "
            "array_1 has shape {a1} but array_2 has shape {a2}".format(
                a1=np.shape(array_1), a2=np.shape(array_2)
            )
        )
    return np.expand_dims(array_1, axis=axis).conj() * np.expand_dims(
        array_2, axis=axis + 1
    )  # utils.py:387-389


def combine_nsamples(nsample_1, nsample_2=None, axis=-1):
    """utils.py:230-278: geometric mean sqrt(n1[j] * n2[i])."""
    if nsample_2 is None:
        nsample_2 = nsample_1.copy()
    if not nsample_1.shape == nsample_2.shape:
        raise ValueError(
            "nsample_1 and nsample_2 must have same shape, "
            "but nsample_1 has shape {d1_s} and "
            "nsample_2 has shape {d2_s}".format(
                d1_s=nsample_1.shape, d2_s=nsample_2.shape
            )
        )
    return np.sqrt(cross_multiply_array(nsample_1, nsample_2, axis=axis))


def remove_auto_correlations(data_array, axes=(0, 1)):
    """utils.py:281-331."""
    if not np.shape(axes)[0] == 2:
        raise ValueError("Shape must be a length 2 tuple/array/list of axis indices.")
    if axes[0] != axes[1] - 1:
        raise ValueError(
            "Axes over which diagonal components are to be remove must be adjacent."
        )
    if data_array.shape[axes[0]] != data_array.shape[axes[1]]:
        raise ValueError(
            "The axes over which diagonal components are to be "
            "removed must have the same shape."
        )
    n_inds = data_array.shape[axes[0]]
    indices = np.logical_not(np.diag(np.ones(n_inds, dtype=bool)))  # utils.py:324
    data_array = np.moveaxis(data_array, axes[0], 0)
    data_array = np.moveaxis(data_array, axes[1], 1)
    data_out = data_array[indices]
    return np.moveaxis(data_out, 0, axes[0])


def generate_noise(noise_power):
    """utils.py:486-506: two full-size draws from the GLOBAL numpy generator,
    all real parts first, then all imaginary parts."""
    noise_power = np.asarray(noise_power)
    noise = noise_power * (
        1 * np.random.normal(size=noise_power.shape)
        + 1j * np.random.normal(size=noise_power.shape)
    )  # utils.py:501-504
    noise /= np.sqrt(2)  # utils.py:505
    return noise


def noise_equivalent_bandwidth(window):
    """utils.py:212-227."""
    return np.sum(window) ** 2 / (np.sum(window**2) * len(window))


def weighted_average(array, uncertainty, weights=None, axis=-1):
    """utils.py:569-645 (value level)."""
    if not array.shape == uncertainty.shape:
        raise ValueError("Input array and uncertainties must have the same shape.")
    if weights is None:
        weights = 1.0 / uncertainty**2
    if np.ndim(weights) == 1 and weights.size != array.shape[axis]:
        raise ValueError(
            "1-Dimenionals weights must have the same shape "
            "as the axis over which the average is taken."
        )
    elif not np.shape(weights) == array.shape:
        raise ValueError("Input array and weights must have the same shape.")
    array_out = np.average(array, weights=weights, axis=axis)
    uncertainty_out = np.sqrt(
        np.sum(uncertainty**2 * np.abs(weights) ** 2, axis=axis)
        / np.abs(np.sum(weights, axis=axis)) ** 2
    )
    return array_out, uncertainty_out


def jy_to_mk(freqs_hz):
    """utils.py:467-483: c^2 / (2 nu^2 k_B) expressed in mK sr / Jy.

    1 Jy = 1e-26 J m^-2, 1 K = 1e3 mK  ->  factor 1e-23.
    """
    freqs_hz = np.asarray(freqs_hz, dtype=np.float64)
    return 1e-23 * C_M_PER_S**2 / (2.0 * freqs_hz**2 * K_B_J_PER_K)


# --------------------------------------------------------------------------
# delay_spectrum.py restatements.  Arrays use the reference's 6-D layout
# (Nspws, Nuv, Npols, Nbls, Ntimes, Nfreqs)  (delay_spectrum.py:159-176).
# --------------------------------------------------------------------------
def baseline_number(ant_1, ant_2):
    """Baseline number convention documented at delay_spectrum.py:269-272
    (pyuvdata antnums_to_baseline for < 2048 antennas)."""
    ant_1 = np.asarray(ant_1, dtype=np.int64)
    ant_2 = np.asarray(ant_2, dtype=np.int64)
    return 2048 * (ant_1 + 1) + (ant_2 + 1) + 2**16


def normalize_spectral_windows(spectral_windows, nfreqs):
    """delay_spectrum.py:3243-3264 -> int array (Nspws, 2) and channel table."""
    if spectral_windows is None:
        spectral_windows = [(0, nfreqs - 1)]
    if not isinstance(spectral_windows, (list, tuple, np.ndarray)):
        spectral_windows = [spectral_windows]
    if not isinstance(spectral_windows[0], (list, tuple, np.ndarray)):
        spectral_windows = [spectral_windows]
    if not all(len(x) == 2 for x in spectral_windows):
        raise ValueError(
            "Spectral window tuples must only have two "
            "elements (stard_index, end_index)"
        )
    spectral_windows = np.asarray(spectral_windows)
    num_freqs = spectral_windows[:, 1] + 1 - spectral_windows[:, 0]
    if not all(num_freqs == num_freqs[0]):
        raise ValueError("Spectral windows must all have the same size.")
    freq_chans = spectral_windows[:, 0, None] + np.arange(num_freqs[0])
    return spectral_windows, freq_chans


def take_spectral_windows_from_data_like_array(data, freq_chans):
    """delay_spectrum.py:3315-3338 (flattened take over Nspws_old*Nfreqs_old)."""
    nspws, nfreqs = freq_chans.shape
    data = np.transpose(data, [1, 2, 3, 4, 0, 5])
    shape = tuple(data.shape[:4]) + (-1,)
    data = data.reshape(shape)
    data = np.take(data, freq_chans.flatten(), axis=4)
    data = data.reshape(shape[:4] + (nspws, nfreqs))
    return np.transpose(data, [4, 0, 1, 2, 3, 5])


def select_spectral_windows(state, spectral_windows=None):
    """delay_spectrum.py:3196-3313 on a plain dict ``state`` (returns a new dict).

    ``state`` keys: data_array, noise_array, nsample_array, flag_array (6-D),
    freq_array (Nspws, Nfreqs) [Hz], trcvr (Nspws, Nfreqs) [K], beam_area,
    beam_sq_area (Nspws, Npols, Nfreqs) [sr].
    """
    out = dict(state)
    nfreqs_old = state["freq_array"].shape[1]
    _, freq_chans = normalize_spectral_windows(spectral_windows, nfreqs_old)
    nspws, nfreqs = freq_chans.shape
    out["freq_array"] = np.take(state["freq_array"], freq_chans)  # 3267
    out["trcvr"] = np.take(state["trcvr"], freq_chans)  # 3268
    flat = freq_chans.flatten()
    for key in ("beam_sq_area", "beam_area"):  # 3273-3283
        npols = state[key].shape[1]
        arr = np.transpose(state[key], [1, 0, 2]).reshape(npols, -1)
        arr = np.take(arr, flat, axis=1).reshape(npols, nspws, nfreqs)
        out[key] = np.transpose(arr, [1, 0, 2])
    for key in ("data_array", "noise_array", "nsample_array", "flag_array"):
        out[key] = take_spectral_windows_from_data_like_array(state[key], freq_chans)
    out["delay_array_ns"] = delay_array_ns(out["freq_array"])
    return out


def delay_array_ns(freq_array_hz):
    """delay_spectrum.py:3301-3304: fftshift(fftfreq(N, d=df)) in ns."""
    ndelays = freq_array_hz.shape[1]
    delays = np.fft.fftfreq(ndelays, d=np.diff(freq_array_hz[0])[0])
    return np.fft.fftshift(delays) * 1e9


def ds_normalized_fourier_transform(
    data_array, noise_array, flag_array, freq_array_hz, taper, inverse=False,
    delay_array_ns_=None,
):
    """delay_spectrum.py:3487-3517: flag mask, delta_x choice, both arrays."""
    if inverse:
        delta_x = np.diff(delay_array_ns_)[0] * 1e-9  # 3500, .si -> seconds
    else:
        delta_x = np.diff(freq_array_hz[0])[0]  # 3502
    float_flags = np.logical_not(flag_array).astype(float)  # 3503
    data_out = normalized_fourier_transform(
        data_array * float_flags, delta_x, axis=-1, taper=taper, inverse=inverse
    )
    noise_out = normalized_fourier_transform(
        noise_array * float_flags, delta_x, axis=-1, taper=taper, inverse=inverse
    )
    return data_out, noise_out


def calculate_noise_power(
    freq_array_hz, trcvr_k, beam_area_sr, integration_time_s, nsample_array
):
    """delay_spectrum.py:3544-3583 -> sigma in Jy, 6-D."""
    nspws, nfreqs = freq_array_hz.shape
    npols = beam_area_sr.shape[1]
    nbls, ntimes = integration_time_s.shape
    tsys = 180.0 * np.power(freq_array_hz * 1e-9 / 0.18, -2.55)  # 3546-3550
    tsys = tsys + trcvr_k  # 3551
    tsys = tsys.reshape(nspws, 1, 1, 1, 1, nfreqs)
    delta_f = np.diff(freq_array_hz[0])[0]  # 3553
    with np.errstate(divide="ignore", invalid="ignore"):
        noise_power = tsys / np.sqrt(
            delta_f * integration_time_s.reshape(1, 1, 1, nbls, ntimes, 1) * nsample_array
        )  # 3564-3570
        noise_power = (
            noise_power * 1e3  # .to("mK")  3576
            * beam_area_sr.reshape(nspws, 1, npols, 1, 1, nfreqs)
            / jy_to_mk(freq_array_hz).reshape(nspws, 1, 1, 1, 1, nfreqs)
        )
    noise_power = np.ma.masked_invalid(noise_power).filled(0)  # 3575-3582
    return noise_power


def npols_noise(polarization_array):
    """delay_spectrum.py:3663-3665."""
    return np.array([2 if p in np.arange(1, 5) else 1 for p in polarization_array])


def calculate_thermal_sensitivity(
    freq_array_hz, trcvr_k, beam_area_sr, integration_time_s, nsample_array,
    polarization_array, taper,
):
    """delay_spectrum.py:3637-3707 -> thermal_power (Nspws,Npols,Nbls,Nbls,Ntimes)
    in K^2 sr^2 Hz^6.  Invalid samples are masked out of the trapezoid (the
    documented intent of np.ma.masked_invalid at :3697; UNPINNED, see header)."""
    nspws, nfreqs = freq_array_hz.shape
    npols = beam_area_sr.shape[1]
    nbls, ntimes = integration_time_s.shape
    nuv = nsample_array.shape[1]
    n2 = nsample_array[:, 0] if nuv == 1 else nsample_array[:, 1]
    samples = combine_nsamples(nsample_array[:, 0], n2, axis=2)  # 3655-3662
    npn = npols_noise(polarization_array).reshape(1, npols, 1, 1, 1, 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        tsys = 180.0 * np.power(freq_array_hz * 1e-9 / 0.18, -2.55)
        tsys = tsys + trcvr_k
        tsys = tsys.reshape(nspws, 1, 1, 1, 1, nfreqs)
        thermal = (
            tsys * 1e3
            * beam_area_sr.reshape(nspws, npols, 1, 1, 1, nfreqs)
            / np.sqrt(
                integration_time_s.reshape(1, 1, 1, nbls, ntimes, 1)
                * samples.reshape(nspws, npols, nbls, nbls, ntimes, nfreqs)
                * npn
                * np.sqrt(2)
            )
        )  # 3674-3694
        integrand = (
            np.ma.masked_invalid(thermal) ** 2
            * taper(nfreqs).reshape(1, 1, 1, 1, 1, nfreqs) ** 2
            * freq_array_hz.reshape(nspws, 1, 1, 1, 1, nfreqs) ** 4
        )
        x = freq_array_hz.reshape(nspws, 1, 1, 1, 1, nfreqs)
        # scipy.integrate.trapz == numpy trapezoid rule (removed from scipy 1.14)
        d = np.diff(x, axis=-1)
        thermal_power = (d * (integrand[..., 1:] + integrand[..., :-1]) / 2.0).sum(-1)
    if isinstance(thermal_power, np.ma.MaskedArray):
        thermal_power = thermal_power.filled(0.0)
    return np.asarray(thermal_power) * 1e-6  # mK^2 -> K^2  (3707)


def calculate_delay_spectrum(
    data_array, noise_array, flag_array, nsample_array, freq_array_hz, trcvr_k,
    beam_area_sr, integration_time_s, polarization_array, taper,
    data_type="frequency",
):
    """delay_spectrum.py:3585-3635, stopping before update_cosmology.

    Returns dict with delay-domain data/noise, power_array, noise_power
    ((Jy Hz)^2 for Jy input) and thermal_power.
    """
    nuv = data_array.shape[1]
    if nuv == 0:
        raise ValueError(
            "No data has be loaded. Add UVData objects before "
            "calling calculate_delay_spectrum."
        )
    if data_type == "frequency":  # 3617-3618
        data_array, noise_array = ds_normalized_fourier_transform(
            data_array, noise_array, flag_array, freq_array_hz, taper
        )
    if nuv == 1:  # 3620-3626
        power = cross_multiply_array(data_array[:, 0], axis=2)
        noise_power = cross_multiply_array(noise_array[:, 0], axis=2)
    else:  # 3627-3633
        power = cross_multiply_array(data_array[:, 0], data_array[:, 1], axis=2)
        noise_power = cross_multiply_array(noise_array[:, 0], noise_array[:, 1], axis=2)
    thermal = calculate_thermal_sensitivity(
        freq_array_hz, trcvr_k, beam_area_sr, integration_time_s, nsample_array,
        polarization_array, taper,
    )
    return dict(
        data_array=data_array, noise_array=noise_array, power_array=power,
        noise_power=noise_power, thermal_power=thermal,
    )


# --------------------------------------------------------------------------
# User-side averaging (examples/simpleDS_tutorial.ipynb cell 41;
# utils.remove_auto_correlations) -- the oracle for the fused average.
# --------------------------------------------------------------------------
def average_power(power, remove_autos=False, over_time=True):
    """(S,P,B,B,T,D) -> (S,P,D) [over_time] or (S,P,T,D): plain mean over the
    ordered pairs (optionally without the diagonal), then over time."""
    if remove_autos:
        p = remove_auto_correlations(power, axes=(2, 3))  # (S,P,B(B-1),T,D)
    else:
        s = power.shape
        p = power.reshape(s[0], s[1], s[2] * s[3], *s[4:])
    out = p.mean(axis=2)
    if over_time:
        out = out.mean(axis=2)
    return out


def pipeline_reference(
    vis, flags, nsample, freq_array_hz, trcvr_k, beam_area_sr, integration_time_s,
    polarization_array, taper=windows.blackmanharris, spectral_windows=None,
    noise_seed=0,
):
    """Whole reference path on one redundant group, as a user of the reference
    would run it: select windows -> generate_noise -> delay_transform ->
    calculate_delay_spectrum (-> tutorial mean).  ``vis`` etc. are 6-D with
    Nspws=1.  Used by bench.py's cpu_baseline leg and the parity tests."""
    state = dict(
        data_array=np.asarray(vis, dtype=np.complex128),
        noise_array=np.zeros(vis.shape, dtype=np.complex128),
        nsample_array=np.asarray(nsample, dtype=np.float64),
        flag_array=np.asarray(flags, dtype=bool),
        freq_array=freq_array_hz, trcvr=trcvr_k, beam_area=beam_area_sr,
        beam_sq_area=beam_area_sr,
    )
    state = select_spectral_windows(state, spectral_windows)
    sigma = calculate_noise_power(
        state["freq_array"], state["trcvr"], state["beam_area"],
        integration_time_s, state["nsample_array"],
    )
    np.random.seed(noise_seed)
    state["noise_array"] = generate_noise(sigma)
    res = calculate_delay_spectrum(
        state["data_array"], state["noise_array"], state["flag_array"],
        state["nsample_array"], state["freq_array"], state["trcvr"],
        state["beam_area"], integration_time_s, polarization_array, taper,
    )
    res["sigma"] = sigma
    res["delay_array_ns"] = state["delay_array_ns"]
    return res
