"""C1 of BASELINE.json: the reference's bundled visibility files through the path.

tests/golden/uvh5_*.npz hold simpleDS/data/test_redundant_array_odd_file.uvh5 (51 baselines,
21 times, 21 channels, one pol) and the 12 lowest baselines of paper_test_file.uvh5 (203
channels, 26.6 % flagged, nsample == 0 where flagged), exactly as stored in the files
(tests/golden/make_uvh5_golden.py).  CPU tests pin the oracle to the literals the reference's
own tests quote for these data; GPU tests run the drop-in DelaySpectrum on them and compare
with the oracle."""
import os

import numpy as np
import pytest
from scipy.signal import windows

from helpers import RTOL_C64, RTOL_C128, assert_close
from oracle import simpleds_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    d = dict(np.load(os.path.join(GOLDEN, name)))
    d["vis_units"] = str(d["vis_units"])
    return d


def extract_loops(d):
    """Independent restatement of the ingest ordering (delay_spectrum.py:913-973) as explicit
    per-baseline loops, the way the reference's extractor tests do (tests/test_utils.py:19-157):
    baselines by rising baseline number, times by rising time."""
    bl = orc.baseline_number(d["ant_1_array"], d["ant_2_array"])
    bls = np.unique(bl)
    times = np.unique(d["time_array"])
    B, T, F, P = len(bls), len(times), d["data_array"].shape[1], d["data_array"].shape[2]
    # the reference's extractors quantise: complex64 data (utils.py:38), float32 nsample
    # (utils.py:78) and float32 integration time (utils.py:156)
    vis = np.zeros((1, 1, P, B, T, F), np.complex64)
    flags = np.zeros((1, 1, P, B, T, F), bool)
    ns = np.zeros((1, 1, P, B, T, F), np.float32)
    itime = np.zeros((B, T), np.float32)
    for i, b in enumerate(bls):
        for j, t in enumerate(times):
            (row,) = np.nonzero((bl == b) & (d["time_array"] == t))
            assert row.size == 1
            vis[0, 0, :, i, j, :] = d["data_array"][row[0]].T
            flags[0, 0, :, i, j, :] = d["flag_array"][row[0]].T
            ns[0, 0, :, i, j, :] = d["nsample_array"][row[0]].T
            itime[i, j] = d["integration_time"][row[0]]
    return dict(vis=vis.astype(complex), flags=flags, nsample=ns.astype(np.float64),
                integration_time_s=itime.astype(np.float64), baselines=bls,
                freq_array_hz=d["freq_array"][None], polarization_array=d["polarization_array"])


# ------------------------------------------------------------------------------- CPU
def test_bundled_file_literals():
    """Constants the reference's tests quote for this file family."""
    d = load("uvh5_odd.npz")
    ex = extract_loops(d)
    assert ex["vis"].shape == (1, 1, 1, 51, 21, 21)
    assert list(ex["baselines"][:3]) == [67611, 69637, 71714]
    assert 69637 in ex["baselines"] and 157697 in ex["baselines"]  # (1,4), (44,0): t_ds.py:1532-1534
    assert orc.baseline_number(np.array([1]), np.array([4]))[0] == 69637
    assert float(d["channel_width"]) == 492610.84375
    delays = orc.delay_array_ns(ex["freq_array_hz"])
    assert np.isclose(delays[5], -483.333327140625, rtol=0, atol=1e-9)  # t_ds.py:1619
    st = orc.select_spectral_windows(dict(
        data_array=ex["vis"], noise_array=np.zeros_like(ex["vis"]), nsample_array=ex["nsample"],
        flag_array=ex["flags"], freq_array=ex["freq_array_hz"], trcvr=np.full((1, 21), 144.0),
        beam_area=np.ones((1, 1, 21)), beam_sq_area=np.ones((1, 1, 21))), [(0, 9), (10, 19)])
    assert st["data_array"].shape == (2, 1, 1, 51, 21, 10)
    assert np.isclose(st["delay_array_ns"][3], -405.999994798125, rtol=0, atol=1e-9)  # t_io.py:294,314
    p = load("uvh5_paper12.npz")
    assert np.isclose(p["flag_array"].mean(), 0.266, atol=0.01)
    assert np.all((p["nsample_array"] == 0) == p["flag_array"])


def test_power_is_cross_multiply_of_transform_on_bundled_file():
    """t_ds.py:1122-1147: power_array == cross_multiply_array(delay_transform(data)[:, 0], axis=2)."""
    ex = extract_loops(load("uvh5_odd.npz"))
    F = 21
    res = orc.calculate_delay_spectrum(
        ex["vis"], np.zeros_like(ex["vis"]), ex["flags"], ex["nsample"], ex["freq_array_hz"],
        np.full((1, F), 144.0), np.full((1, 1, F), 0.76), ex["integration_time_s"],
        ex["polarization_array"], windows.blackmanharris)
    df = ex["freq_array_hz"][0, 1] - ex["freq_array_hz"][0, 0]
    X = orc.normalized_fourier_transform(ex["vis"] * (~ex["flags"]), df)
    want = orc.cross_multiply_array(X[:, 0], axis=2)
    assert res["power_array"].shape == (1, 1, 51, 51, 21, 21)  # t_ds.py:889-909
    assert np.array_equal(res["power_array"], want)
    assert res["thermal_power"].shape == (1, 1, 51, 51, 21)
    assert np.all(np.isfinite(res["thermal_power"])) and np.all(res["thermal_power"] > 0)


# ------------------------------------------------------------------------------- GPU
def _ds_from_fixture(d, precision):
    from simpleds_b200 import DelaySpectrum, UVDataLite

    uv = UVDataLite(d["data_array"], d["flag_array"], d["nsample_array"], d["ant_1_array"], d["ant_2_array"],
                    d["lst_array"], d["integration_time"], d["freq_array"], d["polarization_array"],
                    vis_units=d["vis_units"], uvw_array=d["uvw_array"],
                    Nants_telescope=int(d["Nants_telescope"]))
    return DelaySpectrum(uv=uv, precision=precision)


@pytest.mark.gpu
@pytest.mark.parametrize("name,wins,precision", [
    ("uvh5_odd.npz", [(0, 20)], "complex128"),
    ("uvh5_odd.npz", [(0, 10), (10, 20)], "complex128"),      # overlapping windows of t_ds.py:37
    ("uvh5_odd.npz", [(0, 9), (10, 19)], "complex64"),
    ("uvh5_paper12.npz", [(95, 115)], "complex128"),          # sub-band of t_ds.py:985-986
    ("uvh5_paper12.npz", [(0, 202)], "complex128"),           # 203 = 7 * 29 channels, flags, nsample 0
    ("uvh5_paper12.npz", [(0, 202)], "complex64"),
])
def test_bundled_files_through_delay_spectrum(cuda_device, name, wins, precision):
    d = load(name)
    ex = extract_loops(d)
    ds = _ds_from_fixture(d, precision)
    # ingest: bit-exact ordering and values
    assert np.array_equal(ds.baseline_array, ex["baselines"])
    assert np.array_equal(ds.data_array.numpy().astype(complex), ex["vis"].astype(ds.data_array.numpy().dtype).astype(complex))
    assert np.array_equal(ds.flag_array.cpu().numpy(), ex["flags"])
    assert np.array_equal(ds.nsample_array.cpu().numpy(), ex["nsample"])
    assert np.array_equal(np.asarray(ds.integration_time.value), ex["integration_time_s"])
    F0 = ex["vis"].shape[-1]
    ds.add_trcvr(144.0)
    ds.set_beam_area(0.76)
    ds.select_spectral_windows(wins)
    S, N = len(wins), wins[0][1] - wins[0][0] + 1
    P, B, T = ex["vis"].shape[2:5]
    assert ds.data_array.shape == (S, 1, P, B, T, N)
    rng = np.random.default_rng(3)
    g_re, g_im = rng.standard_normal(ds.data_array.shape), rng.standard_normal(ds.data_array.shape)
    ds.generate_noise(normals=(g_re, g_im))
    ds.calculate_delay_spectrum()
    ds.remove_cosmology()  # t_ds.py:1122-1147 compares the raw products

    state = orc.select_spectral_windows(dict(
        data_array=ex["vis"], noise_array=np.zeros_like(ex["vis"]), nsample_array=ex["nsample"],
        flag_array=ex["flags"], freq_array=ex["freq_array_hz"], trcvr=np.full((1, F0), 144.0),
        beam_area=np.full((1, P, F0), 0.76), beam_sq_area=np.full((1, P, F0), 0.76)), wins)
    sigma = orc.calculate_noise_power(state["freq_array"], state["trcvr"], state["beam_area"],
                                      ex["integration_time_s"], state["nsample_array"])
    noise = sigma * (g_re + 1j * g_im) / np.sqrt(2)
    ref = orc.calculate_delay_spectrum(
        state["data_array"], noise, state["flag_array"], state["nsample_array"], state["freq_array"],
        state["trcvr"], state["beam_area"], ex["integration_time_s"], ex["polarization_array"],
        windows.blackmanharris)
    rtol = RTOL_C128 if precision == "complex128" else RTOL_C64
    assert_close(ds.delay_array.value, state["delay_array_ns"], 1e-12, what="delays")
    assert_close(ds.data_array.numpy(), ref["data_array"], rtol, what="delay data")
    assert_close(ds.noise_array.numpy(), ref["noise_array"], rtol, what="delay noise")
    assert ds.power_array.shape == (S, P, B, B, T, N) and ds.thermal_power.shape == (S, P, B, B, T)
    assert_close(ds.power_array.numpy(), ref["power_array"], 4 * rtol, what="power")
    assert_close(ds.noise_power.numpy(), ref["noise_power"], 4 * rtol, what="noise power")
    assert_close(ds.thermal_power.numpy(), ref["thermal_power"], 1e-11, what="thermal")
    # averaged path on real data == reference tensor then mean (tutorial cell 41)
    av = ds.calculate_averaged_delay_spectrum(remove_autos=True)
    want = orc.average_power(ref["power_array"], remove_autos=True, over_time=True)
    assert_close(av["power"].numpy(), want, 4 * rtol * B, what="averaged power")


@pytest.mark.gpu
@pytest.mark.parametrize("name,wins", [("uvh5_odd.npz", [(0, 20)]), ("uvh5_paper12.npz", [(0, 202)]),
                                       ("uvh5_paper12.npz", [(96, 116), (112, 132)])])
def test_bundled_files_through_fused_pipeline(cuda_device, name, wins):
    """The same real data through RedundantPipeline (one redundant group): 21 and 203 channels are
    not powers of two, so this is the fused chirp-z engine; the row pitch is padded to a multiple
    of 16 channels at ingest (the framework owns the device layout)."""
    import torch

    from oracle import engine_ref
    from simpleds_b200.engine import RedundantPipeline

    ex = extract_loops(load(name))
    vis, flags, ns = (ex[k][0, 0] for k in ("vis", "flags", "nsample"))  # (P,B,T,F)
    P, B, T, F = vis.shape
    Fp = (F + 15) // 16 * 16
    pad = lambda a, v: np.concatenate([a, np.full(a.shape[:-1] + (Fp - F,), v, a.dtype)], axis=-1)  # noqa: E731
    df = ex["freq_array_hz"][0, 1] - ex["freq_array_hz"][0, 0]
    freq = ex["freq_array_hz"][0, 0] + df * np.arange(Fp)
    a = dict(vis=pad(vis.astype(np.complex64), 0), flags=pad(flags, True), nsample=pad(ns.astype(np.float32), 0),
             int_time=ex["integration_time_s"].astype(np.float32))
    go = np.array([0, B])
    pipe = RedundantPipeline(freq, wins, ex["polarization_array"], go, trcvr_k=144.0, beam_area_sr=0.76,
                             precision="complex64", seed=21)
    assert pipe.fused
    pipe.process(*(torch.as_tensor(a[k], device=cuda_device) for k in ("vis", "flags", "nsample", "int_time")),
                 noise="philox", thermal=True)
    res = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in pipe.result().items()}
    N = wins[0][1] - wins[0][0] + 1
    nz = engine_ref.engine_noise_freq_chirpz(pipe.sigma_coef32.cpu().numpy(), a["nsample"], a["int_time"],
                                             pipe.win_start, N, 21)
    ref = engine_ref.group_means(a["vis"], a["flags"], a["nsample"], a["int_time"], go, freq, wins,
                                 ex["polarization_array"], 144.0, 0.76, noise_freq=nz)
    for key, tol in (("power_all", 4 * RTOL_C64), ("power_noautos", 4 * B * RTOL_C64),
                     ("noise_power_all", 1e-4), ("noise_power_noautos", B * 1e-4)):
        assert_close(res[key][0], ref[key][0], tol, what=key)
    np.testing.assert_allclose(res["thermal_all"][0], ref["thermal_all"][0], rtol=4e-5)
    np.testing.assert_allclose(res["thermal_noautos"][0], ref["thermal_noautos"][0], rtol=4e-5 * B)
