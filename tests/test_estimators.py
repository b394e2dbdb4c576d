"""Downstream estimators: bootstrap_array, weighted_average, fold_along_delay.

CPU part: the numpy oracle reproduces, bit for bit, the outputs of the reference's own
``simpleDS/utils.py`` (tests/golden/estimators_golden.npz, made by
tests/golden/make_estimators_golden.py) and the closed-form literals of the reference's tests
(tests/test_utils.py:563-766).  GPU part: the drop-in functions of ``simpleds_b200.utils`` --
which run in libsds.so through the C ABI -- against the oracle and the same literals.
Tolerance: rtol 1e-12 (fp64 sums in a different order than numpy's pairwise reduction);
the gather of bootstrap_array is bit-exact.
"""
import os

import numpy as np
import pytest

from oracle import estimators as orc_est
from oracle import simpleds_oracle as orc

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "estimators_golden.npz"))
RTOL = 1e-12
FOLD_CASES = [(c, k, ax) for c in ("odd", "even", "fft_even") for k in ("real", "cplx") for ax in (2, 0)]


# ---- oracle pinned to the reference --------------------------------------------------------
@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_oracle_weighted_average_matches_reference(name):
    arr, err, w = GOLD[f"wa_{name}_arr"], GOLD[f"wa_{name}_err"], GOLD[f"wa_{name}_w"]
    axis = int(GOLD[f"wa_{name}_axis"])
    o, e = orc.weighted_average(arr, err, axis=axis)
    assert np.array_equal(o, GOLD[f"wa_{name}_ivar_out"]) and np.array_equal(e, GOLD[f"wa_{name}_ivar_err"])
    o, e = orc.weighted_average(arr, err, weights=w, axis=axis)
    assert np.array_equal(o, GOLD[f"wa_{name}_w_out"]) and np.array_equal(e, GOLD[f"wa_{name}_w_err"])


def test_oracle_weighted_average_reference_literals():
    # tests/test_utils.py:578-584
    o, e = orc.weighted_average(np.array([1, 3, 10]), np.array([1, 3, 10]))
    assert np.isclose(o, 1.2784935579781962) and np.isclose(e, 0.9444428250308379)
    # tests/test_utils.py:563-575
    arr = np.arange(20.0).reshape(4, 5)
    o, e = orc.weighted_average(arr, np.ones((4, 5)), weights=np.ones((4, 5)), axis=0)
    assert np.array_equal(o, [7.5, 8.5, 9.5, 10.5, 11.5]) and np.array_equal(e, np.ones(5) / 2.0)
    assert bool(GOLD["wa_1d_weights_nd_array_raises"])
    with pytest.raises(ValueError):
        orc.weighted_average(np.ones((3, 4)), np.ones((3, 4)), weights=np.ones(4), axis=-1)


@pytest.mark.parametrize("case,kind,ax", FOLD_CASES)
def test_oracle_fold_along_delay_matches_reference(case, kind, ax):
    tag = f"fold_{case}_{kind}_ax{ax}"
    axis = -1 if ax == 2 else 0
    d, arr, err, w = GOLD[tag + "_delays"], GOLD[tag + "_arr"], GOLD[tag + "_err"], GOLD[tag + "_w"]
    o, e = orc_est.fold_along_delay(d, arr, err, axis=axis)
    assert np.array_equal(o, GOLD[tag + "_out"]) and np.array_equal(e, GOLD[tag + "_eout"])
    o, e = orc_est.fold_along_delay(d, arr, err, weights=w, axis=axis)
    assert np.array_equal(o, GOLD[tag + "_wout"]) and np.array_equal(e, GOLD[tag + "_weout"])
    if kind == "cplx":
        o, e = orc_est.fold_along_delay(d, arr, err, weights=GOLD[tag + "_wc"], axis=axis)
        assert np.array_equal(o, GOLD[tag + "_wcout"]) and np.array_equal(e, GOLD[tag + "_wceout"])


def test_oracle_fold_reference_quirks():
    assert bool(GOLD["fold_axis1_raises"])  # the reference cannot fold a middle axis
    with pytest.raises(ValueError):
        orc_est.fold_along_delay(np.arange(-10, 11.0), np.ones((2, 21, 3)), np.ones((2, 21, 3)), axis=1)
    with pytest.raises(ValueError):  # tests/test_utils.py:605-611: odd length without a zero bin
        orc_est.fold_along_delay(np.arange(21.0) + 11, np.ones((1, 10, 21)), np.ones((1, 10, 21)))


def test_oracle_bootstrap_matches_reference():
    arr = GOLD["boot_arr"]
    for axis in (0, 1, 2):
        np.random.seed(7 + axis)
        out = orc_est.bootstrap_array(arr, nboot=5, axis=axis)
        assert np.array_equal(out, GOLD[f"boot_out_ax{axis}"])
        assert np.array_equal(np.take(arr, GOLD[f"boot_idx_ax{axis}"], axis=axis), out)
    with pytest.raises(ValueError):
        orc_est.bootstrap_array(arr, axis=3)


# ---- the drop-in (CUDA through the C ABI) --------------------------------------------------
@pytest.fixture(scope="module")
def sds(cuda_device):
    from simpleds_b200 import utils
    from simpleds_b200.units import Quantity
    return utils, Quantity


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_gpu_weighted_average_matches_oracle(sds, name):
    utils, Quantity = sds
    arr, err, w = GOLD[f"wa_{name}_arr"], GOLD[f"wa_{name}_err"], GOLD[f"wa_{name}_w"]
    axis = int(GOLD[f"wa_{name}_axis"])
    o, e = utils.weighted_average(arr, err, axis=axis)
    np.testing.assert_allclose(o, GOLD[f"wa_{name}_ivar_out"], rtol=RTOL)
    np.testing.assert_allclose(e, GOLD[f"wa_{name}_ivar_err"], rtol=RTOL)
    o, e = utils.weighted_average(Quantity(arr, "mK^2 Mpc^3"), Quantity(err, "mK^2 Mpc^3"), weights=w, axis=axis)
    assert str(o.unit) == str(e.unit) == "Mpc^3 mK^2"
    np.testing.assert_allclose(o.value, GOLD[f"wa_{name}_w_out"], rtol=RTOL)
    np.testing.assert_allclose(e.value, GOLD[f"wa_{name}_w_err"], rtol=RTOL)


@pytest.mark.gpu
def test_gpu_weighted_average_reference_tests(sds):
    utils, Quantity = sds
    o, e = utils.weighted_average(np.array([1, 3, 10]), np.array([1, 3, 10]))
    assert np.isclose(o, 1.2784935579781962, rtol=1e-14) and np.isclose(e, 0.9444428250308379, rtol=1e-14)
    arr = Quantity(np.arange(20.0).reshape(4, 5), "m")
    o, e = utils.weighted_average(arr, Quantity(np.ones((4, 5)), "m"), weights=np.ones((4, 5)), axis=0)
    assert np.array_equal(o.value, [7.5, 8.5, 9.5, 10.5, 11.5]) and np.array_equal(e.value, np.ones(5) / 2.0)
    ones = np.ones((3, 3, 2))
    with pytest.raises(ValueError):  # tests/test_utils.py:512-533
        utils.weighted_average(Quantity(ones, "m"), Quantity(ones, "Hz"))
    with pytest.raises(ValueError):
        utils.weighted_average(ones, Quantity(ones, "m"))
    with pytest.raises(ValueError):
        utils.weighted_average(Quantity(ones, "m"), ones)
    with pytest.raises(ValueError):  # tests/test_utils.py:536-560
        utils.weighted_average(Quantity(ones, "m"), Quantity(np.ones((4, 3, 2)), "m"))
    with pytest.raises(ValueError):
        utils.weighted_average(Quantity(ones, "m"), Quantity(ones, "m"), weights=np.ones(4))
    with pytest.raises(ValueError):
        utils.weighted_average(Quantity(ones, "m"), Quantity(ones, "m"), weights=np.ones((5, 3, 2)))


@pytest.mark.gpu
@pytest.mark.parametrize("case,kind,ax", FOLD_CASES)
def test_gpu_fold_along_delay_matches_oracle(sds, case, kind, ax):
    utils, Quantity = sds
    tag = f"fold_{case}_{kind}_ax{ax}"
    axis = -1 if ax == 2 else 0
    d = Quantity(GOLD[tag + "_delays"], "s")
    arr, err = Quantity(GOLD[tag + "_arr"], "mK^2 Mpc^3"), Quantity(GOLD[tag + "_err"], "mK^2 Mpc^3")
    o, e = utils.fold_along_delay(d, arr, err, axis=axis)
    np.testing.assert_allclose(o.value, GOLD[tag + "_out"], rtol=RTOL)
    np.testing.assert_allclose(e.value, GOLD[tag + "_eout"], rtol=RTOL)
    o, e = utils.fold_along_delay(d, arr, err, weights=GOLD[tag + "_w"], axis=axis)
    np.testing.assert_allclose(o.value, GOLD[tag + "_wout"], rtol=RTOL)
    # complex values with REAL weights: the reference casts the weights to complex64
    # (utils.py:784-787) and then squares and sums them in single precision inside
    # weighted_average; the drop-in rounds the weights to float32 the same way (the folded
    # VALUES agree to 1e-12) but propagates the uncertainty in fp64 -> 1e-7 on that output
    np.testing.assert_allclose(e.value, GOLD[tag + "_weout"], rtol=2e-7 if kind == "cplx" else RTOL)
    if kind == "cplx":
        o, e = utils.fold_along_delay(d, arr, err, weights=GOLD[tag + "_wc"], axis=axis)
        np.testing.assert_allclose(o.value, GOLD[tag + "_wcout"], rtol=RTOL)
        np.testing.assert_allclose(e.value, GOLD[tag + "_wceout"], rtol=RTOL)


@pytest.mark.gpu
def test_gpu_fold_along_delay_reference_tests(sds):
    """tests/test_utils.py:587-766 restated on the drop-in."""
    utils, Quantity = sds
    P = "mK^2 Mpc^3"
    s = lambda v: Quantity(np.asarray(v, dtype=float), "s")  # noqa: E731
    with pytest.raises(ValueError):
        utils.fold_along_delay(s(np.arange(20)), Quantity(np.ones((1, 10, 3)), P), Quantity(np.ones((1, 10, 3)), P), axis=2)
    with pytest.raises(ValueError):
        utils.fold_along_delay(s(np.arange(20)), Quantity(np.ones((1, 10, 20)), P), Quantity(np.ones((1, 10, 21)), P), axis=2)
    with pytest.raises(ValueError):
        utils.fold_along_delay(s(np.arange(21) + 11), Quantity(np.ones((1, 10, 21)), P), Quantity(np.ones((1, 10, 21)), P))
    with pytest.raises(TypeError):  # quantity_input
        utils.fold_along_delay(np.arange(21.0), Quantity(np.ones((1, 10, 21)), P), Quantity(np.ones((1, 10, 21)), P))
    # ones stay ones, unit unchanged (odd and even)
    o, e = utils.fold_along_delay(s(np.arange(-10, 11)), Quantity(np.ones((1, 10, 21)), P), Quantity(np.ones((1, 10, 21)), P))
    assert np.allclose(o.value, np.ones((1, 10, 11))) and str(o.unit) == "Mpc^3 mK^2"
    o, e = utils.fold_along_delay(s(np.arange(-10, 10) + 0.5), Quantity(np.ones((1, 10, 20)), P), Quantity(np.ones((1, 10, 20)), P))
    assert np.allclose(o.value, np.ones((1, 10, 10)))
    # amplitude, with and without weights
    arr = np.ones((1, 10, 21))
    arr[:, :, 11:] *= np.sqrt(2)
    arr[:, :, 10] *= 3
    o, e = utils.fold_along_delay(s(np.arange(-10, 11)), Quantity(arr, P), Quantity(np.ones_like(arr), P))
    want = np.ones((1, 10, 11)) * np.mean([np.sqrt(2), 1])
    want[:, :, 0] = 3
    assert np.allclose(o.value, want)
    errs = np.ones_like(arr)
    errs[:, :, 11:] *= 2
    o, e = utils.fold_along_delay(s(np.arange(-10, 11)), Quantity(arr, P), Quantity(errs, P), weights=1.0 / errs**2)
    want = np.ones((1, 10, 11))
    want[:, :, 1:] *= np.average([np.sqrt(2), 1], weights=1.0 / np.array([2.0, 1.0]) ** 2)
    want[:, :, 0] = 3
    assert np.allclose(o.value, want)
    want_e = np.ones((1, 10, 11))
    want_e[:, :, 1:] *= np.sqrt(1.0 / np.sum(1.0 / np.array([2.0, 1.0]) ** 2))
    assert np.allclose(e.value, want_e)
    # complex values; real ("mismatched") weights
    carr = np.ones((1, 10, 21)) * (1 + 1j)
    carr[:, :, 11:].real *= np.sqrt(2)
    carr[:, :, 10].real *= 5
    carr[:, :, 11:].imag *= np.sqrt(3)
    carr[:, :, 10].imag *= 6
    o, e = utils.fold_along_delay(s(np.arange(-10, 11)), Quantity(carr, P), Quantity(np.ones_like(carr), P))
    want = np.ones((1, 10, 11), dtype=complex)
    want[:, :, 1:] *= np.mean([np.sqrt(2), 1]) + 1j * np.mean([np.sqrt(3), 1])
    want[:, :, 0] = 5 + 6j
    assert np.allclose(o.value, want)
    rerrs = np.ones((1, 10, 21))
    rerrs[:, :, 11:] *= 2
    o, e = utils.fold_along_delay(s(np.arange(-10, 11)), Quantity(carr, P), Quantity(rerrs, P), weights=1.0 / rerrs**2)
    want_e = np.ones((1, 10, 11), dtype=complex)
    want_e[:, :, 1:].real *= 1.0 / np.sqrt(np.sum(1.0 / np.array([2.0, 1.0]) ** 2))
    want_e[:, :, 0] = 1 + 0j
    assert np.allclose(e.value, want_e)


@pytest.mark.gpu
def test_gpu_fold_middle_axis_is_the_transposed_fold(sds):
    """The device path folds any axis (the reference only the first or last)."""
    utils, Quantity = sds
    tag = "fold_odd_cplx_ax2"
    arr, err = GOLD[tag + "_arr"], GOLD[tag + "_err"]          # (2, 3, 21)
    d = Quantity(GOLD[tag + "_delays"], "s")
    o, e = utils.fold_along_delay(d, Quantity(np.ascontiguousarray(arr.transpose(0, 2, 1)), "s"),
                                  Quantity(np.ascontiguousarray(err.transpose(0, 2, 1)), "s"), axis=1)
    np.testing.assert_allclose(o.value.transpose(0, 2, 1), GOLD[tag + "_out"], rtol=RTOL)
    np.testing.assert_allclose(e.value.transpose(0, 2, 1), GOLD[tag + "_eout"], rtol=RTOL)


@pytest.mark.gpu
def test_gpu_bootstrap_array(sds):
    utils, Quantity = sds
    arr = GOLD["boot_arr"]
    for axis in (0, 1, 2):
        out = utils.bootstrap_array(arr, nboot=5, axis=axis, indices=GOLD[f"boot_idx_ax{axis}"])
        assert out.dtype == np.complex128 and np.array_equal(out, GOLD[f"boot_out_ax{axis}"])  # bit-exact gather
    real = utils.bootstrap_array(arr.real.copy(), nboot=5, axis=1, indices=GOLD["boot_idx_ax1"])
    assert np.array_equal(real, GOLD["boot_out_ax1"].real)
    with pytest.raises(ValueError):  # tests/test_utils.py:430-435 style
        utils.bootstrap_array(arr, axis=3)
    # device-drawn index table: reproducible, in range, every source row drawn about equally often
    big = np.arange(64.0)[:, None] * np.ones((1, 3))
    a = utils.bootstrap_array(big, nboot=4000, axis=0, seed=11)
    b = utils.bootstrap_array(big, nboot=4000, axis=0, seed=11)
    c = utils.bootstrap_array(big, nboot=4000, axis=0, seed=12)
    assert a.shape == (64, 4000, 3) and np.array_equal(a, b) and not np.array_equal(a, c)
    counts = np.bincount(a[:, :, 0].astype(int).ravel(), minlength=64)
    assert counts.min() > 0 and a.min() >= 0 and a.max() <= 63
    assert np.all(np.abs(counts / counts.mean() - 1.0) < 0.1)  # 4000 draws per bin: sigma = 1.6 %
