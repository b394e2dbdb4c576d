"""Generate golden vectors by EXECUTING the reference's own simpleDS/utils.py.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference package cannot be imported normally (astropy / pyuvdata /
setuptools_scm are absent), so this script loads the single file
``/root/reference/simpleDS/utils.py`` unmodified under a stub ``astropy`` /
``pyuvdata`` namespace that supplies only the names that module touches at
import time.  The functions exercised below contain no unit arithmetic beyond
``delta_x.si`` (a plain float here), so their numerical behaviour is exactly
the reference's.  Outputs go to ``tests/golden/utils_golden.npz`` (committed);
``tests/test_oracle_golden.py`` checks the oracle against them on any machine.

``jy_to_mk`` is NOT exercised: its value comes from astropy unit conversion,
which the stub does not implement.
"""
import importlib.util
import os
import sys
import types

import numpy as np
from scipy.signal import windows

REF_UTILS = "/root/reference/simpleDS/utils.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "utils_golden.npz")


class _Unit:
    """Absorbs any unit algebra done at import time (utils.py:648)."""

    def __init__(self, name="u"):
        self.name = name

    def _same(self, *_a, **_k):
        return self

    __mul__ = __rmul__ = __truediv__ = __rtruediv__ = __pow__ = _same


class _UnitBase(_Unit):
    pass


class Quantity(np.ndarray):
    """Minimal stand-in: an ndarray with a ``unit`` tag whose ``.si`` is itself."""

    def __new__(cls, value, unit=None):
        obj = np.asarray(value, dtype=float).view(cls)
        obj.unit = unit
        return obj

    def __array_finalize__(self, obj):
        self.unit = getattr(obj, "unit", None)

    @property
    def value(self):
        return np.asarray(self)

    @property
    def si(self):
        return float(np.asarray(self)) if self.ndim == 0 else np.asarray(self)


def _quantity_input(*_a, **_k):
    def deco(fn):
        return fn

    return deco


def load_reference_utils():
    astropy = types.ModuleType("astropy")
    units = types.ModuleType("astropy.units")
    units.Quantity = Quantity
    units.UnitBase = _UnitBase
    units.Unit = lambda s: _Unit(s)
    units.quantity_input = _quantity_input
    units.UnitsError = type("UnitsError", (ValueError,), {})
    for name in ("sr", "mK", "Mpc", "Hz", "s", "m", "K", "Jy"):
        setattr(units, name, _Unit(name))
    const = types.ModuleType("astropy.constants")
    cosmology = types.ModuleType("astropy.cosmology")
    cosmo_units = types.ModuleType("astropy.cosmology.units")
    cosmo_units.littleh = _Unit("littleh")
    cosmology.units = cosmo_units
    astropy.units, astropy.constants, astropy.cosmology = units, const, cosmology
    pyuvdata = types.ModuleType("pyuvdata")
    pyuvdata.utils = types.ModuleType("pyuvdata.utils")
    stubs = {
        "astropy": astropy, "astropy.units": units, "astropy.constants": const,
        "astropy.cosmology": cosmology, "astropy.cosmology.units": cosmo_units,
        "pyuvdata": pyuvdata, "pyuvdata.utils": pyuvdata.utils,
    }
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_ref_simpleds_utils", REF_UTILS)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def cplx(rng, shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def main():
    ref = load_reference_utils()
    rng = np.random.Generator(np.random.PCG64(20261019))
    g = {}

    # --- normalized_fourier_transform, forward and "inverse" branch ---------
    # sizes: reference test sizes 21/31 (t_utils.py:379-412), bundled-file sizes
    # 11/21/203, and the pow-2 production sizes.
    for n in (5, 11, 21, 31, 64, 203, 256, 1024):
        x = cplx(rng, (2, 3, n))
        dx = 492610.84375 if n != 256 else 97656.25
        g[f"ft_in_{n}"] = x
        g[f"ft_dx_{n}"] = np.float64(dx)
        g[f"ft_fwd_bh_{n}"] = np.asarray(
            ref.normalized_fourier_transform(x, Quantity(dx), axis=-1)
        )
        g[f"ft_fwd_box_{n}"] = np.asarray(
            ref.normalized_fourier_transform(
                x, Quantity(dx), axis=-1, taper=windows.boxcar
            )
        )
        g[f"ft_inv_bh_{n}"] = np.asarray(
            ref.normalized_fourier_transform(x, Quantity(1e-9), axis=-1, inverse=True)
        )
    # the reference's own KAT (t_utils.py:379-389): delta, boxcar, dx=1
    fake = np.zeros((1, 13, 21))
    fake[0, 7, 11] += 1
    g["ft_kat_delta"] = np.asarray(
        ref.normalized_fourier_transform(fake, Quantity(1.0), taper=windows.boxcar, axis=2)
    )
    # reference quirk: only the LAST axis works -- broadcast_to of the 1-D window
    # to shape (1, N, 1) raises ValueError (utils.py:549-552)
    try:
        ref.normalized_fourier_transform(cplx(rng, (4, 21, 3)), Quantity(2.5), axis=1)
        g["ft_axis1_raises"] = np.bool_(False)
    except ValueError:
        g["ft_axis1_raises"] = np.bool_(True)

    # --- cross_multiply_array ----------------------------------------------
    a1 = cplx(rng, (2, 2, 5, 3, 7))
    a2 = cplx(rng, (2, 2, 5, 3, 7))
    g["xm_a1"], g["xm_a2"] = a1, a2
    g["xm_auto_axis2"] = ref.cross_multiply_array(a1, axis=2)
    g["xm_cross_axis2"] = ref.cross_multiply_array(a1, a2, axis=2)
    g["xm_cross_axis0"] = ref.cross_multiply_array(a1, a2, axis=0)

    # --- combine_nsamples ----------------------------------------------------
    n1 = rng.integers(0, 61, size=(1, 2, 6, 4, 9)).astype(np.float64)
    n2 = rng.integers(0, 61, size=(1, 2, 6, 4, 9)).astype(np.float64)
    g["ns_n1"], g["ns_n2"] = n1, n2
    g["ns_auto_axis2"] = ref.combine_nsamples(n1, axis=2)
    g["ns_cross_axis2"] = ref.combine_nsamples(n1, n2, axis=2)

    # --- remove_auto_correlations -------------------------------------------
    p = cplx(rng, (2, 4, 4, 3, 5))
    g["rac_in"] = p
    g["rac_out_axes12"] = ref.remove_auto_correlations(p, axes=(1, 2))

    # --- generate_noise: global MT19937 stream, reals first then imags -------
    sig = rng.uniform(0.5, 3.0, size=(2, 3, 4, 5))
    np.random.seed(0)
    g["gn_sigma"] = sig
    g["gn_seed0"] = ref.generate_noise(sig)

    # --- small helpers (row f3 of SURVEY 8f) ---------------------------------
    g["neb_bh_203"] = np.float64(
        ref.noise_equivalent_bandwidth(windows.blackmanharris(203))
    )
    arr = rng.standard_normal((3, 8))
    err = rng.uniform(0.5, 2.0, (3, 8))
    wa, we = ref.weighted_average(arr, err, axis=-1)
    g["wavg_arr"], g["wavg_err"], g["wavg_out"], g["wavg_err_out"] = arr, err, wa, we

    np.savez_compressed(OUT, **g)
    print("wrote", OUT, "with", len(g), "arrays")


if __name__ == "__main__":
    main()
