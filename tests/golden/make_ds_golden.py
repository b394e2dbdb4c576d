"""Golden vectors for the DelaySpectrum-level numerics by EXECUTING the reference's own method
bodies: `DelaySpectrum.calculate_noise_power` (delay_spectrum.py:3544-3583),
`DelaySpectrum.calculate_thermal_sensitivity` (3637-3707), `utils.jy_to_mk` (utils.py:467-483)
and `utils.combine_nsamples` (utils.py:230-278, with `cross_multiply_array`, 334-393).

astropy / pyuvdata do not exist in this image, so the functions are cut out of the reference
source files with `ast` (unmodified text) and executed against `miniunits` below: a small,
dimension-checked stand-in for the handful of astropy.units features those bodies use
(Quantity arithmetic, `.to()`, `<<`, `.value`, `.reshape`, `np.sqrt` / `np.power` / `np.diff`
on quantities) plus the CODATA-2018 constants astropy.constants returns (c, k_B).  Unit
conversion factors therefore come from first principles (SI scales), not from the oracle.
The masked-array statement of the thermal integral (3696-3699) is exercised with VALID samples
only -- how a mask survives the product with an astropy Quantity is version dependent and stays
unpinned.

    python tests/golden/make_ds_golden.py      # build container only (needs /root/reference)
"""
import ast
import os
import textwrap
import types

import numpy as np
from scipy.signal import windows

REF = "/root/reference/simpleDS"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ds_golden.npz")

# ------------------------------------------------------------------------------ miniunits
BASE = ("m", "kg", "s", "K", "sr")


class Unit:
    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        """ndarray <op> Unit (also the in-place forms): attach / combine like astropy does."""
        other = inputs[0] if inputs[1] is self else inputs[1]
        if ufunc is np.left_shift and inputs[1] is self:
            return other.to(self) if isinstance(other, Quantity) else Quantity(np.ma.getdata(other), self)
        if ufunc is np.multiply:
            return self * other
        if ufunc in (np.divide, np.true_divide):
            return other / self if inputs[1] is self else self / other
        return NotImplemented

    def __init__(self, scale, dims):
        self.scale, self.dims = float(scale), tuple(dims)

    def __mul__(self, o):
        if isinstance(o, Unit):
            return Unit(self.scale * o.scale, [a + b for a, b in zip(self.dims, o.dims)])
        if isinstance(o, Quantity):
            return Quantity(np.asarray(o), self * o.unit)
        return Quantity(np.asarray(o, dtype=float), self)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Unit):
            return Unit(self.scale / o.scale, [a - b for a, b in zip(self.dims, o.dims)])
        if isinstance(o, Quantity):
            return Quantity(1.0 / np.asarray(o), self / o.unit)
        return Quantity(1.0 / np.asarray(o, dtype=float), self)

    def __rtruediv__(self, o):
        if isinstance(o, (int, float)):  # 1.0 / unit is a unit in astropy
            return Unit(float(o) / self.scale, [-a for a in self.dims])
        return Quantity(np.asarray(o, dtype=float), Unit(1.0 / self.scale, [-a for a in self.dims]))

    def __pow__(self, p):
        return Unit(self.scale ** p, [a * p for a in self.dims])

    def __rlshift__(self, value):  # ndarray << unit : attach
        return Quantity(np.ma.getdata(value), self)

    def same_dims(self, o):
        return all(abs(a - b) < 1e-12 for a, b in zip(self.dims, o.dims))


def _u(scale, **d):
    return Unit(scale, [d.get(b, 0) for b in BASE])


NAMED = {
    "": _u(1), "m": _u(1, m=1), "s": _u(1, s=1), "K": _u(1, K=1), "mK": _u(1e-3, K=1), "sr": _u(1, sr=1),
    "Hz": _u(1, s=-1), "MHz": _u(1e6, s=-1), "GHz": _u(1e9, s=-1), "Jy": _u(1e-26, kg=1, s=-2),
    "J": _u(1, kg=1, m=2, s=-2), "ns": _u(1e-9, s=1),
}


def parse_unit(s):
    """'GHz', '1/s', 'm/s', 'mK*sr/Jy', 'K^2 sr^2 Hz^6' ..."""
    if isinstance(s, Unit):
        return s
    out = NAMED[""]
    for i, part in enumerate(s.replace(" ", "*").split("/")):
        for tok in [t for t in part.split("*") if t and t != "1"]:
            name, _, p = tok.partition("^")
            u = NAMED[name] ** (float(p) if p else 1.0)
            out = out * u if i == 0 else out / u
    return out


class Quantity(np.ndarray):
    __array_priority__ = 1000

    def __new__(cls, value, unit):
        value = np.asarray(value)
        if not np.iscomplexobj(value):
            value = value.astype(float)
        obj = value.view(cls)
        obj.unit = unit
        return obj

    def __array_finalize__(self, obj):
        self.unit = getattr(obj, "unit", NAMED[""])

    @property
    def value(self):
        return np.asarray(self)

    def to(self, unit):
        unit = parse_unit(unit)
        assert self.unit.same_dims(unit), f"cannot convert dims {self.unit.dims} to {unit.dims}"
        return Quantity(np.asarray(self) * (self.unit.scale / unit.scale), unit)

    def __lshift__(self, unit):
        return self.to(unit)

    @property
    def si(self):
        return Quantity(np.asarray(self) * self.unit.scale, Unit(1.0, self.unit.dims))

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        assert method == "__call__", (ufunc, method)
        vals = [np.asarray(x) if isinstance(x, Quantity) else (x.value if isinstance(x, Quantity) else x) for x in inputs]
        units = [x.unit if isinstance(x, Quantity) else (x if isinstance(x, Unit) else None) for x in inputs]
        vals = [1.0 if isinstance(v, Unit) else v for v in vals]
        one = NAMED[""]
        un = [u if u is not None else one for u in units]
        if ufunc in (np.multiply,):
            return Quantity(ufunc(*vals, **kw), un[0] * un[1])
        if ufunc in (np.divide, np.true_divide):
            return Quantity(ufunc(*vals, **kw), un[0] / un[1])
        if ufunc in (np.add, np.subtract):
            assert un[0].same_dims(un[1]), "adding different dimensions"
            b = np.asarray(vals[1]) * (un[1].scale / un[0].scale)
            return Quantity(ufunc(vals[0], b, **kw), un[0])
        if ufunc is np.sqrt:
            return Quantity(ufunc(vals[0], **kw), un[0] ** 0.5)
        if ufunc is np.square:
            return Quantity(ufunc(vals[0], **kw), un[0] ** 2)
        if ufunc is np.power:
            assert units[1] is None
            if all(abs(d) < 1e-12 for d in un[0].dims):  # dimensionless base: fold the scale in
                return Quantity(ufunc(np.asarray(vals[0]) * un[0].scale, vals[1], **kw), one)
            return Quantity(ufunc(vals[0], vals[1], **kw), un[0] ** float(vals[1]))
        if ufunc in (np.isfinite, np.isnan, np.isinf):
            return ufunc(vals[0], **kw)
        if ufunc in (np.negative, np.absolute):
            return Quantity(ufunc(vals[0], **kw), un[0])
        raise NotImplementedError(ufunc)

    def __pow__(self, p):
        return Quantity(np.asarray(self) ** p, self.unit ** p)

    def __iadd__(self, o):
        return np.add(self, o)

    def reshape(self, *shape):
        return Quantity(np.asarray(self).reshape(*shape), self.unit)

    def __getitem__(self, i):
        return Quantity(np.asarray(self)[i], self.unit)


_np_diff = np.diff


class _MaskedQ:
    """np.ma.masked_invalid(quantity).filled(0) of calculate_noise_power (3575-3582)."""

    def __init__(self, q):
        self.q = q

    def filled(self, v):
        a = np.asarray(self.q).copy()
        a[~np.isfinite(a)] = v
        return Quantity(a, self.q.unit)


class _NP:
    """numpy with the three entry points the bodies call on Quantities."""

    class ma:
        @staticmethod
        def masked_invalid(x):
            return _MaskedQ(x) if isinstance(x, Quantity) else np.ma.masked_invalid(x)

    @staticmethod
    def diff(x, *a, **k):
        return Quantity(_np_diff(np.asarray(x), *a, **k), x.unit) if isinstance(x, Quantity) else _np_diff(x, *a, **k)

    def __getattr__(self, name):
        return getattr(np, name)


def units_module():
    m = types.ModuleType("units")
    for k, v in NAMED.items():
        if k:
            setattr(m, k, v)
    m.Unit = parse_unit
    m.Quantity = Quantity
    m.UnitBase = Unit
    m.dimensionless_unscaled = NAMED[""]
    m.quantity_input = lambda *a, **k: (lambda fn: fn)
    return m


def const_module():
    m = types.ModuleType("const")
    m.c = Quantity(299792458.0, parse_unit("m/s"))          # CODATA 2018 (astropy >= 4)
    m.k_B = Quantity(1.380649e-23, parse_unit("J/K"))
    return m


# ------------------------------------------------------------------------------ reference code
def cut(path, name, cls=None):
    """Source text of function `name` (of class `cls`) of a reference file, decorators dropped."""
    src = open(path).read()
    tree = ast.parse(src)
    nodes = tree.body
    if cls:
        nodes = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next(n for n in nodes if isinstance(n, ast.FunctionDef) and n.name == name)
    lines = src.splitlines()[fn.lineno - 1:fn.end_lineno]
    return textwrap.dedent("\n".join(lines))


def load_reference():
    units, const, npx = units_module(), const_module(), _NP()
    utils = types.ModuleType("utils")
    import copy

    g = dict(np=npx, units=units, const=const, copy=copy, windows=windows)
    exec(cut(f"{REF}/utils.py", "jy_to_mk"), g)
    exec(cut(f"{REF}/utils.py", "cross_multiply_array"), g)
    exec(cut(f"{REF}/utils.py", "combine_nsamples"), g)
    utils.jy_to_mk, utils.combine_nsamples = g["jy_to_mk"], g["combine_nsamples"]
    integrate = types.ModuleType("integrate")
    integrate.trapz = lambda y, x=None, axis=-1: np.trapezoid(np.ma.getdata(y), x=x, axis=axis)
    g2 = dict(np=npx, units=units, utils=utils, integrate=integrate)
    exec(cut(f"{REF}/delay_spectrum.py", "calculate_noise_power", "DelaySpectrum"), g2)
    exec(cut(f"{REF}/delay_spectrum.py", "calculate_thermal_sensitivity", "DelaySpectrum"), g2)
    import copy

    uvutils = types.ModuleType("uvutils")
    uvutils._get_iterable = lambda x: x if hasattr(x, "__iter__") else (x,)
    g3 = dict(np=npx, units=units, copy=copy, uvutils=uvutils)
    exec(cut(f"{REF}/delay_spectrum.py", "select_spectral_windows", "DelaySpectrum"), g3)
    exec(cut(f"{REF}/delay_spectrum.py", "_take_spectral_windows_from_data_like_array", "DelaySpectrum"), g3)
    import warnings

    exec(cut(f"{REF}/utils.py", "normalized_fourier_transform"), g)  # needs `windows` default arg
    utils.normalized_fourier_transform, utils.cross_multiply_array = g["normalized_fourier_transform"], g["cross_multiply_array"]
    g4 = dict(np=npx, units=units, utils=utils, warnings=warnings)
    for name in ("normalized_fourier_transform", "delay_transform", "calculate_delay_spectrum"):
        exec(cut(f"{REF}/delay_spectrum.py", name, "DelaySpectrum"), g4)
    utils.ds_methods = {k: g4[k] for k in ("normalized_fourier_transform", "delay_transform",
                                           "calculate_delay_spectrum")}
    utils.thermal = g2["calculate_thermal_sensitivity"]
    utils.select_spectral_windows = g3["select_spectral_windows"]
    utils.take_windows = g3["_take_spectral_windows_from_data_like_array"]
    return g2["calculate_noise_power"], g2["calculate_thermal_sensitivity"], utils


def fake_self(rng, S, U, P, B, T, F, pols, trcvr_inf=False):
    f0 = 100e6 + 20e6 * np.arange(S)[:, None]
    freq = f0 + 97656.25 * np.arange(F)[None]
    trcvr = 100.0 + 50.0 * rng.random((S, F))
    if trcvr_inf:
        trcvr[0, 3] = np.inf
    ns = rng.integers(1, 61, (S, U, P, B, T, F)).astype(float)
    o = types.SimpleNamespace(
        Nspws=S, Nuv=U, Npols=P, Nbls=B, Ntimes=T, Nfreqs=F,
        freq_array=Quantity(freq, NAMED["Hz"]), trcvr=Quantity(trcvr, NAMED["K"]),
        beam_area=Quantity(0.5 + rng.random((S, P, F)), NAMED["sr"]),
        integration_time=Quantity(5.0 + 10.0 * rng.random((B, T)), NAMED["s"]),
        nsample_array=ns, polarization_array=np.array(pols), taper=windows.blackmanharris,
    )
    return o


def main():
    noise_power, thermal, utils = load_reference()
    rng = np.random.Generator(np.random.PCG64(20261020))
    g = {}
    fr = np.array([100e6, 150e6, 199.5e6])
    g["jy_freqs"] = fr
    g["jy_to_mk"] = np.asarray(utils.jy_to_mk(Quantity(fr, NAMED["Hz"])).to("mK*sr/Jy"))
    for tag, (S, U, P, B, T, F, pols, inf) in dict(
            a=(2, 1, 2, 4, 3, 11, [-5, -6], False), b=(1, 2, 3, 3, 2, 21, [1, 2, -7], False),
            c=(1, 1, 1, 3, 2, 8, [-5], True)).items():
        o = fake_self(rng, S, U, P, B, T, F, pols, trcvr_inf=inf)
        for k in ("freq_array", "trcvr", "beam_area", "integration_time"):
            g[f"{tag}_{k}"] = np.asarray(getattr(o, k))
        g[f"{tag}_nsample"] = o.nsample_array
        g[f"{tag}_pols"] = o.polarization_array
        sig = noise_power(o)
        assert sig.unit.same_dims(NAMED["Jy"]) and abs(sig.unit.scale / NAMED["Jy"].scale - 1) < 1e-15
        g[f"{tag}_noise_power_jy"] = np.asarray(sig)
        if not inf:
            thermal(o)
            tp = o.thermal_power
            want_unit = parse_unit("K^2 sr^2 Hz^6")
            assert tp.unit.same_dims(want_unit) and abs(tp.unit.scale / want_unit.scale - 1) < 1e-15
            g[f"{tag}_thermal_power"] = np.asarray(tp)
    # select_spectral_windows + _take_spectral_windows_from_data_like_array (3196-3338): the
    # window gather (integer index math, must be bit-exact) and the delay axis
    for tag, (U, P, B, T, F, wins) in dict(
            w1=(1, 2, 3, 2, 21, [(0, 9), (10, 19)]), w2=(2, 1, 2, 3, 21, [(0, 10), (10, 20)]),
            w3=(1, 1, 2, 2, 203, [(95, 115)])).items():
        o = fake_self(rng, 1, U, P, B, T, F, [-5, -6][:P])
        o.data_array = rng.standard_normal((1, U, P, B, T, F)) + 1j * rng.standard_normal((1, U, P, B, T, F))
        o.noise_array = rng.standard_normal((1, U, P, B, T, F)) + 1j * rng.standard_normal((1, U, P, B, T, F))
        o.flag_array = rng.random((1, U, P, B, T, F)) < 0.2
        o.beam_sq_area = Quantity(rng.random((1, P, F)), NAMED["sr"])
        o.update_cosmology = lambda: None
        o.check = lambda **k: True
        o._take_spectral_windows_from_data_like_array = types.MethodType(utils.take_windows, o)
        for k in ("data_array", "noise_array", "nsample_array", "flag_array"):
            g[f"{tag}_in_{k}"] = np.asarray(getattr(o, k)).copy()
        for k in ("freq_array", "trcvr", "beam_area", "beam_sq_area"):
            g[f"{tag}_in_{k}"] = np.asarray(getattr(o, k)).copy()
        g[f"{tag}_wins"] = np.asarray(wins)
        utils.select_spectral_windows(o, spectral_windows=wins, inplace=True)
        for k in ("data_array", "noise_array", "nsample_array", "flag_array", "freq_array", "trcvr",
                  "beam_area", "beam_sq_area"):
            g[f"{tag}_out_{k}"] = np.asarray(getattr(o, k))
        g[f"{tag}_out_delay_ns"] = np.asarray(o.delay_array.to("ns"))
        assert (o.Nspws, o.Nfreqs, o.Ndelays) == (len(wins), wins[0][1] - wins[0][0] + 1, o.Nfreqs)
    # the driver: DelaySpectrum.normalized_fourier_transform / delay_transform /
    # calculate_delay_spectrum (3487-3542, 3585-3635) calling the real utils functions
    for tag, (U, P, B, T, F) in dict(d1=(1, 2, 4, 3, 21), d2=(2, 1, 3, 2, 16)).items():
        o = fake_self(rng, 1, U, P, B, T, F, [-5, -6][:P])
        shape = (1, U, P, B, T, F)
        o.data_array = Quantity(rng.standard_normal(shape) + 1j * rng.standard_normal(shape), NAMED["Jy"])
        o.noise_array = Quantity(rng.standard_normal(shape) + 1j * rng.standard_normal(shape), NAMED["Jy"])
        o.flag_array = rng.random(shape) < 0.15
        o.data_type = "frequency"
        o.set_delay = lambda: setattr(o, "data_type", "delay")
        o.set_frequency = lambda: setattr(o, "data_type", "frequency")
        o.update_cosmology = lambda **k: None
        for name, fn in utils.ds_methods.items():
            setattr(o, name, types.MethodType(fn, o))
        o.calculate_thermal_sensitivity = types.MethodType(utils.thermal, o)
        for k in ("data_array", "noise_array", "flag_array", "nsample_array", "freq_array", "trcvr",
                  "beam_area", "integration_time"):
            g[f"{tag}_in_{k}"] = np.asarray(getattr(o, k)).copy()
        g[f"{tag}_pols"] = o.polarization_array
        o.calculate_delay_spectrum()
        assert o.data_type == "delay"
        for k in ("data_array", "noise_array", "power_array", "noise_power", "thermal_power"):
            g[f"{tag}_out_{k}"] = np.asarray(getattr(o, k))
    np.savez_compressed(OUT, **g)
    print(OUT, {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
