"""Unit tags.  astropy is not available, so the drop-in carries units as small
immutable tags (name -> power) next to plain tensors; only the bookkeeping the
hot path needs (products, powers, equality, the reference's accepted sets) is
implemented.  No numeric conversion happens here: values are always in the SI
base the reference converts to (Hz, s, K, Jy, sr)."""
from fractions import Fraction


class Unit:
    __slots__ = ("_p",)

    def __init__(self, spec=""):
        if isinstance(spec, Unit):
            self._p = dict(spec._p)
            return
        if isinstance(spec, dict):
            self._p = {k: Fraction(v) for k, v in spec.items() if v != 0}
            return
        p = {}
        for tok in str(spec).replace("*", " ").split():
            if tok == "str":  # pyuvdata spells steradian "str" in vis_units ("K str")
                tok = "sr"
            name, _, power = tok.partition("^")
            p[name] = p.get(name, Fraction(0)) + Fraction(power or 1)
        self._p = {k: v for k, v in p.items() if v != 0}

    def __mul__(self, other):
        other = Unit(other)
        p = dict(self._p)
        for k, v in other._p.items():
            p[k] = p.get(k, Fraction(0)) + v
        return Unit(p)

    def __truediv__(self, other):
        return self * (Unit(other) ** -1)

    def __pow__(self, n):
        return Unit({k: v * Fraction(n) for k, v in self._p.items()})

    def __eq__(self, other):
        try:
            return self._p == Unit(other)._p
        except Exception:
            return NotImplemented

    def __hash__(self):
        return hash(tuple(sorted(self._p.items())))

    def is_equivalent(self, other):
        if isinstance(other, (tuple, list)):
            return any(self == o for o in other)
        return self == other

    def __str__(self):
        if not self._p:
            return ""
        return " ".join(k if v == 1 else f"{k}^{v}" for k, v in sorted(self._p.items()))

    __repr__ = lambda self: f"Unit('{self}')"  # noqa: E731


dimensionless_unscaled = Unit("")
Jy, Hz, K, sr, s, ns, m, mK = (Unit(u) for u in ("Jy", "Hz", "K", "sr", "s", "ns", "m", "mK"))


# scale of the prefixed / non-SI unit names the path meets, to the SI base the reference converts to
# with ``.si`` (utils.py:560, 564); a unit made of other names has scale 1
SI_SCALE = {"ns": (1e-9, "s"), "us": (1e-6, "s"), "ms": (1e-3, "s"), "kHz": (1e3, "Hz"), "MHz": (1e6, "Hz"),
            "GHz": (1e9, "Hz"), "mK": (1e-3, "K"), "km": (1e3, "m"), "Mpc": (3.085677581491367e22, "m")}


def to_si(unit):
    """(scale, SI unit) of a unit tag: value_SI = value * scale  (astropy's ``Quantity.si``)."""
    scale, out = 1.0, {}
    for name, power in Unit(unit)._p.items():
        f, base = SI_SCALE.get(name, (1.0, name))
        scale *= f ** float(power)
        out[base] = out.get(base, Fraction(0)) + power
    return scale, Unit(out)


class UnitConversionError(ValueError):
    """Mirrors astropy.units.UnitConversionError (a ValueError subclass there too)."""


class Quantity:
    """A value (torch tensor or numpy array) with a unit tag.

    Mirrors the slice of astropy.units.Quantity the DelaySpectrum API exposes:
    ``.value``, ``.unit``, ``.shape``, ``.dtype``, indexing, ``numpy()``.
    """

    __slots__ = ("value", "unit")

    def __init__(self, value, unit=""):
        if isinstance(value, Quantity):
            value = value.value
        self.value = value
        self.unit = Unit(unit)

    @property
    def shape(self):
        return tuple(self.value.shape)

    @property
    def dtype(self):
        return self.value.dtype

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        n = 1
        for d in self.shape:
            n *= d
        return n

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        return Quantity(self.value[idx], self.unit)

    def numpy(self):
        v = self.value
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        return v

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __repr__(self):
        return f"<Quantity shape={self.shape} dtype={self.dtype} unit='{self.unit}'>"


def is_quantity(x):
    return isinstance(x, Quantity)
