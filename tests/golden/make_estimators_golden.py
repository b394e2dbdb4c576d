"""Golden vectors for the downstream estimators, made by EXECUTING the reference's own
``simpleDS/utils.py`` (``weighted_average``, ``fold_along_delay``, ``bootstrap_array``).

Build container only (needs /root/reference):  python tests/golden/make_estimators_golden.py

Same loader as ``make_golden.py`` (unmodified utils.py under a stub astropy namespace); the
Quantity stand-in here re-wraps the results of numpy functions (np.split / np.stack / ...
drop ndarray subclasses, and ``fold_along_delay`` asks for ``.imag.value`` afterwards).
Outputs: tests/golden/estimators_golden.npz (committed).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

OUT = os.path.join(HERE, "estimators_golden.npz")


class Q(np.ndarray):
    """ndarray with ``.value`` / ``.unit`` that survives numpy functions."""

    __array_priority__ = 100.0

    def __new__(cls, value, unit=None):
        obj = np.asarray(value).view(cls)
        obj.unit = unit
        return obj

    def __array_finalize__(self, obj):
        self.unit = getattr(obj, "unit", None)

    @property
    def value(self):
        return np.asarray(self)

    def __lshift__(self, unit):  # `uncertainty << array.unit` (utils.py:603)
        return self

    def __array_function__(self, func, types, args, kwargs):
        def strip(a):
            if isinstance(a, Q):
                return np.asarray(a)
            if isinstance(a, (list, tuple)):
                return type(a)(strip(x) for x in a)
            return a

        def wrap(r):
            if isinstance(r, np.ndarray):
                return r.view(Q)
            if isinstance(r, (list, tuple)):
                return type(r)(wrap(x) for x in r)
            return r

        return wrap(func(*strip(args), **{k: strip(v) for k, v in kwargs.items()}))


def main():
    mg.Quantity = Q  # the loader publishes mg.Quantity as astropy.units.Quantity
    ref = mg.load_reference_utils()
    rng = np.random.Generator(np.random.PCG64(20261020))
    g = {}

    # ---- weighted_average ------------------------------------------------------------------
    for name, shape, axis in (("a", (3, 8), -1), ("b", (4, 5, 6), 0), ("c", (2, 7, 3), 1), ("d", (9,), 0)):
        arr = rng.standard_normal(shape)
        err = rng.uniform(0.5, 2.0, shape)
        w = rng.uniform(0.1, 3.0, shape)
        g[f"wa_{name}_arr"], g[f"wa_{name}_err"], g[f"wa_{name}_w"] = arr, err, w
        g[f"wa_{name}_axis"] = np.int64(axis)
        o, e = ref.weighted_average(arr, err, axis=axis)
        g[f"wa_{name}_ivar_out"], g[f"wa_{name}_ivar_err"] = np.asarray(o), np.asarray(e)
        o, e = ref.weighted_average(arr, err, weights=w, axis=axis)
        g[f"wa_{name}_w_out"], g[f"wa_{name}_w_err"] = np.asarray(o), np.asarray(e)
    # a 1-D weight vector is only accepted for 1-D input (utils.py:631-641)
    try:
        ref.weighted_average(np.ones((3, 4)), np.ones((3, 4)), weights=np.ones(4), axis=-1)
        g["wa_1d_weights_nd_array_raises"] = np.bool_(False)
    except ValueError:
        g["wa_1d_weights_nd_array_raises"] = np.bool_(True)

    # ---- fold_along_delay ------------------------------------------------------------------
    cases = {
        "odd": np.arange(-10, 11) * 1e-7,                 # zero bin in the middle
        "even": (np.arange(-10, 10) + 0.5) * 1e-7,        # no zero bin, even size
        "fft_even": np.fft.fftshift(np.fft.fftfreq(16, d=97656.25)),  # zero bin at n/2, even size
    }
    for cname, delays in cases.items():
        n = len(delays)
        for kind in ("real", "cplx"):
            # the reference indexes the 1-D delays with the ARRAY's axis (utils.py:728, 757): only
            # axis = -1 and axis = 0 work
            for axis, shape in ((-1, (2, 3, n)), (0, (n, 2, 3))):
                tag = f"fold_{cname}_{kind}_ax{axis % 3}"
                if kind == "real":
                    arr = rng.standard_normal(shape)
                    err = rng.uniform(0.5, 2.0, shape)
                else:
                    arr = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
                    err = rng.uniform(0.5, 2.0, shape) + 1j * rng.uniform(0.5, 2.0, shape)
                wr = rng.uniform(0.1, 3.0, shape)
                g[tag + "_delays"], g[tag + "_arr"], g[tag + "_err"], g[tag + "_w"] = delays, arr, err, wr
                o, e = ref.fold_along_delay(Q(delays), Q(arr), Q(err), axis=axis)
                g[tag + "_out"], g[tag + "_eout"] = np.asarray(o), np.asarray(e)
                # real weights; for complex values this takes the complex64 path of utils.py:784-787
                o, e = ref.fold_along_delay(Q(delays), Q(arr), Q(err), weights=Q(wr), axis=axis)
                g[tag + "_wout"], g[tag + "_weout"] = np.asarray(o), np.asarray(e)
                if kind == "cplx":
                    wc = wr + 1j * rng.uniform(0.1, 3.0, shape)
                    g[tag + "_wc"] = wc
                    o, e = ref.fold_along_delay(Q(delays), Q(arr), Q(err), weights=Q(wc), axis=axis)
                    g[tag + "_wcout"], g[tag + "_wceout"] = np.asarray(o), np.asarray(e)

    # ---- bootstrap_array -------------------------------------------------------------------
    arr = rng.standard_normal((4, 6, 3)) + 1j * rng.standard_normal((4, 6, 3))
    g["boot_arr"] = arr
    for axis in (0, 1, 2):
        np.random.seed(7 + axis)
        g[f"boot_out_ax{axis}"] = ref.bootstrap_array(arr, nboot=5, axis=axis)
        np.random.seed(7 + axis)
        g[f"boot_idx_ax{axis}"] = np.random.choice(arr.shape[axis], size=(arr.shape[axis], 5), replace=True)

    try:
        ref.fold_along_delay(Q(cases["odd"]), Q(np.ones((2, 21, 3))), Q(np.ones((2, 21, 3))), axis=1)
        g["fold_axis1_raises"] = np.bool_(False)
    except ValueError:  # numpy.exceptions.AxisError is a ValueError
        g["fold_axis1_raises"] = np.bool_(True)

    np.savez_compressed(OUT, **g)
    print("wrote", OUT, "with", len(g), "arrays")


if __name__ == "__main__":
    main()
