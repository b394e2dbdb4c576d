"""Host-side mirror of the reference's ``simpleDS/utils.py`` for the hot path:
same function names, argument meaning and error behaviour; the arithmetic runs
in libsds.so on the GPU.  Citations are to /root/reference/simpleDS/utils.py."""
import itertools

import numpy as np
import torch
from scipy.signal import windows

from . import ops
from .units import Quantity, Unit, is_quantity, to_si

C_M_PER_S = 299792458.0  # astropy.constants.c
K_B_J_PER_K = 1.380649e-23  # astropy.constants.k_B

_noise_calls = itertools.count()


def _device():
    ops._lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _split(x):
    """-> (value, unit, was_numpy)."""
    unit = Unit("")
    if is_quantity(x):
        x, unit = x.value, x.unit
    if isinstance(x, list):
        x = np.asarray(x)
    return x, unit, not isinstance(x, torch.Tensor)


def _to_device(x, complex_=False):
    if isinstance(x, torch.Tensor):
        t = x.to(_device())
    else:
        t = torch.as_tensor(np.ascontiguousarray(x), device=_device())
    if complex_ and not t.is_complex():
        t = t.to(torch.complex128 if t.dtype != torch.float32 else torch.complex64)
    return t


def _wrap(t, unit, as_numpy, force_quantity):
    v = t.cpu().numpy() if as_numpy else t
    return Quantity(v, unit) if force_quantity else v


def normalized_fourier_transform(
    data_array, delta_x, axis=-1, taper=windows.blackmanharris, inverse=False
):
    """utils.py:509-566.  ``delta_x`` must be a Quantity (value in SI units)."""
    if not is_quantity(delta_x):
        raise ValueError(
            "delta_x must be an astropy Quantity object. "
            "value was : {df}".format(df=delta_x)
        )
    value, unit, was_numpy = _split(data_array)
    ndim = len(value.shape)
    if axis not in (-1, ndim - 1):
        # the reference's window broadcast (utils.py:549-552) only works for the
        # last axis; numpy raises ValueError there
        raise ValueError(
            "operands could not be broadcast together: the taper can only be "
            "applied along the last axis"
        )
    t = _to_device(value, complex_=True)
    n = t.shape[-1]
    plan = ops.get_plan(n, ops.math_of(t.dtype), taper)
    # the reference multiplies by delta_x.si (utils.py:560, 564): a step given in MHz or ns scales
    # the result like the same step in Hz or s, and the output carries the SI unit
    si_scale, si_unit = to_si(delta_x.unit)
    dx = float(np.asarray(delta_x.value)) * si_scale
    out = ops.delay_transform(plan, t, None, dx, inverse=inverse)
    return _wrap(out, unit * si_unit, was_numpy, True)


def cross_multiply_array(array_1, array_2=None, axis=0):
    """utils.py:334-393: conj(array_1) x array_2 with a new axis after ``axis``."""
    v1, u1, np1 = _split(array_1)
    if array_2 is None:
        v2, u2 = v1, u1  # array_2 = deepcopy(array_1) keeps its unit (utils.py:364-377)
    else:
        v2, u2, _ = _split(array_2)
    if tuple(v2.shape) != tuple(v1.shape):
        raise ValueError(
            "array_1 and array_2 must have the same shapes. "
            "array_1 has shape {a1} but array_2 has shape {a2}".format(
                a1=tuple(v1.shape), a2=tuple(v2.shape)
            )
        )
    shape = tuple(v1.shape)
    axis = axis % len(shape)
    t1 = _to_device(v1)
    t2 = t1 if array_2 is None else _to_device(v2)
    real_in = not (t1.is_complex() or t2.is_complex())
    real_dtype = torch.promote_types(t1.dtype, t2.dtype) if real_in else None
    c1, c2 = _to_device(t1, complex_=True), _to_device(t2, complex_=True)
    cd = torch.promote_types(c1.dtype, c2.dtype)
    c1, c2 = c1.to(cd), c2.to(cd)
    lead = int(np.prod(shape[:axis], dtype=np.int64))
    trail = int(np.prod(shape[axis + 1:], dtype=np.int64))
    b = shape[axis]
    out = ops.cross_power_full(c1.reshape(lead, 1, b, 1, trail), c2.reshape(lead, 1, b, 1, trail))
    out = out.reshape(shape[:axis] + (b, b) + shape[axis + 1:])
    if real_in:
        out = out.real.to(real_dtype if real_dtype.is_floating_point else torch.float64)
    has_unit = is_quantity(array_1) or is_quantity(array_2)
    return _wrap(out, u1 * u2, np1, has_unit)


def combine_nsamples(nsample_1, nsample_2=None, axis=-1):
    """utils.py:230-278: sqrt of the cross-multiplied sample counts."""
    v1, _, np1 = _split(nsample_1)
    if nsample_2 is None:
        nsample_2 = nsample_1
    v2, _, _ = _split(nsample_2)
    if not tuple(v1.shape) == tuple(v2.shape):
        raise ValueError(
            "nsample_1 and nsample_2 must have same shape, "
            "but nsample_1 has shape {d1_s} and "
            "nsample_2 has shape {d2_s}".format(d1_s=tuple(v1.shape), d2_s=tuple(v2.shape))
        )
    axis = axis % len(v1.shape)
    out = cross_multiply_array(_to_device(v1).to(torch.float64), _to_device(v2).to(torch.float64), axis=axis)
    out = torch.sqrt(out)
    return out.cpu().numpy() if np1 else out


def remove_auto_correlations(data_array, axes=(0, 1)):
    """utils.py:281-331 (pure indexing)."""
    if not np.shape(axes)[0] == 2:
        raise ValueError("Shape must be a length 2 tuple/array/list of axis indices.")
    if axes[0] != axes[1] - 1:
        raise ValueError(
            "Axes over which diagonal components are to be remove must be adjacent."
        )
    value, unit, was_numpy = _split(data_array)
    if value.shape[axes[0]] != value.shape[axes[1]]:
        raise ValueError(
            "The axes over which diagonal components are to be "
            "removed must have the same shape."
        )
    n = value.shape[axes[0]]
    if was_numpy:
        idx = np.logical_not(np.diag(np.ones(n, dtype=bool)))
        moved = np.moveaxis(np.moveaxis(value, axes[0], 0), axes[1], 1)
        out = np.moveaxis(moved[idx], 0, axes[0])
    else:
        idx = ~torch.eye(n, dtype=torch.bool, device=value.device)
        moved = torch.movedim(torch.movedim(value, axes[0], 0), axes[1], 1)
        out = torch.movedim(moved[idx], 0, axes[0])
    return Quantity(out, unit) if is_quantity(data_array) else out


def generate_noise(noise_power, seed=None, normals=None):
    """utils.py:486-506: sigma * (g1 + 1j g2) / sqrt(2).

    The reference draws from numpy's global generator; here the normals come from
    Philox4x32-10 on the device (``seed``; a fresh stream per call when omitted)
    or are injected with ``normals=(g_re, g_im)`` for bit-level parity tests."""
    value, unit, was_numpy = _split(noise_power)
    sig = _to_device(value).to(torch.float64)
    if seed is None:
        seed = 0x5D5 + next(_noise_calls)
    out = ops.generate_noise(sig, seed=seed, normals=normals)
    return _wrap(out, unit, was_numpy, is_quantity(noise_power))


def noise_equivalent_bandwidth(window):
    """utils.py:212-227."""
    window = np.asarray(window)
    return np.sum(window) ** 2 / (np.sum(window**2) * len(window))


def jy_to_mk(freqs):
    """utils.py:467-483: c^2 / (2 nu^2 k_B) in mK sr / Jy; ``freqs`` in Hz."""
    if is_quantity(freqs):
        if not freqs.unit == Unit("Hz"):
            raise ValueError("freqs must be a frequency Quantity in Hz")
        freqs = freqs.value
    freqs = np.asarray(freqs, dtype=np.float64)
    return Quantity(1e-23 * C_M_PER_S**2 / (2.0 * freqs**2 * K_B_J_PER_K), "mK sr Jy^-1")


# ---- downstream estimators on the averaged spectra ---------------------------
def bootstrap_array(array, nboot=100, axis=0, seed=None, indices=None):
    """utils.py:171-209: resample with replacement along ``axis``; the output gains an axis of
    length ``nboot`` right after it.

    The reference draws the (n, nboot) index table from numpy's global generator; here it comes
    from Philox4x32-10 on the device (``seed``; a fresh stream per call when omitted) or is
    injected with ``indices`` for bit-level parity tests.  The gather runs in libsds.so."""
    value, unit, was_numpy = _split(array)
    if axis >= len(value.shape):
        raise ValueError(
            "Specified axis must be shorter than the length "
            "of input array.\n"
            "axis value was {0} but array has {1} dimensions".format(axis, len(value.shape))
        )
    axis = axis % len(value.shape)
    n = int(value.shape[axis])
    t = _to_device(value)
    if not t.is_complex():
        t = t.to(torch.float64)
    elif t.dtype != torch.complex128:
        t = t.to(torch.complex128)
    if indices is None:
        if seed is None:
            seed = 0xB007 + next(_noise_calls)
        idx = ops.bootstrap_indices(n, int(nboot), seed)
    else:
        idx = torch.as_tensor(np.asarray(indices), device=t.device).to(torch.int64)
        if tuple(idx.shape) != (n, int(nboot)):
            raise ValueError(f"indices must have shape ({n}, {int(nboot)}), got {tuple(idx.shape)}")
        if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= n):
            raise ValueError("indices out of range for the resampled axis")
    out = ops.bootstrap_gather(t, idx, axis)
    return _wrap(out, unit, was_numpy, is_quantity(array))


def weighted_average(array, uncertainty, weights=None, axis=-1):
    """utils.py:569-645: weighted average along ``axis`` and the propagated uncertainty
    (inverse-variance weights when ``weights`` is None)."""
    if is_quantity(array):
        if is_quantity(uncertainty):
            if not uncertainty.unit == array.unit:
                raise ValueError(  # astropy raises UnitConversionError (a ValueError) here
                    f"'{uncertainty.unit}' and '{array.unit}' are not convertible"
                )
        else:
            raise ValueError(
                "Input array is a Quantity Object but "
                "uncertainty is not. Either both arrays must be "
                "Quantity objects or neither must be."
            )
    elif is_quantity(uncertainty):
        raise ValueError(
            "Input array is a not Quantity Objcet but "
            "uncertainty is. Either both arrays must be "
            "Quantity objects or neither must be."
        )
    value, unit, was_numpy = _split(array)
    sig, _, _ = _split(uncertainty)
    if not tuple(value.shape) == tuple(sig.shape):
        raise ValueError(
            "Input array and uncertainties must have the same "
            "shape. Array shape: {array}, Uncertainty shape: "
            "{error}".format(array=tuple(value.shape), error=tuple(sig.shape))
        )
    w = None
    if weights is not None:
        w, _, _ = _split(weights)
        if len(w.shape) == 1 and int(np.prod(w.shape)) != value.shape[axis]:
            raise ValueError(
                "1-Dimenionals weights must have the same shape "
                "as the axis over which the average is taken."
            )
        elif not tuple(w.shape) == tuple(value.shape):  # (sic) utils.py:636: also 1-D weights of N-D input
            raise ValueError(
                "Input array and uncertainties must have the same "
                "shape. Array shape: {array}, Uncertainty shape: "
                "{weights}".format(array=tuple(value.shape), weights=tuple(w.shape))
            )
        w = _to_device(w).to(torch.float64)
    axis = axis % len(value.shape)
    out, osig = ops.weighted_average(_to_device(value).to(torch.float64), _to_device(sig).to(torch.float64),
                                     w, axis)
    q = is_quantity(array)
    return _wrap(out, unit, was_numpy, q), _wrap(osig, unit, was_numpy, q)


_FOLD_UNITS = (Unit("mK^2 Mpc^3"), Unit("mK^2 Mpc^3 littleh^-3"), Unit("s"))


def fold_along_delay(delays, array, uncertainty, weights=None, axis=-1):
    """utils.py:651-796: average the bins of equal |delay| along ``axis``.

    ``delays`` and ``array`` must be Quantities (time; power or time), as the reference's
    ``quantity_input`` demands.  The split of the axis follows utils.py:730-770 line by line on
    the host (it is index arithmetic on the 1-D delays); the two-point weighted average and the
    uncertainty propagation of every bin run in libsds.so.  Unlike the reference (which indexes
    the 1-D delays with the array's axis, so that only axis = -1 / 0 work) any axis is folded."""
    if not is_quantity(delays) or not delays.unit == Unit("s"):
        raise TypeError("Argument 'delays' to function 'fold_along_delay' must be a time Quantity")
    if not is_quantity(array) or not array.unit.is_equivalent(_FOLD_UNITS):
        raise TypeError("Argument 'array' to function 'fold_along_delay' must be in one of "
                        "mK^2*Mpc^3, mK^2*Mpc^3/littleh^3, time")
    d = np.asarray(delays.value.cpu() if isinstance(delays.value, torch.Tensor) else delays.value, dtype=np.float64)
    value, unit, was_numpy = _split(array)
    sig, _, _ = _split(uncertainty)
    n = len(d)
    if value.shape[axis] != n:
        raise ValueError(
            "Input array must have same length as the "
            "delays along the specified axis."
            "Axis given was {0}, the array has length {1} "
            "but delays are length {2}".format(axis, value.shape[axis], n)
        )
    if n % 2 != 0 and np.abs(d).min() != 0:
        raise ValueError(
            "Input delays must have either a delay=0 bin "
            "as the central value or have an even size."
        )
    if not tuple(value.shape) == tuple(sig.shape):
        raise ValueError(
            "Input array and uncertainties must have the same "
            "shape. Array shape: {array}, Uncertainty shape: "
            "{error}".format(array=tuple(value.shape), error=tuple(sig.shape))
        )
    axis = axis % len(value.shape)
    x = _to_device(value)
    s = _to_device(sig)
    is_cplx = bool(x.is_complex() and bool(torch.any(x.imag != 0)))
    # split of the axis, utils.py:730-770
    if np.abs(d).min() == 0 and n % 2:
        z = int(np.argmin(np.abs(d)))
        pos0, neg0, nout, zero_bin = z, z, n - z, True
        if n - z != z + 1:
            raise ValueError("all input arrays must have the same shape")  # np.stack, utils.py:772
    else:
        idx = np.where(np.logical_and(d >= 0, np.abs(d) == np.abs(d).min()))[0]
        if idx.size != 1:
            raise ValueError("Input delays must split into a negative and a non-negative half")
        sp = int(idx[0])
        pos0, neg0, nout, zero_bin = sp, sp - 1, n - sp, False
        if n - sp != sp:
            raise ValueError("all input arrays must have the same shape")  # np.stack, utils.py:772
    w = None
    if weights is not None:
        wv, _, _ = _split(weights)
        w = _to_device(wv)
        if tuple(w.shape) != tuple(x.shape):
            raise ValueError(
                "Input array and uncertainties must have the same "
                "shape. Array shape: {array}, Uncertainty shape: "
                "{weights}".format(array=tuple(x.shape), weights=tuple(w.shape))
            )
    if not is_cplx:  # utils.py:776-779: real parts, weights as given
        xr = (x.real if x.is_complex() else x).to(torch.float64)
        sr = (s.real if s.is_complex() else s).to(torch.float64)
        if w is not None and w.is_complex():
            if bool(torch.any(w.imag != 0)):
                raise ValueError("complex weights need complex values (np.average of real values with "
                                 "complex weights is not supported by the device path)")
            w = w.real
        wr = None if w is None else w.to(torch.float64)
        out, osig = ops.fold_along_delay(xr, sr, wr, axis, pos0, neg0, nout, zero_bin)
    else:  # utils.py:780-794: real and imaginary parts with the real / imaginary weights
        xc = x.to(torch.complex128)
        sc = s.to(torch.complex128)
        if w is None:
            wc = None  # the kernel's default is 1 + 1j
        else:
            if w.is_complex():
                wc = w.to(torch.complex128)
                if not bool(torch.any(wc.imag != 0)):
                    wc = torch.complex(wc.real, torch.ones_like(wc.real))
            else:
                # (sic) utils.py:784-787: real weights are cast to complex64 -- single precision
                w32 = w.to(torch.float64).to(torch.float32).to(torch.float64)
                wc = torch.complex(w32, torch.ones_like(w32))
        out, osig = ops.fold_along_delay(xc, sc, wc, axis, pos0, neg0, nout, zero_bin)
    return _wrap(out, unit, was_numpy, True), _wrap(osig, unit, was_numpy, True)
