"""Host-side mirror of the reference's ``simpleDS/utils.py`` for the hot path:
same function names, argument meaning and error behaviour; the arithmetic runs
in libsds.so on the GPU.  Citations are to /root/reference/simpleDS/utils.py."""
import itertools

import numpy as np
import torch
from scipy.signal import windows

from . import ops
from .units import Quantity, Unit, is_quantity

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
    dx = float(np.asarray(delta_x.value))
    out = ops.delay_transform(plan, t, None, dx, inverse=inverse)
    return _wrap(out, unit * delta_x.unit, was_numpy, True)


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
