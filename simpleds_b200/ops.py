"""Tensor-level wrappers over the C ABI (include/sds.h).  Every function takes
and returns CUDA torch tensors and enqueues on torch's current stream; PyTorch
is only the allocator/stream plumbing here -- all arithmetic is in libsds.so."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (SDS_C64, SDS_C128, SDS_F32, SDS_F64, SDS_METHOD_EXPLICIT,
                   SDS_METHOD_FACTORISED, check, complex_dtype_code, ptr, real_dtype_code,
                   stream_ptr)

_PLAN_CACHE = {}


def get_plan(nfreq, math, taper):
    """Plan cache keyed by length, math and the taper's values."""
    values = np.ascontiguousarray(np.asarray(taper(int(nfreq)), dtype=np.float64))
    if values.shape != (int(nfreq),):
        raise ValueError(
            f"taper({nfreq}) must return {nfreq} values, got shape {values.shape}"
        )
    key = (int(nfreq), int(math), torch.cuda.current_device(), values.tobytes())
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if len(_PLAN_CACHE) > 64:
            _PLAN_CACHE.clear()
        plan = _PLAN_CACHE[key] = _lib.Plan(nfreq, math, values)
    return plan


def math_of(dtype):
    return SDS_F64 if dtype in (torch.complex128, torch.float64) else SDS_F32


def take_windows(arr, chans):
    """ds.py:3315-3338.  arr (S0, R..., F0) any 1/4/8/16-byte dtype -> (S, R..., F).

    ``chans`` int array (S, F) of flat channel indices into S0*F0."""
    _lib.require_cuda()
    chans = np.asarray(chans)
    S, F = chans.shape
    S0, F0 = arr.shape[0], arr.shape[-1]
    if chans.size and (chans.min() < 0 or chans.max() >= S0 * F0):
        raise IndexError("spectral window channel index out of range")
    a = arr.contiguous()
    nrows = a.numel() // (S0 * F0) if S0 * F0 else 0
    out = torch.empty((S,) + tuple(a.shape[1:-1]) + (F,), dtype=a.dtype, device=a.device)
    ch = torch.as_tensor(chans.astype(np.int32).ravel(), device=a.device)
    view = a.view(torch.uint8) if a.dtype == torch.bool else a
    oview = out.view(torch.uint8) if out.dtype == torch.bool else out
    check(
        _lib.load().sds_take_windows(ptr(view), ptr(oview), a.element_size(), nrows, S0, F0,
                                     ptr(ch), S, F, stream_ptr()),
        "sds_take_windows",
    )
    return out


def delay_transform(plan, data, flags, delta_x, inverse=False, win_start=None, out_dtype=None):
    """utils.py:509-566 with the flag mask of ds.py:3503-3517 along the last axis.

    data (..., F0) complex; flags same shape (bool/uint8) or None.  Without
    ``win_start`` F0 must equal plan.nfreq and the result has data's shape; with
    it the result is (nwin, ..., plan.nfreq)."""
    _lib.require_cuda()
    d = data.contiguous()
    F0 = d.shape[-1]
    nrows = d.numel() // F0 if F0 else 0
    N = plan.nfreq
    if win_start is None:
        if F0 != N:
            raise ValueError(f"last axis has {F0} channels but the plan is for {N}")
        wins, lead = [0], ()
    else:
        wins, lead = [int(w) for w in win_start], (len(win_start),)
    want = torch.complex128 if plan.math == SDS_F64 else torch.complex64
    if out_dtype is not None and out_dtype != want:
        raise ValueError(f"plan math gives {want}, not {out_dtype}")
    out = torch.empty(lead + tuple(d.shape[:-1]) + (N,), dtype=want, device=d.device)
    f = None
    if flags is not None:
        if tuple(flags.shape) != tuple(d.shape):
            raise ValueError("flags must have the same shape as data")
        f = flags.contiguous()
        f = f.view(torch.uint8) if f.dtype == torch.bool else f.to(torch.uint8)
    w = (C.c_int32 * len(wins))(*wins)
    check(
        _lib.load().sds_delay_transform(
            plan.handle, ptr(d), complex_dtype_code(d), F0, ptr(f), F0, nrows, len(wins), w,
            float(delta_x), int(bool(inverse)), ptr(out),
            SDS_C128 if want == torch.complex128 else SDS_C64, stream_ptr()),
        "sds_delay_transform",
    )
    return out


def noise_sigma(coef, int_time, nsample):
    """ds.py:3544-3583.  coef (S,P,F) f64, int_time (B,T) f64, nsample 6-D -> sigma f64."""
    _lib.require_cuda()
    S, U, P, B, T, F = nsample.shape
    n = nsample.contiguous()
    out = torch.empty(n.shape, dtype=torch.float64, device=n.device)
    coef = coef.to(torch.float64).contiguous()
    it = int_time.to(torch.float64).contiguous()
    if tuple(coef.shape) != (S, P, F) or tuple(it.shape) != (B, T):
        raise ValueError("noise_sigma: coef must be (S,P,F) and int_time (B,T)")
    check(
        _lib.load().sds_noise_sigma(ptr(coef), ptr(it), ptr(n), real_dtype_code(n), S, U, P, B, T,
                                    F, ptr(out), stream_ptr()),
        "sds_noise_sigma",
    )
    return out


def generate_noise(sigma, seed=0, offset=0, normals=None, out_dtype=torch.complex128):
    """utils.py:486-506.  ``normals=(g_re, g_im)`` injects pre-drawn N(0,1) values."""
    _lib.require_cuda()
    s = sigma.to(torch.float64).contiguous()
    out = torch.empty(s.shape, dtype=out_dtype, device=s.device)
    g_re = g_im = None
    if normals is not None:
        g_re = torch.as_tensor(normals[0], dtype=torch.float64, device=s.device).contiguous()
        g_im = torch.as_tensor(normals[1], dtype=torch.float64, device=s.device).contiguous()
        if tuple(g_re.shape) != tuple(s.shape) or tuple(g_im.shape) != tuple(s.shape):
            raise ValueError("normals must have sigma's shape")
    check(
        _lib.load().sds_generate_noise(ptr(s), s.numel(), ptr(g_re), ptr(g_im), int(seed),
                                       int(offset), ptr(out), complex_dtype_code(out),
                                       stream_ptr()),
        "sds_generate_noise",
    )
    return out


def noise_transform(plan, coef, int_time, nsample, flags, delta_f, seed=0, offset=0, normals=None,
                    out_dtype=torch.complex128):
    """sigma -> realisation -> delay transform in one C call (sds_noise_transform).
    Returns (noise_freq, noise_delay), both with nsample's 6-D shape."""
    _lib.require_cuda()
    S, U, P, B, T, F = nsample.shape
    n = nsample.contiguous()
    coef = coef.to(torch.float64).contiguous()
    it = int_time.to(torch.float64).contiguous()
    fl = None if flags is None else flags.contiguous().view(torch.uint8)
    sig = torch.empty(n.shape, dtype=torch.float64, device=n.device)
    nf = torch.empty(n.shape, dtype=out_dtype, device=n.device)
    nd = torch.empty_like(nf)
    g_re = g_im = None
    if normals is not None:
        g_re = torch.as_tensor(normals[0], dtype=torch.float64, device=n.device).contiguous()
        g_im = torch.as_tensor(normals[1], dtype=torch.float64, device=n.device).contiguous()
    check(
        _lib.load().sds_noise_transform(plan.handle, ptr(coef), ptr(it), ptr(n), real_dtype_code(n), S, U, P,
                                        B, T, F, ptr(fl), float(delta_f), int(seed), int(offset), ptr(g_re),
                                        ptr(g_im), ptr(sig), ptr(nf), ptr(nd), complex_dtype_code(nf),
                                        stream_ptr()),
        "sds_noise_transform",
    )
    return nf, nd


def _sp_strides(a):
    """(S,P,B,T,D) view, possibly strided over S and P but contiguous in (B,T,D)."""
    S, P, B, T, D = a.shape
    st = a.stride()
    if (B * T * D) and (st[4] != 1 and D > 1 or st[3] != D and T > 1 or st[2] != T * D and B > 1):
        a = a.contiguous()
        st = a.stride()
    return a, int(st[0]), int(st[1])


def _scale_sp(scale_sp, S, P, dev):
    if scale_sp is None:
        return None
    t = torch.as_tensor(np.asarray(scale_sp, dtype=np.float64), device=dev).contiguous()
    if tuple(t.shape) != (S, P):
        raise ValueError(f"scale_sp must have shape ({S}, {P})")
    return t


def cross_power_full(a1, a2=None, scale_sp=None):
    """utils.py:387-389 with axis=2 on (S,P,B,T,D): out[s,p,i,j,t,d] = a2[i] conj(a1[j]), times
    scale_sp[s,p] when given (the cosmological conversion, applied as the tensor is written)."""
    _lib.require_cuda()
    if a2 is None:
        a2 = a1
    if tuple(a1.shape) != tuple(a2.shape):
        raise ValueError(
            "array_1 and array_2 must have the same shapes. "
            f"array_1 has shape {tuple(a1.shape)} but array_2 has shape {tuple(a2.shape)}"
        )
    if a2.dtype != a1.dtype:
        a2 = a2.to(a1.dtype)
    a1, ss1, sp1 = _sp_strides(a1)
    a2, ss2, sp2 = _sp_strides(a2)
    if (ss1, sp1) != (ss2, sp2):
        a1, a2 = a1.contiguous(), a2.contiguous()
        ss1, sp1 = a1.stride()[0], a1.stride()[1]
    S, P, B, T, D = a1.shape
    out = torch.empty((S, P, B, B, T, D), dtype=a1.dtype, device=a1.device)
    sc = _scale_sp(scale_sp, S, P, a1.device)
    check(
        _lib.load().sds_cross_power_full(ptr(a1), ptr(a2), complex_dtype_code(a1), S, P, B, T, D,
                                         ss1, sp1, ptr(sc), ptr(out), stream_ptr()),
        "sds_cross_power_full",
    )
    return out


def cross_power_sums(a1, a2=None, group_offset=None, method="factorised"):
    """Pair sums without the B x B tensor: returns (sum_all, sum_auto), each
    complex128 (G,S,P,T,D); see sds_cross_power_avg in include/sds.h."""
    _lib.require_cuda()
    if a2 is None:
        a2 = a1
    if tuple(a1.shape) != tuple(a2.shape):
        raise ValueError("array_1 and array_2 must have the same shapes.")
    if a2.dtype != a1.dtype:
        a2 = a2.to(a1.dtype)
    a1, ss1, sp1 = _sp_strides(a1)
    a2, ss2, sp2 = _sp_strides(a2)
    if (ss1, sp1) != (ss2, sp2):
        a1, a2 = a1.contiguous(), a2.contiguous()
        ss1, sp1 = a1.stride()[0], a1.stride()[1]
    S, P, B, T, D = a1.shape
    if group_offset is None:
        group_offset = [0, B]
    go = np.asarray(group_offset, dtype=np.int64)
    if go.ndim != 1 or go.size < 2 or go[0] != 0 or go[-1] != B or np.any(np.diff(go) < 0):
        raise ValueError("group_offset must rise from 0 to Nbls")
    G = go.size - 1
    god = torch.as_tensor(go.astype(np.int32), device=a1.device)
    out_all = torch.empty((G, S, P, T, D), dtype=torch.complex128, device=a1.device)
    out_auto = torch.empty_like(out_all)
    m = {"factorised": SDS_METHOD_FACTORISED, "explicit": SDS_METHOD_EXPLICIT}[method]
    check(
        _lib.load().sds_cross_power_avg(ptr(a1), ptr(a2), complex_dtype_code(a1), S, P, B, T, D,
                                        ss1, sp1, ptr(god), G, m, ptr(out_all), ptr(out_auto),
                                        stream_ptr()),
        "sds_cross_power_avg",
    )
    return out_all, out_auto


def thermal_power_full(n1, n2, coef, freq, int_time, pol_factor, scale_sp=None):
    """ds.py:3637-3707.  n1/n2 (S,P,B,T,F) real; coef (S,P,F); freq (S,F);
    int_time (B,T); pol_factor (P,) -> (S,P,B,B,T) f64, times scale_sp[s,p] when given."""
    _lib.require_cuda()
    if tuple(n1.shape) != tuple(n2.shape):
        raise ValueError(
            "nsample_1 and nsample_2 must have same shape, "
            f"but nsample_1 has shape {tuple(n1.shape)} and nsample_2 has shape {tuple(n2.shape)}"
        )
    if n2.dtype != n1.dtype:
        n2 = n2.to(n1.dtype)
    n1, ss1, sp1 = _sp_strides(n1)
    n2, ss2, sp2 = _sp_strides(n2)
    if (ss1, sp1) != (ss2, sp2):
        n1, n2 = n1.contiguous(), n2.contiguous()
        ss1, sp1 = n1.stride()[0], n1.stride()[1]
    S, P, B, T, F = n1.shape
    dev = n1.device
    coef = torch.as_tensor(coef, dtype=torch.float64, device=dev).contiguous()
    freq = torch.as_tensor(freq, dtype=torch.float64, device=dev).contiguous()
    it = torch.as_tensor(int_time, dtype=torch.float64, device=dev).contiguous()
    pf = torch.as_tensor(pol_factor, dtype=torch.float64, device=dev).contiguous()
    if tuple(coef.shape) != (S, P, F) or tuple(freq.shape) != (S, F) or tuple(it.shape) != (B, T) \
            or tuple(pf.shape) != (P,):
        raise ValueError("thermal_power_full: inconsistent shapes")
    out = torch.empty((S, P, B, B, T), dtype=torch.float64, device=dev)
    sc = _scale_sp(scale_sp, S, P, dev)
    check(
        _lib.load().sds_thermal_power_full(ptr(n1), ptr(n2), real_dtype_code(n1), ss1, sp1,
                                           ptr(coef), ptr(freq), ptr(it), ptr(pf), S, P, B, T, F,
                                           ptr(sc), ptr(out), stream_ptr()),
        "sds_thermal_power_full",
    )
    return out


def thermal_power_sums(n1, n2, coef, freq, int_time, pol_factor, group_offset=None):
    """Pair sums of the thermal estimate (sds_thermal_power_avg): returns
    (sum_all, sum_auto), each f64 (G,S,P,T)."""
    _lib.require_cuda()
    if tuple(n1.shape) != tuple(n2.shape):
        raise ValueError("nsample_1 and nsample_2 must have same shape")
    if n2.dtype != n1.dtype:
        n2 = n2.to(n1.dtype)
    n1, ss1, sp1 = _sp_strides(n1)
    n2, ss2, sp2 = _sp_strides(n2)
    if (ss1, sp1) != (ss2, sp2):
        n1, n2 = n1.contiguous(), n2.contiguous()
        ss1, sp1 = n1.stride()[0], n1.stride()[1]
    S, P, B, T, F = n1.shape
    dev = n1.device
    coef = torch.as_tensor(coef, dtype=torch.float64, device=dev).contiguous()
    freq = torch.as_tensor(freq, dtype=torch.float64, device=dev).contiguous()
    it = torch.as_tensor(int_time, dtype=torch.float64, device=dev).contiguous()
    pf = torch.as_tensor(pol_factor, dtype=torch.float64, device=dev).contiguous()
    if group_offset is None:
        group_offset = [0, B]
    go = np.asarray(group_offset, dtype=np.int64)
    G = go.size - 1
    god = torch.as_tensor(go.astype(np.int32), device=dev)
    out_all = torch.empty((G, S, P, T), dtype=torch.float64, device=dev)
    out_auto = torch.empty_like(out_all)
    check(
        _lib.load().sds_thermal_power_avg(ptr(n1), ptr(n2), real_dtype_code(n1), ss1, sp1, ptr(coef),
                                          ptr(freq), ptr(it), ptr(pf), ptr(god), G, S, P, B, T, F,
                                          ptr(out_all), ptr(out_auto), stream_ptr()),
        "sds_thermal_power_avg",
    )
    return out_all, out_auto


# ---- downstream estimators ------------------------------------------------
def _outer_n_inner(shape, axis):
    outer = int(np.prod(shape[:axis], dtype=np.int64))
    inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
    return outer, int(shape[axis]), inner


def weighted_average(x, sigma, weights, axis):
    """x, sigma: float64 CUDA tensors of one shape; weights: None, 1-D (n,) or the shape of x."""
    x, sigma = x.contiguous(), sigma.contiguous()
    outer, n, inner = _outer_n_inner(x.shape, axis)
    oshape = tuple(x.shape[:axis]) + tuple(x.shape[axis + 1:])
    out = torch.empty(oshape, dtype=torch.float64, device=x.device)
    osig = torch.empty_like(out)
    w1d = 0
    if weights is not None:
        weights = weights.contiguous()
        w1d = 1 if weights.dim() == 1 and x.dim() != 1 else 0
    rc = _lib.load().sds_weighted_average(_lib.ptr(x), _lib.ptr(sigma), _lib.ptr(weights), w1d, outer,
                                          n, inner, _lib.ptr(out), _lib.ptr(osig), _lib.stream_ptr())
    _lib.check(rc, "sds_weighted_average")
    return out, osig


def fold_along_delay(x, sigma, weights, axis, pos0, neg0, nout, zero_bin):
    """x, sigma, weights: CUDA tensors of one shape and dtype (float64 or complex128)."""
    x, sigma = x.contiguous(), sigma.contiguous()
    code = _lib.SDS_C128 if x.is_complex() else _lib.SDS_F64
    outer, n, inner = _outer_n_inner(x.shape, axis)
    oshape = tuple(x.shape[:axis]) + (nout,) + tuple(x.shape[axis + 1:])
    out = torch.empty(oshape, dtype=x.dtype, device=x.device)
    osig = torch.empty_like(out)
    if weights is not None:
        weights = weights.contiguous()
    rc = _lib.load().sds_fold_along_delay(_lib.ptr(x), _lib.ptr(sigma), _lib.ptr(weights), code, outer, n,
                                          inner, int(pos0), int(neg0), int(nout), int(bool(zero_bin)),
                                          _lib.ptr(out), _lib.ptr(osig), _lib.stream_ptr())
    _lib.check(rc, "sds_fold_along_delay")
    return out, osig


def bootstrap_indices(n, nboot, seed):
    idx = torch.empty((n, nboot), dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))
    rc = _lib.load().sds_bootstrap_indices(int(seed) & 0xFFFFFFFFFFFFFFFF, int(n), int(nboot), _lib.ptr(idx),
                                           _lib.stream_ptr())
    _lib.check(rc, "sds_bootstrap_indices")
    return idx


def bootstrap_gather(x, index, axis):
    """x: float64 / complex128 CUDA tensor; index: int64 (n, nboot) CUDA tensor."""
    x, index = x.contiguous(), index.contiguous()
    outer, n, inner = _outer_n_inner(x.shape, axis)
    nboot = int(index.shape[1])
    oshape = tuple(x.shape[:axis]) + (n, nboot) + tuple(x.shape[axis + 1:])
    out = torch.empty(oshape, dtype=x.dtype, device=x.device)
    rc = _lib.load().sds_bootstrap_gather(_lib.ptr(x), _lib.ptr(index), x.element_size(), outer, n, nboot,
                                          inner, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "sds_bootstrap_gather")
    return out


# ---- ingest -----------------------------------------------------------------
def ingest_gather(x, perm):
    """x: (Nblts, Nfreqs, Npols) CUDA tensor of 1-, 4-, 8- or 16-byte elements; perm: int64 (R,)
    CUDA tensor of baseline-time rows.  Returns (Npols, R, Nfreqs): out[p, r, f] = x[perm[r], f, p]."""
    x, perm = x.contiguous(), perm.contiguous()
    nblt, nfreq, npol = (int(v) for v in x.shape)
    out = torch.empty((npol, perm.numel(), nfreq), dtype=x.dtype, device=x.device)
    rc = _lib.load().sds_ingest_gather(ptr(x), ptr(perm), x.element_size(), nblt, perm.numel(), nfreq, npol,
                                       ptr(out), stream_ptr())
    check(rc, "sds_ingest_gather")
    return out
