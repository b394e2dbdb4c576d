"""ctypes binding of libsds.so (see include/sds.h).

There is no CPU fallback: if the shared library is missing or a call fails the
error is raised.  Status codes < 0 (bad argument / unsupported) become
``ValueError`` -- the reference raises ``ValueError`` for every argument
problem on this path -- and CUDA errors become ``RuntimeError``.
"""
import ctypes as C
import os

import numpy as np
import torch

SDS_C64, SDS_C128, SDS_F32, SDS_F64 = 0, 1, 2, 3
SDS_PAIRS_ALL, SDS_PAIRS_NO_AUTOS = 0, 1
SDS_METHOD_FACTORISED, SDS_METHOD_EXPLICIT = 0, 1
SDS_MAX_WINDOWS = 64
SDS_MAX_NFREQ = 4096

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsds.so")

_vp, _i32, _i64, _u64, _dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double


class EngineDesc(C.Structure):
    _fields_ = [
        ("npol", C.c_int32), ("nbl", C.c_int32), ("ntime", C.c_int32), ("nchan", C.c_int32),
        ("nspw", C.c_int32), ("ngroup", C.c_int32), ("nuv", C.c_int32), ("noise_mode", C.c_int32),
        ("win_start", C.c_int32 * SDS_MAX_WINDOWS),
        ("time_offset", C.c_int64), ("ntime_total", C.c_int64),
        ("bl_offset", C.c_int64), ("nbl_total", C.c_int64),
        ("seed", C.c_uint64), ("delta_f", C.c_double),
        ("group_index", C.c_void_p), ("group_bl_start", C.c_void_p),
    ]


SDS_COSMO_MAX_NU = 4


class CosmoParams(C.Structure):
    _fields_ = [
        ("h0_km_s_mpc", C.c_double), ("om0", C.c_double), ("ode0", C.c_double), ("ogamma0", C.c_double),
        ("neff_per_nu", C.c_double), ("nmassless_nu", C.c_double), ("nu_y", C.c_double * SDS_COSMO_MAX_NU),
        ("nmassive_nu", C.c_int32), ("has_massive_nu_or_radiation", C.c_int32),
    ]


_SIGNATURES = {
    "sds_last_error": (C.c_char_p, []),
    "sds_version": (_i32, []),
    "sds_plan_create": (_i32, [C.POINTER(_vp), _i32, _i32, C.POINTER(_dbl)]),
    "sds_plan_destroy": (_i32, [_vp]),
    "sds_plan_nfreq": (_i32, [_vp]),
    "sds_plan_math": (_i32, [_vp]),
    "sds_take_windows": (_i32, [_vp, _vp, _i32, _i64, _i32, _i32, _vp, _i32, _i32, _vp]),
    "sds_delay_transform": (_i32, [_vp, _vp, _i32, _i64, _vp, _i64, _i64, _i32,
                                   C.POINTER(C.c_int32), _dbl, _i32, _vp, _i32, _vp]),
    "sds_noise_sigma": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "sds_generate_noise": (_i32, [_vp, _i64, _vp, _vp, _u64, _u64, _vp, _i32, _vp]),
    "sds_noise_transform": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _dbl,
                                   _u64, _u64, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "sds_cross_power_full": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _vp, _vp, _vp]),
    "sds_thermal_power_full": (_i32, [_vp, _vp, _i32, _i64, _i64, _vp, _vp, _vp, _vp,
                                      _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "sds_thermal_power_avg": (_i32, [_vp, _vp, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i32,
                                     _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "sds_cross_power_avg": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i64, _i64,
                                   _vp, _i32, _i32, _vp, _vp, _vp]),
    "sds_engine_workspace_bytes": (_i64, [_vp, C.POINTER(EngineDesc)]),
    "sds_engine_run": (_i32, [_vp, C.POINTER(EngineDesc)] + [_vp] * 17 + [_vp, C.POINTER(C.c_int32)]),
    "sds_engine_means": (_i32, [_i32, _i32, _i32, _vp, _vp, _dbl] + [_vp] * 12 + [_vp]),
    "sds_weighted_average": (_i32, [_vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp]),
    "sds_fold_along_delay": (_i32, [_vp, _vp, _vp, _i32, _i64, _i64, _i64, _i64, _i64, _i64, _i32,
                                    _vp, _vp, _vp]),
    "sds_bootstrap_gather": (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _i64, _vp, _vp]),
    "sds_bootstrap_indices": (_i32, [_u64, _i64, _i64, _vp, _vp]),
    "sds_ingest_gather": (_i32, [_vp, _vp, _i32, _i64, _i64, _i32, _i32, _vp, _vp]),
    "sds_cosmo_evaluate": (_i32, [C.POINTER(CosmoParams), _vp, _i64, _vp, _vp, _vp]),
    "sds_conversion_integral": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load libsds.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SDS_LIB", LIB_PATH)  # SDS_LIB: a development build of the same library
    if not os.path.exists(path):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m simpleds_b200.build` "
            "(nvcc, sm_100a).  simpleds_b200 has no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc, what):
    if rc == 0:
        return
    msg = load().sds_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg}")


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "simpleds_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback."
        )


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a torch CUDA tensor (or None)."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise ValueError("expected a CUDA tensor")
    return C.c_void_p(t.data_ptr())


def complex_dtype_code(t):
    if t.dtype == torch.complex64:
        return SDS_C64
    if t.dtype == torch.complex128:
        return SDS_C128
    raise ValueError(f"expected a complex64/complex128 tensor, got {t.dtype}")


def real_dtype_code(t):
    if t.dtype == torch.float32:
        return SDS_F32
    if t.dtype == torch.float64:
        return SDS_F64
    raise ValueError(f"expected a float32/float64 tensor, got {t.dtype}")


class Plan:
    """Owns an ``sds_plan*`` (twiddles + taper on the device) for one length/math."""

    def __init__(self, nfreq, math, taper_values):
        require_cuda()
        lib = load()
        taper_values = np.ascontiguousarray(np.asarray(taper_values, dtype=np.float64))
        if taper_values.shape != (nfreq,):
            raise ValueError(
                f"taper({nfreq}) must return {nfreq} values, got shape {taper_values.shape}"
            )
        self.nfreq, self.math = int(nfreq), int(math)
        self._h = C.c_void_p(0)
        rc = lib.sds_plan_create(
            C.byref(self._h), self.nfreq, self.math,
            taper_values.ctypes.data_as(C.POINTER(C.c_double)),
        )
        check(rc, "sds_plan_create")

    @property
    def handle(self):
        return self._h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.sds_plan_destroy(h)
            self._h = C.c_void_p(0)
