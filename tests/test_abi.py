"""CPU-side checks of the boundary: libsds.so loads, exports every symbol that
include/sds.h declares, and the host mirror raises the reference's errors before
touching the device.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sds.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sds_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from simpleds_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), f"libsds.so does not export {name}"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert _lib.load().sds_version() == 100


def test_engine_desc_layout_matches_header():
    from simpleds_b200 import _lib

    # 8 int32 + 64 int32 + 4 int64 + uint64 + double
    assert ctypes.sizeof(_lib.EngineDesc) == 8 * 4 + 64 * 4 + 4 * 8 + 8 + 8 + 2 * 8


def test_plan_argument_errors_surface_as_valueerror():
    from simpleds_b200 import _lib

    lib = _lib.load()
    h = ctypes.c_void_p(0)
    t = (ctypes.c_double * 4)(1, 1, 1, 1)
    rc = lib.sds_plan_create(ctypes.byref(h), 0, _lib.SDS_F64, t)
    assert rc == -1 and b"nfreq" in lib.sds_last_error()
    rc = lib.sds_plan_create(ctypes.byref(h), 4, 7, t)
    assert rc == -1 and b"math" in lib.sds_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "sds_plan_create")


def test_no_cpu_fallback_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from simpleds_b200 import utils
    from simpleds_b200.units import Quantity

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        utils.normalized_fourier_transform(np.zeros((2, 8), complex), Quantity(1.0, "Hz"))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "simpleds_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"


def test_host_errors_match_reference():
    from simpleds_b200 import DelaySpectrum, utils
    from simpleds_b200.units import Quantity, Unit

    ds = DelaySpectrum()
    with pytest.raises(ValueError, match="Input spectral taper must be a function"):
        ds.set_taper("hann")  # ds.py:683-689
    with pytest.raises(ValueError, match="No data has be loaded"):
        ds.calculate_delay_spectrum()  # ds.py:3612-3616, t_ds.py:742-745
    with pytest.raises(ValueError, match="delta_x must be an astropy Quantity"):
        utils.normalized_fourier_transform(np.zeros((1, 13, 21)), delta_x=2.0, axis=2)  # t_utils.py:437-443
    with pytest.raises(ValueError, match="must be 'complex128' or 'complex64'"):
        DelaySpectrum(precision="float16")
    with pytest.raises(ValueError, match="Please Load data before attaching a UVBeam"):
        DelaySpectrum(uvb=object())  # ds.py:583-585
    assert Unit("Jy") * Unit("Hz") == Unit("Jy Hz")
    assert Unit("K str") == Unit("K sr")
    assert (Unit("Jy Hz") ** 2) == Unit("Jy^2 Hz^2")
    q = Quantity(np.ones((2, 3)), "Jy")
    assert q.shape == (2, 3) and q[0].unit == Unit("Jy")
    assert np.isclose(float(utils.jy_to_mk(150e6).value), 1.446590, rtol=1e-6)
    assert 1 == 1.0 / utils.noise_equivalent_bandwidth(np.ones(2000))  # t_utils.py:185-188
