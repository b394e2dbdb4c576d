"""The HDF5 reader the golden-fixture scripts use: since round 2 it lives in the product
(simpleds_b200/hdf5lite.py, the file side of the DelaySpectrum path); this module re-exports it so
that make_uvh5_golden.py keeps working."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from simpleds_b200.hdf5lite import File, lzf_decompress  # noqa: E402,F401
