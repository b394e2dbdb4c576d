"""The package's HDF5 reader / writer (simpleds_b200/hdf5lite.py): round trips of every datatype and
structure the DelaySpectrum file format uses (delay_spectrum.py:2545-2743), hyperslab writes in place
(write_partial), and -- in the build container, where the reference's data directory exists -- the
h5py-written UVH5 files bundled with the reference against the committed fixtures."""
import os

import numpy as np
import pytest

from simpleds_b200 import hdf5lite

REF_DATA = "/root/reference/simpleDS/data"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_round_trip_of_every_datatype_and_attribute(tmp_path):
    rng = np.random.default_rng(0)
    path = str(tmp_path / "a.h5")
    data = {
        "Header/Ntimes": np.int64(21), "Header/channel_width": np.float64(492610.84375),
        "Header/lst_array": rng.random(21), "Header/ant_1_array": rng.integers(0, 64, 51),
        "Header/polarization_array": np.array([-5, -6], dtype=np.int64),
        "Header/freq_array": 1e8 + 1e5 * np.arange(2 * 16, dtype=np.float64).reshape(2, 16),
        "Data/visdata": (rng.standard_normal((2, 1, 2, 3, 4, 16)) + 1j * rng.standard_normal((2, 1, 2, 3, 4, 16))),
        "Data/visdata64": (rng.standard_normal((3, 5)) + 1j * rng.standard_normal((3, 5))).astype(np.complex64),
        "Data/flags": rng.random((2, 1, 2, 3, 4, 16)) < 0.3,
        "Data/nsamples": rng.random((2, 1, 2, 3, 4, 16)).astype(np.float32),
        "Data/i32": rng.integers(-5, 5, (7,)).astype(np.int32), "Data/u8": rng.integers(0, 255, (9,)).astype(np.uint8),
    }
    w = hdf5lite.Writer(path)
    for k, v in data.items():
        w.create_dataset(k, v, attrs={"unit": "Jy Hz"} if k.startswith("Data/vis") else None)
    w.create_dataset("Header/data_type", b"frequency")
    w.create_dataset("Header/taper", b"blackmanharris", attrs={"module": "scipy.signal.windows", "n": np.int64(3)})
    w.create_group("Empty")
    w.close()
    f = hdf5lite.File(path)
    assert f.keys("/") == ["Data", "Empty", "Header"] and f.keys("Empty") == []
    assert set(f.keys("Data")) == {k.split("/")[1] for k in data if k.startswith("Data/")}
    for k, v in data.items():
        got = f.read(k)
        assert got.dtype == np.asarray(v).dtype and got.shape == np.asarray(v).shape, k
        assert np.array_equal(got, v), k
    assert bytes(f.read("Header/data_type").item()) == b"frequency"
    assert f.attrs("Data/visdata") == {"unit": "Jy Hz"}
    a = f.attrs("Header/taper")
    assert a["module"] == "scipy.signal.windows" and int(a["n"]) == 3
    assert "Header/Ntimes" in f and "Header/nope" not in f


def test_hyperslabs_are_written_in_place(tmp_path):
    """initialize_save_file / write_partial: datasets laid out at full size, filled piece by piece."""
    path = str(tmp_path / "p.h5")
    shape = (2, 1, 2, 5, 6, 8)
    rng = np.random.default_rng(1)
    full = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    flags = rng.random(shape) < 0.5
    w = hdf5lite.Writer(path)
    w.create_dataset("Data/visdata", shape=shape, dtype=np.complex128, attrs={"unit": "Jy"})
    w.create_dataset("Data/flags", shape=shape, dtype=np.bool_)
    w.create_dataset("Header/Nbls", np.int64(5))
    w.close()
    assert os.path.getsize(path) >= full.nbytes + flags.size
    assert not hdf5lite.File(path).read("Data/visdata").any()
    for s in range(2):  # by spectral window, polarisation, baseline run and time run
        for p in range(2):
            for b0, b1 in ((0, 2), (2, 5)):
                for t0, t1 in ((0, 4), (4, 6)):
                    idx = (s, 0, p, slice(b0, b1), slice(t0, t1))
                    hdf5lite.write_hyperslab(path, "Data/visdata", idx, full[idx])
                    hdf5lite.write_hyperslab(path, "Data/flags", idx, flags[idx])
    f = hdf5lite.File(path)
    assert np.array_equal(f.read("Data/visdata"), full) and np.array_equal(f.read("Data/flags"), flags)
    hdf5lite.write_hyperslab(path, "Data/visdata", (1,), np.zeros(shape[1:]))  # a whole leading index
    assert not hdf5lite.File(path).read("Data/visdata")[1].any()
    assert int(hdf5lite.File(path).read("Header/Nbls")) == 5


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="the reference's data directory only exists in the build container")
def test_reader_on_the_h5py_written_files_bundled_with_the_reference():
    """Chunked + LZF-filtered data sets, compounds and enums as h5py writes them; the fixtures under
    tests/golden were cut from these files."""
    f = hdf5lite.File(os.path.join(REF_DATA, "test_redundant_array_odd_file.uvh5"))
    g = np.load(os.path.join(GOLDEN, "uvh5_odd.npz"))
    assert np.array_equal(f.read("Data/visdata")[:, 0], g["data_array"])
    assert np.array_equal(f.read("Data/flags")[:, 0], g["flag_array"])
    assert np.array_equal(f.read("Data/nsamples")[:, 0], g["nsample_array"])
    assert np.array_equal(f.read("Header/lst_array"), g["lst_array"])
    assert float(f.read("Header/channel_width")) == 492610.84375
    f = hdf5lite.File(os.path.join(REF_DATA, "paper_test_file.uvh5"))  # float32 pairs, 26.6 % flags, LZF chunks
    assert f.read("Data/visdata").dtype == np.complex64 and f.shape("Data/flags") == (1071, 1, 203, 1)
    assert abs(f.read("Data/flags").mean() - 0.266) < 1e-3


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="the reference's data directory only exists in the build container")
def test_uvdatalite_from_the_bundled_uvh5_file():
    """read_uvh5 on the h5py-written original (no GPU needed to read a file)."""
    from simpleds_b200.delay_spectrum import UVDataLite

    uv = UVDataLite.from_uvh5(os.path.join(REF_DATA, "test_redundant_array_odd_file.uvh5"))
    g = np.load(os.path.join(GOLDEN, "uvh5_odd.npz"))
    assert np.array_equal(uv.data_array, g["data_array"]) and np.array_equal(uv.flag_array, g["flag_array"])
    assert uv.vis_units == "Jy" and uv.Nbls == 51 and uv.Ntimes == 21 and uv.Nants_telescope == int(g["Nants_telescope"])
    assert uv.get_baseline_nums()[0] == 67611  # t_ds.py:271-289
    k = UVDataLite.from_uvh5(os.path.join(REF_DATA, "paper_test_file_k_units.uvh5"))
    assert k.vis_units == "K str" and k.data_array.shape == (1071, 203, 1)
