"""Turn two data files bundled with the reference (simpleDS/data/*.uvh5, read with
mini_hdf5.py because h5py is not installed) into npz fixtures for the C1 configuration of
BASELINE.json ("bundled test visibility file through delay_transform +
calculate_delay_spectrum").  Run once in the build container:

    python tests/golden/make_uvh5_golden.py

Arrays are stored exactly as the files hold them (baseline-time axis time-major, baselines
unsorted within a time), so the ingest ordering rules of delay_spectrum.py:913-973 are part
of what the tests exercise.  paper_test_file.uvh5 is cut to its 12 lowest baseline numbers to
keep the fixture small; it is the file with real flags (26.6 %) and nsample == 0 samples."""
import os

import numpy as np

import mini_hdf5

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = "/root/reference/simpleDS/data"


def load(name, keep_bls=None):
    f = mini_hdf5.File(os.path.join(DATA, name))
    hd = lambda k: f.read("Header/" + k)  # noqa: E731
    a1, a2 = hd("ant_1_array").astype(np.int64), hd("ant_2_array").astype(np.int64)
    bl = 2048 * (a1 + 1) + (a2 + 1) + 2 ** 16
    keep = np.ones(bl.shape, bool) if keep_bls is None else np.isin(bl, np.unique(bl)[:keep_bls])
    vis = f.read("Data/visdata")[:, 0][keep]        # (Nblts, Nfreqs, Npols)
    return dict(
        data_array=vis, flag_array=f.read("Data/flags")[:, 0][keep].astype(bool),
        nsample_array=f.read("Data/nsamples")[:, 0][keep],
        ant_1_array=a1[keep], ant_2_array=a2[keep], lst_array=hd("lst_array")[keep],
        time_array=hd("time_array")[keep], integration_time=hd("integration_time")[keep],
        uvw_array=hd("uvw_array")[keep], freq_array=hd("freq_array")[0],
        polarization_array=hd("polarization_array"), channel_width=np.float64(hd("channel_width")),
        vis_units=np.array(bytes(hd("vis_units").item()).decode()), Nants_telescope=np.int64(hd("Nants_telescope")),
    )


if __name__ == "__main__":
    for name, out, keep in (("test_redundant_array_odd_file.uvh5", "uvh5_odd.npz", None),
                            ("paper_test_file.uvh5", "uvh5_paper12.npz", 12)):
        d = load(name, keep)
        np.savez_compressed(os.path.join(HERE, out), **d)
        print(out, {k: (v.shape, str(v.dtype)) for k, v in d.items() if v.ndim}, os.path.getsize(os.path.join(HERE, out)))
