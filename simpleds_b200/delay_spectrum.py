"""Drop-in ``DelaySpectrum`` for the hot path of simpleDS/delay_spectrum.py.

Same method names, arguments, attribute names/shapes, in-place behaviour,
errors and warnings as the reference for ``select_spectral_windows``,
``set_taper``, ``generate_noise``, ``delay_transform`` and
``calculate_delay_spectrum``; the arrays live in HBM as
torch tensors and every transform / product / estimate is a libsds.so kernel.
astropy and pyuvdata do not exist here, so units are tags
(:mod:`simpleds_b200.units`) and ingest takes a :class:`UVDataLite` or plain
arrays.  Citations "ds.py:N" are to /root/reference/simpleDS/delay_spectrum.py.
"""
import copy
import warnings

import numpy as np
import torch
from scipy.signal import windows

from . import ops, utils
from .io import DelaySpectrumIO, read_uvh5
from .units import Quantity, Unit, UnitConversionError, is_quantity
from .utils import jy_to_mk


def _value(x):
    return x.value if is_quantity(x) else x


def _default_cosmology():
    from . import cosmo as simple_cosmo
    return simple_cosmo.default_cosmology.get()


class UVDataLite:
    """The slice of a pyuvdata UVData object ``add_uvdata`` reads (ds.py:903-932).

    Baseline-time axis is time-major as in pyuvdata.  Arrays:
    data_array/flag_array/nsample_array (Nblts, Nfreqs, Npols); ant_1_array,
    ant_2_array, lst_array [rad], integration_time [s] (Nblts,); freq_array
    (Nfreqs,) [Hz]; polarization_array (Npols,); uvw_array (Nblts, 3) [m]."""

    def __init__(self, data_array, flag_array, nsample_array, ant_1_array, ant_2_array,
                 lst_array, integration_time, freq_array, polarization_array, vis_units="Jy",
                 uvw_array=None, Nants_telescope=None, x_orientation=None):
        self.data_array = np.asarray(data_array)
        self.flag_array = np.asarray(flag_array, dtype=bool)
        self.nsample_array = np.asarray(nsample_array)
        self.ant_1_array = np.asarray(ant_1_array, dtype=np.int64)
        self.ant_2_array = np.asarray(ant_2_array, dtype=np.int64)
        self.lst_array = np.asarray(lst_array, dtype=np.float64)
        self.integration_time = np.asarray(integration_time, dtype=np.float64)
        self.freq_array = np.asarray(freq_array, dtype=np.float64).reshape(-1)
        self.polarization_array = np.asarray(polarization_array, dtype=np.int64)
        self.vis_units = vis_units
        self.uvw_array = None if uvw_array is None else np.asarray(uvw_array, dtype=np.float64)
        self.x_orientation = x_orientation
        # pyuvdata's antnums_to_baseline for < 2048 antennas (ds.py:269-272)
        self.baseline_array = 2048 * (self.ant_1_array + 1) + (self.ant_2_array + 1) + 2**16
        self.Nblts, self.Nfreqs, self.Npols = self.data_array.shape
        self.Nbls = len(np.unique(self.baseline_array))
        self.Ntimes = self.Nblts // max(self.Nbls, 1)
        ants = np.union1d(self.ant_1_array, self.ant_2_array)
        self.Nants_data = len(ants)
        self.Nants_telescope = int(Nants_telescope) if Nants_telescope else self.Nants_data

    @classmethod
    def from_uvh5(cls, path):
        """Read a UVH5 visibility file (what the reference does through pyuvdata's ``UVData.read``)
        with the package's own HDF5 reader; the result feeds ``DelaySpectrum.add_uvdata``."""
        return read_uvh5(path, cls)

    def get_baseline_nums(self):
        return np.unique(self.baseline_array)

    def get_redundancies(self, tol=1.0, include_conjugates=True):
        """Groups of baselines whose (conjugation-normalised) uvw agree within tol."""
        bls = self.get_baseline_nums()
        if self.uvw_array is None:
            return [list(bls)], [np.zeros(3)], [0.0], []
        vecs = np.array([self.uvw_array[self.baseline_array == b].mean(0) for b in bls])
        if include_conjugates:
            flip = (vecs[:, 0] < 0) | ((vecs[:, 0] == 0) & (vecs[:, 1] < 0))
            vecs = np.where(flip[:, None], -vecs, vecs)
        groups, centers = [], []
        for b, v in zip(bls, vecs):
            for gi, c in enumerate(centers):
                if np.linalg.norm(v - c) < tol:
                    groups[gi].append(b)
                    break
            else:
                groups.append([b])
                centers.append(v)
        return groups, centers, [float(np.linalg.norm(c)) for c in centers], []


class DelaySpectrum(DelaySpectrumIO):
    """See module docstring.  ``precision`` = "complex128" (fp64 math, the
    reference's dtype) or "complex64" (fp32 math)."""

    _visdata_params = ["data_array", "nsample_array", "flag_array", "noise_array"]
    _power_params = ["power_array", "noise_power"]
    _thermal_params = ["thermal_power"]

    def __init__(self, uv=None, uvb=None, trcvr=None, taper=None, precision="complex128",
                 noise_seed=0):
        if precision not in ("complex128", "complex64"):
            raise ValueError("precision must be 'complex128' or 'complex64'")
        self.precision = precision
        self.noise_seed = int(noise_seed)
        self._noise_draws = 0
        self.Ntimes = self.Nbls = self.Nfreqs = self.Npols = self.Ndelays = None
        self.Nants_data = self.Nants_telescope = self.Nspws = None
        self.Nuv = 0  # ds.py:128-135
        self.data_type = "frequency"  # ds.py:149-157
        self.vis_units = None
        self.x_orientation = None
        self._data = self._noise = self._nsample = self._flag = None
        self._data_unit = Unit("")
        self.lst_array = self.ant_1_array = self.ant_2_array = self.baseline_array = None
        self.freq_array = self.delay_array = self.polarization_array = self.uvw = None
        self.trcvr = self.beam_area = self.beam_sq_area = self.integration_time = None
        self.power_array = self.noise_power = self.thermal_power = None
        self.redshift = self.k_parallel = self.k_perpendicular = None
        self.unit_conversion = self.thermal_conversion = None
        self.cosmology = _default_cosmology()  # ds.py:551-558
        self.taper = windows.blackmanharris  # ds.py:541-549

        if uv is not None:
            if not isinstance(uv, (list, np.ndarray, tuple)):
                uv = [uv]
            for _uv in uv:
                self.add_uvdata(_uv)
        if uvb is not None:
            if self.Nuv == 0:
                raise ValueError("Please Load data before attaching a UVBeam.")
            self.add_uvbeam(uvb)
        if trcvr is not None:
            self.add_trcvr(trcvr=trcvr)
        if taper is not None:
            self.set_taper(taper=taper)

    # ------------------------------------------------------------------ storage
    @property
    def _cdtype(self):
        return torch.complex128 if self.precision == "complex128" else torch.complex64

    @property
    def _math(self):
        return ops.math_of(self._cdtype)

    @staticmethod
    def _device():
        ops._lib.require_cuda()
        return torch.device("cuda", torch.cuda.current_device())

    def _q(self, t, unit):
        return None if t is None else Quantity(t, unit)

    @property
    def data_array(self):
        return self._q(self._data, self._data_unit)

    @data_array.setter
    def data_array(self, v):
        if is_quantity(v):
            self._data_unit = v.unit
        self._data = None if v is None else self._to_dev(_value(v), self._cdtype)

    @property
    def noise_array(self):
        return self._q(self._noise, self._data_unit)

    @noise_array.setter
    def noise_array(self, v):
        self._noise = None if v is None else self._to_dev(_value(v), self._cdtype)

    @property
    def nsample_array(self):
        return self._nsample

    @nsample_array.setter
    def nsample_array(self, v):
        self._nsample = None if v is None else self._to_dev(_value(v), torch.float64)

    @property
    def flag_array(self):
        return self._flag

    @flag_array.setter
    def flag_array(self, v):
        self._flag = None if v is None else self._to_dev(_value(v), torch.bool)

    def _to_dev(self, x, dtype):
        if isinstance(x, torch.Tensor):
            return x.to(device=self._device(), dtype=dtype)
        return torch.as_tensor(np.ascontiguousarray(x), device=self._device()).to(dtype)

    def copy(self):
        new = copy.copy(self)
        for k, v in self.__dict__.items():
            if isinstance(v, torch.Tensor):
                new.__dict__[k] = v.clone()
            elif isinstance(v, np.ndarray):
                new.__dict__[k] = v.copy()
            elif is_quantity(v):
                vv = v.value
                new.__dict__[k] = Quantity(vv.clone() if isinstance(vv, torch.Tensor) else np.array(vv), v.unit)
        return new

    __deepcopy__ = lambda self, memo: self.copy()  # noqa: E731

    # ------------------------------------------------------------------ small API
    def set_taper(self, taper=None):
        """ds.py:662-691."""
        if taper is None:
            taper = windows.blackmanharris
        if not callable(taper):
            raise ValueError(
                "Input spectral taper must be a function or "
                "callable whose arguments are "
                "the length of the band "
                "over which Fourier Transform is taken."
            )
        self.taper = taper

    def set_delay(self):
        """ds.py:693-714."""
        ok = [Unit("Jy Hz"), Unit("K sr Hz"), Unit("Hz")]
        if not self._data_unit.is_equivalent(ok):
            raise UnitConversionError(
                "Data is not in units consistent "
                "with the delay domain. "
                "Cannot set data_type to delay."
            )
        self.data_type = "delay"

    def set_frequency(self):
        """ds.py:716-733."""
        ok = [Unit("Jy"), Unit("K sr"), Unit("")]
        if not self._data_unit.is_equivalent(ok):
            raise UnitConversionError(
                "Data is not in units consistent "
                "with the frequency domain. "
                "Cannot set data_type to frequency."
            )
        self.data_type = "frequency"

    def add_trcvr(self, trcvr):
        """ds.py:3459-3485.  ``trcvr`` in K: scalar, (Nspws, Nfreqs) array or Quantity."""
        if is_quantity(trcvr):
            if trcvr.unit != Unit("K"):
                raise UnitConversionError("trcvr must be in K")
            trcvr = trcvr.value
        trcvr = np.asarray(trcvr, dtype=np.float64)
        if self.Nspws is None or self.Nfreqs is None:
            # the reference fails inside UVParameter.expected_shape (t_ds.py:736-739)
            raise ValueError("Missing shape parameter Nspws needed to calculate expected shape of parameter")
        if trcvr.size == 1:
            val = np.ones((self.Nspws, self.Nfreqs)) * float(trcvr.reshape(-1)[0])
        elif trcvr.ndim != 2 or trcvr.shape[0] != self.Nspws or trcvr.shape[1] != self.Nfreqs:
            raise ValueError(
                "If input receiver temperature is not a scalar "
                "Quantity, must shape (Nspws, Nfreqs). "
                f"Expected shape was {(self.Nspws, self.Nfreqs)}, but input shape "
                f"was {trcvr.shape}"
            )
        else:
            val = trcvr.copy()
        self.trcvr = Quantity(val, "K")

    def set_beam_area(self, beam_area, beam_sq_area=None):
        """Direct form of add_uvbeam's result (ds.py:3345-3457): (Nspws,Npols,Nfreqs) in sr,
        or a scalar."""
        shape = (self.Nspws, self.Npols, self.Nfreqs)
        ba = np.broadcast_to(np.asarray(_value(beam_area), dtype=np.float64), shape).copy()
        bsq = ba if beam_sq_area is None else np.broadcast_to(
            np.asarray(_value(beam_sq_area), dtype=np.float64), shape).copy()
        self.beam_area, self.beam_sq_area = Quantity(ba, "sr"), Quantity(bsq, "sr")

    def add_uvbeam(self, uvb, no_read_trcvr=False, use_exact=False, bounds_error=False,
                   fill_value="extrapolate", kind="cubic"):
        """ds.py:3345-3457 for a duck-typed beam exposing ``freq_array`` [Hz],
        ``get_beam_area(pol)``, ``get_beam_sq_area(pol)`` and optionally
        ``receiver_temperature_array``; interpolated onto freq_array."""
        from scipy.interpolate import interp1d

        bf = np.asarray(uvb.freq_array, dtype=np.float64).squeeze()
        ba = np.zeros((self.Nspws, self.Npols, self.Nfreqs))
        bsq = np.zeros_like(ba)
        kw = dict(bounds_error=bounds_error, fill_value=fill_value, kind=kind)
        freqs = self.freq_array.value
        for ip, pol in enumerate(self.polarization_array):
            fa = interp1d(bf, np.asarray(uvb.get_beam_area(pol=pol)), **kw)
            fs = interp1d(bf, np.asarray(uvb.get_beam_sq_area(pol=pol)), **kw)
            for s in range(self.Nspws):
                ba[s, ip], bsq[s, ip] = fa(freqs[s]), fs(freqs[s])
        self.beam_area, self.beam_sq_area = Quantity(ba, "sr"), Quantity(bsq, "sr")
        rt = getattr(uvb, "receiver_temperature_array", None)
        if rt is not None and not no_read_trcvr:
            ft = interp1d(bf, np.asarray(rt)[0], **kw)
            self.trcvr = Quantity(np.stack([ft(freqs[s]) for s in range(self.Nspws)]), "K")

    # ------------------------------------------------------------------ ingest
    @classmethod
    def from_arrays(cls, data_array, flag_array, nsample_array, freq_array, integration_time,
                    polarization_array, vis_units="Jy", lst_array=None, baseline_array=None,
                    ant_1_array=None, ant_2_array=None, trcvr=None, beam_area=None, taper=None,
                    precision="complex128", noise_seed=0, uvw=None):
        """Build from arrays already in the reference's 6-D layout
        (Nspws, Nuv, Npols, Nbls, Ntimes, Nfreqs) (ds.py:159-176); ``freq_array`` is
        (Nspws, Nfreqs) in Hz, ``integration_time`` (Nbls, Ntimes) in s."""
        self = cls(precision=precision, noise_seed=noise_seed)
        if uvw is not None:  # baseline vector of the redundant group, metres (ds.py:984)
            self.uvw = Quantity(np.asarray(_value(uvw), dtype=np.float64).reshape(3), "m")
        shape = tuple(data_array.shape)
        if len(shape) != 6:
            raise ValueError("data_array must be (Nspws, Nuv, Npols, Nbls, Ntimes, Nfreqs)")
        self.Nspws, self.Nuv, self.Npols, self.Nbls, self.Ntimes, self.Nfreqs = shape
        if self.Nuv not in (1, 2):
            raise ValueError("Nuv must be 1 or 2")
        self.Ndelays = self.Nfreqs
        self.vis_units = vis_units
        self._data_unit = {"Jy": Unit("Jy"), "K str": Unit("K sr")}.get(vis_units, Unit(""))
        self.data_array = data_array
        self.flag_array = flag_array
        self.nsample_array = nsample_array
        self._noise = torch.zeros_like(self._data)
        self.freq_array = Quantity(np.asarray(_value(freq_array), dtype=np.float64).reshape(self.Nspws, self.Nfreqs), "Hz")
        self.integration_time = Quantity(np.broadcast_to(np.asarray(_value(integration_time), dtype=np.float64), (self.Nbls, self.Ntimes)).copy(), "s")
        self.polarization_array = np.asarray(polarization_array, dtype=np.int64)
        self.lst_array = Quantity(np.arange(self.Ntimes, dtype=np.float64) * 1e-3 if lst_array is None else np.asarray(lst_array, dtype=np.float64), "rad")
        self.baseline_array = np.arange(self.Nbls) if baseline_array is None else np.asarray(baseline_array)
        self.ant_1_array = None if ant_1_array is None else np.asarray(ant_1_array)
        self.ant_2_array = None if ant_2_array is None else np.asarray(ant_2_array)
        inf = np.inf
        self.trcvr = Quantity(np.full((self.Nspws, self.Nfreqs), inf), "K")
        self.beam_area = Quantity(np.full((self.Nspws, self.Npols, self.Nfreqs), inf), "sr")
        self.beam_sq_area = Quantity(np.full((self.Nspws, self.Npols, self.Nfreqs), inf), "sr")
        self._set_delay_array()
        if trcvr is not None:
            self.add_trcvr(trcvr)
        if beam_area is not None:
            self.set_beam_area(beam_area)
        if taper is not None:
            self.set_taper(taper)
        return self

    def add_uvdata(self, uv, spectral_windows=None, tol=1.0):
        """ds.py:861-1131 for a :class:`UVDataLite` (ordering rules of ds.py:913-973:
        baselines sorted by baseline number, times by unwrapped LST)."""
        if not isinstance(uv, UVDataLite):
            raise ValueError("Input data object must be an instance or subclass of UVData.")
        red_groups, uvw_centers = uv.get_redundancies(tol=tol, include_conjugates=True)[:2]
        if len(red_groups) > 1:
            raise ValueError(
                "A DelaySpectrum object can only perform a Fourier "
                "Transform along a single baseline vector. "
                "Downselect the input UVData object to only have "
                "one redundant baseline type."
            )
        if self.Nuv == 2:
            raise ValueError(
                "This DelaySpectrum Object has already "
                "been loaded with two datasets. Create "
                "a new object to cross-multipy different "
                "data."
            )
        bls = uv.get_baseline_nums()  # sorted unique  (ds.py:931)
        P, B, T, F = uv.Npols, uv.Nbls, uv.Ntimes, uv.Nfreqs
        # utils.py:15-168 (get_data_array / get_nsample_array / get_flag_array /
        # get_integration_time) + the LST sort of ds.py:913-916 as ONE row permutation:
        # output row (baseline c, time t) <- baseline-time row perm[c, t] of the UVData arrays;
        # the (Nblts, Nfreqs, Npols) -> (Npols, Nbls, Ntimes, Nfreqs) gather runs on the device
        lst = uv.lst_array[uv.baseline_array == uv.baseline_array[0]]  # ds.py:913-915
        lst_inds = np.argsort(np.unwrap(lst))  # ds.py:916
        perm = np.zeros((B, T), dtype=np.int64)
        ant1 = np.zeros(B, dtype=np.int64)
        ant2 = np.zeros(B, dtype=np.int64)
        for c, b in enumerate(bls):
            inds = np.nonzero(uv.baseline_array == b)[0]
            perm[c] = inds[lst_inds]
            ant1[c], ant2[c] = uv.ant_1_array[inds[0]], uv.ant_2_array[inds[0]]
        itime = np.asarray(uv.integration_time, dtype=np.float32)[perm]  # utils.py:157
        dev = self._device()
        perm_d = torch.as_tensor(perm.reshape(-1), device=dev)

        def gather(arr, dtype):  # -> (1, 1, Npols, Nbls, Ntimes, Nfreqs)
            t = torch.as_tensor(np.ascontiguousarray(arr), device=dev).to(dtype).reshape(-1, F, P)
            return ops.ingest_gather(t, perm_d).reshape(1, 1, P, B, T, F)

        this = DelaySpectrum.from_arrays(
            gather(uv.data_array, torch.complex64),  # utils.py:38
            gather(uv.flag_array, torch.bool),
            gather(uv.nsample_array, torch.float32).to(torch.float64),  # utils.py:78
            uv.freq_array[None], itime.astype(np.float64), uv.polarization_array, vis_units=uv.vis_units,
            lst_array=np.unwrap(lst)[lst_inds], baseline_array=bls, ant_1_array=ant1,
            ant_2_array=ant2, precision=self.precision, noise_seed=self.noise_seed,
            uvw=np.asarray(uvw_centers[0], dtype=np.float64),  # ds.py:984
        )
        this.Nants_data, this.Nants_telescope = uv.Nants_data, uv.Nants_telescope
        this.x_orientation = uv.x_orientation
        this.generate_noise()  # all zeros while beam/trcvr are inf  (ds.py:987)
        if this._data_unit == Unit("K sr"):
            # visibilities in K sr: the noise (drawn in Jy) is converted too (ds.py:989-994)
            j2k = torch.as_tensor(np.asarray(jy_to_mk(this.freq_array.value).value), device=this._noise.device)
            this._noise = this._noise * j2k.reshape(this.Nspws, 1, 1, 1, 1, this.Nfreqs).to(this._noise.real.dtype)
        if this._data_unit == Unit(""):
            warnings.warn(
                "Data is uncalibrated. Unable to covert noise array "
                "to unicalibrated units.",
                UserWarning,
            )
        if self.freq_array is not None:  # ds.py:1002-1005
            this.select_spectral_windows(freqs=self.freq_array)
        else:
            this.select_spectral_windows(spectral_windows=spectral_windows)
        if self._data is not None and self._data_unit != this._data_unit:
            raise UnitConversionError(
                "Input data object is in units "
                "incompatible with saved DelaySpectrum units."
                f"Saved units are: {self._data_unit}, "
                f"input units are: {this._data_unit}."
            )
        if self.Nuv == 0:
            keep = dict(taper=self.taper, precision=self.precision, noise_seed=self.noise_seed)
            self.__dict__.update(this.__dict__)
            self.__dict__.update(keep)
            self.Nuv = 1
            return
        for name in ("Ntimes", "Nbls", "Nfreqs", "Npols", "Nspws", "vis_units"):
            if getattr(self, name) != getattr(this, name):
                raise ValueError(
                    "Input data differs from previously "
                    f"loaded data. Parameter _{name} is not "
                    "the same."
                )
        for name in ("baseline_array", "polarization_array"):
            if not np.array_equal(getattr(self, name), getattr(this, name)):
                raise ValueError(
                    "Input data differs from previously "
                    f"loaded data. Parameter _{name} is not "
                    "the same."
                )
        if not np.allclose(self.freq_array.value, this.freq_array.value):
            raise ValueError(
                "Input data differs from previously "
                "loaded data. Parameter _freq_array is not "
                "the same."
            )
        if not np.allclose(self.lst_array.value, this.lst_array.value):
            diff_min = np.abs(self.lst_array.value - this.lst_array.value).mean() * 12.0 / np.pi * 60.0
            warnings.warn(
                "Input LST arrays differ on average "
                f"by {diff_min} min. Keeping LST array stored from "
                "the first data set read.",
                UserWarning,
            )
        self.Nuv = 2  # ds.py:1091, 1115-1128
        for attr in ("_data", "_noise", "_nsample", "_flag"):
            setattr(self, attr, torch.cat([getattr(self, attr), getattr(this, attr)], dim=1))

    # ------------------------------------------------------------------ windows
    def _set_delay_array(self):
        """ds.py:3301-3304."""
        self.Ndelays = int(self.Nfreqs)
        f0 = self.freq_array.value[0]
        d = (f0[1] - f0[0]) if self.Nfreqs > 1 else 1.0
        delays = np.fft.fftshift(np.fft.fftfreq(self.Ndelays, d=d))
        self.delay_array = Quantity(delays * 1e9, "ns")

    def select_spectral_windows(self, spectral_windows=None, freqs=None, inplace=True):
        """ds.py:3196-3313."""
        this = self if inplace else self.copy()
        if freqs is not None:
            fq = np.asarray(_value(freqs), dtype=np.float64)
            fq = fq.reshape(1, -1) if fq.ndim == 1 else fq
            flat = this.freq_array.value
            spectral_windows = []
            for _fq in fq:
                sp = []
                for _f in _fq:
                    for row in flat:  # every stored window must hold it (ds.py:3233-3240)
                        hit = np.where(_f == row)[0]
                        if hit.size == 0:
                            raise ValueError(
                                f"Frequency {_f / 1e6:f} MHz not found in frequency array"
                            )
                        sp.extend(hit)
                spectral_windows.append([sp[0], sp[-1]])
        if spectral_windows is None:
            spectral_windows = [(0, this.Nfreqs - 1)]
        if not isinstance(spectral_windows, (list, tuple, np.ndarray)):
            spectral_windows = [spectral_windows]
        if not isinstance(spectral_windows[0], (list, tuple, np.ndarray)):
            spectral_windows = [spectral_windows]
        if not all(len(x) == 2 for x in spectral_windows):
            raise ValueError(
                "Spectral window tuples must only have two "
                "elements (stard_index, end_index)"
            )
        spectral_windows = np.asarray(spectral_windows, dtype=np.int64)
        nspws = spectral_windows.shape[0]
        num_freqs = spectral_windows[:, 1] + 1 - spectral_windows[:, 0]
        if not all(num_freqs == num_freqs[0]):
            raise ValueError("Spectral windows must all have the same size.")
        freq_chans = spectral_windows[:, 0, None] + np.arange(num_freqs[0])  # ds.py:3264
        old_total = this.Nspws * this.Nfreqs
        if freq_chans.size and (freq_chans.min() < 0 or freq_chans.max() >= old_total):
            raise IndexError("spectral window channel index out of range")
        this.freq_array = Quantity(np.take(this.freq_array.value, freq_chans), "Hz")
        this.trcvr = Quantity(np.take(this.trcvr.value, freq_chans), "K")
        flat = freq_chans.flatten()
        for name in ("beam_sq_area", "beam_area"):  # ds.py:3273-3283
            arr = np.transpose(getattr(this, name).value, [1, 0, 2]).reshape(this.Npols, -1)
            arr = np.take(arr, flat, axis=1).reshape(this.Npols, nspws, int(num_freqs[0]))
            setattr(this, name, Quantity(np.transpose(arr, [1, 0, 2]).copy(), "sr"))
        for attr in ("_data", "_noise", "_nsample", "_flag"):  # ds.py:3286-3297
            setattr(this, attr, ops.take_windows(getattr(this, attr), freq_chans))
        this.Nspws, this.Nfreqs = int(nspws), int(num_freqs[0])
        this._set_delay_array()
        # like the reference (ds.py:3306-3308): redshift, k_parallel and k_perpendicular exist right
        # after a window selection; the power arrays are left alone
        this.update_cosmology()
        if not inplace:
            return this

    # ------------------------------------------------------------------ noise
    def _tsys(self):
        """180 K (nu / 0.18 GHz)^-2.55 + Trcvr   (ds.py:3546-3551, 3668-3672)."""
        with np.errstate(divide="ignore", invalid="ignore"):
            return 180.0 * np.power(self.freq_array.value * 1e-9 / 0.18, -2.55) + self.trcvr.value

    def calculate_noise_power(self):
        """ds.py:3544-3583: radiometer sigma in Jy, data_array's shape."""
        freq = self.freq_array.value
        delta_f = freq[0, 1] - freq[0, 0] if self.Nfreqs > 1 else np.nan
        with np.errstate(divide="ignore", invalid="ignore"):
            coef = (
                self._tsys()[:, None, :] / np.sqrt(delta_f) * 1e3 * self.beam_area.value
                / utils.jy_to_mk(freq).value[:, None, :]
            )
        dev = self._device()
        sigma = ops.noise_sigma(
            torch.as_tensor(coef, device=dev), torch.as_tensor(self.integration_time.value, device=dev),
            self._nsample,
        )
        return Quantity(sigma, "Jy")

    def generate_noise(self, normals=None, seed=None):
        """ds.py:3340-3343.  ``normals=(g_re, g_im)`` (data_array-shaped N(0,1) draws)
        reproduces the reference's numpy stream exactly; otherwise Philox on device."""
        sigma = self.calculate_noise_power().value
        if seed is None:
            seed = self.noise_seed + self._noise_draws
            self._noise_draws += 1
        self._noise = ops.generate_noise(sigma, seed=seed, normals=normals, out_dtype=self._cdtype)

    # ------------------------------------------------------------------ transform
    def normalized_fourier_transform(self, inverse=False):
        """ds.py:3487-3517: data and noise, masked by ~flag_array, along frequency."""
        if inverse is True:
            d = self.delay_array.value
            delta_x = (d[1] - d[0]) * 1e-9  # .si of ns
            du = Unit("s")
        else:
            f0 = self.freq_array.value[0]
            delta_x = f0[1] - f0[0]
            du = Unit("Hz")
        plan = ops.get_plan(self.Nfreqs, self._math, self.taper)
        self._data = ops.delay_transform(plan, self._data, self._flag, delta_x, inverse=inverse)
        self._noise = ops.delay_transform(plan, self._noise, self._flag, delta_x, inverse=inverse)
        self._data_unit = self._data_unit * du

    def delay_transform(self):
        """ds.py:3519-3542."""
        if self._data_unit == Unit(""):
            warnings.warn(
                "Fourier Transforming uncalibrated data. Units will "
                "not have physical meaning. "
                "Data will be arbitrarily scaled.",
                UserWarning,
            )
        if self.data_type == "frequency":
            self.normalized_fourier_transform()
            self.set_delay()
        elif self.data_type == "delay":
            self.normalized_fourier_transform(inverse=True)
            if self._data_unit.is_equivalent([Unit("Jy Hz s"), Unit("K sr Hz s"), Unit("Hz s")]):
                self._data_unit = self._data_unit / Unit("Hz s")  # Hz * s cancels
            self.set_frequency()
        else:
            raise ValueError(
                f"Unknown data type: {self.data_type}. Unable to perform "
                "delay transformation."
            )

    # ------------------------------------------------------------------ power
    def calculate_delay_spectrum(self, run_check=True, run_check_acceptability=True,
                                 cosmology=None, littleh_units=False):
        """ds.py:3585-3635: delay transform, cross multiplication, thermal estimate and, last,
        ``update_cosmology`` (power in mK^2 Mpc^3; ``remove_cosmology()`` undoes it)."""
        if self.Nuv == 0:
            raise ValueError(
                "No data has be loaded. Add UVData objects before "
                "calling calculate_delay_spectrum."
            )
        if self.data_type == "frequency":
            self.delay_transform()
        j = 0 if self.Nuv == 1 else 1
        # The reference cross-multiplies, estimates the thermal power and then calls update_cosmology,
        # which multiplies the three O(Nbls^2) tensors by a per-(window, pol) factor in a second pass
        # (ds.py:3620-3635, 1306-1311, 1339-1343).  Here the factors are worked out first and ride in
        # the kernels that write the tensors: every tensor is written once, already converted.
        self.power_array = self.noise_power = self.thermal_power = None
        self.unit_conversion = self.thermal_conversion = None
        cv = self._cosmology_factors(cosmology, littleh_units, self._data_unit ** 2, want_thermal=True)
        self.power_array = Quantity(ops.cross_power_full(self._data[:, 0], self._data[:, j],
                                                         scale_sp=cv["power_scale"]), cv["power_unit"])
        self.noise_power = Quantity(ops.cross_power_full(self._noise[:, 0], self._noise[:, j],
                                                         scale_sp=cv["power_scale"]), cv["power_unit"])
        self.calculate_thermal_sensitivity(scale_sp=cv["thermal_scale"], unit=cv["thermal_unit"])

    # ------------------------------------------------------------------ cosmology
    _LITTLEH_POWER = Unit("mK^2 Mpc^3 littleh^-3")

    def _scale_sp(self, q, factor_sp, unit):
        """q * factor[s, p] broadcast over the trailing axes, relabelled."""
        v = q.value
        f = torch.as_tensor(np.asarray(factor_sp, dtype=np.float64), device=v.device)
        f = f.reshape(f.shape + (1,) * (v.dim() - 2))
        return Quantity(v * f, unit)

    def remove_cosmology(self):
        """ds.py:1133-1199: take the cosmological conversion back out of the power estimates."""
        if self.cosmology is None:
            raise ValueError(f"Cannot remove cosmology of type {type(None)}")
        self.k_parallel = None
        self.k_perpendicular = None
        h3 = (self.cosmology.H0 / 100.0) ** 3
        if self.power_array is not None:
            if self.power_array.unit == self._LITTLEH_POWER:
                self.power_array = Quantity(self.power_array.value / h3, "mK^2 Mpc^3")
                self.noise_power = Quantity(self.noise_power.value / h3, "mK^2 Mpc^3")
            if self.unit_conversion is not None and not self.power_array.unit.is_equivalent(
                    (Unit("Jy^2 Hz^2"), Unit("K^2 sr^2 Hz^2"))):
                with np.errstate(divide="ignore", invalid="ignore"):
                    inv = 1.0 / np.asarray(self.unit_conversion.value, dtype=np.float64)
                unit = self.power_array.unit / self.unit_conversion.unit
                self.power_array = self._scale_sp(self.power_array, inv, unit)
                self.noise_power = self._scale_sp(self.noise_power, inv, unit)
        if self.thermal_power is not None:
            if self.thermal_power.unit == self._LITTLEH_POWER:
                self.thermal_power = Quantity(self.thermal_power.value / h3, "mK^2 Mpc^3")
            if self.thermal_conversion is not None and not self.thermal_power.unit.is_equivalent(
                    (Unit("Jy^2 Hz^2"), Unit("K^2 sr^2 Hz^6"))):
                with np.errstate(divide="ignore", invalid="ignore"):
                    inv = 1.0 / np.asarray(self.thermal_conversion.value, dtype=np.float64)
                self.thermal_power = self._scale_sp(self.thermal_power, inv,
                                                    self.thermal_power.unit / self.thermal_conversion.unit)
        self.unit_conversion = None
        self.thermal_conversion = None

    def _conversion_integral(self, x2y, freq_power):
        """trapz_f(freq^pw / X2Y * taper^2 * beam_sq_area) per (window, pol), on the device."""
        dev = self._device()
        freq = torch.as_tensor(np.ascontiguousarray(self.freq_array.value, dtype=np.float64), device=dev)
        x = torch.as_tensor(np.ascontiguousarray(x2y, dtype=np.float64), device=dev)
        tap = torch.as_tensor(np.ascontiguousarray(self.taper(self.Nfreqs), dtype=np.float64), device=dev)
        beam = torch.as_tensor(np.ascontiguousarray(_value(self.beam_sq_area), dtype=np.float64), device=dev)
        out = torch.empty((self.Nspws, self.Npols), dtype=torch.float64, device=dev)
        rc = ops._lib.load().sds_conversion_integral(ops.ptr(freq), ops.ptr(x), ops.ptr(tap), ops.ptr(beam),
                                                     self.Nspws, self.Npols, self.Nfreqs, int(freq_power),
                                                     ops.ptr(out), ops.stream_ptr())
        ops.check(rc, "sds_conversion_integral")
        return out.cpu().numpy()

    def _cosmology_factors(self, cosmology, littleh_units, power_unit, want_thermal):
        """The state part of ds.py:1201-1365 -- cosmology, redshift, k_parallel, k_perpendicular,
        unit_conversion, thermal_conversion -- and the per-(window, pol) factors that take power
        estimates in ``power_unit`` (None: no power array) and the thermal estimate to mK^2 Mpc^3
        (littleh folded in).  Returns dict(power_scale, power_unit, thermal_scale, thermal_unit)."""
        from . import cosmo as simple_cosmo

        if cosmology is not None:
            if not isinstance(cosmology, simple_cosmo.FlatLambdaCDM):
                raise ValueError("Input cosmology must a sub-class of astropy.cosmology.core.Cosmology")
            self.cosmology = cosmology
        else:
            self.cosmology = simple_cosmo.default_cosmology.get()
        freq = np.asarray(self.freq_array.value, dtype=np.float64)
        self.redshift = simple_cosmo.calc_z(self.freq_array).mean(axis=1)
        delay_s = np.asarray(self.delay_array.value, dtype=np.float64) * 1e-9
        self.k_parallel = simple_cosmo.eta2kparr(Quantity(delay_s.reshape(1, self.Ndelays), "s"),
                                                 self.redshift.reshape(self.Nspws, 1), cosmo=self.cosmology)
        self.k_perpendicular = None
        if self.uvw is not None:
            uvw_wave = np.linalg.norm(np.asarray(_value(self.uvw), dtype=np.float64)) / (
                simple_cosmo.C_M_S / freq.mean(axis=1))
            self.k_perpendicular = simple_cosmo.u2kperp(uvw_wave, self.redshift, cosmo=self.cosmology)
        out = dict(power_scale=None, power_unit=None, thermal_scale=None, thermal_unit=None)
        x2y = i4 = None
        if power_unit is not None or want_thermal:
            x2y = simple_cosmo.X2Y(simple_cosmo.calc_z(self.freq_array), cosmo=self.cosmology).value
            with np.errstate(divide="ignore", invalid="ignore"):
                i4 = self._conversion_integral(x2y, 4)
        h3 = (self.cosmology.H0 / 100.0) ** 3 if littleh_units else 1.0
        if power_unit is not None:
            with np.errstate(divide="ignore", invalid="ignore"):
                if power_unit == Unit("Jy^2 Hz^2"):
                    # (c^2 sr / 2 k_B)^2 / integral, SI -> mK^2 Mpc^3 / (Jy^2 Hz^2): K^2 -> mK^2, J -> 1e26 Jy m^2
                    conv = (simple_cosmo.C_M_S**2 / (2 * simple_cosmo.K_B)) ** 2 / i4 * 1e6 * 1e-52
                    self.unit_conversion = Quantity(conv, "mK^2 Mpc^3 Jy^-2 Hz^-2")
                elif power_unit == Unit("K^2 sr^2 Hz^2"):
                    conv = 1.0 / self._conversion_integral(x2y, 0) * 1e6
                    self.unit_conversion = Quantity(conv, "mK^2 Mpc^3 K^-2 sr^-2 Hz^-2")
                else:
                    self.unit_conversion = Quantity(np.ones((self.Nspws, self.Npols)), "")
            unit = Unit(power_unit) * self.unit_conversion.unit
            if not self._data_unit == Unit("Hz"):
                unit = Unit("mK^2 Mpc^3")
            out["power_scale"] = np.asarray(self.unit_conversion.value, dtype=np.float64) * h3
            out["power_unit"] = self._LITTLEH_POWER if littleh_units else unit
        if want_thermal:
            with np.errstate(divide="ignore", invalid="ignore"):
                self.thermal_conversion = Quantity(1.0 / i4 * 1e6, "mK^2 Mpc^3 K^-2 sr^-2 Hz^-6")
            out["thermal_scale"] = np.asarray(self.thermal_conversion.value, dtype=np.float64) * h3
            out["thermal_unit"] = self._LITTLEH_POWER if littleh_units else Unit("mK^2 Mpc^3")
        if littleh_units:
            h = self.cosmology.H0 / 100.0
            self.k_parallel = Quantity(self.k_parallel.value / h, "littleh Mpc^-1")
            if self.k_perpendicular is not None:
                self.k_perpendicular = Quantity(self.k_perpendicular.value / h, "littleh Mpc^-1")
        return out

    def update_cosmology(self, cosmology=None, littleh_units=False):
        """ds.py:1201-1365: redshift, k_parallel, k_perpendicular and the conversion of the power
        estimates to mK^2 Mpc^3 (per (window, pol) integrals of nu^4 / X2Y * taper^2 * beam^2).
        (``calculate_delay_spectrum`` applies the same factors inside the kernels that write the
        tensors; this method re-scales tensors that already exist.)"""
        self.remove_cosmology()
        cv = self._cosmology_factors(cosmology, littleh_units,
                                     None if self.power_array is None else self.power_array.unit,
                                     want_thermal=self.thermal_power is not None)
        if self.power_array is not None:
            self.power_array = self._scale_sp(self.power_array, cv["power_scale"], cv["power_unit"])
            self.noise_power = self._scale_sp(self.noise_power, cv["power_scale"], cv["power_unit"])
        if self.thermal_power is not None:
            self.thermal_power = self._scale_sp(self.thermal_power, cv["thermal_scale"], cv["thermal_unit"])

    def _thermal_inputs(self):
        freq = self.freq_array.value
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            amp = self._tsys()[:, None, :] * 1e3 * self.beam_area.value  # mK sr
            coef = amp**2 * np.asarray(self.taper(self.Nfreqs), dtype=np.float64) ** 2 * freq[:, None, :] ** 4
            coef = np.where(np.isfinite(amp), coef, np.nan)
        npol = np.array([2 if p in np.arange(1, 5) else 1 for p in self.polarization_array])
        return coef, freq, 1.0 / (npol * np.sqrt(2))

    def calculate_thermal_sensitivity(self, scale_sp=None, unit="K^2 sr^2 Hz^6"):
        """ds.py:3637-3707 -> thermal_power (Nspws,Npols,Nbls,Nbls,Ntimes), K^2 sr^2 Hz^6 (times
        ``scale_sp[s, p]`` and labelled ``unit`` when calculate_delay_spectrum passes the conversion)."""
        coef, freq, pf = self._thermal_inputs()
        j = 0 if self.Nuv == 1 else 1
        th = ops.thermal_power_full(self._nsample[:, 0], self._nsample[:, j], coef, freq,
                                    self.integration_time.value, pf, scale_sp=scale_sp)
        self.thermal_power = Quantity(th, unit)

    def calculate_averaged_delay_spectrum(self, remove_autos=False, average_time=True,
                                          method="factorised"):
        """cross multiply + the tutorial's mean over baseline pairs (and time)
        without forming the Nbls x Nbls tensor.  Returns dict of Quantities
        ``power``, ``noise_power`` (Nspws,Npols,[Ntimes,]Ndelays)."""
        if self.Nuv == 0:
            raise ValueError(
                "No data has be loaded. Add UVData objects before "
                "calling calculate_delay_spectrum."
            )
        if self.data_type == "frequency":
            self.delay_transform()
        j = 0 if self.Nuv == 1 else 1
        n = self.Nbls
        out = {}
        for key, arr in (("power", self._data), ("noise_power", self._noise)):
            s_all, s_auto = ops.cross_power_sums(arr[:, 0], arr[:, j], method=method)
            v = (s_all[0] - s_auto[0]) / (n * (n - 1)) if remove_autos else s_all[0] / (n * n)
            if average_time:
                v = v.mean(dim=2)
            out[key] = Quantity(v, self._data_unit ** 2)
        return out
