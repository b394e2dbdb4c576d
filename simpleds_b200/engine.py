"""Batched multi-group front end of the fused path.

The reference handles one redundant group per ``DelaySpectrum`` object and the
caller loops over groups (ds.py:891-900) and averages the (Nbls, Nbls) power
tensor afterwards (tutorial cell 41).  ``RedundantPipeline`` takes a whole array
at once -- visibilities (Npols, Nbls, Ntimes, Nchan) with baselines sorted by
redundant group -- and returns, per group / spectral window / polarisation /
delay, the mean over baseline pairs and times of the data power, of the power of
a matching noise realisation, and of the thermal estimate, i.e. what
``select_spectral_windows -> generate_noise -> calculate_delay_spectrum ->
mean`` gives per group.  One call of ``sds_engine_run`` (include/sds.h) does the
work: power-of-two windows (64..2048) in either precision, any other window
length 8..1024 in complex64 (chirp-z engine); unaligned layouts and complex128 with
other lengths chain the API-level kernels instead.
"""
import ctypes as C

import numpy as np
import torch
from scipy.signal import windows

from . import _lib, ops
from ._lib import SDS_F32, SDS_F64, EngineDesc, check, ptr, stream_ptr

C_M_PER_S = 299792458.0
K_B_J_PER_K = 1.380649e-23


def radiometer_tables(freq_hz, trcvr_k, beam_area_sr, polarization_array, taper_values, delta_f):
    """Per-channel factors of the noise model, evaluated once on the host.

    freq_hz (S,N), trcvr_k (S,N), beam_area_sr (S,P,N).  Returns
    sigma_coef (S,P,N) [ds.py:3546-3581 without the per-sample 1/sqrt(t n)],
    thermal_coef (S,P,N) [ds.py:3668-3699 channel part], pol_factor (P,)."""
    freq_hz = np.asarray(freq_hz, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        tsys = 180.0 * np.power(freq_hz * 1e-9 / 0.18, -2.55) + np.asarray(trcvr_k, dtype=np.float64)
        jy2mk = 1e-23 * C_M_PER_S**2 / (2.0 * freq_hz**2 * K_B_J_PER_K)
        amp = tsys[:, None, :] * 1e3 * np.asarray(beam_area_sr, dtype=np.float64)
        sigma_coef = amp / np.sqrt(delta_f) / jy2mk[:, None, :]
        thermal_coef = amp**2 * np.asarray(taper_values, dtype=np.float64) ** 2 * freq_hz[:, None, :] ** 4
        thermal_coef = np.where(np.isfinite(amp), thermal_coef, np.nan)
    npol = np.array([2 if p in np.arange(1, 5) else 1 for p in polarization_array])
    # numpy keeps the (transposed) memory order of its inputs: the device tables must be C-ordered
    return (np.ascontiguousarray(sigma_coef), np.ascontiguousarray(thermal_coef),
            np.ascontiguousarray(1.0 / (npol * np.sqrt(2))))


class RedundantPipeline:
    """See module docstring.

    Parameters: ``freq_hz`` (Nchan,) channel frequencies; ``spectral_windows`` list of
    (start, end) inclusive channel ranges of equal length (ds.py:3243-3264);
    ``group_offset`` (Ngroups+1,) rising baseline offsets of the redundant groups;
    ``trcvr_k`` scalar or (Nchan,); ``beam_area_sr`` scalar, (Nchan,) or (Npols, Nchan);
    ``precision`` "complex64" (fp32 math) or "complex128" (fp64 math); ``nuv`` = 2: two data
    sets per call (arrays with a leading axis of length 2), cross-multiplied like the reference's
    ``data_array[:, 0]`` x ``data_array[:, 1]`` (ds.py:3620-3633, thermal 3655-3662): the pair means
    are then complex.  Two sets take the fused path in complex64 with power-of-two windows of
    64..1024 channels."""

    def __init__(self, freq_hz, spectral_windows, polarization_array, group_offset, trcvr_k,
                 beam_area_sr, taper=None, precision="complex64", seed=0, device=None, nuv=1):
        _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.taper = windows.blackmanharris if taper is None else taper
        if not callable(self.taper):
            raise ValueError(
                "Input spectral taper must be a function or callable whose arguments are "
                "the length of the band over which Fourier Transform is taken."
            )
        if precision not in ("complex64", "complex128"):
            raise ValueError("precision must be 'complex128' or 'complex64'")
        if nuv not in (1, 2):
            raise ValueError("nuv must be 1 (one data set) or 2 (the cross of two data sets, ds.py:3620-3633)")
        self.nuv = int(nuv)
        self.precision, self.seed = precision, int(seed)
        self.math = SDS_F64 if precision == "complex128" else SDS_F32
        freq_hz = np.asarray(freq_hz, dtype=np.float64).reshape(-1)
        self.nchan = freq_hz.size
        sw = np.asarray(spectral_windows, dtype=np.int64).reshape(-1, 2)
        nf = sw[:, 1] + 1 - sw[:, 0]
        if not np.all(nf == nf[0]):
            raise ValueError("Spectral windows must all have the same size.")
        if sw.min() < 0 or sw[:, 1].max() >= self.nchan:
            raise IndexError("spectral window channel index out of range")
        self.win_start = [int(x) for x in sw[:, 0]]
        self.nspw, self.nfreq = sw.shape[0], int(nf[0])
        self.polarization_array = np.asarray(polarization_array, dtype=np.int64)
        self.npol = self.polarization_array.size
        go = np.asarray(group_offset, dtype=np.int64)
        if go.ndim != 1 or go.size < 2 or go[0] != 0 or np.any(np.diff(go) < 0):
            raise ValueError("group_offset must rise from 0 to Nbls")
        self.group_offset_host = go
        self.ngroup, self.nbl = go.size - 1, int(go[-1])
        self.group_sizes = np.diff(go)
        chans = sw[:, 0, None] + np.arange(self.nfreq)
        self.freq_windows = freq_hz[chans]  # (S, N)
        # delta_x of the transform is the FIRST window's channel width (ds.py:3502)
        self.delta_f = float(self.freq_windows[0, 1] - self.freq_windows[0, 0]) if self.nfreq > 1 else 1.0
        trcvr = np.broadcast_to(np.asarray(trcvr_k, dtype=np.float64), (self.nchan,))[chans]
        beam = np.broadcast_to(np.asarray(beam_area_sr, dtype=np.float64), (self.npol, self.nchan))
        beam = np.transpose(beam[:, chans], (1, 0, 2))  # (S, P, N)
        self.taper_values = np.asarray(self.taper(self.nfreq), dtype=np.float64)
        sc, tc, pf = radiometer_tables(self.freq_windows, trcvr, beam, self.polarization_array,
                                       self.taper_values, self.delta_f)
        dev = self.device
        sc32 = np.ascontiguousarray(np.nan_to_num(sc, nan=np.inf, posinf=np.inf), dtype=np.float32)
        self.sigma_coef64 = torch.as_tensor(sc, device=dev).contiguous()
        self.sigma_coef32 = torch.as_tensor(sc32, device=dev).contiguous()
        self.thermal_coef = torch.as_tensor(tc, device=dev).contiguous()
        self.pol_factor = torch.as_tensor(pf, device=dev).contiguous()
        self.freq_dev = torch.as_tensor(np.ascontiguousarray(self.freq_windows), device=dev).contiguous()
        self.group_offset = torch.as_tensor(go.astype(np.int32), device=dev)
        self.delay_array_ns = np.fft.fftshift(np.fft.fftfreq(self.nfreq, d=self.delta_f)) * 1e9
        self.plan = ops.get_plan(self.nfreq, self.math, self.taper)
        self.fused = self._fused_ok()
        self.kernel_launches = 0
        self.reset()

    def _fused_ok(self):
        """Window lengths the one-pass engine takes: powers of two 64..2048, or any other
        length 8..1024 in complex64 (chirp-z engine); rows and windows 16-sample aligned."""
        n = self.nfreq
        if self.nchan % 16 != 0 or any(w % 16 != 0 for w in self.win_start):
            return False
        if self.nuv == 2:
            return (n & (n - 1)) == 0 and 64 <= n <= 1024 and self.precision == "complex64"
        if (n & (n - 1)) == 0:
            return 64 <= n <= 2048
        npad = (n + 15) // 16 * 16
        return 8 <= n <= 1024 and self.precision == "complex64" and \
            all(w + npad <= self.nchan for w in self.win_start)

    def reset(self):
        dev, shape = self.device, (self.ngroup, self.nspw, self.npol, self.nfreq)
        # with one data set (Nuv = 1) the pair sums are real: sum_ij x_i conj(x_j) = |sum_i x_i|^2;
        # with two they are (sum_i A2_i) conj(sum_j A1_j): complex
        # ONE flat fp64 buffer holds every accumulator (zeroed with one fill, reduced over the ranks in
        # place with one collective: reduce()); the tensors below are views of it
        cplx = self.nuv != 1
        npair = int(np.prod(shape)) * (2 if cplx else 1)
        nth = int(np.prod(shape[:3]))
        self._flat = torch.zeros(4 * npair + 2 * nth + self.ngroup, dtype=torch.float64, device=dev)

        def pair_view(k):
            v = self._flat[k * npair:(k + 1) * npair]
            return torch.view_as_complex(v.view(*shape, 2)) if cplx else v.view(shape)

        self.pow_all, self.pow_auto, self.noise_all, self.noise_auto = (pair_view(k) for k in range(4))
        self.thermal_all = self._flat[4 * npair:4 * npair + nth].view(shape[:3])
        self.thermal_auto = self._flat[4 * npair + nth:4 * npair + 2 * nth].view(shape[:3])
        # times accumulated per group: travels with the sums through reduce(), so that result() is
        # right whichever way the job was sharded (times split over ranks add up, groups dealt to
        # ranks are disjoint)
        self.times_done = self._flat[4 * npair + 2 * nth:]
        self.ntimes_done = 0  # times this process has passed to process() (bookkeeping only)
        self._ws = None

    # ------------------------------------------------------------------ one batch
    def process(self, vis, flags, nsample, int_time, time_offset=0, ntime_total=None, noise="philox",
                thermal=True, noise_in=None, groups=None):
        """Accumulate one batch of times.  vis complex64 (P,B,T,Nchan); flags bool/uint8
        and nsample float32 of the same shape; int_time float32 (B,T).  ``noise``:
        "philox" (in-kernel realisation: the moment-matched discrete draw of
        csrc/sds_engine.cuh in the power-of-two engine), "philox_gauss" (in-kernel
        Box-Muller draw per channel), None, or "inject" with ``noise_in`` (a
        frequency-domain noise array shaped like vis; parity tests).

        ``groups``: rising indices of the redundant groups this call covers (sharding by
        group: every group is its own object in the reference, ds.py:891-900).  The arrays
        then hold only the baselines of those groups, in that order; the other groups' sums
        are not touched and the noise stream stays keyed by the global baseline index.

        A streamed job passes ``time_offset`` (global index of the batch's first time) and
        ``ntime_total``: the in-kernel noise is a function of the global (baseline, time)
        coordinates, so batches and shards of one job realise one and the same noise field."""
        if self.nuv == 2:
            if vis.ndim != 5 or vis.shape[0] != 2:
                raise ValueError("nuv = 2: vis, flags, nsample (and noise_in) carry a leading axis of length 2")
            if not self.fused:
                raise NotImplementedError("two data sets take the fused path only: complex64, aligned power-of-two "
                                          "windows of 64..1024 channels (DelaySpectrum covers the rest)")
            if noise == "philox_gauss":
                raise NotImplementedError("nuv = 2: the in-kernel noise is the production draw (noise='philox')")
        elif vis.ndim != 4:
            raise ValueError("vis must be (Npols, Nbls, Ntimes, Nchan)")
        P, B, T, F0 = vis.shape[-4:]
        sub = None
        if groups is not None:
            sub = self._group_subset(groups)
        nbl_expect = self.nbl if sub is None else sub["nbl"]
        if (P, B, F0) != (self.npol, nbl_expect, self.nchan):
            raise ValueError(f"vis has shape {tuple(vis.shape)}, expected (..., {self.npol}, {nbl_expect}, T, {self.nchan})")
        if tuple(flags.shape) != tuple(vis.shape) or tuple(nsample.shape) != tuple(vis.shape) \
                or tuple(int_time.shape) != (B, T):
            raise ValueError("flags/nsample must match vis and int_time must be (Nbls, Ntimes)")
        if vis.dtype != torch.complex64 or nsample.dtype != torch.float32 or int_time.dtype != torch.float32:
            raise ValueError("vis must be complex64, nsample and int_time float32")
        fl = flags.view(torch.uint8) if flags.dtype == torch.bool else flags
        if fl.dtype != torch.uint8:
            raise ValueError("flags must be bool or uint8")
        vis, fl, nsample, int_time = vis.contiguous(), fl.contiguous(), nsample.contiguous(), int_time.contiguous()
        mode = {None: 0, "philox": 1, "inject": 2, "philox_gauss": 3}[noise]
        if mode == 2:
            if noise_in is None or tuple(noise_in.shape) != tuple(vis.shape) or noise_in.dtype != torch.complex64:
                raise ValueError("noise='inject' needs noise_in complex64 shaped like vis")
            noise_in = noise_in.contiguous()
        time_offset = int(time_offset)
        if ntime_total is None:
            if time_offset != 0:
                raise ValueError("a batch with time_offset > 0 needs ntime_total (the noise counters are keyed "
                                 "by the global time index)")
            ntot = T
        else:
            ntot = int(ntime_total)
        if time_offset < 0 or time_offset + T > ntot:
            raise ValueError(f"times [{time_offset}, {time_offset + T}) do not fit ntime_total = {ntot}")
        if T == 0:  # an empty batch of a streamed job adds nothing
            return
        if sub is not None and len(sub["index"]) == 0:
            return
        if self.fused:
            # two data sets: the thermal sums ride in the noise launch; without noise they go through
            # the chained kernels
            in_kernel = thermal and (self.nuv == 1 or mode != 0)
            self._process_fused(vis, fl, nsample, int_time, mode, noise_in, in_kernel, time_offset, ntot, subset=sub)
            if thermal and not in_kernel:
                self._thermal_two_sets(nsample, int_time, sub)
        else:
            if sub is not None:
                raise NotImplementedError("groups= needs the fused path (aligned windows of a supported length)")
            self._process_unfused(vis, fl, nsample, int_time, mode, noise_in, thermal, time_offset, ntot)
        if sub is None:
            self.times_done += T
        else:
            self.times_done[sub["index_dev"].long()] += T
        self.ntimes_done += T

    def _group_subset(self, groups):
        """Device tables of a rising list of group indices (cached): local group offsets, output
        rows, global first-baseline indices."""
        key = tuple(int(g) for g in groups)
        cache = self.__dict__.setdefault("_subsets", {})
        if key not in cache:
            idx = np.asarray(key, dtype=np.int64)
            if idx.size and (np.any(np.diff(idx) <= 0) or idx[0] < 0 or idx[-1] >= self.ngroup):
                raise ValueError("groups must be rising indices of redundant groups")
            sizes = self.group_sizes[idx] if idx.size else np.zeros(0, np.int64)
            dev = self.device
            cache[key] = dict(
                index=idx, nbl=int(sizes.sum()),
                goff_dev=torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32), device=dev),
                index_dev=torch.as_tensor(idx.astype(np.int32), device=dev),
                bl0_dev=torch.as_tensor(self.group_offset_host[idx].astype(np.int64), device=dev),
            )
        return cache[key]

    def _process_fused(self, vis, fl, nsample, int_time, mode, noise_in, thermal, time_offset, ntot,
                       groups=None, subset=None):
        """One sds_engine_run.  ``groups`` = (g0, g1, goff_dev): the arrays hold only the
        baselines of redundant groups g0..g1-1 (a chunk of process_host); ``subset``: the
        tables of an arbitrary rising list of groups (_group_subset).  The noise counters
        stay keyed by the global baseline index."""
        lib = _lib.load()
        P, B, T, F0 = vis.shape[-4:]
        g0, g1, goff = (0, self.ngroup, self.group_offset) if groups is None else groups
        d = EngineDesc()
        d.npol, d.nbl, d.ntime, d.nchan = P, B, T, F0
        d.nspw, d.ngroup, d.nuv, d.noise_mode = self.nspw, g1 - g0, self.nuv, mode
        if subset is not None:
            g0, g1, goff = 0, self.ngroup, subset["goff_dev"]
            d.ngroup = len(subset["index"])
            d.group_index = subset["index_dev"].data_ptr()
            d.group_bl_start = subset["bl0_dev"].data_ptr()
        for i, w in enumerate(self.win_start):
            d.win_start[i] = w
        d.time_offset, d.ntime_total = int(time_offset), ntot
        d.bl_offset, d.nbl_total = int(self.group_offset_host[g0]), self.nbl
        d.seed, d.delta_f = self.seed, self.delta_f
        need = lib.sds_engine_workspace_bytes(self.plan.handle, C.byref(d))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=self.device)
        launches = C.c_int32(0)
        rc = lib.sds_engine_run(
            self.plan.handle, C.byref(d), ptr(vis), ptr(fl), ptr(nsample), ptr(int_time),
            ptr(goff), ptr(self.sigma_coef32), ptr(self.thermal_coef if thermal else None),
            ptr(self.freq_dev), ptr(self.pol_factor), ptr(noise_in), ptr(self.pow_all[g0:g1]),
            ptr(self.pow_auto[g0:g1]), ptr(self.noise_all[g0:g1]), ptr(self.noise_auto[g0:g1]),
            ptr(self.thermal_all[g0:g1]), ptr(self.thermal_auto[g0:g1]),
            ptr(self._ws), stream_ptr(), C.byref(launches),
        )
        check(rc, "sds_engine_run")
        self.kernel_launches += launches.value

    def _thermal_two_sets(self, nsample, int_time, sub):
        """Thermal estimate of two data sets (ds.py:3655-3662: N_ij = sqrt(n1_j n2_i)): the pair
        sums of sds_thermal_power_avg on the two window-selected nsample arrays."""
        chans = np.asarray(self.win_start)[:, None] + np.arange(self.nfreq)
        n1 = ops.take_windows(nsample[0][None], chans)  # (S,P,B,T,N)
        n2 = ops.take_windows(nsample[1][None], chans)
        go = self.group_offset_host if sub is None else sub["goff_dev"].cpu().numpy().astype(np.int64)
        t_all, t_auto = ops.thermal_power_sums(n1, n2, self.thermal_coef, self.freq_dev, int_time.to(torch.float64),
                                               self.pol_factor, go)
        self.kernel_launches += 3
        if sub is None:
            self.thermal_all += t_all.sum(dim=3)
            self.thermal_auto += t_auto.sum(dim=3)
        else:
            idx = sub["index_dev"].long()
            self.thermal_all[idx] += t_all.sum(dim=3)
            self.thermal_auto[idx] += t_auto.sum(dim=3)

    def _process_unfused(self, vis, fl, nsample, int_time, mode, noise_in, thermal, time_offset, ntot):
        """Any window length / alignment: the API-level kernels chained on device
        (delay transform with window origins -> pair sums; sigma -> draw -> transform)."""
        P, B, T, F0 = vis.shape
        S, N = self.nspw, self.nfreq
        x = ops.delay_transform(self.plan, vis, fl, self.delta_f, win_start=self.win_start)  # (S,P,B,T,N)
        s_all, s_auto = ops.cross_power_sums(x, group_offset=self.group_offset_host)
        self.pow_all += s_all.sum(dim=3).real
        self.pow_auto += s_auto.sum(dim=3).real
        self.kernel_launches += 2
        chans = np.asarray(self.win_start)[:, None] + np.arange(N)
        ns_w = ops.take_windows(nsample[None], chans)  # (S,P,B,T,N)
        self.kernel_launches += 1
        it64 = int_time.to(torch.float64)
        if mode:
            fl_w = ops.take_windows(fl[None], chans)
            if mode in (1, 3):
                sig = ops.noise_sigma(self.sigma_coef64, it64, ns_w[:, None])
                # one Philox stream per batch (the fused path's stream is keyed by
                # global sample coordinates instead; both are valid realisations)
                off = int(time_offset) * S * P * B * N
                nz = ops.generate_noise(sig[:, 0], seed=self.seed, offset=off, out_dtype=x.dtype)
                self.kernel_launches += 2
            else:
                nz = ops.take_windows(noise_in[None], chans)
                self.kernel_launches += 1
            xn = ops.delay_transform(self.plan, nz, fl_w, self.delta_f)
            n_all, n_auto = ops.cross_power_sums(xn, group_offset=self.group_offset_host)
            self.noise_all += n_all.sum(dim=3).real
            self.noise_auto += n_auto.sum(dim=3).real
            self.kernel_launches += 3
        if thermal:
            t_all, t_auto = ops.thermal_power_sums(ns_w, ns_w, self.thermal_coef, self.freq_dev, it64,
                                                   self.pol_factor, self.group_offset_host)
            self.thermal_all += t_all.sum(dim=3)
            self.thermal_auto += t_auto.sum(dim=3)
            self.kernel_launches += 1

    def _host_chunks(self, ntime, chunk_bytes):
        """Cut the group list into runs of whole groups of at most ~chunk_bytes of input."""
        per_bl = 13 * self.npol * ntime * self.nchan
        go, chunks, g0 = self.group_offset_host, [], 0
        while g0 < self.ngroup:
            g1 = g0 + 1
            while g1 < self.ngroup and (go[g1 + 1] - go[g0]) * per_bl <= chunk_bytes:
                g1 += 1
            chunks.append((g0, g1))
            g0 = g1
        return chunks

    def process_host(self, vis, flags, nsample, int_time, time_offset=0, ntime_total=None, noise="philox",
                     thermal=True, chunk_bytes=1 << 29, fetch=True):
        """``process`` for HOST arrays (numpy or CPU torch tensors; pinned memory makes the
        copies asynchronous): the baselines are cut into chunks of whole redundant groups,
        chunk k+1 is uploaded on a copy stream while chunk k runs, and ``result_host()``
        is returned (pinned host buffers viewed as numpy arrays).  A streamed job passes
        ``fetch=False`` for every batch (nothing is copied back, nothing is waited for) and
        calls ``result_host()`` once at its end."""
        if self.nuv != 1:
            raise NotImplementedError("process_host streams one data set (nuv = 1)")
        ins = {}
        for name, a, dt in (("vis", vis, torch.complex64), ("flags", flags, torch.uint8),
                            ("nsample", nsample, torch.float32), ("int_time", int_time, torch.float32)):
            t = torch.as_tensor(a)
            if t.dtype == torch.bool:
                t = t.view(torch.uint8)
            if t.dtype != dt or t.device.type != "cpu":
                raise ValueError(f"{name} must be a host array of {dt}")
            ins[name] = t.contiguous()
        P, B, T, F0 = ins["vis"].shape
        if (P, B, F0) != (self.npol, self.nbl, self.nchan) or tuple(ins["int_time"].shape) != (B, T) \
                or tuple(ins["flags"].shape) != (P, B, T, F0) or tuple(ins["nsample"].shape) != (P, B, T, F0):
            raise ValueError("host arrays do not match the pipeline's (Npols, Nbls, T, Nchan)")
        mode = {None: 0, "philox": 1, "philox_gauss": 3}[noise]
        if ntime_total is None and int(time_offset) != 0:
            raise ValueError("a batch with time_offset > 0 needs ntime_total")
        ntot = T if ntime_total is None else int(ntime_total)
        if int(time_offset) < 0 or int(time_offset) + T > ntot:
            raise ValueError(f"times [{time_offset}, {int(time_offset) + T}) do not fit ntime_total = {ntot}")
        dev = self.device
        if not self.fused:  # any window length: one upload, chained kernels
            d = [ins[k].to(dev, non_blocking=True) for k in ("vis", "flags", "nsample", "int_time")]
            self._process_unfused(*d, mode, None, thermal, time_offset, ntot)
            self.ntimes_done += T
            self.times_done += T
            return self.result_host() if fetch else None

        chunks = self._host_chunks(T, chunk_bytes)
        go = self.group_offset_host
        nb_max = max(int(go[g1] - go[g0]) for g0, g1 in chunks)
        key = (nb_max, T)
        if getattr(self, "_stage_key", None) != key:
            n = P * nb_max * T * F0
            self._stage = [dict(vis=torch.empty(n, dtype=torch.complex64, device=dev),
                                flags=torch.empty(n, dtype=torch.uint8, device=dev),
                                nsample=torch.empty(n, dtype=torch.float32, device=dev),
                                int_time=torch.empty(nb_max * T, dtype=torch.float32, device=dev),
                                ready=torch.cuda.Event(), done=torch.cuda.Event()) for _ in range(2)]
            self._stage_key = key
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._goff_chunks = {}
        cur = torch.cuda.current_stream(dev)
        for st in self._stage:
            st["done"].record(cur)  # the staging buffers are free once earlier work has drained
        for k, (g0, g1) in enumerate(chunks):
            st = self._stage[k % 2]
            b0, b1 = int(go[g0]), int(go[g1])
            nb = b1 - b0
            if (g0, g1) not in self._goff_chunks:
                self._goff_chunks[(g0, g1)] = torch.as_tensor((go[g0:g1 + 1] - b0).astype(np.int32), device=dev)
            views = {name: st[name][:P * nb * T * F0].view(P, nb, T, F0) for name in ("vis", "flags", "nsample")}
            itv = st["int_time"][:nb * T].view(nb, T)
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(st["done"])
                for name in ("vis", "flags", "nsample"):
                    for p in range(P):  # one contiguous host block per polarisation
                        views[name][p].copy_(ins[name][p, b0:b1], non_blocking=True)
                itv.copy_(ins["int_time"][b0:b1], non_blocking=True)
                st["ready"].record(self._copy_stream)
            cur.wait_event(st["ready"])
            self._process_fused(views["vis"], views["flags"], views["nsample"], itv, mode, None, thermal,
                                time_offset, ntot, groups=(g0, g1, self._goff_chunks[(g0, g1)]))
            st["done"].record(cur)
        self.ntimes_done += T
        self.times_done += T
        return self.result_host() if fetch else None

    def result_host(self, ntimes=None):
        """``result()`` copied into (cached) pinned host buffers; numpy views are returned."""
        res = self.result(ntimes)
        if not hasattr(self, "_pinned"):
            self._pinned = {}
        out = {}
        for k, v in res.items():
            if not isinstance(v, torch.Tensor):
                out[k] = v
                continue
            buf = self._pinned.get(k)
            if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                buf = self._pinned[k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            buf.copy_(v, non_blocking=True)
            out[k] = buf
        torch.cuda.current_stream(self.device).synchronize()
        return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}

    def reduce(self, dst=0, all_ranks=False):
        """Sum the ranks' partial accumulators and their per-group time counts (one
        collective on one flat buffer; see shard.reduce_partials).  Afterwards ``result()``
        on ``dst`` (every rank with ``all_ranks``) is the result of the whole job, whether
        the ranks split the times or the groups."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        # the accumulators are views of one flat fp64 buffer (reset()): one collective, in place
        # (shard.reduce_partials does the same for a dict of separate tensors)
        if all_ranks:
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
        else:
            dist.reduce(self._flat, dst=dst, op=dist.ReduceOp.SUM)

    # ------------------------------------------------------------------ results
    def sums(self):
        """Raw accumulators (what ranks reduce over NVLink)."""
        return dict(pow_all=self.pow_all, pow_auto=self.pow_auto, noise_all=self.noise_all,
                    noise_auto=self.noise_auto, thermal_all=self.thermal_all,
                    thermal_auto=self.thermal_auto, times_done=self.times_done)

    def result(self, ntimes=None):
        """Means over ordered baseline pairs and times, per (group, spw, pol[, delay]).

        ``*_all``: every ordered pair (i, j), autos included (tutorial cell 41);
        ``*_noautos``: i != j (utils.remove_auto_correlations); NaN for 1-baseline groups."""
        # one launch for the six means (sds_engine_means; the per-group time counts are read on the device)
        lib = _lib.load()
        if getattr(self, "_group_sizes_dev", None) is None:
            self._group_sizes_dev = torch.as_tensor(self.group_sizes.astype(np.int32), device=self.device)
        out = {}
        names = ("power_all", "power_noautos", "noise_power_all", "noise_power_noautos")
        for k in names:
            out[k] = torch.empty_like(self.pow_all)
        out["thermal_all"] = torch.empty_like(self.thermal_all)
        out["thermal_noautos"] = torch.empty_like(self.thermal_all)
        row_len = self.nfreq * (2 if self.pow_all.is_complex() else 1)
        check(lib.sds_engine_means(
            self.ngroup, self.nspw * self.npol, row_len, ptr(self._group_sizes_dev), ptr(self.times_done),
            float(ntimes) if ntimes is not None else 0.0,
            ptr(self.pow_all), ptr(self.pow_auto), ptr(self.noise_all), ptr(self.noise_auto),
            ptr(self.thermal_all), ptr(self.thermal_auto),
            ptr(out["power_all"]), ptr(out["power_noautos"]), ptr(out["noise_power_all"]),
            ptr(out["noise_power_noautos"]), ptr(out["thermal_all"]), ptr(out["thermal_noautos"]),
            stream_ptr()), "sds_engine_means")
        out["delay_array_ns"] = self.delay_array_ns
        return out

    # work and byte counts used by bench.py
    def pair_spectra(self, ntimes):
        return int((self.group_sizes.astype(np.int64) ** 2).sum()) * self.nspw * self.npol * int(ntimes)

    def algorithmic_bytes(self, ntimes):
        distinct = len(set(c for w in self.win_start for c in range(w, w + self.nfreq)))
        return 13 * self.nuv * self.npol * self.nbl * int(ntimes) * distinct
