"""Oracle for the batched / fused path.  TEST INFRASTRUCTURE ONLY.

For every redundant group the reference's own sequence is run through the
restated functions of ``simpleds_oracle`` -- select_spectral_windows ->
(noise) -> delay_transform -> calculate_delay_spectrum, materialising the
(Nspws, Npols, Nbls, Nbls, Ntimes, Ndelays) tensors -- and the user-side mean of
the tutorial (cell 41) / ``utils.remove_auto_correlations`` is taken with numpy.
Also restates, in numpy float32, the noise stream the fused kernel draws
(sds_engine.cu) so that the in-kernel realisation can be checked value by value.
"""
import numpy as np
from scipy.signal import windows

from . import simpleds_oracle as orc
from .philox import philox4x32

_LO = np.uint64(0xFFFFFFFF)


def group_means(vis, flags, nsample, int_time, group_offset, freq_hz, spectral_windows,
                polarization_array, trcvr_k, beam_area_sr, taper=windows.blackmanharris,
                noise_freq=None):
    """vis/flags/nsample (P,B,T,F0) -- or (2,P,B,T,F0) for two data sets (Nuv = 2); int_time (B,T);
    noise_freq: frequency-domain noise (S,P,B,T,N) already window-selected ((S,2,P,B,T,N) for two
    sets), or None (sigma * 0).  Returns dict of arrays indexed [group] -> (S,P,D) means (all pairs /
    no autos; complex for two sets) and thermal (S,P)."""
    two = vis.ndim == 5
    if not two:
        vis, flags, nsample = vis[None], flags[None], nsample[None]
        noise_freq = None if noise_freq is None else noise_freq[:, None]
    U, P, B, T, F0 = vis.shape
    freq_hz = np.asarray(freq_hz, dtype=np.float64).reshape(1, F0)
    trcvr = np.broadcast_to(np.asarray(trcvr_k, dtype=np.float64), (F0,)).reshape(1, F0)
    beam = np.broadcast_to(np.asarray(beam_area_sr, dtype=np.float64), (P, F0)).reshape(1, P, F0)
    out = {k: [] for k in ("power_all", "power_noautos", "noise_power_all", "noise_power_noautos",
                           "thermal_all", "thermal_noautos", "sigma")}
    for g in range(len(group_offset) - 1):
        b0, b1 = int(group_offset[g]), int(group_offset[g + 1])
        sl = slice(b0, b1)
        state = dict(
            data_array=vis[None, :, :, sl].astype(np.complex128),
            noise_array=np.zeros((1, U, P, b1 - b0, T, F0), complex),
            nsample_array=nsample[None, :, :, sl].astype(np.float64),
            flag_array=flags[None, :, :, sl].astype(bool),
            freq_array=freq_hz, trcvr=trcvr, beam_area=beam, beam_sq_area=beam,
        )
        state = orc.select_spectral_windows(state, spectral_windows)
        it = int_time[sl].astype(np.float64)
        sigma = orc.calculate_noise_power(state["freq_array"], state["trcvr"], state["beam_area"],
                                          it, state["nsample_array"])
        if noise_freq is not None:
            state["noise_array"] = noise_freq[:, :, :, sl].astype(np.complex128)
        res = orc.calculate_delay_spectrum(
            state["data_array"], state["noise_array"], state["flag_array"], state["nsample_array"],
            state["freq_array"], state["trcvr"], state["beam_area"], it, polarization_array, taper,
        )
        n = b1 - b0
        for key, arr in (("power", res["power_array"]), ("noise_power", res["noise_power"])):
            out[key + "_all"].append(orc.average_power(arr, remove_autos=False))
            out[key + "_noautos"].append(
                orc.average_power(arr, remove_autos=True) if n > 1 else np.full(arr.shape[:2] + arr.shape[-1:], np.nan)
            )
        th = res["thermal_power"]  # (S,P,B,B,T)
        out["thermal_all"].append(th.reshape(th.shape[0], th.shape[1], -1).mean(-1))
        if n > 1:
            off = ~np.eye(n, dtype=bool)
            out["thermal_noautos"].append(th[:, :, off].reshape(th.shape[0], th.shape[1], -1).mean(-1))
        else:
            out["thermal_noautos"].append(np.full(th.shape[:2], np.nan))
        out["sigma"].append(sigma[:, 0])
    return out


ENGINE_PHILOX_ROUNDS = 7


def engine_noise_freq(sigma_coef32, nsample, int_time, win_start, nfreq, seed, time_offset=0,
                      ntime_total=None, bl_offset=0, nbl_total=None):
    """The frequency-domain noise the fused kernel realises (before mask and taper):
    sigma * sqrt(-ln u1) * exp(2 pi i u2), in float32 like the kernel (sds_engine_f32.cu).

    sigma_coef32 (S,P,N) float32; nsample (P,B,T,F0) float32; int_time (B,T) float32.
    A row (s,p,b,t) of N samples consumes 3N/8 Philox4x32-7 calls, 128-bit counter =
    (row_id low word, row_id high word, 3 tau + k, 0) (k = 0..2), row_id = ((s*P + p)*Bt + b)*Tt + t; thread tau
    owns channels tau + (N/8) r, r = 0..7, and cuts its 12 words W into eight 46-bit draws
    (23-bit radius field, 23-bit angle field): for the pair (r = 2m, 2m + 1), (w0, w1, w2) = W[3m..3m+2],
      even: rad = w0[22:0],   ang = w1[22:0]
      odd : rad = w2[22:0],   ang = w2[30:24] : w1[31:24] : w0[31:24]
    Each field becomes the float f = 1 + field 2^-23 in [1, 2); u1 = 2 - f, angle = 2 pi f."""
    P, B, T, _ = nsample.shape
    S, N = len(win_start), nfreq
    tpr = N // 8
    Tt = T if ntime_total is None else ntime_total
    Bt = B if nbl_total is None else nbl_total
    s_i, p_i, b_i, t_i, tau_i = np.meshgrid(np.arange(S), np.arange(P), np.arange(B), np.arange(T),
                                            np.arange(tpr), indexing="ij")
    row_id = ((s_i.astype(np.uint64) * np.uint64(P) + p_i.astype(np.uint64)) * np.uint64(Bt)
              + (b_i + bl_offset).astype(np.uint64)) * np.uint64(Tt) + (t_i + time_offset).astype(np.uint64)
    words = []
    for k in range(3):
        ctr = row_id
        words += list(philox4x32((ctr & _LO).astype(np.uint32), (ctr >> np.uint64(32)).astype(np.uint32),
                                 (3 * tau_i + k).astype(np.uint32), np.zeros(ctr.shape, np.uint32),
                                 np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF),
                                 rounds=ENGINE_PHILOX_ROUNDS))
    W = [w.astype(np.uint32) for w in words]  # 12 arrays (S,P,B,T,tpr)
    rad23 = np.zeros((S, P, B, T, N), np.uint32)
    ang23 = np.zeros((S, P, B, T, N), np.uint32)
    u = np.uint32
    for m in range(4):
        w0, w1, w2 = W[3 * m], W[3 * m + 1], W[3 * m + 2]
        ch_even = np.arange(tpr) + tpr * (2 * m)
        ch_odd = np.arange(tpr) + tpr * (2 * m + 1)
        rad23[..., ch_even] = w0 & u(0x7FFFFF)
        ang23[..., ch_even] = w1 & u(0x7FFFFF)
        rad23[..., ch_odd] = w2 & u(0x7FFFFF)
        ang23[..., ch_odd] = (((w2 >> u(24)) & u(0x7F)) << u(16)) | ((w1 >> u(24)) << u(8)) | (w0 >> u(24))
    f32 = np.float32
    u1 = (f32(2.0) - (f32(1.0) + rad23.astype(f32) * f32(2.0 ** -23))).astype(f32)  # exact
    w = (-np.log(u1.astype(np.float64))).astype(f32)
    ang = ((f32(1.0) + ang23.astype(f32) * f32(2.0 ** -23)) * f32(6.2831853071795865)).astype(f32)
    chans = np.asarray(win_start)[:, None] + np.arange(N)
    ns_w = np.stack([nsample[..., chans[s]] for s in range(S)])  # (S,P,B,T,N)
    tn = int_time[None, None, :, :, None].astype(f32) * ns_w.astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        amp = sigma_coef32[:, :, None, None, :].astype(f32) * np.sqrt(w) / np.sqrt(tn)
    amp = np.where(tn > 0, amp, f32(0)).astype(f32)
    return (amp * np.cos(ang.astype(np.float64)) + 1j * amp * np.sin(ang.astype(np.float64))).astype(np.complex64)


def engine_noise_freq_mm(sigma_coef32, nsample, int_time, win_start, nfreq, seed, time_offset=0,
                         ntime_total=None, bl_offset=0, nbl_total=None, nuv=1, uv=0):
    """The frequency-domain noise of the power-of-two engine's production draw (noise_mode 1,
    csrc/sds_engine.cuh ``engine_word_mm`` / ``engine_signs_mm``), before mask and taper:
    ``sigma * B * (s1 + i s2)`` with B in {0, 1}, s1, s2 in {-1, +1} -- zero mean, circular,
    E|n|^2 = sigma^2 and E|n|^4 = 2 sigma^4 like the complex Gaussian of utils.py:501-505.

    One Philox4x32-7 call serves thread tau (channels tau + (N/8) r, r = 0..7) of four consecutive
    baselines: counter = (blk low word, blk high word, tau, 2),
    blk = (((s*P + p) * nuv + uv) * ceil(Bt / 4) + (b >> 2)) * Tt + t with global b, t (uv = index of
    the data set, nuv = 1 or 2); baseline b takes word b & 3; bit r of the word is B, bit 8 + r the
    sign of the real and bit 16 + r of the imaginary part."""
    P, B, T, _ = nsample.shape
    S, N = len(win_start), nfreq
    tpr = N // 8
    Tt = T if ntime_total is None else ntime_total
    Bt = B if nbl_total is None else nbl_total
    nblk = (Bt + 3) // 4
    s_i, p_i, b_i, t_i, tau_i = np.meshgrid(np.arange(S), np.arange(P), np.arange(B), np.arange(T),
                                            np.arange(tpr), indexing="ij")
    bg = (b_i + bl_offset).astype(np.uint64)
    blk = (((s_i.astype(np.uint64) * np.uint64(P) + p_i.astype(np.uint64)) * np.uint64(nuv) + np.uint64(uv))
           * np.uint64(nblk) + (bg >> np.uint64(2))) * np.uint64(Tt) + (t_i + time_offset).astype(np.uint64)
    W = philox4x32((blk & _LO).astype(np.uint32), (blk >> np.uint64(32)).astype(np.uint32),
                   tau_i.astype(np.uint32), np.full(blk.shape, 2, np.uint32),
                   np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF),
                   rounds=ENGINE_PHILOX_ROUNDS)
    sel = (bg & np.uint64(3)).astype(np.int64)
    word = np.choose(sel, [w.astype(np.uint32) for w in W])  # (S,P,B,T,tpr)
    unit = np.zeros((S, P, B, T, N), np.complex128)
    for r in range(8):
        on = ((word >> np.uint32(r)) & np.uint32(1)).astype(np.float64)
        s_re = 1.0 - 2.0 * ((word >> np.uint32(8 + r)) & np.uint32(1)).astype(np.float64)
        s_im = 1.0 - 2.0 * ((word >> np.uint32(16 + r)) & np.uint32(1)).astype(np.float64)
        unit[..., np.arange(tpr) + tpr * r] = on * (s_re + 1j * s_im)
    chans = np.asarray(win_start)[:, None] + np.arange(N)
    ns_w = np.stack([nsample[..., chans[s]] for s in range(S)]).astype(np.float64)  # (S,P,B,T,N)
    tn = int_time[None, None, :, :, None].astype(np.float64) * ns_w
    with np.errstate(divide="ignore", invalid="ignore"):
        amp = sigma_coef32[:, :, None, None, :].astype(np.float64) / np.sqrt(tn)
    amp = np.where(tn > 0, amp, 0.0)
    return (amp * unit).astype(np.complex64)


def engine_noise_freq_chirpz(sigma_coef32, nsample, int_time, win_start, nfreq, seed, time_offset=0,
                             ntime_total=None, bl_offset=0, nbl_total=None):
    """The noise of the chirp-z engine (window lengths that are not powers of two,
    sds_engine_blue.cu): convolution length M = smallest power of two >= max(64, 2N - 1), row
    team of TPR = M / 8 threads; thread tau draws channels tau + TPR r (r < 4) from two
    Philox4x32-7 calls, counter = (row_id low word, row_id high word, 2 tau + k, 0): the low 23
    bits of word 2r are the radius field, those of word 2r + 1 the angle field (f = 1 + field 2^-23,
    u1 = 2 - f, angle = 2 pi f, as in engine_noise_freq)."""
    P, B, T, _ = nsample.shape
    S, N = len(win_start), nfreq
    M = 64
    while M < 2 * N - 1:
        M *= 2
    tpr = M // 8
    Tt = T if ntime_total is None else ntime_total
    Bt = B if nbl_total is None else nbl_total
    s_i, p_i, b_i, t_i, tau_i = np.meshgrid(np.arange(S), np.arange(P), np.arange(B), np.arange(T),
                                            np.arange(tpr), indexing="ij")
    row_id = ((s_i.astype(np.uint64) * np.uint64(P) + p_i.astype(np.uint64)) * np.uint64(Bt)
              + (b_i + bl_offset).astype(np.uint64)) * np.uint64(Tt) + (t_i + time_offset).astype(np.uint64)
    W = []
    for k in range(2):
        ctr = row_id
        W += list(philox4x32((ctr & _LO).astype(np.uint32), (ctr >> np.uint64(32)).astype(np.uint32),
                             (2 * tau_i + k).astype(np.uint32), np.zeros(ctr.shape, np.uint32),
                             np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF),
                             rounds=ENGINE_PHILOX_ROUNDS))
    rad = np.zeros((S, P, B, T, 4 * tpr), np.uint32)
    ang = np.zeros((S, P, B, T, 4 * tpr), np.uint32)
    for r in range(4):
        ch = np.arange(tpr) + tpr * r
        rad[..., ch] = W[2 * r]
        ang[..., ch] = W[2 * r + 1]
    rad, ang = rad[..., :N], ang[..., :N]
    f32 = np.float32
    m23 = np.uint32(0x7FFFFF)
    u1 = (f32(2.0) - (f32(1.0) + (rad & m23).astype(f32) * f32(2.0 ** -23))).astype(f32)  # exact
    w = (-np.log(u1.astype(np.float64))).astype(f32)
    a = ((f32(1.0) + (ang & m23).astype(f32) * f32(2.0 ** -23)) * f32(6.2831853071795865)).astype(f32).astype(np.float64)
    chans = np.asarray(win_start)[:, None] + np.arange(N)
    ns_w = np.stack([nsample[..., chans[s]] for s in range(S)])
    tn = int_time[None, None, :, :, None].astype(f32) * ns_w.astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        amp = sigma_coef32[:, :, None, None, :].astype(f32) * np.sqrt(w) / np.sqrt(tn)
    amp = np.where(tn > 0, amp, f32(0)).astype(f32)
    return (amp * np.cos(a) + 1j * amp * np.sin(a)).astype(np.complex64)
