/*
 * sds.h -- C ABI of libsds.so: the B200 (sm_100a) implementation of the simpleDS
 * delay-spectrum hot path.
 *
 * The reference (rasg-affiliates/simpleDS) is pure Python/numpy and has no FFI
 * seam; the seam is its Python API.  Every entry point below names the
 * reference statements it replaces ("ds.py" = simpleDS/delay_spectrum.py,
 * "utils.py" = simpleDS/utils.py).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer owned by the caller unless the
 *     parameter is documented as "host";
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as
 *     void*); nothing is allocated except inside sds_plan_create;
 *   - return value: 0 ok, <0 invalid argument / unsupported, >0 a cudaError_t;
 *     sds_last_error() returns a thread-local message for the last failure;
 *   - complex numbers are interleaved (re, im); arrays are C-ordered.
 */
#ifndef SDS_H
#define SDS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDS_VERSION 100
#define SDS_MAX_WINDOWS 64
#define SDS_MAX_NFREQ 4096

enum sds_dtype { SDS_C64 = 0, SDS_C128 = 1, SDS_F32 = 2, SDS_F64 = 3 };
enum sds_status { SDS_OK = 0, SDS_EINVAL = -1, SDS_ENOTSUP = -2, SDS_ENOMEM = -3 };

/* pair-selection mode of the averaged cross power (utils.py:281-331 removes the
 * diagonal; the tutorial's plain mean keeps it) */
enum sds_pair_mode { SDS_PAIRS_ALL = 0, SDS_PAIRS_NO_AUTOS = 1 };
/* how the pair sum is evaluated: factorised |sum x|^2 identities, or the
 * explicit warp-shuffle pair loop */
enum sds_pair_method { SDS_METHOD_FACTORISED = 0, SDS_METHOD_EXPLICIT = 1 };

typedef struct sds_plan sds_plan;
typedef void *sds_stream; /* cudaStream_t */

const char *sds_last_error(void);
int sds_version(void);

/* Plan for delay transforms of length `nfreq` (any 1 <= nfreq <= SDS_MAX_NFREQ;
 * power-of-two lengths 64..2048 take the fast path).  `math` = SDS_F32 or
 * SDS_F64 selects the arithmetic; `taper_host` is taper(nfreq) evaluated by the
 * caller (host, fp64) -- replaces the window broadcast at utils.py:548-552 and
 * the taper argument of ds.py:3504-3517.  Uploads twiddles and the taper. */
int sds_plan_create(sds_plan **plan, int nfreq, int math, const double *taper_host);
int sds_plan_destroy(sds_plan *plan);
int sds_plan_nfreq(const sds_plan *plan);
int sds_plan_math(const sds_plan *plan);

/* Spectral-window gather, ds.py:3315-3338 (_take_spectral_windows_from_data_like_array):
 * in  (nspw_in,  nrows, nfreq_in), out (nspw_out, nrows, nfreq_out),
 * out[s, r, f] = in[c / nfreq_in, r, c % nfreq_in], c = chans[s * nfreq_out + f]
 * (`chans`: device int32, the flattened freq_chans of ds.py:3264-3270).
 * elem_bytes in {1, 4, 8, 16}. */
int sds_take_windows(const void *in, void *out, int elem_bytes, int64_t nrows,
                     int nspw_in, int nfreq_in, const int32_t *chans, int nspw_out,
                     int nfreq_out, sds_stream stream);

/* Delay transform, utils.py:509-566 + the flag mask of ds.py:3503-3517.
 * forward: out[w, r, k'] = delta_x * sum_n in[r, win_start[w] + n] * (1 - flag) *
 *          taper[n] * exp(-2 pi i (k' - N/2) n / N)         (fft, fftshift, * dx)
 * inverse: ifftshift(ifft(in * (1 - flag))) / taper * delta_x  (utils.py:561-564,
 *          including its quirks).
 * in: `in_dtype` SDS_C64|SDS_C128, rows `in_row_stride` elements apart;
 * flags: uint8 (numpy bool), nullable, rows `flag_row_stride` elements apart;
 * win_start: HOST int32[nwin], nwin <= SDS_MAX_WINDOWS;
 * out: [nwin][nrows][N], SDS_C64 for an SDS_F32 plan, SDS_C128 for SDS_F64. */
int sds_delay_transform(const sds_plan *plan, const void *in, int in_dtype,
                        int64_t in_row_stride, const uint8_t *flags,
                        int64_t flag_row_stride, int64_t nrows, int nwin,
                        const int32_t *win_start, double delta_x, int inverse,
                        void *out, int out_dtype, sds_stream stream);

/* Radiometer sigma in Jy, ds.py:3544-3583 (+ utils.py:467-483):
 * sigma[s,u,p,b,t,f] = coef[s,p,f] / sqrt(int_time[b,t] * nsample[s,u,p,b,t,f]),
 * non-finite -> 0.  `coef` (device fp64 [S][P][F]) = 1e3 * Tsys * beam_area /
 * (sqrt(delta_f) * jy_to_mk) is the per-channel part evaluated by the host.
 * nsample dtype SDS_F32|SDS_F64; sigma_out fp64. */
int sds_noise_sigma(const double *coef, const double *int_time, const void *nsample,
                    int nsample_dtype, int S, int U, int P, int B, int T, int F,
                    double *sigma_out, sds_stream stream);

/* Complex Gaussian realisation, utils.py:486-506: out = sigma * (g1 + i g2) / sqrt(2).
 * g_re/g_im (device fp64[n]) inject pre-drawn normals (parity mode); when both
 * are NULL the normals come from Philox4x32-10 keyed by `seed`, counter =
 * (`offset` + flat index) / 2, Box-Muller.  out dtype SDS_C64|SDS_C128. */
int sds_generate_noise(const double *sigma, int64_t n, const double *g_re,
                       const double *g_im, uint64_t seed, uint64_t offset, void *out,
                       int out_dtype, sds_stream stream);

/* The noise leg of the unfused API in one call: sds_noise_sigma -> sds_generate_noise ->
 * sds_delay_transform (forward, flag mask applied) enqueued back to back -- ds.py:3544-3583,
 * utils.py:486-506 and the noise half of ds.py:3503-3517.  Arrays are the 6-D
 * (S,U,P,B,T,F) layout with F == the plan's nfreq; `sigma_scratch` fp64 and `noise_freq`
 * (the frequency-domain realisation the reference keeps as noise_array) are caller-provided,
 * `noise_delay` receives the transformed noise; `dtype` SDS_C64 (SDS_F32 plan) or SDS_C128
 * (SDS_F64 plan) for both complex arrays.  g_re / g_im / seed / offset as in sds_generate_noise. */
int sds_noise_transform(const sds_plan *plan, const double *coef, const double *int_time,
                        const void *nsample, int nsample_dtype, int S, int U, int P, int B, int T,
                        int F, const uint8_t *flags, double delta_f, uint64_t seed, uint64_t offset,
                        const double *g_re, const double *g_im, double *sigma_scratch,
                        void *noise_freq, void *noise_delay, int dtype, sds_stream stream);

/* Materialised cross power, utils.py:387-389 as called at ds.py:3620-3633:
 * out[s,p,i,j,t,d] = a2[s,p,i,t,d] * conj(a1[s,p,j,t,d]).  a1/a2 have element
 * strides (stride_s, stride_p) over their first two axes and are contiguous in
 * (B,T,D); dtype SDS_C64|SDS_C128 for inputs and output alike.
 * scale_sp: NULL, or fp64 [S][P]: out is multiplied by scale_sp[s][p] as it is written (the
 * cosmological unit conversion of delay_spectrum.py:1306-1311, which the reference applies in a
 * second pass over the tensor). */
int sds_cross_power_full(const void *a1, const void *a2, int dtype, int S, int P,
                         int B, int T, int D, int64_t stride_s, int64_t stride_p,
                         const double *scale_sp, void *out, sds_stream stream);

/* Materialised thermal estimate, ds.py:3637-3707 + utils.py:230-278.
 * n1 = nsample of the conj ("j") side, n2 = of the "i" side (same pointer when
 * Nuv == 1), strides as above, contiguous in (B,T,F).
 * coef   fp64 [S][P][F] = (1e3 * Tsys * beam_area)^2 * taper^2 * nu^4  (host-made)
 * freq   fp64 [S][F] in Hz (trapezoid abscissa), int_time fp64 [B][T] (j side),
 * pol_factor fp64 [P] = 1 / (npols_noise * sqrt(2)).
 * out fp64 [S][P][B][B][T] in K^2 sr^2 Hz^6; invalid samples are masked out of
 * the trapezoid (their panels contribute 0).  scale_sp: NULL, or fp64 [S][P] multiplied in as the
 * tensor is written (the thermal conversion of delay_spectrum.py:1339-1343). */
int sds_thermal_power_full(const void *n1, const void *n2, int nsample_dtype,
                           int64_t stride_s, int64_t stride_p, const double *coef,
                           const double *freq, const double *int_time,
                           const double *pol_factor, int S, int P, int B, int T, int F,
                           const double *scale_sp, double *out, sds_stream stream);

/* Pair sums of the thermal estimate per (group, spw, pol, time), same inputs as
 * sds_thermal_power_full plus the group table of sds_cross_power_avg:
 *   out_all [g,s,p,t] = sum_{i,j in g} thermal[s,p,i,j,t]
 *   out_auto[g,s,p,t] = sum_{i in g}   thermal[s,p,i,i,t]      (fp64) */
int sds_thermal_power_avg(const void *n1, const void *n2, int nsample_dtype,
                          int64_t stride_s, int64_t stride_p, const double *coef,
                          const double *freq, const double *int_time,
                          const double *pol_factor, const int32_t *group_offset, int ngroup,
                          int S, int P, int B, int T, int F, double *out_all,
                          double *out_auto, sds_stream stream);

/* Averaged cross power (cross multiply + the user-side mean of the tutorial,
 * utils.py:387-389 + utils.py:281-331), never materialising the B x B tensor.
 * Baselines [group_offset[g], group_offset[g+1]) form redundant group g
 * (device int32[ngroup+1]).  For every (g,s,p,t,d):
 *   out_all [g,s,p,t,d] = sum_{i,j in g}      a2[i] conj(a1[j])
 *   out_auto[g,s,p,t,d] = sum_{i in g}        a2[i] conj(a1[i])
 * (complex128, interleaved).  Means follow on the host as out_all / n^2 or
 * (out_all - out_auto) / (n (n - 1)).  method = sds_pair_method. */
int sds_cross_power_avg(const void *a1, const void *a2, int dtype, int S, int P, int B,
                        int T, int D, int64_t stride_s, int64_t stride_p,
                        const int32_t *group_offset, int ngroup, int method,
                        double *out_all, double *out_auto, sds_stream stream);

/* ------------------------------------------------------------------------
 * (Window lengths: powers of two 64..2048 in either arithmetic; any other length 8..1024 with
 * an SDS_F32 plan through the chirp-z engine.  Rows and window starts must be multiples of 16
 * samples and a window must leave ceil16(nfreq) samples before the end of its row; otherwise
 * SDS_ENOTSUP and the caller chains the kernels above.)
 * Fused engine: window select + flag mask + taper + delay transform + noise
 * realisation + its transform + pair sums + thermal sums in ONE pass over the
 * 13 B/sample inputs (ds.py:3196-3338, 3487-3517, 3544-3583, 3585-3707 and
 * utils.py:387-389, 486-506, 509-560 fused; the B x B tensors of ds.py:399-450
 * are never formed).
 * ------------------------------------------------------------------------ */
typedef struct sds_engine_desc {
  int32_t npol;        /* polarisations in the batch                            */
  int32_t nbl;         /* baselines in the batch (all groups, group-sorted)     */
  int32_t ntime;       /* times in the batch                                    */
  int32_t nchan;       /* channels per input row (row pitch, samples)           */
  int32_t nspw;        /* spectral windows, each plan->nfreq channels           */
  int32_t ngroup;      /* redundant groups                                      */
  int32_t nuv;         /* 1: auto (one data set); 2: cross of two data sets     */
  int32_t noise_mode;  /* 0 none, 1 in-kernel draw (below), 2 injected (noise_in) */
  int32_t win_start[SDS_MAX_WINDOWS];
  int64_t time_offset; /* global index of the batch's first time (RNG counter)  */
  int64_t ntime_total; /* global number of times             (RNG counter)      */
  int64_t bl_offset;   /* global index of the batch's first baseline            */
  int64_t nbl_total;   /* global number of baselines                            */
  uint64_t seed;
  double delta_f;      /* Hz, scales the transform (utils.py:560)               */
  /* A sub-list of the redundant groups (sharding by group, delay_spectrum.py:891-900: every group
   * is its own object): the arrays hold only the baselines of `ngroup` chosen groups.  Device
   * pointers, both NULL for the whole list:
   *   group_index[g]    row of the output arrays group g accumulates into,
   *   group_bl_start[g] index of the group's first baseline in the whole array (RNG counter). */
  const int32_t *group_index;
  const int64_t *group_bl_start;
} sds_engine_desc;

/* In-kernel noise, replacing utils.py:486-506 + ds.py:3544-3583 (the reference draws
 * sigma (g1 + i g2) / sqrt(2) per sample from numpy's global generator; parity is statistical).
 * Counter-based: Philox4x32-7 words (Salmon et al. 2011) keyed by `seed`, the 128-bit counter a
 * GLOBAL sample coordinate, so the realisation does not depend on batching or sharding (exact
 * word-to-channel maps: oracle/engine_ref.py).
 *   noise_mode 1 (production; power-of-two windows): the moment-matched discrete law
 *     n = sigma * B * (s1 + i s2), B in {0, 1} (p = 1/2), s1, s2 = +-1: E n = 0, E n^2 = 0,
 *     E |n|^2 = sigma^2, E |n|^4 = 2 sigma^4 as for the Gaussian; three bits per sample, one call per
 *     thread and block of four baselines, counter = (blk low word, blk high word, tau, 2),
 *     blk = ((s npol + p) ceil(nbl_total / 4) + (b >> 2)) ntime_total + t with global b, t.  The
 *     engine only ever returns DELAY-domain pair sums, i.e. sums over the >= 64 tapered channels of a
 *     row, whose law the per-channel moments fix (csrc/sds_engine.cuh, tests/test_gpu_noise_stats.py).
 *     Windows that are not powers of two (chirp-z engine) draw Box-Muller in this mode too.
 *   noise_mode 3: sigma * sqrt(-ln u1) * exp(2 pi i u2) per sample (Box-Muller), u1 and u2 23-bit
 *     uniforms; counter (row_id low word, row_id high word, 3 tau + k, 0),
 *     row_id = ((s * npol + p) * nbl_total + bl_offset + b) * ntime_total + time_offset + t.
 *   noise_mode 2: injected frequency-domain noise (noise_in), for exact parity tests.
 *
 * Inputs: vis  c64 [nuv][npol][nbl][ntime][nchan]; flags u8, nsample f32 same
 * shape; int_time f32 [nbl][ntime]; group_offset int32[ngroup+1];
 * sigma_coef fp32 [nspw][npol][N] (as in sds_noise_sigma);
 * thermal_coef fp64 [nspw][npol][N] (as in sds_thermal_power_full, with the
 * trapezoid weights NOT folded in), freq fp64 [nspw][N];
 * noise_in c64 same shape as vis (noise_mode 2 only).
 * Outputs (fp64, accumulated INTO -- zero them first; sums over the batch's
 * times): pow_all/pow_auto/noise_all/noise_auto real [ngroup][nspw][npol][N] (with
 * nuv == 1 the pair sums sum_ij x_i conj(x_j) = |sum_i x_i|^2 are real),
 * thermal_all/thermal_auto real [ngroup][nspw][npol] (already including the
 * 1/(npol sqrt2) and 1e-6 factors and the per-baseline 1/t_int).
 * workspace: sds_engine_workspace_bytes() bytes of device scratch.
 * Device resources: besides registers and shared memory the power-of-two kernels of this call allocate
 * TENSOR MEMORY (tcgen05.alloc, 32-256 columns per CTA, freed before the CTA exits) for the exchanges
 * of the transform and for their accumulators; two such kernels on one SM wait for each other's
 * columns, they do not fail. */
int64_t sds_engine_workspace_bytes(const sds_plan *plan, const sds_engine_desc *desc);
int sds_engine_run(const sds_plan *plan, const sds_engine_desc *desc, const void *vis,
                   const uint8_t *flags, const float *nsample, const float *int_time,
                   const int32_t *group_offset, const float *sigma_coef,
                   const double *thermal_coef, const double *freq,
                   const double *pol_factor, const void *noise_in, double *pow_all,
                   double *pow_auto, double *noise_all, double *noise_auto,
                   double *thermal_all, double *thermal_auto, void *workspace,
                   sds_stream stream, int32_t *launches_out);
/* launches_out (host pointer, may be NULL): number of kernel launches the call enqueued.
 *
 * Two data sets (desc->nuv == 2; delay_spectrum.py:3620-3633: power[i, j] = A2_i conj(A1_j) with
 * A1 / A2 the transforms of data_array[:, 0] / [:, 1]): vis, flags, nsample and noise_in carry a
 * leading axis of length 2 ((2, Npols, Nbls, Ntimes, Nchan)); pow_* and noise_* are COMPLEX,
 * interleaved (re, im) doubles of twice the size:
 *   *_all  += (sum_i A2_i) conj(sum_j A1_j),    *_auto += sum_i A2_i conj(A1_i);
 * the in-kernel noise of set u is keyed by blk = (((s npol + p) 2 + u) ceil(nbl_total / 4) + (b >> 2))
 * ntime_total + t.  SDS_F32 plans with power-of-two windows of 64..1024 channels.  The thermal estimate
 * (delay_spectrum.py:3655-3707: N_ij = sqrt(n1_j n2_i), integration time of baseline j) rides in the
 * noise launch: thermal_coef != NULL needs noise_mode 1 or 2. */

/* Means over ordered baseline pairs and times from the sums sds_engine_run accumulates (tutorial cell 41:
 * power.mean over (i, j, t); utils.py:281-331 remove_auto_correlations for the i != j variant), the six
 * outputs in ONE launch:
 *   x_all     = all  / (n n T),
 *   x_noautos = (all - auto) / (n (n - 1) T)      (NaN for a group of one baseline: no cross pairs)
 * with n = group_size[g] and T = times_done[g] (the group's count of accumulated times, which travels
 * with the sums through a multi-GPU reduce; a group that saw no time gives NaN) or ntimes when > 0.
 * rows_per_group = nspw * npol; row_len = values per row of the pair sums (N, or 2 N for the complex
 * sums of two data sets).  Any input / output pair may be NULL together. */
int sds_engine_means(int32_t ngroup, int32_t rows_per_group, int32_t row_len, const int32_t *group_size,
                     const double *times_done, double ntimes, const double *pow_all,
                     const double *pow_auto, const double *noise_all, const double *noise_auto,
                     const double *thermal_all, const double *thermal_auto, double *power_all,
                     double *power_noautos, double *noise_power_all, double *noise_power_noautos,
                     double *thermal_mean_all, double *thermal_mean_noautos, sds_stream stream);

/* ---- downstream estimators on the (averaged) spectra: SURVEY 8(f3) ----------------------
 * All arrays are C-ordered and viewed as (outer, n, inner) around the reduced / folded /
 * resampled axis; values are fp64 (real) or complex128 (interleaved, dtype = SDS_F64 / SDS_C128). */

/* utils.py:569-645 weighted_average: out[o,i] = sum_k w x / sum_k w,
 * out_sigma[o,i] = sqrt(sum_k sigma^2 |w|^2 / |sum_k w|^2), real fp64.
 * weights: NULL -> inverse variance 1 / sigma^2 (utils.py:628-629); weights_1d != 0 -> one weight
 * per position of the averaged axis (n values), else the shape of x. */
int sds_weighted_average(const double *x, const double *sigma, const double *weights,
                         int weights_1d, int64_t outer, int64_t n, int64_t inner, double *out,
                         double *out_sigma, sds_stream stream);

/* utils.py:651-796 fold_along_delay: output bin j (0 <= j < nout) is the weighted average of the
 * input bins pos0 + j and neg0 - j along the folded axis (the host side derives pos0 / neg0 from
 * the delays exactly as utils.py:721-761 split the array); with zero_bin != 0 bin j = 0 is the
 * shared delay = 0 bin, whose uncertainty is scaled by sqrt(2) before the average
 * (utils.py:733-737).  dtype SDS_F64: real values, uncertainties and weights.  dtype SDS_C128:
 * real and imaginary parts are folded separately with the real / imaginary parts of the
 * weights (utils.py:778-794).  weights == NULL -> ones (real) / 1 + 1j (complex). */
int sds_fold_along_delay(const void *x, const void *sigma, const void *weights, int dtype,
                         int64_t outer, int64_t n, int64_t inner, int64_t pos0, int64_t neg0,
                         int64_t nout, int zero_bin, void *out, void *out_sigma,
                         sds_stream stream);

/* utils.py:171-209 bootstrap_array, the gather (np.take with an (n, nboot) index table):
 * out[o, i, b, :] = in[o, index[i * nboot + b], :], elem_bytes bytes per element (8 or 16). */
int sds_bootstrap_gather(const void *in, const int64_t *index, int elem_bytes, int64_t outer,
                         int64_t n, int64_t nboot, int64_t inner, void *out, sds_stream stream);
/* ... and the index table itself, index[i * nboot + b] uniform in [0, n): Philox4x32-10 keyed by
 * `seed`, counter = flat position / 4 (the reference draws from numpy's global MT19937,
 * utils.py:205-207; injected indices give bit-level parity, this gives a reproducible stream). */
int sds_bootstrap_indices(uint64_t seed, int64_t n, int64_t nboot, int64_t *index,
                          sds_stream stream);

/* ---- ingest: SURVEY 8(f2) -------------------------------------------------------------------
 * utils.py:15-168 (get_data_array / get_nsample_array / get_flag_array: one np.transpose per
 * baseline) with the ordering rules of delay_spectrum.py:913-973 (baselines by baseline number,
 * times by unwrapped LST) folded into ONE row permutation computed by the caller:
 *   out[p][r][f] = in[perm[r]][f][p],   in (nblt, nfreq, npol), out (npol, nrow, nfreq),
 * elem_bytes 1 (flags), 4 (nsample), 8 (complex64 visibilities) or 16.  Bit-exact (a copy). */
int sds_ingest_gather(const void *in, const int64_t *perm, int elem_bytes, int64_t nblt,
                      int64_t nrow, int nfreq, int npol, void *out, sds_stream stream);

/* ---- cosmological normalisation: SURVEY 8(f1) ---------------------------------------------
 * The flat-LambdaCDM background the reference gets from astropy (cosmo.py:20: Planck15), as plain
 * numbers: Ode0 = 1 - Om0 - Ogamma0 - Onu0 already evaluated by the caller. */
#define SDS_COSMO_MAX_NU 4
typedef struct sds_cosmo_params {
  double h0_km_s_mpc, om0, ode0, ogamma0;
  double neff_per_nu;       /* Neff / number of neutrino species                          */
  double nmassless_nu;      /* number of massless species                                 */
  double nu_y[SDS_COSMO_MAX_NU]; /* m_nu c^2 / (k_B T_nu0) of the massive species          */
  int32_t nmassive_nu;
  int32_t has_massive_nu_or_radiation; /* 0: Tcmb0 == 0, no radiation terms at all        */
} sds_cosmo_params;
/* E(z) (astropy FlatLambdaCDM.efunc, used at cosmo.py:131, 157, 183) and the line-of-sight
 * comoving distance in Mpc (= comoving_transverse_distance for a flat universe, cosmo.py:77,
 * 105, 182) at n redshifts; either output may be NULL. */
int sds_cosmo_evaluate(const sds_cosmo_params *params, const double *z, int64_t n,
                       double *efunc_out, double *comoving_distance_mpc_out, sds_stream stream);
/* delay_spectrum.py:1260-1268 / 1286-1297 / 1318-1330: out[s,p] = trapz over the window's
 * frequencies of  freq^freq_power / x2y * taper^2 * beam_sq_area,  freq and x2y (nspw, nfreq),
 * taper (nfreq), beam_sq_area (nspw, npol, nfreq); freq_power = 4 (Jy^2 Hz^2 power and the thermal
 * estimate) or 0 (K^2 sr^2 Hz^2 power). */
int sds_conversion_integral(const double *freq, const double *x2y, const double *taper,
                            const double *beam_sq_area, int nspw, int npol, int nfreq,
                            int freq_power, double *out, sds_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* SDS_H */
