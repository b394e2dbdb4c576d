// K1 fast path: forward delay transform of power-of-two lengths 64..2048 with the
// register-radix Stockham schedule of sds_radix.cuh / sds_packed.cuh.
//
// utils.py:548-560 + the mask multiply of ds.py:3503-3517 in one pass: every thread
// loads its 8 samples of a row straight from global memory (element tau + (N/8) r, so a
// warp reads contiguous 256 B / 512 B runs), multiplies by (1 - flag) * taper * delta_x,
// transforms in registers with shared-memory exchanges, and stores the spectrum with the
// fftshift as an index rotation (again contiguous per warp).  HBM-bound: 9 (17) B in +
// 8 (16) B out per sample for complex64 (complex128).
#include "sds_packed.cuh"

namespace sds {

template <typename T>
struct Pow2Params {
  const void *in;
  const uint8_t *flags;
  void *out;
  const T *taper;
  const cpx<T> *tw;
  int64_t in_row_stride, flag_row_stride, nrows, total;
  T delta_x;
  int in_dtype, out_dtype, nwin;
  int win_start[SDS_MAX_WINDOWS];
};

template <typename T, int N> struct Pow2Cfg {
  static constexpr int TPR = N / 8;
  static constexpr int TEAM = TPR < 32 ? 32 : TPR;
  static constexpr int ROWS = TEAM / TPR;
  static constexpr int THREADS = TEAM < 256 ? 256 : TEAM;
  static constexpr int NTEAM = THREADS / TEAM;
  static constexpr int TW_BYTES = ((fft_tw_total(N) * (int)sizeof(cpx<T>) + 127) / 128) * 128;
  static constexpr int SMEM = TW_BYTES + NTEAM * ROWS * 2 * N * (int)sizeof(cpx<T>);
};

template <typename T, int N>
__global__ void __launch_bounds__(Pow2Cfg<T, N>::THREADS)
delay_fft_pow2_kernel(const __grid_constant__ Pow2Params<T> P) {
  using C = Pow2Cfg<T, N>;
  constexpr int TPR = C::TPR, TEAM = C::TEAM, ROWS = C::ROWS;
  extern __shared__ __align__(128) unsigned char smem[];
  cpx<T> *tw_s = reinterpret_cast<cpx<T> *>(smem);
  for (int i = threadIdx.x; i < fft_tw_total(N); i += C::THREADS) tw_s[i] = P.tw[i];
  const int team = threadIdx.x / TEAM, tl = threadIdx.x % TEAM;
  const int sub = tl / TPR, tau = tl % TPR, bar_id = 1 + team;
  cpx<T> *bufA = reinterpret_cast<cpx<T> *>(smem + C::TW_BYTES) + (size_t)((team * ROWS + sub) * 2) * N;
  cpx<T> *bufB = bufA + N;
  __syncthreads();

  T tapdx[8];  // taper * delta_x at this thread's channels (utils.py:552, 560)
#pragma unroll
  for (int r = 0; r < 8; ++r) tapdx[r] = P.taper[tau + TPR * r] * P.delta_x;

  const int64_t nstep = (P.total + ROWS - 1) / ROWS;  // team steps of ROWS rows
  const int64_t stride = (int64_t)gridDim.x * C::NTEAM;
  // software pipeline: the loads of the team's NEXT row are in flight while the current one
  // is transformed (one row per team at a time would leave the kernel latency-bound)
  cpx<T> nxt[8];
  uint32_t nxt_flag = 0;
  auto fetch = [&](int64_t stp) {
    const int64_t rw = stp * ROWS + sub;
    const bool active = stp < nstep && rw < P.total;
    const int w = active ? (int)(rw / P.nrows) : 0;
    const int64_t row = active ? rw % P.nrows : 0;
    const int64_t ibase = row * P.in_row_stride + P.win_start[w];
    const int64_t fbase = row * P.flag_row_stride + P.win_start[w];
    nxt_flag = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      nxt[r] = active ? load_cpx<T>(P.in, P.in_dtype, ibase + tau + TPR * r) : cpx<T>{(T)0, (T)0};
      const bool f = !active || (P.flags && P.flags[fbase + tau + TPR * r]);
      nxt_flag |= (f ? 1u : 0u) << r;
    }
  };
  int64_t stp = (int64_t)blockIdx.x * C::NTEAM + team;
  fetch(stp);
  for (; stp < nstep; stp += stride) {
    const int64_t rw = stp * ROWS + sub;
    const bool active = rw < P.total;
    const int w = active ? (int)(rw / P.nrows) : 0;
    const int64_t row = active ? rw % P.nrows : 0;
    cpx<T> v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      // data * (~flag) * taper * delta_x   (ds.py:3503-3505, utils.py:558-560)
      const T m = (nxt_flag >> r) & 1u ? (T)0 : tapdx[r];
      v[r] = {nxt[r].x * m, nxt[r].y * m};
    }
    fetch(stp + stride);
    if constexpr (sizeof(T) == sizeof(float)) {
      f2 p[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) p[r] = f2_make(v[r].x, v[r].y);
      TeamFftP<N>::run(p, reinterpret_cast<f2 *>(bufA), reinterpret_cast<f2 *>(bufB),
                       reinterpret_cast<const f2 *>(tw_s), tau, bar_id);
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = {f2_lo(p[r]), f2_hi(p[r])};
    } else {
      TeamFft<T, N>::run(v, bufA, bufB, tw_s, tau, bar_id);
    }
    if (active) {
      const int64_t obase = ((int64_t)w * P.nrows + row) * N;
#pragma unroll
      for (int r = 0; r < 8; ++r)  // fftshift (utils.py:559): bin k lands at (k + N/2) mod N
        store_cpx<T>(P.out, P.out_dtype, obase + ((tau + TPR * r + N / 2) % N), v[r]);
    }
    team_sync<TEAM>(bar_id);  // the exchange buffers are reused by the next row
  }
}

template <typename T, int N>
static int launch_pow2_n(const Pow2Params<T> &P, cudaStream_t st) {
  using C = Pow2Cfg<T, N>;
  static DeviceLatch configured;
  int cfg_dev = 0;
  if (configured.need(&cfg_dev)) {
    SDS_CUDA(cudaFuncSetAttribute(delay_fft_pow2_kernel<T, N>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    SDS_CUDA(cudaFuncSetAttribute(delay_fft_pow2_kernel<T, N>,
                                  cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    configured.set(cfg_dev);
  }
  int dev = 0, nsm = 148, per_sm = 1;
  SDS_CUDA(cudaGetDevice(&dev));
  SDS_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  SDS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, delay_fft_pow2_kernel<T, N>,
                                                         C::THREADS, C::SMEM));
  if (per_sm < 1) return SDS_ENOTSUP;
  const int64_t nstep = (P.total + C::ROWS - 1) / C::ROWS;
  int64_t blocks = (nstep + C::NTEAM - 1) / C::NTEAM;
  if (blocks > (int64_t)nsm * per_sm) blocks = (int64_t)nsm * per_sm;  // persistent, grid-stride
  delay_fft_pow2_kernel<T, N><<<(unsigned)blocks, C::THREADS, C::SMEM, st>>>(P);
  SDS_LAUNCH_CHECK("delay_fft_pow2_kernel");
  return SDS_OK;
}

template <typename T>
static int launch_pow2_t(int N, const Pow2Params<T> &P, cudaStream_t st) {
  switch (N) {
    case 64: return launch_pow2_n<T, 64>(P, st);
    case 128: return launch_pow2_n<T, 128>(P, st);
    case 256: return launch_pow2_n<T, 256>(P, st);
    case 512: return launch_pow2_n<T, 512>(P, st);
    case 1024: return launch_pow2_n<T, 1024>(P, st);
    case 2048: return launch_pow2_n<T, 2048>(P, st);
    default: return SDS_ENOTSUP;
  }
}

int launch_pow2(const Plan &pl, const void *in, int in_dtype, int64_t in_row_stride,
                const uint8_t *flags, int64_t flag_row_stride, int64_t nrows, int nwin,
                const int32_t *win_start, double delta_x, void *out, int out_dtype,
                cudaStream_t st) {
  if (!pl.is_pow2 || pl.nfreq < 64 || pl.nfreq > 2048 || !pl.team_tw_f || !pl.team_tw_d)
    return SDS_ENOTSUP;
  auto fill = [&](auto &P) {
    P.in = in; P.flags = flags; P.out = out;
    P.in_row_stride = in_row_stride; P.flag_row_stride = flag_row_stride;
    P.nrows = nrows; P.total = nrows * nwin;
    P.in_dtype = in_dtype; P.out_dtype = out_dtype; P.nwin = nwin;
    for (int i = 0; i < SDS_MAX_WINDOWS; ++i) P.win_start[i] = i < nwin ? win_start[i] : 0;
  };
  if (pl.math == SDS_F64) {
    Pow2Params<double> P;
    fill(P);
    P.taper = pl.taper_d; P.tw = pl.team_tw_d; P.delta_x = delta_x;
    return launch_pow2_t<double>(pl.nfreq, P, st);
  }
  Pow2Params<float> P;
  fill(P);
  P.taper = pl.taper_f; P.tw = pl.team_tw_f; P.delta_x = (float)delta_x;
  return launch_pow2_t<float>(pl.nfreq, P, st);
}

}  // namespace sds
