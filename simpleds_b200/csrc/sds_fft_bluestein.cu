// K1 for lengths that are not powers of two (21, 203 = 7 x 29 channels of the reference's
// bundled files, ...): Bluestein's chirp-z form on top of the register-radix power-of-two
// transform.  With c[n] = exp(-i pi n^2 / N),
//     X[k] = c[k] * sum_n (x[n] c[n]) conj(c)[k - n]
// is a linear convolution, evaluated as a circular one of length M >= 2N - 1 (M a power of
// two): two length-M transforms per row (the inverse one as conj(FFT(conj(.))) / M) and
// three pointwise multiplies, all in registers; the chirp and the transform of the wrapped
// conj(c) come from the plan (computed in long double).  Same fusion as the other K1 kernels:
// flag mask, taper, delta_x and the fftshift index rotation (utils.py:548-560,
// ds.py:3503-3517).
#include "sds_packed.cuh"

namespace sds {

template <typename T>
struct BlueParams {
  const void *in;
  const uint8_t *flags;
  void *out;
  const T *taper;
  const cpx<T> *chirp;  // [N]
  const cpx<T> *bhat;   // [M]
  const cpx<T> *tw;     // team twiddles of length M
  int64_t in_row_stride, flag_row_stride, nrows, total;
  T delta_x;
  int in_dtype, out_dtype, nwin, N;
  int win_start[SDS_MAX_WINDOWS];
};

template <typename T, int M> struct BlueCfg {
  static constexpr int TPR = M / 8;
  static constexpr int TEAM = TPR < 32 ? 32 : TPR;
  static constexpr int ROWS = TEAM / TPR;
  static constexpr int THREADS = TEAM < 256 ? 256 : TEAM;
  static constexpr int NTEAM = THREADS / TEAM;
  static constexpr int TW_BYTES = ((fft_tw_total(M) * (int)sizeof(cpx<T>) + 127) / 128) * 128;
  static constexpr int SMEM = TW_BYTES + NTEAM * ROWS * 2 * M * (int)sizeof(cpx<T>);
};

template <typename T>
__device__ __forceinline__ cpx<T> cconj(cpx<T> a) { return {a.x, -a.y}; }

// one length-M transform of the team's row (packed fp32x2 for float)
template <typename T, int M>
__device__ __forceinline__ void team_fft(cpx<T> (&v)[8], cpx<T> *bufA, cpx<T> *bufB,
                                         const cpx<T> *tw_s, int tau, int bar_id) {
  if constexpr (sizeof(T) == sizeof(float)) {
    f2 p[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) p[r] = f2_make(v[r].x, v[r].y);
    TeamFftP<M>::run(p, reinterpret_cast<f2 *>(bufA), reinterpret_cast<f2 *>(bufB),
                     reinterpret_cast<const f2 *>(tw_s), tau, bar_id);
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = {f2_lo(p[r]), f2_hi(p[r])};
  } else {
    TeamFft<T, M>::run(v, bufA, bufB, tw_s, tau, bar_id);
  }
}

template <typename T, int M>
__global__ void __launch_bounds__(BlueCfg<T, M>::THREADS)
delay_fft_bluestein_kernel(const __grid_constant__ BlueParams<T> P) {
  using C = BlueCfg<T, M>;
  constexpr int TPR = C::TPR, TEAM = C::TEAM, ROWS = C::ROWS;
  extern __shared__ __align__(128) unsigned char smem[];
  cpx<T> *tw_s = reinterpret_cast<cpx<T> *>(smem);
  for (int i = threadIdx.x; i < fft_tw_total(M); i += C::THREADS) tw_s[i] = P.tw[i];
  const int team = threadIdx.x / TEAM, tl = threadIdx.x % TEAM;
  const int sub = tl / TPR, tau = tl % TPR, bar_id = 1 + team;
  cpx<T> *bufA = reinterpret_cast<cpx<T> *>(smem + C::TW_BYTES) + (size_t)((team * ROWS + sub) * 2) * M;
  cpx<T> *bufB = bufA + M;
  __syncthreads();

  const int N = P.N;
  // per-thread constants at its 8 positions n = k = tau + TPR r of the padded sequence
  cpx<T> cin[8], cout[8], bh[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int n = tau + TPR * r;
    const cpx<T> c = n < N ? P.chirp[n] : cpx<T>{(T)0, (T)0};
    const T a = n < N ? P.taper[n] * P.delta_x : (T)0;  // taper * delta_x (utils.py:552, 560)
    cin[r] = {c.x * a, c.y * a};
    cout[r] = {c.x / (T)M, c.y / (T)M};                  // output chirp and the 1/M of the inverse
    bh[r] = P.bhat[n];
  }

  const int64_t nstep = (P.total + ROWS - 1) / ROWS;
  for (int64_t stp = (int64_t)blockIdx.x * C::NTEAM + team; stp < nstep;
       stp += (int64_t)gridDim.x * C::NTEAM) {
    const int64_t rw = stp * ROWS + sub;
    const bool active = rw < P.total;
    const int w = active ? (int)(rw / P.nrows) : 0;
    const int64_t row = active ? rw % P.nrows : 0;
    const int64_t ibase = row * P.in_row_stride + P.win_start[w];
    const int64_t fbase = row * P.flag_row_stride + P.win_start[w];
    cpx<T> v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int n = tau + TPR * r;
      v[r] = {(T)0, (T)0};
      if (active && n < N && !(P.flags && P.flags[fbase + n]))  // data * (~flag)  ds.py:3503-3505
        v[r] = cmul(load_cpx<T>(P.in, P.in_dtype, ibase + n), cin[r]);
    }
    team_fft<T, M>(v, bufA, bufB, tw_s, tau, bar_id);
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = cconj(cmul(v[r], bh[r]));
    team_sync<TEAM>(bar_id);  // the exchange buffers are reused by the second transform
    team_fft<T, M>(v, bufA, bufB, tw_s, tau, bar_id);
    if (active) {
      const int64_t obase = ((int64_t)w * P.nrows + row) * N;
      const int c0 = N / 2;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int k = tau + TPR * r;
        if (k < N) {  // fftshift (utils.py:559): bin k lands at (k + N // 2) mod N
          int kk = k + c0;
          if (kk >= N) kk -= N;
          store_cpx<T>(P.out, P.out_dtype, obase + kk, cmul(cconj(v[r]), cout[r]));
        }
      }
    }
    team_sync<TEAM>(bar_id);
  }
}

template <typename T, int M>
static int launch_blue_m(const BlueParams<T> &P, cudaStream_t st) {
  using C = BlueCfg<T, M>;
  static DeviceLatch configured;
  int cfg_dev = 0;
  if (configured.need(&cfg_dev)) {
    SDS_CUDA(cudaFuncSetAttribute(delay_fft_bluestein_kernel<T, M>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    SDS_CUDA(cudaFuncSetAttribute(delay_fft_bluestein_kernel<T, M>,
                                  cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    configured.set(cfg_dev);
  }
  int dev = 0, nsm = 148, per_sm = 1;
  SDS_CUDA(cudaGetDevice(&dev));
  SDS_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  SDS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, delay_fft_bluestein_kernel<T, M>,
                                                         C::THREADS, C::SMEM));
  if (per_sm < 1) return SDS_ENOTSUP;
  const int64_t nstep = (P.total + C::ROWS - 1) / C::ROWS;
  int64_t blocks = (nstep + C::NTEAM - 1) / C::NTEAM;
  if (blocks > (int64_t)nsm * per_sm) blocks = (int64_t)nsm * per_sm;
  delay_fft_bluestein_kernel<T, M><<<(unsigned)blocks, C::THREADS, C::SMEM, st>>>(P);
  SDS_LAUNCH_CHECK("delay_fft_bluestein_kernel");
  return SDS_OK;
}

template <typename T>
static int launch_blue_t(int M, const BlueParams<T> &P, cudaStream_t st) {
  switch (M) {
    case 64: return launch_blue_m<T, 64>(P, st);
    case 128: return launch_blue_m<T, 128>(P, st);
    case 256: return launch_blue_m<T, 256>(P, st);
    case 512: return launch_blue_m<T, 512>(P, st);
    case 1024: return launch_blue_m<T, 1024>(P, st);
    case 2048: return launch_blue_m<T, 2048>(P, st);
    default: return SDS_ENOTSUP;
  }
}

// forward transform of a non-power-of-two length; SDS_ENOTSUP when the plan has no tables
int launch_bluestein(const Plan &pl, const void *in, int in_dtype, int64_t in_row_stride,
                     const uint8_t *flags, int64_t flag_row_stride, int64_t nrows, int nwin,
                     const int32_t *win_start, double delta_x, void *out, int out_dtype,
                     cudaStream_t st) {
  if (pl.blue_m == 0) return SDS_ENOTSUP;
  auto fill = [&](auto &P) {
    P.in = in; P.flags = flags; P.out = out;
    P.in_row_stride = in_row_stride; P.flag_row_stride = flag_row_stride;
    P.nrows = nrows; P.total = nrows * nwin;
    P.in_dtype = in_dtype; P.out_dtype = out_dtype; P.nwin = nwin; P.N = pl.nfreq;
    for (int i = 0; i < SDS_MAX_WINDOWS; ++i) P.win_start[i] = i < nwin ? win_start[i] : 0;
  };
  if (pl.math == SDS_F64) {
    BlueParams<double> P;
    fill(P);
    P.taper = pl.taper_d; P.chirp = pl.blue_chirp_d; P.bhat = pl.blue_bhat_d; P.tw = pl.blue_tw_d;
    P.delta_x = delta_x;
    return launch_blue_t<double>(pl.blue_m, P, st);
  }
  BlueParams<float> P;
  fill(P);
  P.taper = pl.taper_f; P.chirp = pl.blue_chirp_f; P.bhat = pl.blue_bhat_f; P.tw = pl.blue_tw_f;
  P.delta_x = (float)delta_x;
  return launch_blue_t<float>(pl.blue_m, P, st);
}

}  // namespace sds
