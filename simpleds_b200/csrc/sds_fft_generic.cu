// K1/K4 any-length path: batched shared-memory Stockham transform with runtime
// radices.  Correct for every 1 <= N <= SDS_MAX_NFREQ; it serves the inverse branch
// (utils.py:561-564) and the lengths the fast paths do not take.  sds_delay_transform below
// routes forward transforms of power-of-two lengths to sds_fft_pow2.cu and of other lengths
// 8..1024 (21, 203, ... of the reference's bundled files) to sds_fft_bluestein.cu.
//
// Replaces utils.py:548-566 and the mask multiply of ds.py:3503-3517: flag mask,
// taper, transform, fftshift (an index rotation on the store) and the delta_x
// scale happen in one pass; the reference makes five passes over complex128.
#include "sds_common.cuh"

namespace sds {

template <typename T>
struct GenericFftParams {
  const void *in;
  const uint8_t *flags;
  void *out;
  const T *taper;
  const cpx<T> *tw;
  int64_t in_row_stride, flag_row_stride, nrows, total;
  T delta_x;
  int in_dtype, out_dtype, inverse, N, nradix, nwin, rows_per_block, tpr;
  int radix[16];
  int win_start[SDS_MAX_WINDOWS];
};

template <typename T>
__global__ void __launch_bounds__(256)
delay_fft_generic_kernel(const __grid_constant__ GenericFftParams<T> P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cpx<T> *buf = reinterpret_cast<cpx<T> *>(smem_raw);
  const int N = P.N;
  const int team = threadIdx.x / P.tpr, lt = threadIdx.x % P.tpr;
  const int64_t rw = (int64_t)blockIdx.x * P.rows_per_block + team;
  const bool active = team < P.rows_per_block && rw < P.total;
  const int w = active ? (int)(rw / P.nrows) : 0;
  const int64_t r = active ? rw % P.nrows : 0;
  cpx<T> *A = buf + (size_t)team * 2 * N, *B = A + N;

  if (active) {
    const int64_t ibase = r * P.in_row_stride + P.win_start[w];
    const int64_t fbase = r * P.flag_row_stride + P.win_start[w];
    for (int n = lt; n < N; n += P.tpr) {
      cpx<T> v = load_cpx<T>(P.in, P.in_dtype, ibase + n);
      if (P.flags) {  // data * (~flag).astype(float)   ds.py:3503-3505
        T m = P.flags[fbase + n] ? (T)0 : (T)1;
        v.x *= m;
        v.y *= m;
      }
      if (!P.inverse) {  // data * win   utils.py:558
        T wn = P.taper[n];
        v.x *= wn;
        v.y *= wn;
      }
      A[n] = v;
    }
  }
  __syncthreads();

  int Ns = 1;
  for (int s = 0; s < P.nradix; ++s) {
    const int rdx = P.radix[s];
    const int span = N / rdx;             // input stride between the rdx legs
    const int tstep = N / (Ns * rdx);     // twiddle step of this stage
    if (active) {
      for (int o = lt; o < N; o += P.tpr) {
        const int k = o % Ns, p = (o / Ns) % rdx, jq = o / (Ns * rdx);
        const int j = jq * Ns + k;
        int base = k * tstep + p * span;
        if (base >= N) base -= N;
        int idx = 0;
        cpx<T> acc = {(T)0, (T)0};
        for (int q = 0; q < rdx; ++q) {
          cpx<T> v = A[j + q * span];
          cpx<T> t = P.tw[idx];
          if (P.inverse) t.y = -t.y;
          acc.x += v.x * t.x - v.y * t.y;
          acc.y += v.x * t.y + v.y * t.x;
          idx += base;
          if (idx >= N) idx -= N;
        }
        B[o] = acc;
      }
    }
    __syncthreads();
    cpx<T> *tmp = A; A = B; B = tmp;
    Ns *= rdx;
  }

  if (active) {
    const int c0 = N / 2;
    const int64_t obase = ((int64_t)w * P.nrows + r) * N;
    if (!P.inverse) {
      for (int k = lt; k < N; k += P.tpr) {  // fftshift + * delta_x  utils.py:559-560
        int kk = k + c0;
        if (kk >= N) kk -= N;
        cpx<T> v = A[k];
        v.x *= P.delta_x;
        v.y *= P.delta_x;
        store_cpx<T>(P.out, P.out_dtype, obase + kk, v);
      }
    } else {
      const T inv_n = (T)1 / (T)N;
      for (int n = lt; n < N; n += P.tpr) {  // ifftshift, / win * delta_x  utils.py:563-564
        int nn = n + N - c0;
        if (nn >= N) nn -= N;
        cpx<T> v = A[n];
        const T wn = P.taper[nn];
        v.x = v.x * inv_n / wn * P.delta_x;
        v.y = v.y * inv_n / wn * P.delta_x;
        store_cpx<T>(P.out, P.out_dtype, obase + nn, v);
      }
    }
  }
}

template <typename T>
static int launch_generic(const Plan &pl, const void *in, int in_dtype, int64_t in_row_stride,
                          const uint8_t *flags, int64_t flag_row_stride, int64_t nrows, int nwin,
                          const int32_t *win_start, double delta_x, int inverse, void *out,
                          int out_dtype, const T *taper, const cpx<T> *tw, cudaStream_t st) {
  GenericFftParams<T> P;
  P.in = in; P.flags = flags; P.out = out; P.taper = taper; P.tw = tw;
  P.in_row_stride = in_row_stride; P.flag_row_stride = flag_row_stride;
  P.nrows = nrows; P.total = nrows * nwin; P.delta_x = (T)delta_x;
  P.in_dtype = in_dtype; P.out_dtype = out_dtype; P.inverse = inverse;
  P.N = pl.nfreq; P.nradix = pl.nradix; P.nwin = nwin;
  for (int i = 0; i < 16; ++i) P.radix[i] = i < pl.nradix ? pl.radix[i] : 1;
  for (int i = 0; i < SDS_MAX_WINDOWS; ++i) P.win_start[i] = i < nwin ? win_start[i] : 0;
  int np2 = 32;
  while (np2 < pl.nfreq && np2 < 256) np2 <<= 1;
  P.tpr = np2;                       // threads per row (32..256)
  P.rows_per_block = 256 / P.tpr;
  size_t smem = (size_t)P.rows_per_block * 2 * pl.nfreq * sizeof(cpx<T>);
  static thread_local size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    SDS_CUDA(cudaFuncSetAttribute(delay_fft_generic_kernel<T>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  int64_t blocks = (P.total + P.rows_per_block - 1) / P.rows_per_block;
  SDS_CHECK_ARG(blocks < (1ll << 31), "sds_delay_transform: too many rows (%lld)", (long long)P.total);
  delay_fft_generic_kernel<T><<<(unsigned)blocks, 256, smem, st>>>(P);
  SDS_LAUNCH_CHECK("delay_fft_generic_kernel");
  return SDS_OK;
}

// fast path (sds_fft_pow2.cu); returns SDS_ENOTSUP when the shape is not covered
int launch_pow2(const Plan &pl, const void *in, int in_dtype, int64_t in_row_stride,
                const uint8_t *flags, int64_t flag_row_stride, int64_t nrows, int nwin,
                const int32_t *win_start, double delta_x, void *out, int out_dtype,
                cudaStream_t st);

// complex64 fast path with bulk-TMA staging (sds_fft_tma.cu); SDS_ENOTSUP when the layout is not
// 16-byte aligned or the plan is not SDS_F32
int launch_pow2_tma(const Plan &pl, const void *in, int in_dtype, int64_t in_row_stride,
                    const uint8_t *flags, int64_t flag_row_stride, int64_t nrows, int nwin,
                    const int32_t *win_start, double delta_x, void *out, int out_dtype,
                    cudaStream_t st);

// non-power-of-two lengths (sds_fft_bluestein.cu); SDS_ENOTSUP when the plan has no tables
int launch_bluestein(const Plan &pl, const void *in, int in_dtype, int64_t in_row_stride,
                     const uint8_t *flags, int64_t flag_row_stride, int64_t nrows, int nwin,
                     const int32_t *win_start, double delta_x, void *out, int out_dtype,
                     cudaStream_t st);

}  // namespace sds

extern "C" int sds_delay_transform(const sds_plan *plan, const void *in, int in_dtype,
                                   int64_t in_row_stride, const uint8_t *flags,
                                   int64_t flag_row_stride, int64_t nrows, int nwin,
                                   const int32_t *win_start, double delta_x, int inverse,
                                   void *out, int out_dtype, sds_stream stream) {
  SDS_CHECK_ARG(plan && in && out, "sds_delay_transform: NULL plan/in/out");
  SDS_CHECK_ARG(in_dtype == SDS_C64 || in_dtype == SDS_C128,
                "sds_delay_transform: in_dtype must be SDS_C64 or SDS_C128");
  SDS_CHECK_ARG(nwin >= 1 && nwin <= SDS_MAX_WINDOWS && win_start,
                "sds_delay_transform: nwin=%d outside [1, %d] or win_start NULL", nwin,
                SDS_MAX_WINDOWS);
  SDS_CHECK_ARG(nrows >= 0, "sds_delay_transform: nrows < 0");
  const sds::Plan &pl = plan->p;
  const int want = pl.math == SDS_F64 ? SDS_C128 : SDS_C64;
  SDS_CHECK_ARG(out_dtype == want,
                "sds_delay_transform: out_dtype %d does not match the plan's math (want %d)",
                out_dtype, want);
  for (int w = 0; w < nwin; ++w)
    SDS_CHECK_ARG(win_start[w] >= 0 && (int64_t)win_start[w] + pl.nfreq <= in_row_stride,
                  "sds_delay_transform: window %d [%d, %d) exceeds the row stride %lld", w,
                  win_start[w], win_start[w] + pl.nfreq, (long long)in_row_stride);
  if (nrows == 0) return SDS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (!inverse && pl.is_pow2) {
    int rc = sds::launch_pow2_tma(pl, in, in_dtype, in_row_stride, flags, flag_row_stride, nrows,
                                  nwin, win_start, delta_x, out, out_dtype, st);
    if (rc != SDS_ENOTSUP) return rc;
    rc = sds::launch_pow2(pl, in, in_dtype, in_row_stride, flags, flag_row_stride, nrows, nwin,
                          win_start, delta_x, out, out_dtype, st);
    if (rc != SDS_ENOTSUP) return rc;
  }
  if (!inverse && pl.blue_m) {
    int rc = sds::launch_bluestein(pl, in, in_dtype, in_row_stride, flags, flag_row_stride, nrows,
                                   nwin, win_start, delta_x, out, out_dtype, st);
    if (rc != SDS_ENOTSUP) return rc;
  }
  if (pl.math == SDS_F64)
    return sds::launch_generic<double>(pl, in, in_dtype, in_row_stride, flags, flag_row_stride,
                                       nrows, nwin, win_start, delta_x, inverse, out, out_dtype,
                                       pl.taper_d, (const sds::cpx<double> *)pl.twiddle_d, st);
  return sds::launch_generic<float>(pl, in, in_dtype, in_row_stride, flags, flag_row_stride, nrows,
                                    nwin, win_start, delta_x, inverse, out, out_dtype, pl.taper_f,
                                    (const sds::cpx<float> *)pl.twiddle_f, st);
}
