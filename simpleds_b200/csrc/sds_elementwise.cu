// Gather / elementwise pieces of the path: spectral-window take, radiometer
// sigma, complex Gaussian realisation.
#include "sds_common.cuh"

namespace sds {

// ---- spectral-window gather (ds.py:3315-3338) --------------------------------
template <typename E>
__global__ void take_windows_kernel(const E *__restrict__ in, E *__restrict__ out,
                                    const int32_t *__restrict__ chans, int64_t nrows,
                                    int nfreq_in, int nspw_out, int nfreq_out) {
  const int64_t total = (int64_t)nspw_out * nrows * nfreq_out;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(o % nfreq_out);
    const int64_t sr = o / nfreq_out;
    const int64_t r = sr % nrows;
    const int s = (int)(sr / nrows);
    const int c = chans[s * nfreq_out + f];
    out[o] = in[((int64_t)(c / nfreq_in) * nrows + r) * nfreq_in + (c % nfreq_in)];
  }
}

// ---- sigma (ds.py:3544-3583) -------------------------------------------------
__global__ void noise_sigma_kernel(const double *__restrict__ coef,
                                   const double *__restrict__ int_time, const void *nsample,
                                   int ns_dtype, int S, int U, int P, int B, int T, int F,
                                   double *__restrict__ out) {
  const int64_t total = (int64_t)S * U * P * B * T * F;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (int64_t)gridDim.x * blockDim.x) {
    int64_t q = o;
    const int f = (int)(q % F); q /= F;
    const int t = (int)(q % T); q /= T;
    const int b = (int)(q % B); q /= B;
    const int p = (int)(q % P); q /= P;
    q /= U;
    const int s = (int)q;
    const double n = load_real(nsample, ns_dtype, o);
    // Tsys / sqrt(df * t_int * nsample) * 1e3 * Omega / jy_to_mk; coef carries
    // the channel-only factors, including 1/sqrt(df)
    double sig = coef[((int64_t)s * P + p) * F + f] / sqrt(int_time[(int64_t)b * T + t] * n);
    if (!isfinite(sig)) sig = 0.0;  // np.ma.masked_invalid(...).filled(0)  ds.py:3575-3582
    out[o] = sig;
  }
}

// ---- noise draw (utils.py:486-506) -------------------------------------------
// Box-Muller on Philox4x32-10: counter = (flat index >> 1), two complex samples
// per 128-bit block; sample parity picks the word pair.
__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint64_t q, double &g1,
                                                   double &g2) {
  const uint64_t ctr = q >> 1;
  u32x4 c = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  u32x4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t a = (q & 1) ? r.z : r.x, b = (q & 1) ? r.w : r.y;
  const double u1 = ((double)a + 0.5) * 2.3283064365386963e-10;  // (0,1)
  const double u2 = ((double)b + 0.5) * 2.3283064365386963e-10;
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  g1 = rad * cs;
  g2 = rad * sn;
}

__global__ void generate_noise_kernel(const double *__restrict__ sigma, int64_t n,
                                      const double *__restrict__ g_re,
                                      const double *__restrict__ g_im, uint64_t seed,
                                      uint64_t offset, void *out, int out_dtype) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n;
       o += (int64_t)gridDim.x * blockDim.x) {
    double g1, g2;
    if (g_re) {
      g1 = g_re[o];
      g2 = g_im[o];
    } else {
      philox_normal_pair(seed, offset + (uint64_t)o, g1, g2);
    }
    const double s = sigma[o];
    // noise = sigma * (g1 + 1j g2); noise /= sqrt(2)   utils.py:501-505
    // (numpy divides a complex by a real scalar as a multiply by its reciprocal)
    const double scl = 1.0 / 1.4142135623730951;
    cpx<double> v = {s * g1 * scl, s * g2 * scl};
    store_cpx<double>(out, out_dtype, o, v);
  }
}

static inline unsigned grid_for(int64_t total, int block) {
  int64_t g = (total + block - 1) / block;
  const int64_t cap = 148 * 16;  // grid-stride beyond 16 CTAs per SM
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace sds

extern "C" int sds_take_windows(const void *in, void *out, int elem_bytes, int64_t nrows,
                                int nspw_in, int nfreq_in, const int32_t *chans, int nspw_out,
                                int nfreq_out, sds_stream stream) {
  SDS_CHECK_ARG(in && out && chans, "sds_take_windows: NULL pointer");
  SDS_CHECK_ARG(nrows >= 0 && nspw_in >= 1 && nfreq_in >= 1 && nspw_out >= 1 && nfreq_out >= 1,
                "sds_take_windows: bad dimensions");
  const int64_t total = (int64_t)nspw_out * nrows * nfreq_out;
  if (total == 0) return SDS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = sds::grid_for(total, 256);
  switch (elem_bytes) {
    case 1:
      sds::take_windows_kernel<uint8_t><<<g, 256, 0, st>>>((const uint8_t *)in, (uint8_t *)out, chans, nrows, nfreq_in, nspw_out, nfreq_out);
      break;
    case 4:
      sds::take_windows_kernel<uint32_t><<<g, 256, 0, st>>>((const uint32_t *)in, (uint32_t *)out, chans, nrows, nfreq_in, nspw_out, nfreq_out);
      break;
    case 8:
      sds::take_windows_kernel<uint2><<<g, 256, 0, st>>>((const uint2 *)in, (uint2 *)out, chans, nrows, nfreq_in, nspw_out, nfreq_out);
      break;
    case 16:
      sds::take_windows_kernel<uint4><<<g, 256, 0, st>>>((const uint4 *)in, (uint4 *)out, chans, nrows, nfreq_in, nspw_out, nfreq_out);
      break;
    default:
      sds::set_error("sds_take_windows: elem_bytes=%d not in {1,4,8,16}", elem_bytes);
      return SDS_EINVAL;
  }
  SDS_LAUNCH_CHECK("take_windows_kernel");
  return SDS_OK;
}

extern "C" int sds_noise_sigma(const double *coef, const double *int_time, const void *nsample,
                               int nsample_dtype, int S, int U, int P, int B, int T, int F,
                               double *sigma_out, sds_stream stream) {
  SDS_CHECK_ARG(coef && int_time && nsample && sigma_out, "sds_noise_sigma: NULL pointer");
  SDS_CHECK_ARG(nsample_dtype == SDS_F32 || nsample_dtype == SDS_F64,
                "sds_noise_sigma: nsample dtype must be SDS_F32 or SDS_F64");
  SDS_CHECK_ARG(S >= 0 && U >= 0 && P >= 0 && B >= 0 && T >= 0 && F >= 0,
                "sds_noise_sigma: negative dimension");
  const int64_t total = (int64_t)S * U * P * B * T * F;
  if (total == 0) return SDS_OK;
  sds::noise_sigma_kernel<<<sds::grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      coef, int_time, nsample, nsample_dtype, S, U, P, B, T, F, sigma_out);
  SDS_LAUNCH_CHECK("noise_sigma_kernel");
  return SDS_OK;
}

extern "C" int sds_generate_noise(const double *sigma, int64_t n, const double *g_re,
                                  const double *g_im, uint64_t seed, uint64_t offset, void *out,
                                  int out_dtype, sds_stream stream) {
  SDS_CHECK_ARG(sigma && out, "sds_generate_noise: NULL pointer");
  SDS_CHECK_ARG((g_re == nullptr) == (g_im == nullptr),
                "sds_generate_noise: g_re and g_im must both be given or both be NULL");
  SDS_CHECK_ARG(out_dtype == SDS_C64 || out_dtype == SDS_C128,
                "sds_generate_noise: out dtype must be SDS_C64 or SDS_C128");
  SDS_CHECK_ARG(n >= 0, "sds_generate_noise: n < 0");
  if (n == 0) return SDS_OK;
  sds::generate_noise_kernel<<<sds::grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      sigma, n, g_re, g_im, seed, offset, out, out_dtype);
  SDS_LAUNCH_CHECK("generate_noise_kernel");
  return SDS_OK;
}
