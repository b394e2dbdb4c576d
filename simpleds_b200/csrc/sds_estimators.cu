// Downstream estimators on the averaged spectra: weighted average with
// uncertainty propagation, fold along the delay axis, bootstrap resampling.  The arrays
// here are the kilobyte-to-megabyte outputs of the fused engine, so the kernels are
// plain coalesced passes (one thread per output element, consecutive threads along the
// contiguous `inner` axis; a warp per output when the reduced axis is the contiguous one).
#include "sds_common.cuh"

namespace sds {

// ---- weighted_average (utils.py:569-645) -------------------------------------------------
struct WavgAcc {
  double swx, sw, ss;  // sum w x, sum w, sum sigma^2 w^2
};
__device__ __forceinline__ void wavg_add(WavgAcc &a, double x, double s, double w) {
  a.swx += w * x;
  a.sw += w;
  a.ss += s * s * (w * w);
}

// inner > 1: thread per (o, i), coalesced along i
__global__ void wavg_inner_kernel(const double *__restrict__ x, const double *__restrict__ sig,
                                  const double *__restrict__ w, int w1d, int64_t outer, int64_t n,
                                  int64_t inner, double *__restrict__ out, double *__restrict__ osig) {
  const int64_t total = outer * inner;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = e / inner, i = e % inner;
    WavgAcc a = {0.0, 0.0, 0.0};
    int64_t k = 0;
    for (; k + 4 <= n; k += 4) {  // four independent loads in flight per array
      double xs[4], ss[4], ws[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t idx = (o * n + k + u) * inner + i;
        xs[u] = x[idx];
        ss[u] = sig[idx];
        ws[u] = w ? (w1d ? w[k + u] : w[idx]) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) wavg_add(a, xs[u], ss[u], w ? ws[u] : 1.0 / (ss[u] * ss[u]));
    }
    for (; k < n; ++k) {
      const int64_t idx = (o * n + k) * inner + i;
      const double s = sig[idx];
      const double wk = w ? (w1d ? w[k] : w[idx]) : 1.0 / (s * s);
      wavg_add(a, x[idx], s, wk);
    }
    out[e] = a.swx / a.sw;
    osig[e] = sqrt(a.ss / (fabs(a.sw) * fabs(a.sw)));
  }
}
// inner == 1: warp per output, lanes stride the contiguous reduced axis
__global__ void wavg_last_kernel(const double *__restrict__ x, const double *__restrict__ sig,
                                 const double *__restrict__ w, int w1d, int64_t outer, int64_t n,
                                 double *__restrict__ out, double *__restrict__ osig) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t o = warp; o < outer; o += nwarp) {
    WavgAcc a = {0.0, 0.0, 0.0};
    int64_t k = lane;
    for (; k + 96 < n; k += 128) {  // four independent loads in flight per array
      double xs[4], ss[4], ws[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t idx = o * n + k + 32 * u;
        xs[u] = x[idx];
        ss[u] = sig[idx];
        ws[u] = w ? (w1d ? w[k + 32 * u] : w[idx]) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) wavg_add(a, xs[u], ss[u], w ? ws[u] : 1.0 / (ss[u] * ss[u]));
    }
    for (; k < n; k += 32) {
      const int64_t idx = o * n + k;
      const double s = sig[idx];
      const double wk = w ? (w1d ? w[k] : w[idx]) : 1.0 / (s * s);
      wavg_add(a, x[idx], s, wk);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a.swx += __shfl_xor_sync(0xffffffffu, a.swx, off);
      a.sw += __shfl_xor_sync(0xffffffffu, a.sw, off);
      a.ss += __shfl_xor_sync(0xffffffffu, a.ss, off);
    }
    if (lane == 0) {
      out[o] = a.swx / a.sw;
      osig[o] = sqrt(a.ss / (fabs(a.sw) * fabs(a.sw)));
    }
  }
}

// ---- fold_along_delay (utils.py:651-796) ---------------------------------------------------
// NC = 1 real, 2 complex (component c of x / sigma / weights folded on its own)
template <int NC>
__global__ void fold_kernel(const double *__restrict__ x, const double *__restrict__ sig,
                            const double *__restrict__ w, int64_t outer, int64_t n, int64_t inner,
                            int64_t pos0, int64_t neg0, int64_t nout, int zero_bin,
                            double *__restrict__ out, double *__restrict__ osig) {
  const int64_t total = outer * nout * inner;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e % inner, j = (e / inner) % nout, o = e / (inner * nout);
    const int64_t ip = ((o * n + pos0 + j) * inner + i) * NC;
    const int64_t in = ((o * n + neg0 - j) * inner + i) * NC;
    const double zs = (zero_bin && j == 0) ? 1.4142135623730951 : 1.0;  // utils.py:737
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double wp = w ? w[ip + c] : 1.0, wn = w ? w[in + c] : 1.0;
      const double sp = sig[ip + c] * zs, sn = sig[in + c] * zs;
      const double sw = wp + wn;
      out[e * NC + c] = (x[ip + c] * wp + x[in + c] * wn) / sw;
      osig[e * NC + c] = sqrt((sp * sp * (wp * wp) + sn * sn * (wn * wn)) / (fabs(sw) * fabs(sw)));
    }
  }
}

// ---- bootstrap_array (utils.py:171-209) ---------------------------------------------------
template <typename V>
__global__ void boot_gather_kernel(const V *__restrict__ in, const int64_t *__restrict__ index,
                                   int64_t outer, int64_t n, int64_t nboot, int64_t inner,
                                   V *__restrict__ out) {
  const int64_t total = outer * n * nboot * inner;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e % inner;
    const int64_t ib = (e / inner) % (n * nboot);
    const int64_t o = e / (inner * n * nboot);
    out[e] = in[(o * n + __ldg(index + ib)) * inner + i];
  }
}
__global__ void boot_index_kernel(uint32_t k0, uint32_t k1, int64_t n, int64_t count,
                                  int64_t *__restrict__ index) {
  const int64_t quads = (count + 3) / 4;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < quads;
       q += (int64_t)gridDim.x * blockDim.x) {
    const u32x4 r = philox4x32_10({(uint32_t)q, (uint32_t)(q >> 32), 0u, 0u}, k0, k1);
    const uint32_t v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (4 * q + k < count)  // floor(u n / 2^32): uniform on [0, n) up to n / 2^32
        index[4 * q + k] = (int64_t)(((uint64_t)v[k] * (uint64_t)n) >> 32);
  }
}

static int grid_for(int64_t work, int block) {
  int64_t g = (work + block - 1) / block;
  const int64_t cap = 148 * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace sds

using namespace sds;

extern "C" int sds_weighted_average(const double *x, const double *sigma, const double *weights,
                                    int weights_1d, int64_t outer, int64_t n, int64_t inner,
                                    double *out, double *out_sigma, sds_stream stream) {
  SDS_CHECK_ARG(x && sigma && out && out_sigma, "sds_weighted_average: null pointer");
  SDS_CHECK_ARG(outer >= 0 && n >= 1 && inner >= 1, "sds_weighted_average: bad dims");
  if (outer == 0) return SDS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (inner == 1) {
    wavg_last_kernel<<<grid_for(outer * 32, 256), 256, 0, st>>>(x, sigma, weights, weights_1d, outer,
                                                               n, out, out_sigma);
  } else {
    wavg_inner_kernel<<<grid_for(outer * inner, 256), 256, 0, st>>>(x, sigma, weights, weights_1d,
                                                                   outer, n, inner, out, out_sigma);
  }
  SDS_LAUNCH_CHECK("sds_weighted_average");
  return SDS_OK;
}

extern "C" int sds_fold_along_delay(const void *x, const void *sigma, const void *weights, int dtype,
                                    int64_t outer, int64_t n, int64_t inner, int64_t pos0,
                                    int64_t neg0, int64_t nout, int zero_bin, void *out,
                                    void *out_sigma, sds_stream stream) {
  SDS_CHECK_ARG(x && sigma && out && out_sigma, "sds_fold_along_delay: null pointer");
  SDS_CHECK_ARG(dtype == SDS_F64 || dtype == SDS_C128, "sds_fold_along_delay: dtype must be SDS_F64 or SDS_C128");
  SDS_CHECK_ARG(outer >= 0 && n >= 1 && inner >= 1 && nout >= 1, "sds_fold_along_delay: bad dims");
  SDS_CHECK_ARG(pos0 >= 0 && pos0 + nout <= n && neg0 < n && neg0 - (nout - 1) >= 0,
                "sds_fold_along_delay: fold range [%lld..], [..%lld] x %lld leaves the axis of %lld",
                (long long)pos0, (long long)neg0, (long long)nout, (long long)n);
  if (outer == 0) return SDS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(outer * nout * inner, 256);
  if (dtype == SDS_F64)
    fold_kernel<1><<<grid, 256, 0, st>>>((const double *)x, (const double *)sigma, (const double *)weights,
                                         outer, n, inner, pos0, neg0, nout, zero_bin, (double *)out,
                                         (double *)out_sigma);
  else
    fold_kernel<2><<<grid, 256, 0, st>>>((const double *)x, (const double *)sigma, (const double *)weights,
                                         outer, n, inner, pos0, neg0, nout, zero_bin, (double *)out,
                                         (double *)out_sigma);
  SDS_LAUNCH_CHECK("sds_fold_along_delay");
  return SDS_OK;
}

extern "C" int sds_bootstrap_gather(const void *in, const int64_t *index, int elem_bytes,
                                    int64_t outer, int64_t n, int64_t nboot, int64_t inner, void *out,
                                    sds_stream stream) {
  SDS_CHECK_ARG(in && index && out, "sds_bootstrap_gather: null pointer");
  SDS_CHECK_ARG(elem_bytes == 8 || elem_bytes == 16, "sds_bootstrap_gather: elem_bytes must be 8 or 16");
  SDS_CHECK_ARG(outer >= 0 && n >= 1 && nboot >= 1 && inner >= 1, "sds_bootstrap_gather: bad dims");
  if (outer == 0) return SDS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(outer * n * nboot * inner, 256);
  if (elem_bytes == 8)
    boot_gather_kernel<double><<<grid, 256, 0, st>>>((const double *)in, index, outer, n, nboot, inner,
                                                     (double *)out);
  else
    boot_gather_kernel<double2><<<grid, 256, 0, st>>>((const double2 *)in, index, outer, n, nboot,
                                                      inner, (double2 *)out);
  SDS_LAUNCH_CHECK("sds_bootstrap_gather");
  return SDS_OK;
}

extern "C" int sds_bootstrap_indices(uint64_t seed, int64_t n, int64_t nboot, int64_t *index,
                                     sds_stream stream) {
  SDS_CHECK_ARG(index, "sds_bootstrap_indices: null pointer");
  SDS_CHECK_ARG(n >= 1 && nboot >= 1 && n < ((int64_t)1 << 32), "sds_bootstrap_indices: bad dims");
  const int64_t count = n * nboot;
  boot_index_kernel<<<grid_for((count + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      (uint32_t)seed, (uint32_t)(seed >> 32), n, count, index);
  SDS_LAUNCH_CHECK("sds_bootstrap_indices");
  return SDS_OK;
}
