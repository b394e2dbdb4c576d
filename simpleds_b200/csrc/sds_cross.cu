// K2: cross power of redundant baselines.
//   * cross_power_full : the reference's materialised (S,P,B,B,T,D) tensor
//     (utils.py:387-389 via ds.py:3620-3633) -- write-bandwidth bound.
//   * cross_power_avg  : pair sums per (group, spw, pol, time, delay) without the
//     tensor -- factorised (O(B)) or explicit warp-shuffle pairs (O(B^2)).
#include "sds_common.cuh"

namespace sds {

// out[sp, i, j, e] = a2[sp, i, e] * conj(a1[sp, j, e]),  e = (t, d) flattened
template <typename T>
__global__ void __launch_bounds__(256)
cross_power_full_kernel(const cpx<T> *__restrict__ a1, const cpx<T> *__restrict__ a2, int P,
                        int B, int64_t TD, int64_t stride_s, int64_t stride_p,
                        const double *__restrict__ scale_sp, cpx<T> *__restrict__ out) {
  const int sp = blockIdx.z, i = blockIdx.y;
  // per-(window, pol) factor (the cosmological conversion, delay_spectrum.py:1306-1311): applied
  // to the i side once per element, in fp64, so the tensor is written once and already converted
  const double sc = scale_sp ? scale_sp[sp] : 1.0;
  const int64_t in_base = (int64_t)(sp / P) * stride_s + (int64_t)(sp % P) * stride_p;
  const cpx<T> *a1p = a1 + in_base, *a2p = a2 + in_base + (int64_t)i * TD;
  cpx<T> *op = out + ((int64_t)sp * B + i) * B * TD;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < TD;
       e += (int64_t)gridDim.x * blockDim.x) {
    cpx<T> x = a2p[e];
    if (scale_sp) x = {(T)((double)x.x * sc), (T)((double)x.y * sc)};
#pragma unroll 4
    for (int j = 0; j < B; ++j) op[(int64_t)j * TD + e] = cmul_conj(x, a1p[(int64_t)j * TD + e]);
  }
}

// factorised pair sums: all = (sum_i a2_i) conj(sum_j a1_j), auto = sum_i a2_i conj(a1_i)
template <typename T>
__global__ void __launch_bounds__(256)
cross_power_avg_fact_kernel(const cpx<T> *__restrict__ a1, const cpx<T> *__restrict__ a2, int P,
                            int64_t TD, int64_t stride_s, int64_t stride_p,
                            const int32_t *__restrict__ goff, int ngroup, int SP,
                            double2 *__restrict__ out_all, double2 *__restrict__ out_auto) {
  const int sp = blockIdx.y, g = blockIdx.z;
  const int64_t in_base = (int64_t)(sp / P) * stride_s + (int64_t)(sp % P) * stride_p;
  const int b0 = goff[g], b1 = goff[g + 1];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < TD;
       e += (int64_t)gridDim.x * blockDim.x) {
    cpx<double> s1 = {0, 0}, s2 = {0, 0}, au = {0, 0};
    for (int b = b0; b < b1; ++b) {
      const cpx<T> y = a1[in_base + (int64_t)b * TD + e], x = a2[in_base + (int64_t)b * TD + e];
      const cpx<double> yd = {(double)y.x, (double)y.y}, xd = {(double)x.x, (double)x.y};
      s1 = cadd(s1, yd);
      s2 = cadd(s2, xd);
      au = cadd(au, cmul_conj(xd, yd));
    }
    const cpx<double> al = cmul_conj(s2, s1);
    const int64_t o = ((int64_t)g * SP + sp) * TD + e;
    out_all[o] = make_double2(al.x, al.y);
    out_auto[o] = make_double2(au.x, au.y);
  }
}

// explicit pairs: lanes own 32 baselines j of a tile at one (t,d); x_i of the i
// tile is broadcast lane by lane with __shfl_sync; every ordered pair is formed.
// Block = 8 warps over a tile of 32 consecutive e = (t,d) elements.
template <typename T>
__global__ void __launch_bounds__(256)
cross_power_avg_pairs_kernel(const cpx<T> *__restrict__ a1, const cpx<T> *__restrict__ a2, int P,
                             int64_t TD, int64_t stride_s, int64_t stride_p,
                             const int32_t *__restrict__ goff, int ngroup, int SP,
                             double2 *__restrict__ out_all, double2 *__restrict__ out_auto) {
  __shared__ T xs_re[32][33], xs_im[32][33], ys_re[32][33], ys_im[32][33];
  const int sp = blockIdx.y, g = blockIdx.z;
  const int64_t in_base = (int64_t)(sp / P) * stride_s + (int64_t)(sp % P) * stride_p;
  const int b0 = goff[g], b1 = goff[g + 1], nb = b1 - b0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t e0 = (int64_t)blockIdx.x * 32;
  double all_re[4] = {0, 0, 0, 0}, all_im[4] = {0, 0, 0, 0};
  double au_re[4] = {0, 0, 0, 0}, au_im[4] = {0, 0, 0, 0};

  for (int it = 0; it < nb; it += 32) {
    for (int jt = 0; jt < nb; jt += 32) {
      __syncthreads();
      // coalesced tile loads: thread (warp*4+q, lane) -> baseline row, element e0+lane
      for (int rr = warp; rr < 32; rr += 8) {
        const int64_t e = e0 + lane;
        cpx<T> x = {0, 0}, y = {0, 0};
        if (e < TD) {
          if (it + rr < nb) x = a2[in_base + (int64_t)(b0 + it + rr) * TD + e];
          if (jt + rr < nb) y = a1[in_base + (int64_t)(b0 + jt + rr) * TD + e];
        }
        xs_re[rr][lane] = x.x; xs_im[rr][lane] = x.y;
        ys_re[rr][lane] = y.x; ys_im[rr][lane] = y.y;
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int el = warp * 4 + q;  // element of the tile handled by this warp
        const T xr = xs_re[lane][el], xi = xs_im[lane][el];  // lane <-> baseline i
        const T yr = ys_re[lane][el], yi = ys_im[lane][el];  // lane <-> baseline j
        double sr = 0, si = 0;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const T bxr = __shfl_sync(0xffffffffu, xr, r), bxi = __shfl_sync(0xffffffffu, xi, r);
          sr += (double)bxr * yr + (double)bxi * yi;   // x_i * conj(y_lane)
          si += (double)bxi * yr - (double)bxr * yi;
        }
        all_re[q] += sr; all_im[q] += si;
        if (it == jt) {
          au_re[q] += (double)xr * yr + (double)xi * yi;
          au_im[q] += (double)xi * yr - (double)xr * yi;
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    double v[4] = {all_re[q], all_im[q], au_re[q], au_im[q]};
#pragma unroll
    for (int m = 16; m > 0; m >>= 1)
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] += __shfl_xor_sync(0xffffffffu, v[c], m);
    const int64_t e = e0 + warp * 4 + q;
    if (lane == 0 && e < TD) {
      const int64_t o = ((int64_t)g * SP + sp) * TD + e;
      out_all[o] = make_double2(v[0], v[1]);
      out_auto[o] = make_double2(v[2], v[3]);
    }
  }
}

template <typename T>
static int run_full(const void *a1, const void *a2, int S, int P, int B, int64_t TD,
                    int64_t stride_s, int64_t stride_p, const double *scale_sp, void *out, cudaStream_t st) {
  int64_t gx = (TD + 255) / 256;
  if (gx > 4096) gx = 4096;
  dim3 grid((unsigned)gx, B, S * P);
  cross_power_full_kernel<T><<<grid, 256, 0, st>>>((const cpx<T> *)a1, (const cpx<T> *)a2, P, B,
                                                   TD, stride_s, stride_p, scale_sp, (cpx<T> *)out);
  SDS_LAUNCH_CHECK("cross_power_full_kernel");
  return SDS_OK;
}

template <typename T>
static int run_avg(const void *a1, const void *a2, int S, int P, int64_t TD, int64_t stride_s,
                   int64_t stride_p, const int32_t *goff, int ngroup, int method, double *out_all,
                   double *out_auto, cudaStream_t st) {
  if (method == SDS_METHOD_FACTORISED) {
    int64_t gx = (TD + 255) / 256;
    if (gx > 65535) gx = 65535;
    dim3 grid((unsigned)gx, S * P, ngroup);
    cross_power_avg_fact_kernel<T><<<grid, 256, 0, st>>>(
        (const cpx<T> *)a1, (const cpx<T> *)a2, P, TD, stride_s, stride_p, goff, ngroup, S * P,
        (double2 *)out_all, (double2 *)out_auto);
    SDS_LAUNCH_CHECK("cross_power_avg_fact_kernel");
  } else {
    int64_t gx = (TD + 31) / 32;
    SDS_CHECK_ARG(gx < (1ll << 31), "sds_cross_power_avg: T*D too large");
    dim3 grid((unsigned)gx, S * P, ngroup);
    cross_power_avg_pairs_kernel<T><<<grid, 256, 0, st>>>(
        (const cpx<T> *)a1, (const cpx<T> *)a2, P, TD, stride_s, stride_p, goff, ngroup, S * P,
        (double2 *)out_all, (double2 *)out_auto);
    SDS_LAUNCH_CHECK("cross_power_avg_pairs_kernel");
  }
  return SDS_OK;
}

}  // namespace sds

extern "C" int sds_cross_power_full(const void *a1, const void *a2, int dtype, int S, int P, int B,
                                    int T, int D, int64_t stride_s, int64_t stride_p,
                                    const double *scale_sp, void *out, sds_stream stream) {
  SDS_CHECK_ARG(a1 && a2 && out, "sds_cross_power_full: NULL pointer");
  SDS_CHECK_ARG(dtype == SDS_C64 || dtype == SDS_C128,
                "sds_cross_power_full: dtype must be SDS_C64 or SDS_C128");
  SDS_CHECK_ARG(S >= 0 && P >= 0 && B >= 0 && T >= 0 && D >= 0, "sds_cross_power_full: negative dimension");
  SDS_CHECK_ARG((int64_t)S * P <= 65535 && B <= 65535, "sds_cross_power_full: S*P or B exceeds 65535");
  if ((int64_t)S * P * B * T * D == 0) return SDS_OK;
  const int64_t TD = (int64_t)T * D;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == SDS_C64 ? sds::run_full<float>(a1, a2, S, P, B, TD, stride_s, stride_p, scale_sp, out, st)
                          : sds::run_full<double>(a1, a2, S, P, B, TD, stride_s, stride_p, scale_sp, out, st);
}

extern "C" int sds_cross_power_avg(const void *a1, const void *a2, int dtype, int S, int P, int B,
                                   int T, int D, int64_t stride_s, int64_t stride_p,
                                   const int32_t *group_offset, int ngroup, int method,
                                   double *out_all, double *out_auto, sds_stream stream) {
  SDS_CHECK_ARG(a1 && a2 && out_all && out_auto && group_offset, "sds_cross_power_avg: NULL pointer");
  SDS_CHECK_ARG(dtype == SDS_C64 || dtype == SDS_C128,
                "sds_cross_power_avg: dtype must be SDS_C64 or SDS_C128");
  SDS_CHECK_ARG(method == SDS_METHOD_FACTORISED || method == SDS_METHOD_EXPLICIT,
                "sds_cross_power_avg: unknown method %d", method);
  SDS_CHECK_ARG(S >= 0 && P >= 0 && B >= 0 && T >= 0 && D >= 0 && ngroup >= 0,
                "sds_cross_power_avg: negative dimension");
  SDS_CHECK_ARG((int64_t)S * P <= 65535 && ngroup <= 65535,
                "sds_cross_power_avg: S*P or ngroup exceeds 65535");
  if ((int64_t)S * P * T * D == 0 || ngroup == 0) return SDS_OK;
  const int64_t TD = (int64_t)T * D;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == SDS_C64
             ? sds::run_avg<float>(a1, a2, S, P, TD, stride_s, stride_p, group_offset, ngroup,
                                   method, out_all, out_auto, st)
             : sds::run_avg<double>(a1, a2, S, P, TD, stride_s, stride_p, group_offset, ngroup,
                                    method, out_all, out_auto, st);
}
