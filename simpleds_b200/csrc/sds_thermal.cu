// K3: thermal-noise estimate (ds.py:3637-3707, utils.py:230-278).
//
// thermal[s,p,i,j,t] = 1e-6 * trapz_nu( theta^2 * w^2 * nu^4 ),
// theta = 1e3 Tsys Omega / sqrt(t_int[j,t] * sqrt(n1[j] n2[i]) * npol * sqrt2).
// Every factor separates into an i part, a j part and a channel part, and so
// does the validity mask of np.ma.masked_invalid (a trapezoid panel counts only
// when both of its end samples are valid for both baselines), so
//   thermal[i,j,t] = 1e-6 polf[p] / t_int[j,t] *
//        sum_k  L[i,k] R[j,k] wl[k] + L'[i,k] R'[j,k] wr[k]
// with a = nsample^-1/2, e = valid(k) & valid(k+1), L = e a(k), L' = e a(k+1),
// wl = dnu_k/2 * coef[k], wr = dnu_k/2 * coef[k+1].  The full kernel is a small
// fp64 tiled contraction over nu; the reference builds the (S,P,B,B,T,F) temp.
#include "sds_common.cuh"

namespace sds {

constexpr int TH_TILE = 16;
constexpr int TH_KC = 32;

__device__ __forceinline__ bool ns_valid(double n) { return n > 0.0; }  // finite theta

__global__ void __launch_bounds__(TH_TILE *TH_TILE)
thermal_full_kernel(const void *n1, const void *n2, int ns_dtype, int64_t stride_s,
                    int64_t stride_p, const double *__restrict__ coef,
                    const double *__restrict__ freq, const double *__restrict__ int_time,
                    const double *__restrict__ pol_factor, const double *__restrict__ scale_sp, int P,
                    int B, int T, int F, double *__restrict__ out) {
  __shared__ double Ll[TH_TILE][TH_KC + 1], Lr[TH_TILE][TH_KC + 1];
  __shared__ double Rl[TH_TILE][TH_KC + 1], Rr[TH_TILE][TH_KC + 1];
  const int ntile = (B + TH_TILE - 1) / TH_TILE;
  const int it = blockIdx.x / ntile, jt = blockIdx.x % ntile;
  const int t = blockIdx.y, sp = blockIdx.z, s = sp / P, p = sp % P;
  const int tx = threadIdx.x % TH_TILE, ty = threadIdx.x / TH_TILE;
  const int64_t base = (int64_t)s * stride_s + (int64_t)p * stride_p;
  const double *cf = coef + (int64_t)sp * F, *nu = freq + (int64_t)s * F;
  double acc = 0.0;

  for (int k0 = 0; k0 < F - 1; k0 += TH_KC) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < TH_TILE * TH_KC; idx += TH_TILE * TH_TILE) {
      const int row = idx / TH_KC, kk = idx % TH_KC, k = k0 + kk;
      double ll = 0, lr = 0, rl = 0, rr = 0;
      if (k < F - 1) {
        const double c0 = cf[k], c1 = cf[k + 1];
        const bool cv = isfinite(c0) && isfinite(c1);
        const double hd = 0.5 * (nu[k + 1] - nu[k]);
        const int bi = it * TH_TILE + row, bj = jt * TH_TILE + row;
        if (cv && bi < B) {  // i side: n2 (ds.py:3660-3661, utils.py:387-389 index order)
          const int64_t o = base + ((int64_t)bi * T + t) * F + k;
          const double a = load_real(n2, ns_dtype, o), b = load_real(n2, ns_dtype, o + 1);
          if (ns_valid(a) && ns_valid(b)) {
            ll = rsqrt(a) * hd * c0;
            lr = rsqrt(b) * hd * c1;
          }
        }
        if (cv && bj < B) {  // j side: n1
          const int64_t o = base + ((int64_t)bj * T + t) * F + k;
          const double a = load_real(n1, ns_dtype, o), b = load_real(n1, ns_dtype, o + 1);
          if (ns_valid(a) && ns_valid(b)) {
            rl = rsqrt(a);
            rr = rsqrt(b);
          }
        }
      }
      Ll[row][kk] = ll; Lr[row][kk] = lr; Rl[row][kk] = rl; Rr[row][kk] = rr;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < TH_KC; ++kk)
      acc += Ll[ty][kk] * Rl[tx][kk] + Lr[ty][kk] * Rr[tx][kk];
  }
  const int i = it * TH_TILE + ty, j = jt * TH_TILE + tx;
  if (i < B && j < B) {
    const double tj = int_time[(int64_t)j * T + t];
    const double v = (tj > 0.0) ? 1e-6 * pol_factor[p] * acc / tj : 0.0;
    // optional per-(window, pol) factor: the thermal conversion of delay_spectrum.py:1339-1343
    out[(((int64_t)sp * B + i) * B + j) * T + t] = scale_sp ? v * scale_sp[sp] : v;
  }
}

// Pair sums of the thermal estimate per (group, spw, pol, time) without the B x B
// array: all = sum_k wl A B + wr A' B', auto = sum_k wl C + wr C' (header comment).
__global__ void __launch_bounds__(128)
thermal_avg_kernel(const void *n1, const void *n2, int ns_dtype, int64_t stride_s,
                   int64_t stride_p, const double *__restrict__ coef,
                   const double *__restrict__ freq, const double *__restrict__ int_time,
                   const double *__restrict__ pol_factor, const int32_t *__restrict__ goff, int P,
                   int T, int F, int SP, double *__restrict__ out_all,
                   double *__restrict__ out_auto) {
  __shared__ double red[2][4];
  const int t = blockIdx.x, sp = blockIdx.y, g = blockIdx.z, s = sp / P, p = sp % P;
  const int b0 = goff[g], b1 = goff[g + 1];
  const int64_t base = (int64_t)s * stride_s + (int64_t)p * stride_p;
  const double *cf = coef + (int64_t)sp * F, *nu = freq + (int64_t)s * F;
  double all = 0.0, au = 0.0;
  for (int k = threadIdx.x; k < F - 1; k += blockDim.x) {
    const double c0 = cf[k], c1 = cf[k + 1];
    if (!(isfinite(c0) && isfinite(c1))) continue;
    const double hd = 0.5 * (nu[k + 1] - nu[k]);
    double A = 0, Ap = 0, Bv = 0, Bp = 0, Cv = 0, Cp = 0;
    for (int b = b0; b < b1; ++b) {
      const int64_t o = base + ((int64_t)b * T + t) * F + k;
      const double tj = int_time[(int64_t)b * T + t];
      const double it = tj > 0.0 ? 1.0 / tj : 0.0;
      const double i0 = load_real(n2, ns_dtype, o), i1 = load_real(n2, ns_dtype, o + 1);
      const double j0 = load_real(n1, ns_dtype, o), j1 = load_real(n1, ns_dtype, o + 1);
      const bool ei = ns_valid(i0) && ns_valid(i1), ej = ns_valid(j0) && ns_valid(j1);
      const double l0 = ei ? rsqrt(i0) : 0.0, l1 = ei ? rsqrt(i1) : 0.0;
      const double r0 = ej ? rsqrt(j0) * it : 0.0, r1 = ej ? rsqrt(j1) * it : 0.0;
      A += l0; Ap += l1; Bv += r0; Bp += r1; Cv += l0 * r0; Cp += l1 * r1;
    }
    all += hd * (c0 * A * Bv + c1 * Ap * Bp);
    au += hd * (c0 * Cv + c1 * Cp);
  }
  for (int m = 16; m > 0; m >>= 1) {
    all += __shfl_xor_sync(0xffffffffu, all, m);
    au += __shfl_xor_sync(0xffffffffu, au, m);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = all; red[1][threadIdx.x >> 5] = au; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double f = 1e-6 * pol_factor[p];
    const int64_t o = ((int64_t)g * SP + sp) * T + t;
    out_all[o] = f * (red[0][0] + red[0][1] + red[0][2] + red[0][3]);
    out_auto[o] = f * (red[1][0] + red[1][1] + red[1][2] + red[1][3]);
  }
}

}  // namespace sds

extern "C" int sds_thermal_power_avg(const void *n1, const void *n2, int nsample_dtype,
                                     int64_t stride_s, int64_t stride_p, const double *coef,
                                     const double *freq, const double *int_time,
                                     const double *pol_factor, const int32_t *group_offset,
                                     int ngroup, int S, int P, int B, int T, int F,
                                     double *out_all, double *out_auto, sds_stream stream) {
  SDS_CHECK_ARG(n1 && n2 && coef && freq && int_time && pol_factor && group_offset && out_all &&
                    out_auto, "sds_thermal_power_avg: NULL pointer");
  SDS_CHECK_ARG(nsample_dtype == SDS_F32 || nsample_dtype == SDS_F64,
                "sds_thermal_power_avg: nsample dtype must be SDS_F32 or SDS_F64");
  SDS_CHECK_ARG(S >= 0 && P >= 0 && B >= 0 && T >= 0 && F >= 0 && ngroup >= 0,
                "sds_thermal_power_avg: negative dimension");
  SDS_CHECK_ARG((int64_t)S * P <= 65535 && ngroup <= 65535,
                "sds_thermal_power_avg: S*P or ngroup exceeds 65535");
  if ((int64_t)S * P * T == 0 || ngroup == 0) return SDS_OK;
  dim3 grid(T, S * P, ngroup);
  sds::thermal_avg_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(
      n1, n2, nsample_dtype, stride_s, stride_p, coef, freq, int_time, pol_factor, group_offset, P,
      T, F, S * P, out_all, out_auto);
  SDS_LAUNCH_CHECK("thermal_avg_kernel");
  return SDS_OK;
}

extern "C" int sds_thermal_power_full(const void *n1, const void *n2, int nsample_dtype,
                                      int64_t stride_s, int64_t stride_p, const double *coef,
                                      const double *freq, const double *int_time,
                                      const double *pol_factor, int S, int P, int B, int T, int F,
                                      const double *scale_sp, double *out, sds_stream stream) {
  SDS_CHECK_ARG(n1 && n2 && coef && freq && int_time && pol_factor && out,
                "sds_thermal_power_full: NULL pointer");
  SDS_CHECK_ARG(nsample_dtype == SDS_F32 || nsample_dtype == SDS_F64,
                "sds_thermal_power_full: nsample dtype must be SDS_F32 or SDS_F64");
  SDS_CHECK_ARG(S >= 0 && P >= 0 && B >= 0 && T >= 0 && F >= 0,
                "sds_thermal_power_full: negative dimension");
  SDS_CHECK_ARG(T <= 65535 && (int64_t)S * P <= 65535,
                "sds_thermal_power_full: T or S*P exceeds 65535");
  if ((int64_t)S * P * B * T == 0) return SDS_OK;
  const int ntile = (B + sds::TH_TILE - 1) / sds::TH_TILE;
  dim3 grid(ntile * ntile, T, S * P);
  sds::thermal_full_kernel<<<grid, sds::TH_TILE * sds::TH_TILE, 0, (cudaStream_t)stream>>>(
      n1, n2, nsample_dtype, stride_s, stride_p, coef, freq, int_time, pol_factor, scale_sp, P, B, T, F, out);
  SDS_LAUNCH_CHECK("thermal_full_kernel");
  return SDS_OK;
}
