// Shared device/host helpers for libsds.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/sds.h"

namespace sds {

// ---- error channel (thread-local, no exceptions cross the ABI) ------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *where);

#define SDS_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      sds::set_error(__VA_ARGS__);      \
      return SDS_EINVAL;                \
    }                                   \
  } while (0)

#define SDS_CUDA(call)                                   \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return sds::cuda_fail(e__, #call); \
  } while (0)

#define SDS_LAUNCH_CHECK(name)                                   \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return sds::cuda_fail(e__, name);    \
  } while (0)

// ---- per-device one-time kernel configuration ----------------------------------------------
// cudaFuncSetAttribute is per device: a process that drives several GPUs must set the dynamic
// shared-memory limit on each of them.  One latch per (kernel instantiation, device); the race
// of two host threads setting the same attribute twice is benign.
struct DeviceLatch {
  bool done[64] = {};
  bool need(int *dev) {
    if (cudaGetDevice(dev) != cudaSuccess || *dev < 0 || *dev >= 64) return true;
    return !done[*dev];
  }
  void set(int dev) {
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
};

// Resident CTAs per SM of a kernel that allocates tensor memory.  cudaOccupancyMaxActiveBlocksPerMultiprocessor
// answers 1 for such a kernel whatever it allocates, although CTAs whose columns fit the SM's 512 are resident
// together (scripts/ubench/tmem_xchg.cu), so the budget is worked out from the kernel's own resources:
// registers (allocated per warp in units of 8 per thread), shared memory (+ 1 KB the system reserves per CTA),
// tensor-memory columns, 32 CTAs / 64 warps per SM.
template <typename K>
inline int tmem_ctas_per_sm(K kernel, int threads, int dyn_smem, int tm_cols) {
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return 1;
  int dev = 0, regs_sm = 65536, smem_sm = 233472;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
  cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  const int warps = (threads + 31) / 32;
  const int regs_cta = ((fa.numRegs + 7) / 8 * 8) * 32 * warps;
  int n = regs_cta > 0 ? regs_sm / regs_cta : 32;
  const int smem_cta = dyn_smem + (int)fa.sharedSizeBytes + 1024;
  if (smem_sm / smem_cta < n) n = smem_sm / smem_cta;
  if (tm_cols > 0 && 512 / tm_cols < n) n = 512 / tm_cols;
  if (64 / warps < n) n = 64 / warps;
  if (n > 32) n = 32;
  return n < 1 ? 1 : n;
}

// ---- complex helpers -------------------------------------------------------
template <typename T> struct cpx { T x, y; };
template <typename T> struct vec2;
template <> struct vec2<float> { using type = float2; };
template <> struct vec2<double> { using type = double2; };

template <typename T>
__device__ __forceinline__ cpx<T> cmul(cpx<T> a, cpx<T> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T>
__device__ __forceinline__ cpx<T> cmul_conj(cpx<T> a, cpx<T> b) {  // a * conj(b)
  return {a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y};
}
template <typename T>
__device__ __forceinline__ cpx<T> cadd(cpx<T> a, cpx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T>
__device__ __forceinline__ cpx<T> csub(cpx<T> a, cpx<T> b) { return {a.x - b.x, a.y - b.y}; }

// load a complex element of runtime dtype (SDS_C64 / SDS_C128) as cpx<T>
template <typename T>
__device__ __forceinline__ cpx<T> load_cpx(const void *p, int dtype, int64_t i) {
  if (dtype == SDS_C64) {
    float2 v = reinterpret_cast<const float2 *>(p)[i];
    return {(T)v.x, (T)v.y};
  }
  double2 v = reinterpret_cast<const double2 *>(p)[i];
  return {(T)v.x, (T)v.y};
}
template <typename T>
__device__ __forceinline__ void store_cpx(void *p, int dtype, int64_t i, cpx<T> v) {
  if (dtype == SDS_C64) {
    reinterpret_cast<float2 *>(p)[i] = make_float2((float)v.x, (float)v.y);
  } else {
    reinterpret_cast<double2 *>(p)[i] = make_double2((double)v.x, (double)v.y);
  }
}
__device__ __forceinline__ double load_real(const void *p, int dtype, int64_t i) {
  return dtype == SDS_F32 ? (double)reinterpret_cast<const float *>(p)[i]
                          : reinterpret_cast<const double *>(p)[i];
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based -----------------------
struct u32x4 { uint32_t x, y, z, w; };
__host__ __device__ __forceinline__ u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
    u32x4 n;
    n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
    n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
    n.w = (uint32_t)p0;
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// ---- plan --------------------------------------------------------------------
struct Plan {
  int nfreq;
  int math;               // SDS_F32 | SDS_F64
  int nradix;             // generic Stockham factorisation
  int radix[16];
  double *taper_d;        // [N] fp64
  float *taper_f;         // [N] fp32
  double2 *twiddle_d;     // [N] exp(-2 pi i k / N)
  float2 *twiddle_f;
  cpx<float> *team_tw_f;   // sds_radix.cuh tables (power-of-two lengths 64..4096)
  cpx<double> *team_tw_d;
  int is_pow2;
  // Bluestein (chirp-z) tables for other lengths: the length-N transform as a circular
  // convolution of power-of-two length blue_m >= 2N - 1 (sds_fft_bluestein.cu)
  int blue_m;                 // 0: not available
  cpx<float> *blue_chirp_f;   // [N]  exp(-i pi n^2 / N)
  cpx<double> *blue_chirp_d;
  cpx<float> *blue_bhat_f;    // [M]  FFT_M of the wrapped exp(+i pi m^2 / N), |m| < N
  cpx<double> *blue_bhat_d;
  cpx<float> *blue_tw_f;      // team twiddles of length M
  cpx<double> *blue_tw_d;
};

}  // namespace sds

struct sds_plan { sds::Plan p; };
