// Declarations shared by the fused-engine translation units (sds_engine.cu: C ABI,
// setup kernel, fp64 kernel; sds_engine_f32.cu: the packed-fp32 kernel).
#pragma once
#include "sds_radix.cuh"

namespace sds {

constexpr int ENG_STAGES = 3;
#ifndef ENG_MINBLOCKS
#define ENG_MINBLOCKS 1
#endif
constexpr int LEG_DATA = 1, LEG_NOISE = 2, LEG_THERMAL = 4;

struct EngineParams {
  const float2 *vis;
  const uint8_t *flags;
  const float *nsample;
  const float *int_time;
  const int32_t *goff;
  const float *sigma_coef;
  const float2 *noise_in;
  const double *pol_factor;
  const void *taper;      // T[N]
  const void *tw;         // cpx<T>[fft_tw_total(N)]
  const float *taper_f;   // float versions for the noise leg
  const float2 *tw_f;
  // workspace tables (engine_setup_kernel)
  const int32_t *item_start;  // [ngroup + 1]
  const double *wl, *wr;      // [nspw][npol][N] trapezoid-weighted thermal coefficients
  int *counter;
  double *pow_all, *pow_auto, *noise_all, *noise_auto, *th_all, *th_auto;
  int64_t time_offset, ntime_total, bl_offset, nbl_total;
  double delta_f;
  uint32_t seed_lo, seed_hi;
  int npol, nbl, ntime, nchan, nspw, ngroup, noise_mode, target_rows;
  int win_start[SDS_MAX_WINDOWS];
};

// ---- PTX helpers (mbarrier + 1-D bulk TMA) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rsqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <typename T> __device__ __forceinline__ T rsqrt_t(float x);
template <> __device__ __forceinline__ float rsqrt_t<float>(float x) { return fast_rsqrt(x); }
template <> __device__ __forceinline__ double rsqrt_t<double>(float x) { return rsqrt((double)x); }

// chunk(g) = clamp(target_rows / n_g, 1, ntime); items(g) = ceil(ntime / chunk) nspw npol
__device__ __host__ __forceinline__ int eng_chunk(int ng, int ntime, int target_rows) {
  int c = ng > 0 ? target_rows / ng : ntime;
  return c < 1 ? 1 : (c > ntime ? ntime : c);
}

// fp32 engine (sds_engine_f32.cu); returns an sds_status / cudaError_t, counts its launches
int run_engine_f32(int N, const EngineParams &P, int legs, cudaStream_t st, int *launches);

}  // namespace sds
