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
constexpr int LEG_THERMAL64 = 8;  // fp32 engine only: thermal sums in fp64 (complex128 mode)

struct EngineParams {
  const float2 *vis;
  const uint8_t *flags;
  const float *nsample;
  const float *int_time;
  const int32_t *goff;
  const int32_t *group_id;    // [ngroup] output index of each group, or NULL (identity)
  const int64_t *group_bl0;   // [ngroup] global index of each group's first baseline, or NULL (bl_offset + goff)
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
  uint32_t pkey[16];          // Philox round keys: pkey[2 r] = seed_lo + r W0, pkey[2 r + 1] = seed_hi + r W1
  int npol, nbl, ntime, nchan, nspw, ngroup, noise_mode, target_rows;
  int zero;                   // always 0: lets a kernel form addresses ptxas cannot prove warp-uniform
  int noise_gauss;            // noise_mode 1: 0 = moment-matched discrete draw (engine_draw_mm), 1 = Box-Muller (engine_draw8)
  int win_start[SDS_MAX_WINDOWS];
  // windows whose length is not a power of two (sds_engine_blue.cu): chirp-z tables of the plan
  int nfreq;                  // window length N
  const float2 *blue_chirp;   // [N]
  const float2 *blue_bhat;    // [M]
  const float2 *blue_tw;      // team twiddles of length M
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
// the same on shared-space addresses computed once (no generic-to-shared conversion per use)
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d_a(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
// one lane of a converged warp (elect.sync): unlike `lane == 0`, the compiler knows exactly one
// thread runs the guarded code, so bulk-copy operands stay in uniform registers
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
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
// fp64 x^-1/2 of a positive fp32 value: the fp32 approximation (2 ulp) refined by two Newton
// steps y <- y (1.5 - 0.5 x y^2) in fp64 (relative error 1e-7 -> 2e-14 -> 1e-16); a third of the
// instructions of the library rsqrt(double), which also handles specials this path never sees
template <> __device__ __forceinline__ double rsqrt_t<double>(float x) {
  const double xd = (double)x, hx = 0.5 * xd;
  double y = (double)fast_rsqrt(x);
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}

// ---- the fused engines' noise draw ------------------------------------------------
#ifndef ENG_PHILOX_ROUNDS
#define ENG_PHILOX_ROUNDS 7  // smallest Crush-resistant Philox4x32 (Salmon et al. 2011, table 2)
#endif
// Philox4x32 with the round keys taken from the kernel parameters (constant bank operands:
// no key-schedule arithmetic in the loop)
template <int ROUNDS>
__device__ __forceinline__ u32x4 philox4x32_keyed(u32x4 c, const uint32_t (&key)[16]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
    u32x4 n;
    n.x = (uint32_t)(p1 >> 32) ^ c.y ^ key[2 * r];
    n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ c.w ^ key[2 * r + 1];
    n.w = (uint32_t)p0;
    c = n;
  }
  return c;
}
__host__ __device__ __forceinline__ void philox_fill_keys(uint32_t seed_lo, uint32_t seed_hi,
                                                          uint32_t (&key)[16]) {
  for (int r = 0; r < 8; ++r) {
    key[2 * r] = seed_lo + (uint32_t)r * 0x9E3779B9u;
    key[2 * r + 1] = seed_hi + (uint32_t)r * 0xBB67AE85u;
  }
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sin(float x) {
  float y;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_cos(float x) {
  float y;
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Eight unit draws of the thread that owns channels tau + TPR r (r = 0..7) of row `row_id`:
// rad2 = sqrt(-log2 u1) (the caller folds sqrt(ln 2) into its amplitude), (cs, sn) =
// exp(2 pi i u2).  A row of N = 8 TPR draws takes 3N/8 Philox calls; the 128-bit counter of call
// k = 0..2 of thread tau is (row_id low word, row_id high word, 3 tau + k, 0) -- no per-row 64-bit
// counter arithmetic, and the first round's product with the row words serves all three calls.
// The thread's 12 words are cut into eight 46-bit draws (23-bit radius uniform, 23-bit angle).  Each field is OR-ed under the exponent of 1.0f (one LOP3 builds
// the float f in [1, 2), no integer conversion): u1 = 2 - f in (0, 1], angle = 2 pi f (mod 2 pi).
// Per word triple (w0, w1, w2) = W[3m .. 3m+2], draws r = 2m (even) and 2m + 1 (odd):
//   even: radius field w0[22:0],  angle field w1[22:0]
//   odd : radius field w2[22:0],  angle field w2[30:24] : w1[31:24] : w0[31:24]   (two PRMT)
// IMAD.WIDE, 4 FMA-pipe cycles on B200, is the most expensive instruction of the kernel, so a row
// takes three Philox calls per thread and not four.
constexpr float ENG_SQRT_LN2 = 0.83255461115769775635f;
__device__ __forceinline__ float unit_float_23(uint32_t bits) {  // [1, 2) from the low 23 bits
  return __uint_as_float((bits & 0x007FFFFFu) | 0x3F800000u);
}
template <int TPR>
__device__ __forceinline__ void engine_draw8(uint64_t row_id, int tau, const uint32_t (&key)[16],
                                             float (&rad2)[8], float (&cs)[8], float (&sn)[8]) {
  uint32_t W[12];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const u32x4 rn = philox4x32_keyed<ENG_PHILOX_ROUNDS>(
        {(uint32_t)row_id, (uint32_t)(row_id >> 32), (uint32_t)(3 * tau + k), 0u}, key);
    W[4 * k] = rn.x; W[4 * k + 1] = rn.y; W[4 * k + 2] = rn.z; W[4 * k + 3] = rn.w;
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const uint32_t w0 = W[3 * m], w1 = W[3 * m + 1], w2 = W[3 * m + 2];
    const uint32_t hi = __byte_perm(__byte_perm(w0, w1, 0x0073), w2, 0x0710);
    const float fr[2] = {unit_float_23(w0), unit_float_23(w2)};
    const float fa[2] = {unit_float_23(w1), unit_float_23(hi)};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      rad2[2 * m + h] = fast_sqrt(-fast_lg2(2.0f - fr[h]));
      const float ang = fa[h] * 6.2831853071795865f;
      cs[2 * m + h] = fast_cos(ang);
      sn[2 * m + h] = fast_sin(ang);
    }
  }
}

// ---- the production draw of the power-of-two engine: moment-matched discrete noise ------------
// The engine never exposes the frequency-domain noise; what leaves it are delay-domain pair sums,
// i.e. sums over >= 64 tapered channels of every row.  By the central limit theorem the law of the
// transformed noise is fixed by the per-channel moments, so the per-channel draw is the cheapest law
// that matches the complex Gaussian sigma (g1 + i g2) / sqrt(2) of utils.py:501-505 in every moment
// the pair sums see at leading order:
//     n[c] = sigma[c] * B * (s1 + i s2),   B in {0, 1} (p = 1/2),  s1, s2 in {-1, +1}
//   E n = 0,  E n^2 = 0 (circular),  E |n|^2 = sigma^2,  E |n|^4 = 2 sigma^4 (= the Gaussian value;
//   a uniform or a constant-modulus draw has 1.4 and 1), so that mean AND variance of every |X[k]|^2
//   and of every cross product agree with the Gaussian realisation up to O(1 / N_eff^2)
//   (tests/test_gpu_noise_stats.py measures them).  The covariance between delay bins -- taper, flag
//   comb and nsample profile of the row -- is exact, because the per-channel amplitudes are.
// Three random bits per sample and no arithmetic: the on/off bit rides on the flag select, the two
// signs are XOR-ed into the amplitude.  One Philox4x32-7 call serves the same thread of FOUR
// consecutive baselines: counter = (blk low word, blk high word, tau, 2) with
//   blk = ((s npol + p) ceil(nbl_total / 4) + (b >> 2)) ntime_total + t      (b, t global indices)
// and baseline b takes word b & 3 of the call.  Bits of the word, for the thread's channel tau + TPR r:
//   bit r: on / off,  bit 8 + r: sign of the real part,  bit 16 + r: sign of the imaginary part.
// the call of block `blk`, and word `sel` of it
template <int ROUNDS>
__device__ __forceinline__ u32x4 engine_call_mm(uint64_t blk, int tau, const uint32_t (&key)[16]) {
  return philox4x32_keyed<ROUNDS>({(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)tau, 2u}, key);
}
__device__ __forceinline__ uint32_t engine_pick_mm(const u32x4 &w, uint32_t sel) {
  const uint32_t a = (sel & 1u) ? w.y : w.x, b = (sel & 1u) ? w.w : w.z;
  return (sel & 2u) ? b : a;
}
// the two multipliers of a sample: dsel = unflagged ? a : 0 (data), nsel = unflagged and on-bit ? b : 0
// (noise); one predicate chain (flag test, bit test AND-ed into it) instead of three selects
__device__ __forceinline__ void engine_selects_mm(uint32_t flag, uint32_t word, uint32_t bit, float a, float b,
                                                  float &dsel, float &nsel) {
  asm("{\n"
      ".reg .pred pk, pn;\n"
      ".reg .b32 t;\n"
      "setp.eq.u32 pk, %2, 0;\n"
      "and.b32 t, %3, %4;\n"
      "setp.ne.and.u32 pn, t, 0, pk;\n"
      "selp.f32 %0, %5, 0f00000000, pk;\n"
      "selp.f32 %1, %6, 0f00000000, pn;\n"
      "}\n"
      : "=f"(dsel), "=f"(nsel)
      : "r"(flag), "r"(word), "r"(bit), "f"(a), "f"(b));
}
// (re, im) = (+-amp, +-amp) for sample r of the word
__device__ __forceinline__ unsigned long long engine_signs_mm(float amp, uint32_t word, int r) {
  const uint32_t ab = __float_as_uint(amp);
  const uint32_t re = ab ^ ((word << (23 - r)) & 0x80000000u);
  const uint32_t im = ab ^ ((word << (15 - r)) & 0x80000000u);
  return (unsigned long long)re | ((unsigned long long)im << 32);
}

// chunk(g) = clamp(target_rows / n_g, 1, ntime); items(g) = ceil(ntime / chunk) nspw npol
__device__ __host__ __forceinline__ int eng_chunk(int ng, int ntime, int target_rows) {
  int c = ng > 0 ? target_rows / ng : ntime;
  return c < 1 ? 1 : (c > ntime ? ntime : c);
}

// fp32 engine (sds_engine_f32.cu); returns an sds_status / cudaError_t, counts its launches
int run_engine_f32(int N, const EngineParams &P, int legs, cudaStream_t st, int *launches);
// fp32 engine for window lengths that are not powers of two (sds_engine_blue.cu), M = the plan's
// convolution length
int run_engine_blue(int M, const EngineParams &P, int legs, cudaStream_t st, int *launches);
// fp32 engine for two data sets (sds_engine_uv2.cu): leg = LEG_DATA or LEG_NOISE, one launch each
int run_engine_uv2(int N, const EngineParams &P, int leg, cudaStream_t st, int *launches);

// Four unit draws of the thread that owns channels tau + TPR r (r = 0..3) of row `row_id` in the
// chirp-z engine (N <= 4 TPR): two Philox calls, counter = (row_id low word, row_id high word,
// 2 tau + k, 0); draw r takes word 2r for the radius field and word 2r + 1 for the angle field
// (low 23 bits each, under the exponent of 1.0f as in engine_draw8).
template <int TPR>
__device__ __forceinline__ void engine_draw4(uint64_t row_id, int tau, const uint32_t (&key)[16],
                                             float (&rad2)[4], float (&cs)[4], float (&sn)[4]) {
  uint32_t W[8];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const u32x4 rn = philox4x32_keyed<ENG_PHILOX_ROUNDS>(
        {(uint32_t)row_id, (uint32_t)(row_id >> 32), (uint32_t)(2 * tau + k), 0u}, key);
    W[4 * k] = rn.x; W[4 * k + 1] = rn.y; W[4 * k + 2] = rn.z; W[4 * k + 3] = rn.w;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    rad2[r] = fast_sqrt(-fast_lg2(2.0f - unit_float_23(W[2 * r])));
    const float ang = unit_float_23(W[2 * r + 1]) * 6.2831853071795865f;
    cs[r] = fast_cos(ang);
    sn[r] = fast_sin(ang);
  }
}

}  // namespace sds
