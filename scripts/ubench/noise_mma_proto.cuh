// PROTOTYPE kept as evidence (profiles/r2_mma_dft_microbench.txt): the engine-integrated form of the tensor-pipe noise
// transform that was measured and NOT adopted (an HMMA holds the issue port for 8 clk on B200).  Not part of libsds.so.
// The noise leg's length-256 transform on the tensor pipe (the one pipe the fused engine leaves
// idle): 256 = 16 x 16, two DFT-16 stages as warp-level matrix products (mma.sync m16n8k16, bf16
// operands, fp32 accumulation) with the W256 twiddles applied to the fp32 accumulators in between.
// One row per warp, operands and results in registers: the accumulator fragment of stage 1 IS the
// A fragment of stage 2, so the transform needs no shared-memory exchange at all.
//
// Why the legacy mma.sync and not tcgen05: the operands are generated in registers (a Philox draw
// per row) and the results are consumed in registers (pair sums); tcgen05.mma takes its operands
// from shared memory and leaves its result in tensor memory, i.e. two more round trips per row for
// 65 kflop of work.  Measured on B200 (profiles/r2_mma_dft_microbench.txt): HMMA.16816 issues every
// 8 clk per SM sub-partition whatever the operand type, next to FFMA2 work on the FP32 pipe.
//
// Precision: the noise realisation is statistical.  bf16 samples keep the fp32
// exponent range (no scaling contract on sigma); the constant DFT-16 matrix is split into a bf16 hi
// and a bf16 lo part (|entry|^2 = 1 to 1e-5 -- a single bf16 matrix biases the noise power by 2e-3,
// measured), so a stage is 8 + 8 HMMA.
//
// Lane (g = lane / 4, t = lane % 4) and the m16n8k16 fragment layouts (PTX ISA):
//   A (16x16): a0 (row g, k 2t..2t+1), a1 (row g+8, same k), a2 (row g, k 2t+8..2t+9), a3 (row g+8)
//   B (16x8):  b0 (k 2t..2t+1, col g), b1 (k 2t+8..2t+9, col g)
//   C (16x8):  c0 (row g, col 2t), c1 (row g, col 2t+1), c2 / c3 the same for row g+8
// Index maps (x[c], c = 16 n1 + n2;  X[k], k = k1 + 16 k2):
//   stage 1   Y[k1][n2] = sum_n1 F[k1][n1] x[n1][n2]        A = F, B = x
//   twiddle   Z[k1][n2] = Y[k1][n2] W256^(k1 n2)
//   stage 2   X[k1][k2] = sum_n2 Z[k1][n2] F[n2][k2]        A = Z, B = F
// The k slots of stage 1 are permuted (slot 2t + a + 8b holds n1 = t + 4a + 8b) and the columns of
// its data operand likewise (column g + 8j holds n2 = (g >> 1) + 4 (g & 1) + 8j).  With that
// (1) a thread's eight channels are  cb + 8j + 64a + 128b,  cb = 16t + (g >> 1) + 4 (g & 1): the
// reads of the staged row are two-way bank conflicted instead of four-way, and (2) because F is
// symmetric the A fragments of stage 1 are exactly the B fragments of stage 2 (16 registers of
// constants in all).  A thread's eight delay bins are  g + 8 (e >> 1) + 16 (2t + (e & 1) + 8j).
#pragma once
#include "../../simpleds_b200/csrc/sds_packed.cuh"

namespace sds {

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (lo, hi) -> bf16x2, round to nearest even
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

struct NoiseMma256 {
  // shared-memory table written once per CTA: [8 chunks][32 lanes] of 16 bytes
  //   chunks 0..3: Re hi, Re lo, Im hi, Im lo fragments of exp(-2 pi i k n / 16) (4 registers each)
  //   chunks 4..7: the thread's 8 twiddles W256^(k1 n2), two (re, im) pairs per chunk
  static constexpr int TABLE_BYTES = 8 * 32 * 16;
  static constexpr int N = 256;

  // channel offset of input slot r = 4j + a + 2b relative to cb(lane)
  static __host__ __device__ constexpr int chan_off(int r) {
    return 8 * (r >> 2) + 64 * (r & 1) + 128 * ((r >> 1) & 1);
  }
  static __device__ __forceinline__ int chan_base(int lane) {
    return 16 * (lane & 3) + (lane >> 3) + 4 * ((lane >> 2) & 1);
  }
  // delay bin of output slot q = 4j + e relative to bin_base(lane)
  static __host__ __device__ constexpr int bin_off(int q) {
    return 8 * ((q >> 1) & 1) + 16 * ((q & 1) + 8 * (q >> 2));
  }
  static __device__ __forceinline__ int bin_base(int lane) { return (lane >> 2) + 32 * (lane & 3); }

  // one warp fills the table (call before a __syncthreads)
  static __device__ void fill_table(uint4 *tab, int lane) {
    const int g = lane >> 2, t = lane & 3;
    uint32_t rh[4], rl[4], ih[4], il[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int row = g + 8 * (q & 1), n0 = t + 8 * (q >> 1), n1 = n0 + 4;
      float c0, s0, c1, s1;
      sincospif(-(float)((row * n0) & 15) / 8.0f, &s0, &c0);
      sincospif(-(float)((row * n1) & 15) / 8.0f, &s1, &c1);
      rh[q] = pack_bf16(c0, c1);
      ih[q] = pack_bf16(s0, s1);
      rl[q] = pack_bf16(c0 - bf16_lo(rh[q]), c1 - bf16_hi(rh[q]));
      il[q] = pack_bf16(s0 - bf16_lo(ih[q]), s1 - bf16_hi(ih[q]));
    }
    tab[0 * 32 + lane] = make_uint4(rh[0], rh[1], rh[2], rh[3]);
    tab[1 * 32 + lane] = make_uint4(rl[0], rl[1], rl[2], rl[3]);
    tab[2 * 32 + lane] = make_uint4(ih[0], ih[1], ih[2], ih[3]);
    tab[3 * 32 + lane] = make_uint4(il[0], il[1], il[2], il[3]);
    float tw[16];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {  // stage-1 output (j, e): k1 = g + 8 (e >> 1), n2 = t + 4 (e & 1) + 8j
        const int k1 = g + 8 * (e >> 1), n2 = t + 4 * (e & 1) + 8 * j;
        sincospif(-(float)((k1 * n2) & 255) / 128.0f, &tw[2 * (4 * j + e) + 1], &tw[2 * (4 * j + e)]);
      }
#pragma unroll
    for (int c = 0; c < 4; ++c)
      tab[(4 + c) * 32 + lane] = make_uint4(__float_as_uint(tw[4 * c]), __float_as_uint(tw[4 * c + 1]),
                                            __float_as_uint(tw[4 * c + 2]), __float_as_uint(tw[4 * c + 3]));
  }

  struct Frags {
    uint32_t rh[4], rl[4], ih[4], il[4];
  };
  static __device__ __forceinline__ void load_frags(Frags &F, const uint4 *tab, int lane) {
    const uint4 c0 = tab[lane], c1 = tab[32 + lane], c2 = tab[64 + lane], c3 = tab[96 + lane];
    F.rh[0] = c0.x; F.rh[1] = c0.y; F.rh[2] = c0.z; F.rh[3] = c0.w;
    F.rl[0] = c1.x; F.rl[1] = c1.y; F.rl[2] = c1.z; F.rl[3] = c1.w;
    F.ih[0] = c2.x; F.ih[1] = c2.y; F.ih[2] = c2.z; F.ih[3] = c2.w;
    F.il[0] = c3.x; F.il[1] = c3.y; F.il[2] = c3.z; F.il[3] = c3.w;
  }
  // The transform in three pieces, so that the caller can issue other (FP32-pipe) work of the same
  // warp between them while the tensor pipe runs:
  //   stage1: xr[j][h], xi[j][h] (bf16x2 fragments of the row: column tile j; h = 0: k slots 2t, 2t+1,
  //           h = 1: 2t+8, 2t+9; low half = the even slot)  ->  yr / yi [j][e] fp32 accumulators
  //   twiddle_pack: W256 twiddles on the accumulators, rounded to the bf16 A fragments of stage 2
  //   stage2: -> vr / vi [j][e], the row's transform at bins bin_base + bin_off(4j + e)
  static __device__ __forceinline__ void stage1(const uint32_t (&xr)[2][2], const uint32_t (&xi)[2][2],
                                                const Frags &F, float (&yr)[2][4], float (&yi)[2][4]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) yr[j][e] = yi[j][e] = 0.f;
      const uint32_t nx0 = xi[j][0] ^ 0x80008000u, nx1 = xi[j][1] ^ 0x80008000u;
      mma_bf16_16816(yr[j], F.rh, xr[j][0], xr[j][1]);
      mma_bf16_16816(yi[j], F.ih, xr[j][0], xr[j][1]);
      mma_bf16_16816(yr[j], F.ih, nx0, nx1);
      mma_bf16_16816(yi[j], F.rh, xi[j][0], xi[j][1]);
      mma_bf16_16816(yr[j], F.rl, xr[j][0], xr[j][1]);
      mma_bf16_16816(yi[j], F.il, xr[j][0], xr[j][1]);
      mma_bf16_16816(yr[j], F.il, nx0, nx1);
      mma_bf16_16816(yi[j], F.rl, xi[j][0], xi[j][1]);
    }
  }
  static __device__ __forceinline__ void twiddle_pack(const float (&yr)[2][4], const float (&yi)[2][4],
                                                      const uint4 *tab, int lane, uint32_t (&zr)[4],
                                                      uint32_t (&zi)[4]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint4 ta = tab[(4 + 2 * j) * 32 + lane], tb = tab[(5 + 2 * j) * 32 + lane];
      const f2 w[4] = {f2_make(__uint_as_float(ta.x), __uint_as_float(ta.y)),
                       f2_make(__uint_as_float(ta.z), __uint_as_float(ta.w)),
                       f2_make(__uint_as_float(tb.x), __uint_as_float(tb.y)),
                       f2_make(__uint_as_float(tb.z), __uint_as_float(tb.w))};
      f2 z[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) z[e] = f2_cmul(f2_make(yr[j][e], yi[j][e]), w[e]);
      zr[2 * j] = pack_bf16(f2_lo(z[0]), f2_lo(z[1]));
      zr[2 * j + 1] = pack_bf16(f2_lo(z[2]), f2_lo(z[3]));
      zi[2 * j] = pack_bf16(f2_hi(z[0]), f2_hi(z[1]));
      zi[2 * j + 1] = pack_bf16(f2_hi(z[2]), f2_hi(z[3]));
    }
  }
  static __device__ __forceinline__ void stage2(const uint32_t (&zr)[4], const uint32_t (&zi)[4],
                                                const Frags &F, float (&vr)[2][4], float (&vi)[2][4]) {
    uint32_t nzi[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) nzi[q] = zi[q] ^ 0x80008000u;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) vr[j][e] = vi[j][e] = 0.f;
      mma_bf16_16816(vr[j], zr, F.rh[j], F.rh[2 + j]);
      mma_bf16_16816(vi[j], zr, F.ih[j], F.ih[2 + j]);
      mma_bf16_16816(vr[j], nzi, F.ih[j], F.ih[2 + j]);
      mma_bf16_16816(vi[j], zi, F.rh[j], F.rh[2 + j]);
      mma_bf16_16816(vr[j], zr, F.rl[j], F.rl[2 + j]);
      mma_bf16_16816(vi[j], zr, F.il[j], F.il[2 + j]);
      mma_bf16_16816(vr[j], nzi, F.il[j], F.il[2 + j]);
      mma_bf16_16816(vi[j], zi, F.rl[j], F.rl[2 + j]);
    }
  }
};

// Eight complex draws per thread and row for the tensor formulation, from ONE Philox4x32-7 call
// (counter = (row_id low word, row_id high word, lane, 1); the fourth word separates this stream
// from engine_draw8's).  The transform sums >= 92 effective channels (Blackman-Harris, N = 256), so
// the delay-domain noise is Gaussian by the central limit theorem whatever the per-channel law
// (excess kurtosis -0.015; tests/test_gpu_noise_stats.py): each component is  +-(1 + m / 128),
// m = 7 random bits, sign = 1 random bit -- a bf16 value built by ONE LOP3 (no conversion, no
// MUFU), symmetric about zero, second moment NOISE_MMA_U2.  Word i of the call gives the bf16x2
// registers d[2i] (bits 0-6, 15 | 16-22, 31) and d[2i+1] (the same fields of the word rotated right
// by 8); d[4j + 2h] is the real and d[4j + 2h + 1] the imaginary fragment of column tile j, half h.
constexpr double NOISE_MMA_U2 = 2.321624755859375;               // mean of (1 + m/128)^2, m = 0..127
constexpr float NOISE_MMA_NORM = 0.4640758717303815f;            // 1 / sqrt(2 NOISE_MMA_U2)
template <int ROUNDS>
__device__ __forceinline__ void noise_draw_bf16(uint64_t row_id, int lane, const uint32_t (&key)[16],
                                                uint32_t (&d)[8]) {
  const u32x4 rn = philox4x32_keyed<ROUNDS>(
      {(uint32_t)row_id, (uint32_t)(row_id >> 32), (uint32_t)lane, 1u}, key);
  const uint32_t w[4] = {rn.x, rn.y, rn.z, rn.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    d[2 * i] = (w[i] & 0x807F807Fu) | 0x3F803F80u;
    d[2 * i + 1] = (__funnelshift_r(w[i], w[i], 8) & 0x807F807Fu) | 0x3F803F80u;
  }
}

}  // namespace sds
