// Packed fp32x2 complex arithmetic for sm_100a and the packed version of the
// 8-elements-per-thread Stockham schedule of sds_radix.cuh.
//
// Blackwell issues add/mul/fma on PAIRS of fp32 held in one 64-bit register
// (PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2).  A complex sample
// (re, im) is such a pair, so a complex add is ONE instruction and a complex
// multiply TWO; lane swaps, per-lane negation and scalar broadcast are operand
// modifiers of FADD2/FFMA2 (.LO_HI, .NP, .F32), so "multiply by -i" costs nothing.
// The fused engine is bound by instruction issue, not by the FMA pipe, which is
// what makes the packed forms pay.
#pragma once
#include "sds_radix.cuh"

namespace sds {

typedef unsigned long long f2;  // (lo, hi) = (re, im)

__device__ __forceinline__ f2 f2_make(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float f2_lo(f2 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float f2_hi(f2 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) {
  f2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f2 f2_scale(f2 a, float s) { return f2_mul(a, f2_make(s, s)); }
// a * (-i) = (im, -re): folds into the .LO_HI.NP operand modifier of the consumer
__device__ __forceinline__ f2 f2_mul_mi(f2 a) { return f2_make(f2_hi(a), -f2_lo(a)); }
// complex product a * w
__device__ __forceinline__ f2 f2_cmul(f2 a, f2 w) {
  const float wx = f2_lo(w), wy = f2_hi(w);
  return f2_fma(f2_make(-f2_hi(a), f2_lo(a)), f2_make(wy, wy), f2_mul(a, f2_make(wx, wx)));
}
// acc + (re^2, im^2): |a|^2 is the sum of the two lanes of the accumulator
__device__ __forceinline__ f2 f2_norm_acc(f2 acc, f2 a) { return f2_fma(a, a, acc); }
__device__ __forceinline__ float f2_hsum(f2 a) { return f2_lo(a) + f2_hi(a); }
__device__ __forceinline__ f2 f2_shfl_xor(f2 v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }

// In-place decimation-in-frequency butterflies (natural order in, bit-reversed out)
template <int R> struct RadixP;
template <> struct RadixP<1> {
  static __device__ __forceinline__ void run(f2 *) {}
};
template <> struct RadixP<2> {
  static __device__ __forceinline__ void run(f2 *a) {
    const f2 t = a[0];
    a[0] = f2_add(t, a[1]);
    a[1] = f2_sub(t, a[1]);
  }
};
template <> struct RadixP<4> {
  static __device__ __forceinline__ void run(f2 *a) {
    const f2 u0 = f2_add(a[0], a[2]), d0 = f2_sub(a[0], a[2]);
    const f2 u1 = f2_add(a[1], a[3]), d1 = f2_mul_mi(f2_sub(a[1], a[3]));
    a[0] = f2_add(u0, u1);
    a[1] = f2_sub(u0, u1);
    a[2] = f2_add(d0, d1);
    a[3] = f2_sub(d0, d1);
  }
};
template <> struct RadixP<8> {
  static __device__ __forceinline__ void run(f2 *a) {
    const float h = 0.70710678118654752440f;
    f2 d[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      d[n] = f2_sub(a[n], a[n + 4]);
      a[n] = f2_add(a[n], a[n + 4]);
    }
    a[4] = d[0];
    a[5] = f2_scale(f2_add(d[1], f2_mul_mi(d[1])), h);   // * W8^1
    a[6] = f2_mul_mi(d[2]);                              // * W8^2
    a[7] = f2_scale(f2_sub(f2_mul_mi(d[3]), d[3]), h);   // * W8^3
    RadixP<4>::run(a);
    RadixP<4>::run(a + 4);
  }
};

// Packed twin of TeamFft<float, N> (sds_radix.cuh): same thread/element ownership, same
// twiddle table ([m][q-1][tau] of (re, im) pairs), same swizzled exchange buffers.
template <int N> struct TeamFftP {
  static constexpr int TPR = N / 8;
  static constexpr int NPASS = fft_npass(N);
  static constexpr int WORKER = TPR < 32 ? 32 : TPR;

  template <int P>
  static __device__ __forceinline__ void pass(f2 (&v)[8], f2 *bufA, f2 *bufB,
                                              const f2 *__restrict__ tw, int tau, int bar) {
    constexpr int R = fft_radix(N, P), NB = 8 / R, NS = fft_ns(N, P), LG = ilog2(R);
    if constexpr (P > 0) {
      constexpr int OFF = fft_tw_offset(N, P);
#pragma unroll
      for (int m = 0; m < NB; ++m)
#pragma unroll
        for (int q = 1; q < R; ++q)
          v[m + NB * q] = f2_cmul(v[m + NB * q], tw[OFF + (m * (R - 1) + (q - 1)) * TPR + tau]);
    }
    f2 *out = (P & 1) ? bufB : bufA;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
      f2 a[R];
#pragma unroll
      for (int q = 0; q < R; ++q) a[q] = v[m + NB * q];
      RadixP<R>::run(a);
      if constexpr (P < NPASS - 1) {
        const int j = tau + TPR * m;
        const int base = (j / NS) * (NS * R) + (j % NS);
#pragma unroll
        for (int p = 0; p < R; ++p) out[fft_pad(base + p * NS)] = a[brev(p, LG)];
      } else {
#pragma unroll
        for (int p = 0; p < R; ++p) v[m + NB * p] = a[brev(p, LG)];
      }
    }
    if constexpr (P < NPASS - 1) {
      team_sync<WORKER>(bar);
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = out[fft_pad(tau + TPR * r)];
      pass<P + 1>(v, bufA, bufB, tw, tau, bar);
    }
  }

  static __device__ __forceinline__ void run(f2 (&v)[8], f2 *bufA, f2 *bufB,
                                             const f2 *__restrict__ tw, int tau, int bar) {
    pass<0>(v, bufA, bufB, tw, tau, bar);
  }

  // Two independent transforms in lockstep (one team sync per pass serves both): twice
  // the instruction-level parallelism for the same number of barriers.  Each transform
  // has ONE exchange buffer (ba, bb); a second team sync after the reads of a pass makes
  // the buffer reusable by the next pass (cheap for warp-sized teams, and it keeps the
  // shared-memory footprint small enough for three CTAs per SM).
  template <int P>
  static __device__ __forceinline__ void pass2(f2 (&a)[8], f2 (&b)[8], f2 *ba, f2 *bb,
                                               const f2 *__restrict__ tw, int tau, int bar) {
    constexpr int R = fft_radix(N, P), NB = 8 / R, NS = fft_ns(N, P), LG = ilog2(R);
    if constexpr (P > 0) {
      constexpr int OFF = fft_tw_offset(N, P);
#pragma unroll
      for (int m = 0; m < NB; ++m)
#pragma unroll
        for (int q = 1; q < R; ++q) {
          const f2 w = tw[OFF + (m * (R - 1) + (q - 1)) * TPR + tau];
          a[m + NB * q] = f2_cmul(a[m + NB * q], w);
          b[m + NB * q] = f2_cmul(b[m + NB * q], w);
        }
    }
#pragma unroll
    for (int m = 0; m < NB; ++m) {
      f2 x[R], y[R];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        x[q] = a[m + NB * q];
        y[q] = b[m + NB * q];
      }
      RadixP<R>::run(x);
      RadixP<R>::run(y);
      if constexpr (P < NPASS - 1) {
        const int j = tau + TPR * m;
        const int base = (j / NS) * (NS * R) + (j % NS);
#pragma unroll
        for (int p = 0; p < R; ++p) {
          ba[fft_pad(base + p * NS)] = x[brev(p, LG)];
          bb[fft_pad(base + p * NS)] = y[brev(p, LG)];
        }
      } else {
#pragma unroll
        for (int p = 0; p < R; ++p) {
          a[m + NB * p] = x[brev(p, LG)];
          b[m + NB * p] = y[brev(p, LG)];
        }
      }
    }
    if constexpr (P < NPASS - 1) {
      team_sync<WORKER>(bar);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        a[r] = ba[fft_pad(tau + TPR * r)];
        b[r] = bb[fft_pad(tau + TPR * r)];
      }
      if constexpr (P < NPASS - 2) team_sync<WORKER>(bar);  // reads done before the next pass writes
      pass2<P + 1>(a, b, ba, bb, tw, tau, bar);
    }
  }
  // NOTE: the caller must team-sync between two run2 calls on the same buffers (the engine's
  // step loop does, at the top of every step).
  static __device__ __forceinline__ void run2(f2 (&a)[8], f2 (&b)[8], f2 *ba, f2 *bb,
                                              const f2 *__restrict__ tw, int tau, int bar) {
    pass2<0>(a, b, ba, bb, tw, tau, bar);
  }
};

// ---- TeamFftX: the same schedule with raw shared-memory addresses ---------------------------
// fft_pad is linear over XOR (a shift and a mask), and every index of the schedule is a per-thread
// part plus a compile-time part on disjoint bits (base + p NS, tau + TPR r), so
//   fft_pad(base + p NS) = fft_pad(base) ^ fft_pad(p NS).
// With the exchange buffers aligned to their own size (8 N bytes, a power of two) the byte address
// of an access is  (per-thread address, computed once per kernel) ^ (constant): ONE LOP3 per
// access, shared by the two lockstep transforms (their buffers are 8 N bytes apart: an immediate).
__device__ __forceinline__ void sts_f2(uint32_t addr, f2 v) {
  asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ f2 lds_f2(uint32_t addr) {
  f2 v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
// two packed values in one 128-bit access (16-byte aligned address)
__device__ __forceinline__ void sts_f2x2(uint32_t addr, f2 v, f2 w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"((uint32_t)v), "r"((uint32_t)(v >> 32)),
               "r"((uint32_t)w), "r"((uint32_t)(w >> 32))
               : "memory");
}
__device__ __forceinline__ void lds_f2x2(uint32_t addr, f2 &v, f2 &w) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v), "=l"(w) : "r"(addr) : "memory");
}
template <int OFF> __device__ __forceinline__ void sts_f2x2_off(uint32_t addr, f2 v, f2 w) {
  asm volatile("st.shared.v4.b32 [%0 + %5], {%1, %2, %3, %4};" ::"r"(addr), "r"((uint32_t)v),
               "r"((uint32_t)(v >> 32)), "r"((uint32_t)w), "r"((uint32_t)(w >> 32)), "n"(OFF)
               : "memory");
}
template <int OFF> __device__ __forceinline__ void lds_f2x2_off(uint32_t addr, f2 &v, f2 &w) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2 + %3];" : "=l"(v), "=l"(w) : "r"(addr), "n"(OFF) : "memory");
}
template <int OFF> __device__ __forceinline__ void sts_f2_off(uint32_t addr, f2 v) {
  asm volatile("st.shared.b64 [%0 + %2], %1;" ::"r"(addr), "l"(v), "n"(OFF) : "memory");
}
template <int OFF> __device__ __forceinline__ f2 lds_f2_off(uint32_t addr) {
  f2 v;
  asm volatile("ld.shared.b64 %0, [%1 + %2];" : "=l"(v) : "r"(addr), "n"(OFF) : "memory");
  return v;
}

// ---- tensor memory (TMEM) as an exchange medium between the threads of a warp ------------------
// tcgen05.st.32x32b writes thread t's registers to TMEM lane t; tcgen05.ld.16x256b reads the
// accumulator-fragment layout (thread t <- lanes t/4 and t/4 + 8, column pairs t%4 + 4j).  Store with
// the one, load with the other, and (thread, register) index bits are swapped -- the exchange an FFT
// pass needs -- without one wavefront of the shared-memory data pipe (which bounds the fused engine)
// and beside it: scripts/ubench/tmem_xchg.cu measures 14 clk per 4 KB warp exchange and SM against
// 32 clk through shared memory, and the two overlap (profiles/r2_tmem_exchange_microbench.txt).
__device__ __forceinline__ void tm_alloc32(uint32_t smem_dst) {  // one warp; 32 columns
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_dst) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_dealloc32(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// thread t -> lane t, columns [col, col + 16): eight packed values
__device__ __forceinline__ void tm_st8(uint32_t addr, f2 v0, f2 v1, f2 v2, f2 v3, f2 v4, f2 v5, f2 v6, f2 v7) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
      "r"((uint32_t)v0), "r"((uint32_t)(v0 >> 32)), "r"((uint32_t)v1), "r"((uint32_t)(v1 >> 32)), "r"((uint32_t)v2),
      "r"((uint32_t)(v2 >> 32)), "r"((uint32_t)v3), "r"((uint32_t)(v3 >> 32)), "r"((uint32_t)v4), "r"((uint32_t)(v4 >> 32)),
      "r"((uint32_t)v5), "r"((uint32_t)(v5 >> 32)), "r"((uint32_t)v6), "r"((uint32_t)(v6 >> 32)), "r"((uint32_t)v7),
      "r"((uint32_t)(v7 >> 32))
      : "memory");
}
// 16 lanes from the address's lane, 16 columns: thread t gets, as packed values,
//   d0 = (lane t/4, pair t%4), d1 = (lane t/4 + 8, pair t%4), d2 = (lane t/4, pair 4 + t%4), d3 = (lane t/4 + 8, pair 4 + t%4)
__device__ __forceinline__ void tm_ld4(uint32_t addr, f2 &d0, f2 &d1, f2 &d2, f2 &d3) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr)
               : "memory");
  d0 = (f2)r[0] | ((f2)r[1] << 32);
  d1 = (f2)r[2] | ((f2)r[3] << 32);
  d2 = (f2)r[4] | ((f2)r[5] << 32);
  d3 = (f2)r[6] | ((f2)r[7] << 32);
}

// thread-private tensor-memory columns (32x32b: thread t <-> lane t): register-file overflow space that
// costs ONE instruction per 8 / 16 / 32 words -- the fused engine keeps its accumulators there
template <int COLS> __device__ __forceinline__ void tm_alloc(uint32_t smem_dst) {  // one warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tm_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tm_ld_w8(uint32_t addr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void tm_ld_w16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void tm_ld_w32(uint32_t addr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void tm_st_w8(uint32_t addr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tm_st_w16(uint32_t addr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tm_st_w32(uint32_t addr, const uint32_t (&r)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ f2 f2_from_words(uint32_t lo, uint32_t hi) { return (f2)lo | ((f2)hi << 32); }

template <int N> struct TeamFftX {
  static constexpr int TPR = N / 8;
  static constexpr int NPASS = fft_npass(N);
  static constexpr int WORKER = TPR < 32 ? 32 : TPR;
  static constexpr int BUFB = N * 8;               // bytes per exchange buffer (and their alignment)
  static constexpr int NST = NPASS > 1 ? NPASS - 1 : 1;
  struct Addr {
    uint32_t st[NST];  // byte address of this thread's first store of pass p (every storing pass has radix 8)
    uint32_t ld;       // byte address of this thread's first load after a pass
  };
  // xa: shared-space byte address of this row's buffer pair (aligned to BUFB)
  static __device__ __forceinline__ Addr make_addr(uint32_t xa, int tau) {
    Addr A;
#pragma unroll
    for (int p = 0; p < NST; ++p) {
      const int NS = fft_ns(N, p), R = fft_radix(N, p);
      const int base = (tau / NS) * (NS * R) + (tau % NS);
      A.st[p] = xa + 8u * (uint32_t)fft_pad(base);
    }
    A.ld = xa + 8u * (uint32_t)fft_pad(tau);
    return A;
  }

  // twiddles of pass P (P > 0) into registers; called BEFORE the exchange that precedes the pass, so
  // that the loads do not queue behind the sixteen loads of the exchange (the table is read-only)
  template <int P>
  static __device__ __forceinline__ void load_tw(f2 (&w)[7], const f2 *__restrict__ tw, int tau) {
    constexpr int R = fft_radix(N, P), NB = 8 / R, NS = fft_ns(N, P);
    constexpr int OFF = fft_tw_offset(N, P);
    // the twiddle of butterfly j depends on j % NS only: with NS < TPR the lanes of a warp read
    // NS distinct words (a broadcast: fewer shared-memory wavefronts than TPR distinct ones)
    const int tt = (NB == 1 && NS < TPR) ? (tau & (NS - 1)) : tau;
#pragma unroll
    for (int m = 0; m < NB; ++m)
#pragma unroll
      for (int q = 1; q < R; ++q) w[m * (R - 1) + (q - 1)] = tw[OFF + (m * (R - 1) + (q - 1)) * TPR + tt];
  }
  // TM1 (one transform only): the LAST exchange through tensor memory, as in pass_il below (16 columns)
  template <int P, bool TWO, bool TM1 = false>
  static __device__ __forceinline__ void pass(f2 (&a)[8], f2 (&b)[8], const Addr &A,
                                              const f2 *__restrict__ tw, int tau, int bar, const f2 (&wp)[7],
                                              uint32_t tm = 0) {
    static_assert(!(TM1 && TWO), "the pair goes through pass_il");
    constexpr int R = fft_radix(N, P), NB = 8 / R, NS = fft_ns(N, P), LG = ilog2(R);
    static_assert(P == NPASS - 1 || (R == 8 && NB == 1), "storing passes have radix 8");
    if constexpr (P > 0) {
#pragma unroll
      for (int m = 0; m < NB; ++m)
#pragma unroll
        for (int q = 1; q < R; ++q) {
          const f2 w = wp[m * (R - 1) + (q - 1)];
          a[m + NB * q] = f2_cmul(a[m + NB * q], w);
          if (TWO) b[m + NB * q] = f2_cmul(b[m + NB * q], w);
        }
    }
#pragma unroll
    for (int m = 0; m < NB; ++m) {
      f2 x[R], y[R];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        x[q] = a[m + NB * q];
        if (TWO) y[q] = b[m + NB * q];
      }
      RadixP<R>::run(x);
      if (TWO) RadixP<R>::run(y);
      if constexpr (TM1 && P == NPASS - 2) {
        tm_st8(tm, x[brev(0, 3)], x[brev(1, 3)], x[brev(2, 3)], x[brev(3, 3)], x[brev(4, 3)], x[brev(5, 3)],
               x[brev(6, 3)], x[brev(7, 3)]);
      } else if constexpr (TM1 && P < NPASS - 1) {
        // (TM1: 8-byte slots padded j + (j >> 4): additive like the pair's layout, immediates only)
        tm1_store<P, 0>(A.st[P], x);
      } else if constexpr (P < NPASS - 1) {
#pragma unroll
        for (int p = 0; p < R; ++p) {
          const uint32_t ad = A.st[P] ^ (8u * (uint32_t)fft_pad(p * NS));
          sts_f2(ad, x[brev(p, LG)]);
          if (TWO) sts_f2_off<BUFB>(ad, y[brev(p, LG)]);
        }
      } else {
#pragma unroll
        for (int p = 0; p < R; ++p) {
          a[m + NB * p] = x[brev(p, LG)];
          if (TWO) b[m + NB * p] = y[brev(p, LG)];
        }
      }
    }
    if constexpr (TM1 && P == NPASS - 2) {
      f2 wn[7];
      load_tw<P + 1>(wn, tw, tau);  // (the last pass's section is stored permuted: tm_tw_src)
      tm_wait_st();
      tm_ld4(tm, a[0], a[2], a[1], a[3]);  // (slot renaming: see pass_il)
      tm_ld4(tm + (16u << 16), a[4], a[6], a[5], a[7]);
      tm_wait_ld();
      pass<P + 1, TWO, TM1>(a, b, A, tw, tau, bar, wn, tm);
    } else if constexpr (TM1 && P < NPASS - 1) {
      f2 wn[7];
      load_tw<P + 1>(wn, tw, tau);
      team_sync<WORKER>(bar);
      tm1_load<0>(A.ld, a);
      if constexpr (P < NPASS - 2) team_sync<WORKER>(bar);
      pass<P + 1, TWO, TM1>(a, b, A, tw, tau, bar, wn, tm);
    } else if constexpr (P < NPASS - 1) {
      f2 wn[7];
      load_tw<P + 1>(wn, tw, tau);
      team_sync<WORKER>(bar);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint32_t ad = A.ld ^ (8u * (uint32_t)fft_pad(TPR * r));
        a[r] = lds_f2(ad);
        if (TWO) b[r] = lds_f2_off<BUFB>(ad);
      }
      if constexpr (P < NPASS - 2) team_sync<WORKER>(bar);  // reads done before the next pass writes
      pass<P + 1, TWO, TM1>(a, b, A, tw, tau, bar, wn, tm);
    }
  }
  // NOTE: the caller must team-sync between two calls on the same buffers (the engine's step
  // loop does, at the top of every step).
  static __device__ __forceinline__ void run2(f2 (&a)[8], f2 (&b)[8], const Addr &A,
                                              const f2 *__restrict__ tw, int tau, int bar) {
    const f2 w0[7] = {};
    pass<0, true>(a, b, A, tw, tau, bar, w0);
  }
  static __device__ __forceinline__ void run(f2 (&a)[8], const Addr &A, const f2 *__restrict__ tw,
                                             int tau, int bar) {
    const f2 w0[7] = {};
    pass<0, false>(a, a, A, tw, tau, bar, w0);
  }
  // One transform whose last exchange goes through tensor memory: its shared-memory exchanges use 8-byte
  // slots at j + (j >> 4) -- conflict-free per half warp for the stores and loads of the schedule (checked
  // for N = 256) and additive over the index split, so every access is [register + immediate].
  static __host__ __device__ constexpr int p1_pad(int j) { return j + (j >> 4); }
  static constexpr int P1_BYTES = 8 * (N + N / 16);
  static __device__ __forceinline__ Addr make_addr_p1(uint32_t xa, int tau) {
    Addr A;
#pragma unroll
    for (int p = 0; p < NST; ++p) {
      const int NS = fft_ns(N, p), R = fft_radix(N, p);
      A.st[p] = xa + 8u * (uint32_t)p1_pad((tau / NS) * (NS * R) + (tau % NS));
    }
    A.ld = xa + 8u * (uint32_t)p1_pad(tau);
    return A;
  }
  template <int P, int p> static __device__ __forceinline__ void tm1_store(uint32_t st, const f2 (&x)[8]) {
    if constexpr (p < 8) {
      sts_f2_off<8 * p1_pad(p * fft_ns(N, P))>(st, x[brev(p, 3)]);
      tm1_store<P, p + 1>(st, x);
    }
  }
  template <int r> static __device__ __forceinline__ void tm1_load(uint32_t ld, f2 (&a)[8]) {
    if constexpr (r < 8) {
      a[r] = lds_f2_off<8 * p1_pad(TPR * r)>(ld);
      tm1_load<r + 1>(ld, a);
    }
  }
  // A: make_addr_p1; tm: 16 tensor-memory columns; bins of lane l on return: tm_tau(l) + TPR r
  static __device__ __forceinline__ void run_tm1(f2 (&a)[8], const Addr &A, const f2 *__restrict__ tw,
                                                 int tau, int bar, uint32_t tm) {
    const f2 w0[7] = {};
    pass<0, false, true>(a, a, A, tw, tau, bar, w0, tm);
  }

  // ---- interleaved pair: the two lockstep transforms share ONE buffer of 16-byte slots -------
  // Slot j holds (a[j], b[j]), so an exchange access of both transforms is one 128-bit
  // instruction (half the LDS / STS of run2, the same bytes).  A 128-bit access is served per
  // quarter warp.  The buffer is PADDED, not XOR-swizzled: slot j sits at j + (j >> 3), which keeps
  // the eight slots of every quarter warp on distinct 16-byte bank groups for all store and load
  // patterns of the schedule (N = 64 ... 2048, checked exhaustively) and is ADDITIVE over the
  // schedule's index split (per-thread part + compile-time part, no carries between them) -- so
  // every access is [per-thread address + immediate]: no address instruction at all, and no
  // alignment requirement beyond 16 bytes.  A row's region is IL_BYTES = 18 N bytes.
  static constexpr int IL_BYTES = 16 * (N + N / 8);
  static __host__ __device__ constexpr int il_pad(int j) { return j + (j >> 3); }
  static __device__ __forceinline__ Addr make_addr_il(uint32_t xa, int tau) {
    Addr A;
#pragma unroll
    for (int p = 0; p < NST; ++p) {
      const int NS = fft_ns(N, p), R = fft_radix(N, p);
      const int base = (tau / NS) * (NS * R) + (tau % NS);
      A.st[p] = xa + 16u * (uint32_t)il_pad(base);
    }
    A.ld = xa + 16u * (uint32_t)il_pad(tau);
    return A;
  }
  // TM: the LAST exchange goes through tensor memory instead (one-warp teams whose last pass has radix
  // 4, i.e. N = 256).  In the Stockham schedule that exchange sends output t of thread tau to thread
  // (t1 t0 tau2 tau1 tau0), slot (tau4 tau3 t2) [index bits]; the TMEM store / load pair sends it to
  // thread (tau2 tau1 tau0 t1 t0), slot (tau4 t2 tau3): the same exchange up to a fixed renaming.  So
  // after it physical lane l ACTS AS team member tm_tau(l) (twiddle lookups of the last pass, ownership
  // of the spectrum: bins tm_tau(l) + TPR r) and the loaded slots are renamed (low two bits swapped).
  static constexpr bool TM_OK = NPASS >= 2 && fft_radix(N, NPASS - 1) == 4 && TPR == 32;
  static __device__ __forceinline__ int tm_tau(int lane) { return ((lane & 3) << 3) | (lane >> 2); }
  // The last pass looks its twiddles up as team member tm_tau(lane): read through that index the lanes of
  // a warp would hit shared-memory banks 2 ways (8-byte entries; 4 ways for 16-byte complex128 entries),
  // so kernels that use the tensor-memory exchange keep the LAST pass's section of the table permuted --
  // entry [k][lane] holds the twiddle of member tm_tau(lane) -- and every lane reads its own column.
  // Table index i of the shared copy <- index tm_tw_src(i) of the plan's table.
  static __host__ __device__ constexpr int tm_tw_src(int i) {
    constexpr int OFFL = fft_tw_offset(N, NPASS - 1);
    if (i < OFFL) return i;
    const int k = (i - OFFL) / TPR, lane = (i - OFFL) % TPR;
    return OFFL + k * TPR + (((lane & 3) << 3) | (lane >> 2));
  }
  // free_bar (TM only): an mbarrier on which lane 0 arrives once the warp has made its last read of the
  // exchange region (the four-step schedule lets the team overwrite the region as soon as all warps have)
  template <int P, bool TM = false>
  static __device__ __forceinline__ void pass_il(f2 (&a)[8], f2 (&b)[8], const Addr &A,
                                                 const f2 *__restrict__ tw, int tau, int bar, const f2 (&wp)[7],
                                                 uint32_t tm = 0, uint32_t free_bar = 0) {
    constexpr int R = fft_radix(N, P), NB = 8 / R, NS = fft_ns(N, P), LG = ilog2(R);
    static_assert(P == NPASS - 1 || (R == 8 && NB == 1), "storing passes have radix 8");
    if constexpr (P > 0) {
#pragma unroll
      for (int m = 0; m < NB; ++m)
#pragma unroll
        for (int q = 1; q < R; ++q) {
          const f2 w = wp[m * (R - 1) + (q - 1)];
          a[m + NB * q] = f2_cmul(a[m + NB * q], w);
          b[m + NB * q] = f2_cmul(b[m + NB * q], w);
        }
    }
#pragma unroll
    for (int m = 0; m < NB; ++m) {
      f2 x[R], y[R];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        x[q] = a[m + NB * q];
        y[q] = b[m + NB * q];
      }
      RadixP<R>::run(x);
      RadixP<R>::run(y);
      if constexpr (TM && P == NPASS - 2) {
        // data in columns [0, 16), noise in [16, 32) of the warp's 32 lanes
        tm_st8(tm, x[brev(0, 3)], x[brev(1, 3)], x[brev(2, 3)], x[brev(3, 3)], x[brev(4, 3)], x[brev(5, 3)],
               x[brev(6, 3)], x[brev(7, 3)]);
        tm_st8(tm + 16u, y[brev(0, 3)], y[brev(1, 3)], y[brev(2, 3)], y[brev(3, 3)], y[brev(4, 3)], y[brev(5, 3)],
               y[brev(6, 3)], y[brev(7, 3)]);
      } else if constexpr (P < NPASS - 1) {
        il_store<P, 0>(A.st[P], x, y);
      } else {
#pragma unroll
        for (int p = 0; p < R; ++p) {
          a[m + NB * p] = x[brev(p, LG)];
          b[m + NB * p] = y[brev(p, LG)];
        }
      }
    }
    if constexpr (TM && P == NPASS - 2) {
      f2 wn[7];
      load_tw<P + 1>(wn, tw, tau);  // (the last pass's section is stored permuted: tm_tw_src)
      tm_wait_st();
      // loaded slot (tau4 t2 tau3) is logical slot (tau4 tau3 t2): 1 <-> 2, 5 <-> 6
      tm_ld4(tm, a[0], a[2], a[1], a[3]);
      tm_ld4(tm + (16u << 16), a[4], a[6], a[5], a[7]);
      tm_ld4(tm + 16u, b[0], b[2], b[1], b[3]);
      tm_ld4(tm + (16u << 16) + 16u, b[4], b[6], b[5], b[7]);
      tm_wait_ld();
      pass_il<P + 1, TM>(a, b, A, tw, tau, bar, wn, tm, free_bar);
    } else if constexpr (P < NPASS - 1) {
      f2 wn[7];
      load_tw<P + 1>(wn, tw, tau);
      team_sync<WORKER>(bar);
      il_load<0>(A.ld, a, b);
      if constexpr (P < NPASS - 2) team_sync<WORKER>(bar);  // reads done before the next pass writes
      if constexpr (TM && P == NPASS - 3) {
        if (free_bar != 0u && (tau & 31) == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(free_bar) : "memory");
      }
      pass_il<P + 1, TM>(a, b, A, tw, tau, bar, wn, tm, free_bar);
    }
  }
  // the eight stores of a radix-8 pass / the eight loads after it, offsets as immediates
  template <int P, int p>
  static __device__ __forceinline__ void il_store(uint32_t st, const f2 (&x)[8], const f2 (&y)[8]) {
    if constexpr (p < 8) {
      sts_f2x2_off<16 * il_pad(p * fft_ns(N, P))>(st, x[brev(p, 3)], y[brev(p, 3)]);
      il_store<P, p + 1>(st, x, y);
    }
  }
  template <int r>
  static __device__ __forceinline__ void il_load(uint32_t ld, f2 (&a)[8], f2 (&b)[8]) {
    if constexpr (r < 8) {
      lds_f2x2_off<16 * il_pad(TPR * r)>(ld, a[r], b[r]);
      il_load<r + 1>(ld, a, b);
    }
  }
  static __device__ __forceinline__ void run2_il(f2 (&a)[8], f2 (&b)[8], const Addr &A,
                                                 const f2 *__restrict__ tw, int tau, int bar) {
    const f2 w0[7] = {};
    pass_il<0>(a, b, A, tw, tau, bar, w0);
  }
  // tm: the warp's tensor-memory address (its 32 lanes, 32 columns); spectrum bins of lane l on return:
  // tm_tau(l) + TPR r
  static __device__ __forceinline__ void run2_tm(f2 (&a)[8], f2 (&b)[8], const Addr &A,
                                                 const f2 *__restrict__ tw, int tau, int bar, uint32_t tm,
                                                 uint32_t free_bar = 0) {
    const f2 w0[7] = {};
    pass_il<0, true>(a, b, A, tw, tau, bar, w0, tm, free_bar);
  }
};

// ---- N = R x 256 (R = 2, 4, 8: N = 512, 1024, 2048), "four-step" -----------------------------------
// The generic schedule runs a long row on a team of R warps with a team-wide barrier per pass (three of
// them at N = 2048).  Here the team makes ONE radix-R pass (thread tau holds x[tau + 32 R r]: Q = 8 / R
// butterflies over n1, for n2 = tau + 32 R j) and ONE team-wide exchange -- a transposition after which
// warp w owns the 256 samples of residue k1 = w -- and then every warp runs its own 256-point transform
// on the warp-local machinery above (interleaved exchange in the warp's region of the same buffer, tensor
// memory for the last exchange), with no further team barrier:
//   X[k1 + R k2] = sum_n2 W_256^(n2 k2) [ W_N^(n2 k1) sum_n1 x[n2 + 256 n1] W_R^(n1 k1) ].
// On return lane l of warp w holds bins  w + R (tm_tau(l) + 32 r)  =  (w + R tm_tau(l)) + 32 R r.
template <int R> struct TeamFftRx256 {
  static_assert(R == 2 || R == 4 || R == 8, "N = 512, 1024 or 2048");
  using F256 = TeamFftX<256>;
  static constexpr int N = 256 * R, TEAM = 32 * R, Q = 8 / R, LGR = ilog2(R);
  static constexpr int REGION = F256::IL_BYTES;            // 4608 B: one warp's padded exchange region
  static constexpr int XBYTES = R * REGION;                // the row's exchange bytes (= 18 N)
  static constexpr int TW_A = N;                           // [k1][n2] W_N^(n2 k1)
  static constexpr int TW_TOTAL = TW_A + fft_tw_total(256);  // + the team table of length 256
  // tables computed by the CTA itself (sincospi of exact arguments), no plan storage
  static __device__ __forceinline__ void fill_tables(f2 *tab, int tid, int nthreads) {
    for (int i = tid; i < TW_A; i += nthreads) {
      const int k1 = i >> 8, n2 = i & 255;
      float sn, cs;
      sincospif(-2.0f * (float)((n2 * k1) & (N - 1)) / (float)N, &sn, &cs);
      tab[i] = f2_make(cs, sn);
    }
    constexpr int NP = fft_npass(256);
#pragma unroll
    for (int p = 1; p < NP; ++p) {
      const int RR = fft_radix(256, p), NB = 8 / RR, NS = fft_ns(256, p), off = fft_tw_offset(256, p);
      for (int i = tid; i < NB * (RR - 1) * 32; i += nthreads) {
        // (last pass: the column of lane l holds the twiddle of team member tm_tau(l), see tm_tw_src)
        const int tau = p == NP - 1 ? F256::tm_tau(i & 31) : (i & 31);
        const int mq = i >> 5, m = mq / (RR - 1), q = mq % (RR - 1) + 1;
        const int k = (tau + 32 * m) % NS;
        float sn, cs;
        sincospif(-2.0f * (float)(q * k) / (float)(NS * RR), &sn, &cs);
        tab[TW_A + off + i] = f2_make(cs, sn);
      }
    }
  }
  // the radix-R pass: slot j + Q k1 holds output k1 of butterfly j afterwards
  static __device__ __forceinline__ void radix_pass(f2 (&a)[8]) {
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      f2 x[R];
#pragma unroll
      for (int n1 = 0; n1 < R; ++n1) x[n1] = a[j + Q * n1];
      RadixP<R>::run(x);
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) a[j + Q * k1] = x[brev(k1, LGR)];
    }
  }
  // value (k1, n2 = tau + TEAM j) -> region k1, slot n2; warp w then reads slots lane + 32 r of region w
  template <int s>
  static __device__ __forceinline__ void store_t(uint32_t st, const f2 (&a)[8], const f2 (&b)[8]) {
    if constexpr (s < 8) {
      constexpr int j = s % Q, k1 = s / Q;
      sts_f2x2_off<k1 * REGION + 16 * TEAM * j>(st, a[s], b[s]);
      store_t<s + 1>(st, a, b);
    }
  }
  template <int r>
  static __device__ __forceinline__ void load_t(uint32_t ld, f2 (&a)[8], f2 (&b)[8]) {
    if constexpr (r < 8) {
      lds_f2x2_off<512 * r>(ld, a[r], b[r]);
      load_t<r + 1>(ld, a, b);
    }
  }
  // xT: the row's exchange buffer; A: make_addr_il(xT + (tau >> 5) REGION, tau & 31); tw: the tables above.
  // front: the radix-R pass; the caller then waits until the team has released the buffer (free_bar of
  // the previous row), calls store, makes ONE team-wide synchronisation, and calls back -- which releases
  // the buffer (free_bar) as soon as the warp has made its last read of its region.  Between back and the
  // next store a warp runs free of the team.
  static __device__ __forceinline__ void front(f2 (&a)[8], f2 (&b)[8]) {
    radix_pass(a);
    radix_pass(b);
  }
  static __device__ __forceinline__ void store(const f2 (&a)[8], const f2 (&b)[8], uint32_t xT, int tau) {
    store_t<0>(xT + 16u * (uint32_t)tau, a, b);
  }
  // ---- one transform per row (data without noise; the noise leg of the complex128 mode) ----------------
  // 8-byte slots: region k1 of REGION1 bytes, the warp-local transform on F256::run_tm1 in the same region
  static constexpr int REGION1 = F256::P1_BYTES;
  template <int s> static __device__ __forceinline__ void store_t1(uint32_t st, const f2 (&a)[8]) {
    if constexpr (s < 8) {
      constexpr int j = s % Q, k1 = s / Q;
      sts_f2_off<k1 * REGION1 + 8 * TEAM * j>(st, a[s]);
      store_t1<s + 1>(st, a);
    }
  }
  template <int r> static __device__ __forceinline__ void load_t1(uint32_t ld, f2 (&a)[8]) {
    if constexpr (r < 8) {
      a[r] = lds_f2_off<256 * r>(ld);
      load_t1<r + 1>(ld, a);
    }
  }
  // the twiddles W_N^(n2 k1) of the values this thread will own after the transposition (loaded in
  // the shadow of the team-wide exchange)
  static __device__ __forceinline__ void twiddles(f2 (&t)[8], const f2 *__restrict__ tw, int tau) {
    const int w = tau >> 5, lane = tau & 31;
#pragma unroll
    for (int r = 0; r < 8; ++r) t[r] = tw[w * 256 + lane + 32 * r];
  }
  // A: make_addr_p1(xT + (tau >> 5) REGION1, tau & 31).  The caller team-syncs between two calls (the step
  // loop does, at the top of every step).
  static __device__ __forceinline__ void run1(f2 (&a)[8], uint32_t xT, const typename F256::Addr &A,
                                              const f2 *__restrict__ tw, int tau, int bar, uint32_t tm) {
    radix_pass(a);
    store_t1<0>(xT + 8u * (uint32_t)tau, a);
    f2 t[8];
    twiddles(t, tw, tau);
    team_sync<TEAM>(bar);
    const int w = tau >> 5, lane = tau & 31;
    load_t1<0>(xT + (uint32_t)w * REGION1 + 8u * (uint32_t)lane, a);
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r] = f2_cmul(a[r], t[r]);
    __syncwarp();  // the warp has read its region; its own exchange may overwrite it
    F256::run_tm1(a, A, tw + TW_A, lane, bar, tm);
  }
  static __device__ __forceinline__ void back(f2 (&a)[8], f2 (&b)[8], const f2 (&t)[8], uint32_t xT,
                                              const typename F256::Addr &A, const f2 *__restrict__ tw, int tau,
                                              int bar, uint32_t tm, uint32_t free_bar) {
    const int w = tau >> 5, lane = tau & 31;
    load_t<0>(xT + (uint32_t)w * REGION + 16u * (uint32_t)lane, a, b);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      a[r] = f2_cmul(a[r], t[r]);
      b[r] = f2_cmul(b[r], t[r]);
    }
    __syncwarp();  // the warp has read its region; its own exchange may overwrite it
    F256::run2_tm(a, b, A, tw + TW_A, lane, bar, tm, free_bar);
  }
};
using TeamFft8x256 = TeamFftRx256<8>;

// ---- small helpers shared by the fused engines ------------------------------------------------
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// team-wide OR of a predicate (one vote for a warp-sized team, a reducing named barrier otherwise)
template <int TEAM_THREADS>
__device__ __forceinline__ bool team_any(bool pred, int barrier_id) {
  if constexpr (TEAM_THREADS <= 32) {
    return __any_sync(0xffffffffu, pred);
  } else {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 p, %1, 0;\n"
        "barrier.red.or.pred q, %2, %3, p;\n"
        "selp.u32 %0, 1, 0, q;\n"
        "}\n"
        : "=r"(r)
        : "r"((uint32_t)pred), "r"(barrier_id), "n"(TEAM_THREADS)
        : "memory");
    return r != 0;
  }
}

}  // namespace sds
