// Register-resident radix-2/4/8 butterflies and the 8-elements-per-thread
// Stockham schedule shared by the power-of-two fast paths (K1 pow2, fused engine).
//
// A transform of length N = 8 * TPR is carried by TPR threads ("a row team").
// Thread tau always owns logical elements { tau + TPR * r : r = 0..7 } -- on
// input (samples) and on output (spectrum bins) alike.  Pass p has radix R_p
// (8, 8, ..., then N's remaining factor 2 or 4) and NB = 8 / R_p butterflies per
// thread; butterfly m of thread tau is Stockham butterfly j = tau + TPR * m and
// uses register slots { m + NB * q }.  Between passes values go through shared
// memory (two ping-pong buffers, one team sync per pass), XOR-swizzled to keep the
// strided writes conflict-free.
#pragma once
#include "sds_common.cuh"

namespace sds {

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

// radix list of length N: 8s while they fit, then the remaining factor
__host__ __device__ constexpr int fft_npass(int N) { return (ilog2(N) + 2) / 3; }
__host__ __device__ constexpr int fft_radix(int N, int p) {
  int lg = ilog2(N) - 3 * p;
  return lg >= 3 ? 8 : (1 << lg);
}
__host__ __device__ constexpr int fft_ns(int N, int p) {  // product of radices before pass p
  int ns = 1;
  for (int i = 0; i < p; ++i) ns *= fft_radix(N, i);
  return ns;
}
// twiddle table (one per length): for pass p >= 1, entries [m][q-1][tau]
__host__ __device__ constexpr int fft_tw_count(int N, int p) {
  return p == 0 ? 0 : (8 / fft_radix(N, p)) * (fft_radix(N, p) - 1) * (N / 8);
}
__host__ __device__ constexpr int fft_tw_offset(int N, int p) {
  int off = 0;
  for (int i = 0; i < p; ++i) off += fft_tw_count(N, i);
  return off;
}
__host__ __device__ constexpr int fft_tw_total(int N) { return fft_tw_offset(N, fft_npass(N)); }
// XOR swizzle of the exchange buffers: conflict-free for every write and read pattern
// of the schedule below for N >= 128, cpx<float> and cpx<double> (exhaustive search over
// shift/mask pairs, checked per 128-byte wavefront)
__host__ __device__ constexpr int fft_pad(int o) { return o ^ ((o >> 3) & 15); }
__host__ __device__ constexpr int fft_buf_elems(int N) { return N; }

__host__ __device__ constexpr int brev(int i, int bits) {
  int r = 0;
  for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
  return r;
}

template <typename T> __device__ __forceinline__ cpx<T> mul_mi(cpx<T> a) { return {a.y, -a.x}; }  // * (-i)

// In-place decimation-in-frequency butterflies: natural order in, bit-reversed out
// (slot i holds X[brev(i)]).  Forward sign exp(-2 pi i nk / R).
template <typename T, int R> struct Radix;

template <typename T> struct Radix<T, 1> {
  static __device__ __forceinline__ void run(cpx<T> *) {}
};
template <typename T> struct Radix<T, 2> {
  static __device__ __forceinline__ void run(cpx<T> *a) {
    const cpx<T> t = a[0];
    a[0] = cadd(t, a[1]);
    a[1] = csub(t, a[1]);
  }
};
template <typename T> struct Radix<T, 4> {
  static __device__ __forceinline__ void run(cpx<T> *a) {
    const cpx<T> u0 = cadd(a[0], a[2]), d0 = csub(a[0], a[2]);
    const cpx<T> u1 = cadd(a[1], a[3]), d1 = mul_mi(csub(a[1], a[3]));
    a[0] = cadd(u0, u1);
    a[1] = csub(u0, u1);
    a[2] = cadd(d0, d1);
    a[3] = csub(d0, d1);
  }
};
template <typename T> struct Radix<T, 8> {
  static __device__ __forceinline__ void run(cpx<T> *a) {
    const T h = (T)0.70710678118654752440;
    cpx<T> d[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      d[n] = csub(a[n], a[n + 4]);
      a[n] = cadd(a[n], a[n + 4]);
    }
    a[4] = d[0];
    a[5] = {(d[1].x + d[1].y) * h, (d[1].y - d[1].x) * h};    // * W8^1
    a[6] = mul_mi(d[2]);                                       // * W8^2
    a[7] = {(d[3].y - d[3].x) * h, -(d[3].x + d[3].y) * h};   // * W8^3
    Radix<T, 4>::run(a);
    Radix<T, 4>::run(a + 4);
  }
};

// One team sync: warp-level for teams that fit a warp, a named barrier otherwise.
template <int TEAM_THREADS>
__device__ __forceinline__ void team_sync(int barrier_id) {
  if constexpr (TEAM_THREADS <= 32) {
    __syncwarp();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(barrier_id), "n"(TEAM_THREADS) : "memory");
  }
}

// The per-thread part of a length-N forward transform.  v[r] = x[tau + TPR r] on
// entry, X[tau + TPR r] on return.  bufA/bufB: this row's exchange buffers
// (fft_buf_elems(N) each); tw: the length's twiddle table (fft_tw_total(N)).
template <typename T, int N> struct TeamFft {
  static constexpr int TPR = N / 8;
  static constexpr int NPASS = fft_npass(N);
  static constexpr int WORKER = TPR < 32 ? 32 : TPR;

  template <int P>
  static __device__ __forceinline__ void pass(cpx<T> (&v)[8], cpx<T> *bufA, cpx<T> *bufB,
                                              const cpx<T> *__restrict__ tw, int tau, int bar) {
    constexpr int R = fft_radix(N, P), NB = 8 / R, NS = fft_ns(N, P), LG = ilog2(R);
    if constexpr (P > 0) {
      constexpr int OFF = fft_tw_offset(N, P);
#pragma unroll
      for (int m = 0; m < NB; ++m)
#pragma unroll
        for (int q = 1; q < R; ++q)
          v[m + NB * q] = cmul(v[m + NB * q], tw[OFF + (m * (R - 1) + (q - 1)) * TPR + tau]);
    }
    cpx<T> *out = (P & 1) ? bufB : bufA;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
      cpx<T> a[R];
#pragma unroll
      for (int q = 0; q < R; ++q) a[q] = v[m + NB * q];
      Radix<T, R>::run(a);
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

  static __device__ __forceinline__ void run(cpx<T> (&v)[8], cpx<T> *bufA, cpx<T> *bufB,
                                             const cpx<T> *__restrict__ tw, int tau, int bar) {
    pass<0>(v, bufA, bufB, tw, tau, bar);
  }
};

// Host side: fill the twiddle table of length N (called from sds_plan_create).
template <typename T>
void fill_team_twiddles(int N, cpx<T> *tab) {
  const int tpr = N / 8, np = fft_npass(N);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int p = 1; p < np; ++p) {
    const int R = fft_radix(N, p), NB = 8 / R, NS = fft_ns(N, p), off = fft_tw_offset(N, p);
    for (int m = 0; m < NB; ++m)
      for (int q = 1; q < R; ++q)
        for (int tau = 0; tau < tpr; ++tau) {
          const int j = tau + tpr * m, k = j % NS;
          // W_N^(q k N / (NS R)) = exp(-2 pi i q k / (NS R))
          const long double ang = -two_pi * (long double)(q * k) / (long double)(NS * R);
          tab[off + (m * (R - 1) + (q - 1)) * tpr + tau] = {(T)cosl(ang), (T)sinl(ang)};
        }
  }
}

}  // namespace sds
