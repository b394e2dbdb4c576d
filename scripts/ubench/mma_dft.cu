// Micro-benchmark / prototype: the noise leg's length-256 transform as two 16x16 DFT stages on the
// legacy tensor path (mma.sync m16n8k16, bf16 operands with a hi/lo split of the constant DFT-16
// matrix, fp32 accumulation), one row per warp, no shared-memory exchange.
//   part 1: throughput of HMMA (f16 / bf16 / tf32), HMUL2.BF16, cvt.bf16x2, PRMT, and of HMMA
//           issued next to FFMA2 (do the two pipes overlap?)
//   part 2: correctness of the two-stage transform against a double-precision DFT
//   part 3: cycles per row of {generate a row, transform, accumulate} for the tensor formulation
//           and for the register/shared-memory Stockham formulation it would replace
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o mma_dft mma_dft.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../simpleds_b200/csrc/sds_engine.cuh"
#include "../../simpleds_b200/csrc/sds_packed.cuh"

using namespace sds;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t hmul2_bf16(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// ---------------------------------------------------------------- part 1: pipe rates
template <int MODE>
__global__ void rate_kernel(int iters, float *out, long long *cyc) {
  float d[4][4];
  uint32_t a[4], b[4], u[8];
  f2 p[8];
  for (int i = 0; i < 4; ++i) {
    a[i] = 0x3F803F80u + threadIdx.x * 0x00010001u * i;
    b[i] = 0x3F003F00u + threadIdx.x;
    for (int j = 0; j < 4; ++j) d[i][j] = (float)(i + j);
  }
  for (int i = 0; i < 8; ++i) { u[i] = threadIdx.x * 77 + i * 0x01010101u; p[i] = (f2)(threadIdx.x + i) * 0x3f8000013f800001ull; }
  const f2 cb = 0x3f8000003f800000ull;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mma_bf16(d[i], a, b[i & 1], b[2 + (i & 1)]);
      }
      if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mma_f16(d[i], a, b[i & 1], b[2 + (i & 1)]);
      }
      if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mma_tf32(d[i], a, b[i & 1], b[2 + (i & 1)]);
      }
      if (MODE == 3) {  // 4 HMMA + 8 FFMA2
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          mma_bf16(d[i], a, b[i & 1], b[2 + (i & 1)]);
          p[2 * i] = f2_fma(p[2 * i], cb, cb);
          p[2 * i + 1] = f2_fma(p[2 * i + 1], cb, cb);
        }
      }
      if (MODE == 4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = hmul2_bf16(u[i], 0x3F813F81u);
      }
      if (MODE == 5) {
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = pack_bf16(__uint_as_float(u[i]), __uint_as_float(u[(i + 1) & 7]));
      }
      if (MODE == 6) {
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = __byte_perm(u[i], u[(i + 1) & 7], 0x7614);
      }
      if (MODE == 7) {  // 2 dependent HMMA chains only (latency-bound issue)
#pragma unroll
        for (int i = 0; i < 4; ++i) mma_bf16(d[i & 1], a, b[i & 1], b[2 + (i & 1)]);
      }
    }
  }
  long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) acc += d[i][j];
  for (int i = 0; i < 8; ++i) acc += (float)(u[i] & 0xff) + (float)(p[i] & 0xff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run_rate(const char *name, int per_rep, int threads) {
  float *out; long long *cyc;
  CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
  const int iters = 2000;
  rate_kernel<MODE><<<148, threads>>>(10, out, cyc); CK(cudaDeviceSynchronize());
  rate_kernel<MODE><<<148, threads>>>(iters, out, cyc); CK(cudaDeviceSynchronize());
  long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double wi = (double)iters * 4 * per_rep * (threads / 32) / 4.0;  // warp instructions per SMSP
  printf("%-30s warps/SM %2d: %.4f warp-instr/clk/SMSP = %.2f clk per warp-instr\n", name, threads / 32, wi / c, c / wi);
  cudaFree(out); cudaFree(cyc);
}

// ---------------------------------------------------------------- the two-stage transform
// Lane (g = lane / 4, t = lane % 4).  The K slots of the first stage are permuted (slot 2t + a + 8b holds
// n1 = t + 4a + 8b) and so are the columns of its data operand (column g + 8j holds
// n2 = (g >> 1) + 4 (g & 1) + 8j), which makes (1) the staged-row reads two-way instead of four-way bank
// conflicted and (2) the constant fragments of stage 1 (A operand) serve as the B operand of stage 2
// (the DFT-16 matrix is symmetric).
struct Dft16Frags {
  uint32_t rh[4], rl[4], ih[4], il[4];  // Re / Im of exp(-2 pi i k n / 16), hi and lo bf16 parts
};
__device__ __forceinline__ void make_frags(Dft16Frags &F, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int row = g + 8 * (q & 1), n0 = t + 8 * (q >> 1), n1 = n0 + 4;
    float c0, s0, c1, s1;
    sincospif(-(float)((row * n0) & 15) / 8.0f, &s0, &c0);
    sincospif(-(float)((row * n1) & 15) / 8.0f, &s1, &c1);
    const uint32_t rh = pack_bf16(c0, c1), ih = pack_bf16(s0, s1);
    F.rh[q] = rh; F.ih[q] = ih;
    F.rl[q] = pack_bf16(c0 - bf16_lo(rh), c1 - bf16_hi(rh));
    F.il[q] = pack_bf16(s0 - bf16_lo(ih), s1 - bf16_hi(ih));
  }
}
// twiddles W256^(k1 n2) of the thread's 8 stage-1 outputs: index j * 4 + e, k1 = g + 8 (e >> 1),
// n2 = t + 4 (e & 1) + 8 j
__device__ __forceinline__ void make_twiddles(f2 (&tw)[8], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k1 = g + 8 * (e >> 1), n2 = t + 4 * (e & 1) + 8 * j;
      float s, c;
      sincospif(-(float)((k1 * n2) & 255) / 128.0f, &s, &c);
      tw[j * 4 + e] = f2_make(c, s);
    }
}
// channel of input slot (j, q): j = column tile, q = 0..3 <-> (a = q & 1, b = q >> 1)
__device__ __forceinline__ int in_channel(int lane, int j, int q) {
  const int g = lane >> 2, t = lane & 3;
  const int n1 = t + 4 * (q & 1) + 8 * (q >> 1), n2 = (g >> 1) + 4 * (g & 1) + 8 * j;
  return 16 * n1 + n2;
}
// delay bin of output slot (j, e)
__device__ __forceinline__ int out_bin(int lane, int j, int e) {
  const int g = lane >> 2, t = lane & 3;
  return g + 8 * (e >> 1) + 16 * (2 * t + (e & 1) + 8 * j);
}

// xr[j][h], xi[j][h]: bf16x2 data fragments of column tile j (h = 0: K slots 2t, 2t+1; h = 1: 2t+8, 2t+9)
template <bool SPLIT>
__device__ __forceinline__ void dft256_mma(const uint32_t (&xr)[2][2], const uint32_t (&xi)[2][2],
                                           const Dft16Frags &F, const f2 (&tw)[8], f2 (&X)[8]) {
  uint32_t zr[4], zi[4], nzi[4];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float yr[4] = {0.f, 0.f, 0.f, 0.f}, yi[4] = {0.f, 0.f, 0.f, 0.f};
    const uint32_t nx0 = xi[j][0] ^ 0x80008000u, nx1 = xi[j][1] ^ 0x80008000u;
    mma_bf16(yr, F.rh, xr[j][0], xr[j][1]);
    mma_bf16(yi, F.ih, xr[j][0], xr[j][1]);
    mma_bf16(yr, F.ih, nx0, nx1);
    mma_bf16(yi, F.rh, xi[j][0], xi[j][1]);
    if (SPLIT) {
      mma_bf16(yr, F.rl, xr[j][0], xr[j][1]);
      mma_bf16(yi, F.il, xr[j][0], xr[j][1]);
      mma_bf16(yr, F.il, nx0, nx1);
      mma_bf16(yi, F.rl, xi[j][0], xi[j][1]);
    }
    f2 z[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) z[e] = f2_cmul(f2_make(yr[e], yi[e]), tw[j * 4 + e]);
    zr[2 * j] = pack_bf16(f2_lo(z[0]), f2_lo(z[1]));
    zr[2 * j + 1] = pack_bf16(f2_lo(z[2]), f2_lo(z[3]));
    zi[2 * j] = pack_bf16(f2_hi(z[0]), f2_hi(z[1]));
    zi[2 * j + 1] = pack_bf16(f2_hi(z[2]), f2_hi(z[3]));
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) nzi[q] = zi[q] ^ 0x80008000u;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float vr[4] = {0.f, 0.f, 0.f, 0.f}, vi[4] = {0.f, 0.f, 0.f, 0.f};
    mma_bf16(vr, zr, F.rh[j], F.rh[2 + j]);
    mma_bf16(vi, zr, F.ih[j], F.ih[2 + j]);
    mma_bf16(vr, nzi, F.ih[j], F.ih[2 + j]);
    mma_bf16(vi, zi, F.rh[j], F.rh[2 + j]);
    if (SPLIT) {
      mma_bf16(vr, zr, F.rl[j], F.rl[2 + j]);
      mma_bf16(vi, zr, F.il[j], F.il[2 + j]);
      mma_bf16(vr, nzi, F.il[j], F.il[2 + j]);
      mma_bf16(vi, zi, F.rl[j], F.rl[2 + j]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) X[j * 4 + e] = f2_make(vr[e], vi[e]);
  }
}

// ---------------------------------------------------------------- part 2: correctness
template <bool SPLIT>
__global__ void dft_check_kernel(const float2 *x, float2 *y, int nrows) {
  const int lane = threadIdx.x & 31, row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  Dft16Frags F; f2 tw[8];
  make_frags(F, lane); make_twiddles(tw, lane);
  uint32_t xr[2][2], xi[2][2];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float2 a = x[row * 256 + in_channel(lane, j, 2 * h)], b = x[row * 256 + in_channel(lane, j, 2 * h + 1)];
      xr[j][h] = pack_bf16(a.x, b.x);
      xi[j][h] = pack_bf16(a.y, b.y);
    }
  f2 X[8];
  dft256_mma<SPLIT>(xr, xi, F, tw, X);
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) y[row * 256 + out_bin(lane, j, e)] = make_float2(f2_lo(X[j * 4 + e]), f2_hi(X[j * 4 + e]));
}

static float bf16_round_host(float v) {
  uint32_t u; memcpy(&u, &v, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  u &= 0xFFFF0000u;
  float r; memcpy(&r, &u, 4);
  return r;
}

template <bool SPLIT> void check_dft(const char *name) {
  const int nrows = 2048;
  std::vector<float2> hx(nrows * 256), hy(nrows * 256);
  srand(1);
  for (auto &v : hx) {  // +-[1, 2) magnitudes times a smooth amplitude profile, like the kernel's noise rows
    v.x = bf16_round_host((rand() % 2 ? 1.f : -1.f) * (1.f + (rand() % 128) / 128.f));
    v.y = bf16_round_host((rand() % 2 ? 1.f : -1.f) * (1.f + (rand() % 128) / 128.f));
  }
  float2 *dx, *dy;
  CK(cudaMalloc(&dx, hx.size() * 8)); CK(cudaMalloc(&dy, hy.size() * 8));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 8, cudaMemcpyHostToDevice));
  dft_check_kernel<SPLIT><<<nrows / 4, 128>>>(dx, dy, nrows);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hy.data(), dy, hy.size() * 8, cudaMemcpyDeviceToHost));
  // reference: double DFT of the first 64 rows for the error, all rows' power per bin for the variance ratio
  double err2 = 0, ref2 = 0, maxerr = 0;
  std::vector<double> pw_got(256, 0.0), pw_ref(256, 0.0);
  std::vector<double> cs(256), sn(256);
  for (int m = 0; m < 256; ++m) { cs[m] = cos(2 * M_PI * m / 256.0); sn[m] = -sin(2 * M_PI * m / 256.0); }
  for (int r = 0; r < nrows; ++r)
    for (int k = 0; k < 256; ++k) {
      double re = 0, im = 0;
      for (int c = 0; c < 256; ++c) {
        const int m = (k * c) & 255;
        re += hx[r * 256 + c].x * cs[m] - hx[r * 256 + c].y * sn[m];
        im += hx[r * 256 + c].x * sn[m] + hx[r * 256 + c].y * cs[m];
      }
      const double dr = hy[r * 256 + k].x - re, di = hy[r * 256 + k].y - im;
      err2 += dr * dr + di * di; ref2 += re * re + im * im;
      maxerr = fmax(maxerr, sqrt(dr * dr + di * di));
      pw_got[k] += (double)hy[r * 256 + k].x * hy[r * 256 + k].x + (double)hy[r * 256 + k].y * hy[r * 256 + k].y;
      pw_ref[k] += re * re + im * im;
    }
  double rmin = 1e9, rmax = 0, rsum = 0;
  for (int k = 0; k < 256; ++k) { const double q = pw_got[k] / pw_ref[k]; rmin = fmin(rmin, q); rmax = fmax(rmax, q); rsum += q; }
  printf("%-22s rel. rms error %.3e  max abs err / rms %.3e  power ratio per bin: mean %.7f min %.7f max %.7f\n", name,
         sqrt(err2 / ref2), maxerr / sqrt(ref2 / (nrows * 256.0)), rsum / 256, rmin, rmax);
  cudaFree(dx); cudaFree(dy);
}

// ---------------------------------------------------------------- part 3: cycles per row
// one Philox call per thread and row: 128 bits -> 8 complex +-[1, 2) bf16 draws (7 mantissa bits + sign each)
__device__ __forceinline__ void draw_bf16(uint64_t row_id, int lane, const uint32_t (&key)[16], uint32_t (&d)[8]) {
  const u32x4 rn = philox4x32_keyed<ENG_PHILOX_ROUNDS>({(uint32_t)row_id, (uint32_t)(row_id >> 32), (uint32_t)lane, 1u}, key);
  const uint32_t w[4] = {rn.x, rn.y, rn.z, rn.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    d[2 * i] = (w[i] & 0x807F807Fu) | 0x3F803F80u;
    d[2 * i + 1] = (__funnelshift_r(w[i], w[i], 8) & 0x807F807Fu) | 0x3F803F80u;
  }
}

struct KeyParams { uint32_t key[16]; };

template <int MODE>
__global__ void __launch_bounds__(128, 3) row_kernel(int rows_per_warp, const __grid_constant__ KeyParams K, const float *amp_tab,
                                                      float *out, long long *cyc) {
  __shared__ __align__(128) unsigned char xbuf[2048 + 4 * 2 * 2048 + 4096];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  Dft16Frags F; f2 tw[8];
  make_frags(F, lane); make_twiddles(tw, lane);
  float sig[8];
  for (int r = 0; r < 8; ++r) sig[r] = amp_tab[(lane + 32 * r) & 255];
  f2 S1[8]; float S2[8];
  for (int r = 0; r < 8; ++r) { S1[r] = 0ull; S2[r] = 0.f; }
  const uint64_t row0 = ((uint64_t)blockIdx.x * 4 + wid) * (uint64_t)rows_per_warp;
  // Stockham formulation state
  const uint32_t s0 = smem_u32(xbuf), s1 = (s0 + 2047u) & ~2047u;
  f2 *tw_s = reinterpret_cast<f2 *>(xbuf + (s1 - s0) + 4 * 2 * 2048);
  const uint32_t xa = s1 + wid * 2 * 2048;
  const TeamFftX<256>::Addr fa = TeamFftX<256>::make_addr(xa, lane);
  static_assert(fft_tw_total(256) * 8 <= 4096, "twiddle table");
  for (int i = threadIdx.x; i < fft_tw_total(256); i += 128) tw_s[i] = f2_make(1.f, 0.f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < rows_per_warp; ++it) {
    const float rt = 0.3f + 1e-6f * (float)it;
    if (MODE == 0 || MODE == 1 || MODE == 3) {  // tensor formulation (1: without the hi/lo split; 3: generation only)
      uint32_t d[8];
      draw_bf16(row0 + it, lane, K.key, d);
      uint32_t xr[2][2], xi[2][2];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t a2 = pack_bf16(sig[4 * j + 2 * h] * rt, sig[4 * j + 2 * h + 1] * rt);
          xr[j][h] = hmul2_bf16(d[4 * j + 2 * h], a2);
          xi[j][h] = hmul2_bf16(d[4 * j + 2 * h + 1], a2);
        }
      f2 X[8];
      if (MODE == 3) {
#pragma unroll
        for (int q = 0; q < 8; ++q) X[q] = f2_make(bf16_lo(xr[q >> 2][(q >> 1) & 1]), bf16_hi(xi[q >> 2][(q >> 1) & 1]));
      } else {
        dft256_mma<MODE == 0>(xr, xi, F, tw, X);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        S1[q] = f2_add(S1[q], X[q]);
        S2[q] = fmaf(f2_lo(X[q]), f2_lo(X[q]), fmaf(f2_hi(X[q]), f2_hi(X[q]), S2[q]));
      }
    } else {  // MODE 2: what the engine does today: 3 Philox calls, Box-Muller, Stockham in registers + shared memory
      float rad[8], cs[8], sn[8];
      engine_draw8<32>(row0 + it, lane, K.key, rad, cs, sn);
      f2 v[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = f2_scale(f2_make(cs[r], sn[r]), sig[r] * rt * rad[r]);
      __syncwarp();
      TeamFftX<256>::run(v, fa, tw_s, lane, 1);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        S1[q] = f2_add(S1[q], v[q]);
        S2[q] = fmaf(f2_lo(v[q]), f2_lo(v[q]), fmaf(f2_hi(v[q]), f2_hi(v[q]), S2[q]));
      }
    }
  }
  long long t1 = clock64();
  float acc = 0;
  for (int r = 0; r < 8; ++r) acc += f2_lo(S1[r]) + f2_hi(S1[r]) + S2[r];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run_rows(const char *name) {
  float *out, *amp; long long *cyc;
  const int grid = 148 * 3, rows = 4000;
  CK(cudaMalloc(&out, grid * 128 * 4)); CK(cudaMalloc(&cyc, grid * 8)); CK(cudaMalloc(&amp, 256 * 4));
  std::vector<float> ha(256);
  for (int i = 0; i < 256; ++i) ha[i] = 0.35875f - 0.48829f * cosf(2 * M_PI * i / 255.f) + 0.14128f * cosf(4 * M_PI * i / 255.f) - 0.01168f * cosf(6 * M_PI * i / 255.f);
  CK(cudaMemcpy(amp, ha.data(), 1024, cudaMemcpyHostToDevice));
  KeyParams K; philox_fill_keys(123u, 456u, K.key);
  CK(cudaFuncSetAttribute(row_kernel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  row_kernel<MODE><<<grid, 128>>>(10, K, amp, out, cyc); CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  row_kernel<MODE><<<grid, 128>>>(rows, K, amp, out, cyc);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(grid); CK(cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
  double c = 0; for (auto v : h) c += v; c /= grid;
  // 12 warps per SM = 3 per SMSP, each doing `rows` rows in c cycles
  printf("%-44s %.1f clk per row and SMSP (12 warps/SM), %.3f ms, %.2f Grows/s\n", name, c / (3.0 * rows), ms,
         (double)grid * 4 * rows / (ms * 1e6));
  cudaFree(out); cudaFree(cyc); cudaFree(amp);
}

int main() {
  for (int threads : {128, 384, 512}) {
    run_rate<0>("HMMA.16816.F32.BF16", 4, threads);
    run_rate<1>("HMMA.16816.F32 (f16)", 4, threads);
    run_rate<2>("HMMA.1688.F32.TF32", 4, threads);
    run_rate<7>("HMMA bf16, 2 chains", 4, threads);
    run_rate<3>("4 HMMA + 8 FFMA2", 12, threads);
    run_rate<4>("HMUL2.BF16", 8, threads);
    run_rate<5>("F2FP.BF16.PACK", 8, threads);
    run_rate<6>("PRMT", 8, threads);
  }
  check_dft<true>("bf16 hi/lo split");
  check_dft<false>("bf16 single");
  run_rows<0>("tensor DFT, hi/lo (32 HMMA/row)");
  run_rows<1>("tensor DFT, single (16 HMMA/row)");
  run_rows<3>("tensor path, generation + sums only");
  run_rows<2>("Box-Muller + Stockham (today's noise leg)");
  return 0;
}
