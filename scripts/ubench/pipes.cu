// Micro-benchmark: issue/pipe throughput of FFMA, FFMA2, FADD2, IMAD.WIDE, LOP3, MUFU and mixes on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 fma2(f2 a,f2 b,f2 c){f2 r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ f2 add2(f2 a,f2 b){f2 r; asm volatile("add.rn.f32x2 %0,%1,%2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ float ffma(float a,float b,float c){float r; asm volatile("fma.rn.f32 %0,%1,%2,%3;":"=f"(r):"f"(a),"f"(b),"f"(c)); return r;}
__device__ __forceinline__ unsigned long long imadw(unsigned a, unsigned b){unsigned long long r; asm volatile("mul.wide.u32 %0,%1,%2;":"=l"(r):"r"(a),"r"(b)); return r;}
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b, unsigned c){unsigned r; asm volatile("lop3.b32 %0,%1,%2,%3,0x96;":"=r"(r):"r"(a),"r"(b),"r"(c)); return r;}
__device__ __forceinline__ float mufu(float a){float r; asm volatile("ex2.approx.ftz.f32 %0,%1;":"=f"(r):"f"(a)); return r;}

template <int MODE>
__global__ void k(int iters, float *out, long long *cyc) {
  f2 p[8]; float s[8]; unsigned u[8];
  for (int i = 0; i < 8; ++i) { p[i] = (f2)(threadIdx.x + i) * 0x3f8000013f800001ull; s[i] = 1.0f + i + threadIdx.x; u[i] = threadIdx.x * 77 + i; }
  const f2 cb = 0x3f8000003f800000ull; const float cs = 1.0000001f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) s[i] = ffma(s[i], cs, cs);
        if (MODE == 1) p[i] = fma2(p[i], cb, cb);
        if (MODE == 2) p[i] = add2(p[i], cb);
        if (MODE == 3) { unsigned long long w = imadw(u[i], 0xD2511F53u); u[i] = (unsigned)(w >> 32) ^ (unsigned)w; }  // IMAD.WIDE + LOP3
        if (MODE == 4) u[i] = lop(u[i], u[(i + 1) & 7], 0x9E3779B9u);
        if (MODE == 5) s[i] = mufu(s[i]);
        if (MODE == 6) { p[i] = fma2(p[i], cb, cb); s[i] = ffma(s[i], cs, cs); }          // FFMA2 + FFMA
        if (MODE == 7) { p[i] = fma2(p[i], cb, cb); u[i] = lop(u[i], u[(i + 1) & 7], 0x9E3779B9u); }  // FFMA2 + LOP3
        if (MODE == 8) { p[i] = fma2(p[i], cb, cb); unsigned long long w = imadw(u[i], 0xD2511F53u); u[i] = (unsigned)(w >> 32) + (unsigned)w; }  // FFMA2 + IMAD.WIDE + IADD
        if (MODE == 9) { s[i] = ffma(s[i], cs, cs); u[i] = lop(u[i], u[(i + 1) & 7], 0x9E3779B9u); }  // FFMA + LOP3
        if (MODE == 10) { unsigned long long w = imadw(u[i], 0xD2511F53u); u[i] = (unsigned)(w >> 32); }  // IMAD.WIDE only
        if (MODE == 11) { s[i] = ffma(s[i], cs, cs); unsigned long long w = imadw(u[i], 0xD2511F53u); u[i] = (unsigned)(w >> 32); }  // FFMA + IMAD.WIDE
      }
    }
  }
  long long t1 = clock64();
  float acc = 0; for (int i = 0; i < 8; ++i) acc += s[i] + (float)(p[i] & 0xff) + (float)(u[i] & 0xff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, int per_iter, int threads) {
  float *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int iters = 4000;
  k<MODE><<<148, threads>>>(10, out, cyc); cudaDeviceSynchronize();
  k<MODE><<<148, threads>>>(iters, out, cyc); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double warp_inst_per_smsp = (double)iters * 32 * per_iter * (threads / 32) / 4.0;
  printf("%-28s warps/SM %2d: %.3f warp-instr/clk/SMSP (%d instr per slot-group)\n", name, threads / 32, warp_inst_per_smsp / c, per_iter);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int threads : {128, 256, 512}) {
    run<0>("FFMA", 1, threads); run<1>("FFMA2", 1, threads); run<2>("FADD2", 1, threads);
    run<10>("IMAD.WIDE", 1, threads); run<3>("IMAD.WIDE+LOP3", 2, threads); run<4>("LOP3", 1, threads); run<5>("MUFU.EX2", 1, threads);
    run<6>("FFMA2+FFMA", 2, threads); run<7>("FFMA2+LOP3", 2, threads); run<8>("FFMA2+IMAD.WIDE+IADD", 3, threads);
    run<9>("FFMA+LOP3", 2, threads); run<11>("FFMA+IMAD.WIDE", 2, threads);
  }
  return 0;
}
