// Micro-benchmark: what does one legacy HMMA.16816 cost the other pipes of its SM sub-partition?
// Per loop body: 1 HMMA + K independent instructions of one type; 12 warps / SM.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 fma2(f2 a,f2 b,f2 c){f2 r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ float ffma(float a,float b,float c){float r; asm volatile("fma.rn.f32 %0,%1,%2,%3;":"=f"(r):"f"(a),"f"(b),"f"(c)); return r;}
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b, unsigned c){unsigned r; asm volatile("lop3.b32 %0,%1,%2,%3,0x96;":"=r"(r):"r"(a),"r"(b),"r"(c)); return r;}
__device__ __forceinline__ void mma(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// TYPE 0 FFMA, 1 FFMA2, 2 LOP3, 3 LDS; NM = HMMAs per body (0 or 1)
template <int TYPE, int K, int NM>
__global__ void k(int iters, float *out, long long *cyc) {
  __shared__ float sm[1024];
  float d[4][4]; unsigned a[4], b[4], u[16]; f2 p[16]; float s[16];
  for (int i = 0; i < 4; ++i) { a[i] = 0x3F803F80u + threadIdx.x * i; b[i] = 0x3F003F00u + threadIdx.x; for (int j = 0; j < 4; ++j) d[i][j] = i + j; }
  for (int i = 0; i < 16; ++i) { u[i] = threadIdx.x * 77 + i; p[i] = (f2)(threadIdx.x + i) * 0x3f8000013f800001ull; s[i] = 1.f + i; }
  sm[threadIdx.x] = threadIdx.x; __syncthreads();
  const f2 cb = 0x3f8000003f800000ull;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
      if (NM) mma(d[rep], a, b[rep & 1], b[2 + (rep & 1)]);
#pragma unroll
      for (int i = 0; i < K; ++i) {
        if (TYPE == 0) s[i] = ffma(s[i], 1.0000001f, 1.0000001f);
        if (TYPE == 1) p[i] = fma2(p[i], cb, cb);
        if (TYPE == 2) u[i] = lop(u[i], u[(i + 1) & 15], 0x9E3779B9u);
        if (TYPE == 3) s[i] += sm[(threadIdx.x + 32 * i + it) & 1023];
      }
    }
  }
  long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) acc += d[i][j];
  for (int i = 0; i < 16; ++i) acc += s[i] + (float)(p[i] & 0xff) + (float)(u[i] & 0xff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int TYPE, int K, int NM> void run(const char *name) {
  float *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000, threads = 384;
  k<TYPE, K, NM><<<148, threads>>>(10, out, cyc); cudaDeviceSynchronize();
  k<TYPE, K, NM><<<148, threads>>>(iters, out, cyc); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  // per SMSP: 3 warps, each iters*4 bodies
  printf("%-6s K=%2d HMMA=%d: %.2f clk per body and SMSP-warp slot (3 warps/SMSP) -> %.2f clk per body per SMSP\n", name, K, NM,
         c / (iters * 4.0), c / (iters * 4.0) / 1.0 * 1.0 / 1.0);
  cudaFree(out); cudaFree(cyc);
}
template <int TYPE> void sweep(const char *name) {
  run<TYPE, 2, 0>(name); run<TYPE, 2, 1>(name); run<TYPE, 4, 0>(name); run<TYPE, 4, 1>(name);
  run<TYPE, 8, 0>(name); run<TYPE, 8, 1>(name); run<TYPE, 16, 0>(name); run<TYPE, 16, 1>(name);
}
int main() {
  run<0, 0, 1>("none");
  sweep<0>("FFMA"); sweep<1>("FFMA2"); sweep<2>("LOP3"); sweep<3>("LDS");
  return 0;
}
