// Micro-benchmark: tensor memory (TMEM) as a register-exchange medium between the threads of a warp.
// A warp stores with tcgen05.st.32x32b (thread t -> TMEM lane t, 16 consecutive columns) and loads with
// tcgen05.ld.16x256b (thread t <- lanes t/4 and t/4 + 8, column pairs t%4 + 4j): the two fragment
// layouts differ, so store + load is a fixed permutation of (thread, register) -- the kind of index-bit
// swap an FFT pass needs -- that never touches the shared-memory data pipe.
//   mode 0: print the permutation;  mode 1: TMEM exchange throughput;  mode 2: shared-memory exchange
//   (4 x STS.128 + 4 x LDS.128 per thread, the fused engine's interleaved exchange) throughput;
//   mode 3: both in the same loop (do the two pipes overlap?)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_xchg tmem_xchg.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// 16 lanes x 256 bit, x2: regs [0,1] row t/4 cols 2(t%4)+{0,1}; [2,3] row t/4+8; [4,5] row t/4 cols 8+..; [6,7] row t/4+8
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t addr, uint32_t *d) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr int COLS = 32;

__global__ void __launch_bounds__(128) k(int mode, int iters, uint32_t *out, long long *cyc) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase_s)), "n"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tbase_s + ((uint32_t)(warp * 32) << 16);

  uint32_t v[16], d[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = ((uint32_t)lane << 8) | (uint32_t)j;
  if (mode == 0) {
    tmem_st16(tb, v);
    tmem_wait_st();
    tmem_ld_16x256b_x2(tb, d);
    tmem_ld_16x256b_x2(tb + (16u << 16), d + 8);
    tmem_wait_ld();
    if (blockIdx.x == 0 && warp == 1)
      for (int j = 0; j < 16; ++j) out[lane * 16 + j] = d[j];
  } else {
    const uint32_t sa = smem_u32(dsm) + (uint32_t)warp * 4608u;
    const uint32_t st_a = sa + 16u * (uint32_t)(9 * lane);          // padded slots 8 lane + p -> 9 lane + p
    const uint32_t ld_a = sa + 16u * (uint32_t)(lane + (lane >> 3));  // slots lane + 32 r -> + 36 r
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 1 || mode == 3) {
        tmem_st16(tb, v);
        tmem_wait_st();
        tmem_ld_16x256b_x2(tb, d);
        tmem_ld_16x256b_x2(tb + (16u << 16), d + 8);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = d[j] + 1u;
      }
      if (mode == 2 || mode == 3) {
#pragma unroll
        for (int p = 0; p < 4; ++p)
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(st_a + 32u * p), "r"(v[4 * p]), "r"(v[4 * p + 1]),
                       "r"(v[4 * p + 2]), "r"(v[4 * p + 3])
                       : "memory");
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 4; ++p)
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(d[4 * p]), "=r"(d[4 * p + 1]), "=r"(d[4 * p + 2]), "=r"(d[4 * p + 3])
                       : "r"(ld_a + 16u * 36u * p)
                       : "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = d[j] + 1u;
      }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) acc ^= v[j];
    if (acc == 0x12345678u) out[0] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase_s), "n"(COLS));
}

int main() {
  uint32_t *out;
  long long *cyc;
  const int nblk = 148 * 3;
  cudaMalloc(&out, 512 * 4);
  cudaMalloc(&cyc, nblk * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  k<<<1, 128, 70 * 1024>>>(0, 0, out, cyc);
  uint32_t h[512];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  printf("permutation: dest thread t', regs 0..15 = (source thread, source reg)\n");
  for (int t = 0; t < 32; ++t) {
    printf("t'=%2d:", t);
    for (int j = 0; j < 16; ++j) printf(" (%2u,%2u)", h[t * 16 + j] >> 8, h[t * 16 + j] & 255u);
    printf("\n");
  }
  const int iters = 20000;
  for (int mode = 1; mode <= 3; ++mode) {
    k<<<nblk, 128, 70 * 1024>>>(mode, 100, out, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<nblk, 128, 70 * 1024>>>(mode, iters, out, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long hc[nblk];
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < nblk; ++i) mean += (double)hc[i];
    mean /= nblk;
    // per SM: 12 warps, each moves 2 KB out and 2 KB back per iteration
    // are the three CTAs of an SM resident together?  kernel time / time of one CTA: 1 if so, 3 if serialised
    printf("mode %d: %.1f clk per iteration per warp -> %.1f clk per warp-exchange per SM if 12 warps/SM are resident; "
           "kernel %.3f ms = %.2f x one CTA's time (at 1.9 GHz), err %s\n",
           mode, mean / iters, mean / iters / 12.0, ms, ms * 1.9e6 / mean, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
