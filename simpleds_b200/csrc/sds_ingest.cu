// Ingest gather: UVData's (Nblts, Nfreqs, Npols) arrays -> the DelaySpectrum
// layout (Npols, Nbls * Ntimes, Nfreqs), rows permuted by the caller's (baseline number, LST)
// order -- utils.py:15-168 and delay_spectrum.py:913-973 as one pass.  A block takes one output
// row: it reads that baseline-time's Nfreqs * Npols elements once, contiguously (coalesced,
// staged in shared memory), and writes Npols contiguous runs of Nfreqs elements.  Pure copy:
// HBM-bound, 2 x elem_bytes per element.
#include "sds_common.cuh"

namespace sds {

template <typename V>
__global__ void ingest_gather_kernel(const V *__restrict__ in, const int64_t *__restrict__ perm,
                                     int64_t nrow, int nfreq, int npol, V *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V *tile = reinterpret_cast<V *>(smem_raw);
  const int chunk = nfreq * npol;
  for (int64_t r = blockIdx.x; r < nrow; r += gridDim.x) {
    const V *src = in + perm[r] * (int64_t)chunk;
    for (int e = threadIdx.x; e < chunk; e += blockDim.x) tile[e] = src[e];
    __syncthreads();
    for (int o = threadIdx.x; o < chunk; o += blockDim.x) {
      const int p = o / nfreq, f = o - p * nfreq;
      out[((int64_t)p * nrow + r) * nfreq + f] = tile[f * npol + p];
    }
    __syncthreads();
  }
}

template <typename V>
static int launch_ingest(const void *in, const int64_t *perm, int64_t nrow, int nfreq, int npol,
                         void *out, cudaStream_t st) {
  const size_t smem = (size_t)nfreq * npol * sizeof(V);
  if (smem > 48 * 1024)
    SDS_CUDA(cudaFuncSetAttribute(ingest_gather_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  const int64_t cap = 148 * 8;
  ingest_gather_kernel<V><<<(int)(nrow < cap ? nrow : cap), 256, smem, st>>>(
      (const V *)in, perm, nrow, nfreq, npol, (V *)out);
  SDS_LAUNCH_CHECK("ingest_gather_kernel");
  return SDS_OK;
}

}  // namespace sds

extern "C" int sds_ingest_gather(const void *in, const int64_t *perm, int elem_bytes, int64_t nblt,
                                 int64_t nrow, int nfreq, int npol, void *out, sds_stream stream) {
  SDS_CHECK_ARG(in && perm && out, "sds_ingest_gather: null pointer");
  SDS_CHECK_ARG(nblt >= 1 && nrow >= 0 && nfreq >= 1 && npol >= 1, "sds_ingest_gather: bad dims");
  SDS_CHECK_ARG((int64_t)nfreq * npol * elem_bytes <= 200 * 1024,
                "sds_ingest_gather: one baseline-time of %d x %d x %d bytes does not fit shared memory",
                nfreq, npol, elem_bytes);
  if (nrow == 0) return SDS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  switch (elem_bytes) {
    case 1: return sds::launch_ingest<uint8_t>(in, perm, nrow, nfreq, npol, out, st);
    case 4: return sds::launch_ingest<uint32_t>(in, perm, nrow, nfreq, npol, out, st);
    case 8: return sds::launch_ingest<uint64_t>(in, perm, nrow, nfreq, npol, out, st);
    case 16: return sds::launch_ingest<uint4>(in, perm, nrow, nfreq, npol, out, st);
    default: break;
  }
  sds::set_error("sds_ingest_gather: elem_bytes must be 1, 4, 8 or 16, got %d", elem_bytes);
  return SDS_EINVAL;
}
