// Ingest gather: UVData's (Nblts, Nfreqs, Npols) arrays -> the DelaySpectrum
// layout (Npols, Nbls * Ntimes, Nfreqs), rows permuted by the caller's (baseline number, LST)
// order -- utils.py:15-168 and delay_spectrum.py:913-973 as one pass.  Pure copy, HBM-bound:
// 2 x elem_bytes per element.  The pol axis is the fastest input axis and the slowest output
// axis; with Npols in {1, 2, 4} the transposition happens in registers: a thread loads the
// Npols values of one channel (or, for 1-byte flags, of 16 / Npols channels) as one aligned
// vector and stores one value (vector) per polarisation, so both sides are fully coalesced.
// Other Npols / odd Nfreqs take a shared-memory staged row.
#include "sds_common.cuh"

namespace sds {

template <typename V, int P> struct alignas((sizeof(V) * P > 16) ? 16 : sizeof(V) * P) PolPack {
  V v[P];
};

// 4-, 8- and 16-byte elements: thread = (row, channel)
template <typename V, int P>
__global__ void ingest_vec_kernel(const V *__restrict__ in, const int64_t *__restrict__ perm,
                                  int64_t nrow, int nfreq, V *__restrict__ out) {
  const int64_t total = nrow * nfreq;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / nfreq;
    const int f = (int)(e - r * nfreq);
    const PolPack<V, P> pk =
        *reinterpret_cast<const PolPack<V, P> *>(in + (__ldg(perm + r) * nfreq + f) * P);
#pragma unroll
    for (int p = 0; p < P; ++p) out[((int64_t)p * nrow + r) * nfreq + f] = pk.v[p];
  }
}

// 1-byte elements (flags): thread = (row, 16 / P consecutive channels), 16 bytes in, P vectors out
template <int P>
__global__ void ingest_bytes_kernel(const uint8_t *__restrict__ in, const int64_t *__restrict__ perm,
                                    int64_t nrow, int nfreq, uint8_t *__restrict__ out) {
  constexpr int CH = 16 / P;  // channels per thread
  const int per_row = nfreq / CH;
  const int64_t total = nrow * per_row;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / per_row;
    const int f0 = (int)(e - r * per_row) * CH;
    const uint4 w = *reinterpret_cast<const uint4 *>(in + (__ldg(perm + r) * nfreq + f0) * P);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int p = 0; p < P; ++p) {
      uint32_t o[CH / 4 > 0 ? CH / 4 : 1] = {};
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int b = k * P + p;  // byte index in the 16-byte load
        o[k / 4] |= ((ws[b / 4] >> (8 * (b % 4))) & 0xFFu) << (8 * (k % 4));
      }
      uint8_t *dst = out + ((int64_t)p * nrow + r) * nfreq + f0;
      if constexpr (CH == 4) *reinterpret_cast<uint32_t *>(dst) = o[0];
      else if constexpr (CH == 8) *reinterpret_cast<uint2 *>(dst) = make_uint2(o[0], o[1]);
      else *reinterpret_cast<uint4 *>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// any Npols / Nfreqs: a block stages one row in shared memory
template <typename V>
__global__ void ingest_row_kernel(const V *__restrict__ in, const int64_t *__restrict__ perm,
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

static int grid_cap(int64_t work, int block) {
  const int64_t g = (work + block - 1) / block, cap = 148 * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <typename V>
static int launch_ingest(const void *in, const int64_t *perm, int64_t nrow, int nfreq, int npol,
                         void *out, cudaStream_t st) {
  const V *i = (const V *)in;
  V *o = (V *)out;
  const int grid = grid_cap(nrow * nfreq, 256);
  if (sizeof(V) > 1 && npol == 1) {
    ingest_vec_kernel<V, 1><<<grid, 256, 0, st>>>(i, perm, nrow, nfreq, o);
  } else if (sizeof(V) > 1 && npol == 2) {
    ingest_vec_kernel<V, 2><<<grid, 256, 0, st>>>(i, perm, nrow, nfreq, o);
  } else if (sizeof(V) > 1 && npol == 4) {
    ingest_vec_kernel<V, 4><<<grid, 256, 0, st>>>(i, perm, nrow, nfreq, o);
  } else {
    const size_t smem = (size_t)nfreq * npol * sizeof(V);
    if (smem > 48 * 1024)
      SDS_CUDA(cudaFuncSetAttribute(ingest_row_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    const int64_t cap = 148 * 8;
    ingest_row_kernel<V><<<(int)(nrow < cap ? nrow : cap), 256, smem, st>>>(i, perm, nrow, nfreq, npol, o);
  }
  SDS_LAUNCH_CHECK("ingest_gather");
  return SDS_OK;
}

static int launch_ingest_bytes(const void *in, const int64_t *perm, int64_t nrow, int nfreq, int npol,
                               void *out, cudaStream_t st) {
  const uint8_t *i = (const uint8_t *)in;
  uint8_t *o = (uint8_t *)out;
  const bool vec = (npol == 1 || npol == 2 || npol == 4) && nfreq % (16 / npol) == 0 &&
                   (reinterpret_cast<uintptr_t>(in) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
                   ((int64_t)nrow * nfreq) % 16 == 0;
  if (!vec) return launch_ingest<uint8_t>(in, perm, nrow, nfreq, npol, out, st);
  const int grid = grid_cap(nrow * (nfreq / (16 / npol)), 256);
  if (npol == 1) ingest_bytes_kernel<1><<<grid, 256, 0, st>>>(i, perm, nrow, nfreq, o);
  else if (npol == 2) ingest_bytes_kernel<2><<<grid, 256, 0, st>>>(i, perm, nrow, nfreq, o);
  else ingest_bytes_kernel<4><<<grid, 256, 0, st>>>(i, perm, nrow, nfreq, o);
  SDS_LAUNCH_CHECK("ingest_gather");
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
    case 1: return sds::launch_ingest_bytes(in, perm, nrow, nfreq, npol, out, st);
    case 4: return sds::launch_ingest<uint32_t>(in, perm, nrow, nfreq, npol, out, st);
    case 8: return sds::launch_ingest<uint64_t>(in, perm, nrow, nfreq, npol, out, st);
    case 16: return sds::launch_ingest<uint4>(in, perm, nrow, nfreq, npol, out, st);
    default: break;
  }
  sds::set_error("sds_ingest_gather: elem_bytes must be 1, 4, 8 or 16, got %d", elem_bytes);
  return SDS_EINVAL;
}
