// Power-of-two fast path of K1 (placeholder until the register-radix kernels land).
#include "sds_common.cuh"
namespace sds {
int launch_pow2(const Plan &, const void *, int, int64_t, const uint8_t *, int64_t, int64_t, int,
                const int32_t *, double, void *, int, cudaStream_t) {
  return SDS_ENOTSUP;
}
}  // namespace sds
