// K1 fast path, complex64: forward delay transform of power-of-two lengths 64..2048 with the machinery
// of the fused engine (sds_engine_f32.cu) -- rows staged by 1-D bulk TMA through an mbarrier ring per
// row team, packed fp32x2 butterflies, XOR-addressed exchange (TeamFftX) -- and the spectrum stored
// straight from registers, fftshift as an index rotation (a warp writes contiguous 256-byte runs).
//
// utils.py:548-560 + the mask multiply of ds.py:3503-3517 in one pass; HBM-bound: 9 B in (8 B without
// flags) + 8 B out per sample.  Needs 16-byte aligned rows and window origins (bulk copies); everything
// else falls back to delay_fft_pow2_kernel (sds_fft_pow2.cu).
#include "sds_engine.cuh"
#include "sds_packed.cuh"

namespace sds {

constexpr int K1_STAGES = 3;

struct K1Params {
  const float2 *in;
  const uint8_t *flags;  // nullable
  float2 *out;
  const float *taper;
  const float2 *tw;
  int64_t in_row_stride, flag_row_stride, nrows, total;
  float delta_x;
  int nwin;
  int win_start[SDS_MAX_WINDOWS];
};

template <int N> struct K1Cfg {
  static constexpr int TPR = N / 8;
  static constexpr int TEAM = TPR < 32 ? 32 : TPR;
  static constexpr int ROWS = TEAM / TPR;
  static constexpr int THREADS = TEAM < 128 ? 128 : TEAM;
  static constexpr int NTEAM = THREADS / TEAM;
  static constexpr int MINB = 512 / THREADS < 1 ? 1 : 512 / THREADS;  // 16 warps per SM
  static constexpr int ROWB = 9 * N;  // vis 8N | flags N
  static constexpr int SLOT = ROWS * ROWB;
  static constexpr int BUF = N * 8;
  static constexpr int XREGION = NTEAM * ROWS * BUF;  // one exchange buffer per row, aligned to BUF
  static constexpr int TW_BYTES = ((fft_tw_total(N) * 8 + 127) / 128) * 128;
  static constexpr int TEAM_BYTES = 64 + K1_STAGES * SLOT;
  static constexpr int SMEM = BUF + XREGION + TW_BYTES + NTEAM * TEAM_BYTES;
};

template <int N, bool FLAGS>
__global__ void __launch_bounds__(K1Cfg<N>::THREADS, K1Cfg<N>::MINB)
delay_fft_tma_kernel(const __grid_constant__ K1Params P) {
  using C = K1Cfg<N>;
  using FFT = TeamFftX<N>;
  constexpr int TPR = C::TPR, TEAM = C::TEAM, ROWS = C::ROWS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t s0 = smem_u32(smem_raw);
  unsigned char *smem = smem_raw + (((s0 + C::BUF - 1u) & ~(uint32_t)(C::BUF - 1)) - s0);
  f2 *tw_s = reinterpret_cast<f2 *>(smem + C::XREGION);
  for (int i = threadIdx.x; i < fft_tw_total(N); i += C::THREADS) tw_s[i] = reinterpret_cast<const f2 *>(P.tw)[i];
  const int team = threadIdx.x / TEAM, tl = threadIdx.x % TEAM;
  const int sub = tl / TPR, tau = tl % TPR, bar_id = 1 + team;
  // TeamFftX addresses a PAIR of buffers (the second one BUF bytes above the first) for two lockstep
  // transforms; the single transform only ever touches the first
  const uint32_t xa = smem_u32(smem) + (uint32_t)(team * ROWS + sub) * C::BUF;
  const typename FFT::Addr fa = FFT::make_addr(xa, tau);
  unsigned char *tbase = smem + C::XREGION + C::TW_BYTES + (size_t)team * C::TEAM_BYTES;
  uint64_t *full = reinterpret_cast<uint64_t *>(tbase);
  unsigned char *stages = tbase + 64;
  const uint32_t full_a = smem_u32(full), stages_a = smem_u32(stages);
  if (tl == 0) {
    for (int i = 0; i < K1_STAGES; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  float tapdx[8];  // taper * delta_x at this thread's channels (utils.py:552, 560)
#pragma unroll
  for (int r = 0; r < 8; ++r) tapdx[r] = __ldg(P.taper + tau + TPR * r) * P.delta_x;

  const int64_t nstep = (P.total + ROWS - 1) / ROWS;  // team steps of ROWS (window, row) pairs
  const int64_t stride = (int64_t)gridDim.x * C::NTEAM;
  const bool issuer_warp = TEAM == 32 || __shfl_sync(0xffffffffu, tl, 0) == 0;
  uint32_t slot_c = 0, par_c = 0, slot_p = 0;
  int64_t stp_p = (int64_t)blockIdx.x * C::NTEAM + team;  // next step to stage
  auto issue = [&]() {
    if (stp_p < nstep && issuer_warp && elect_one()) {
      const int64_t rw0 = stp_p * ROWS;
      const int nact = (int)min((int64_t)ROWS, P.total - rw0);
      const uint32_t bar = full_a + 8u * slot_p;
      mbar_expect_tx_a(bar, (uint32_t)nact * (uint32_t)(FLAGS ? 9 * N : 8 * N));
      uint32_t dst = stages_a + slot_p * (uint32_t)C::SLOT;
      for (int rr = 0; rr < nact; ++rr, dst += C::ROWB) {
        const int64_t rw = rw0 + rr;
        const int w = (int)(rw / P.nrows);
        const int64_t row = rw - (int64_t)w * P.nrows;
        tma_load_1d_a(dst, P.in + row * P.in_row_stride + P.win_start[w], 8 * N, bar);
        if (FLAGS) tma_load_1d_a(dst + 8 * N, P.flags + row * P.flag_row_stride + P.win_start[w], N, bar);
      }
    }
    slot_p = slot_p + 1 == K1_STAGES ? 0 : slot_p + 1;
    stp_p += stride;
  };
  for (int k = 0; k < K1_STAGES - 1; ++k) issue();
  for (int64_t stp = (int64_t)blockIdx.x * C::NTEAM + team; stp < nstep; stp += stride) {
    team_sync<TEAM>(bar_id);  // everyone is done with the slot about to be refilled (and with the exchange buffer)
    issue();
    const int64_t rw = stp * ROWS + sub;
    const bool active = ROWS == 1 || rw < P.total;
    mbar_wait_a(full_a + 8u * slot_c, par_c);
    const unsigned char *row_s = stages + slot_c * C::SLOT + sub * C::ROWB;
    const f2 *vis_s = reinterpret_cast<const f2 *>(row_s) + tau;
    const unsigned char *fl_s = row_s + 8 * N + tau;
    f2 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {  // data * (~flag) * taper * delta_x   (ds.py:3503-3505, utils.py:558-560)
      const bool keep = active && (!FLAGS || fl_s[TPR * r] == 0);
      v[r] = f2_scale(active ? vis_s[TPR * r] : 0ull, keep ? tapdx[r] : 0.f);
    }
    FFT::run(v, fa, tw_s, tau, bar_id);
    if (active) {
      const int w = (int)(rw / P.nrows);
      const int64_t row = rw - (int64_t)w * P.nrows;
      f2 *op = reinterpret_cast<f2 *>(P.out) + ((int64_t)w * P.nrows + row) * N;
#pragma unroll
      for (int r = 0; r < 8; ++r)  // fftshift (utils.py:559): bin k lands at (k + N/2) mod N
        op[(tau + TPR * r + N / 2) % N] = v[r];
    }
    slot_c = slot_c + 1 == K1_STAGES ? 0 : slot_c + 1;
    if (slot_c == 0) par_c ^= 1u;
  }
}

template <int N, bool FLAGS>
static int launch_k1_n(const K1Params &P, cudaStream_t st) {
  using C = K1Cfg<N>;
  static DeviceLatch configured;
  int cfg_dev = 0;
  if (configured.need(&cfg_dev)) {
    SDS_CUDA(cudaFuncSetAttribute(delay_fft_tma_kernel<N, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    SDS_CUDA(cudaFuncSetAttribute(delay_fft_tma_kernel<N, FLAGS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    configured.set(cfg_dev);
  }
  int dev = 0, nsm = 148, per_sm = 1;
  SDS_CUDA(cudaGetDevice(&dev));
  SDS_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  SDS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, delay_fft_tma_kernel<N, FLAGS>, C::THREADS, C::SMEM));
  if (per_sm < 1) return SDS_ENOTSUP;
  const int64_t nstep = (P.total + C::ROWS - 1) / C::ROWS;
  int64_t blocks = (nstep + C::NTEAM - 1) / C::NTEAM;
  if (blocks > (int64_t)nsm * per_sm) blocks = (int64_t)nsm * per_sm;  // persistent, grid-stride
  delay_fft_tma_kernel<N, FLAGS><<<(unsigned)blocks, C::THREADS, C::SMEM, st>>>(P);
  SDS_LAUNCH_CHECK("delay_fft_tma_kernel");
  return SDS_OK;
}

template <int N>
static int launch_k1_f(const K1Params &P, cudaStream_t st) {
  return P.flags ? launch_k1_n<N, true>(P, st) : launch_k1_n<N, false>(P, st);
}

// SDS_ENOTSUP when the layout does not allow bulk copies (the caller falls back)
int launch_pow2_tma(const Plan &pl, const void *in, int in_dtype, int64_t in_row_stride, const uint8_t *flags,
                    int64_t flag_row_stride, int64_t nrows, int nwin, const int32_t *win_start, double delta_x,
                    void *out, int out_dtype, cudaStream_t st) {
  if (!pl.is_pow2 || pl.nfreq < 64 || pl.nfreq > 2048 || !pl.team_tw_f || pl.math != SDS_F32 ||
      in_dtype != SDS_C64 || out_dtype != SDS_C64)
    return SDS_ENOTSUP;
  // 16-byte aligned bulk-copy sources: complex64 rows at even element offsets, flag rows at multiples of 16
  if (((uintptr_t)in % 16) != 0 || (in_row_stride % 2) != 0 || ((uintptr_t)out % 8) != 0) return SDS_ENOTSUP;
  if (flags && (((uintptr_t)flags % 16) != 0 || (flag_row_stride % 16) != 0)) return SDS_ENOTSUP;
  for (int w = 0; w < nwin; ++w)
    if (win_start[w] % (flags ? 16 : 2) != 0) return SDS_ENOTSUP;
  K1Params P;
  P.in = (const float2 *)in; P.flags = flags; P.out = (float2 *)out;
  P.taper = pl.taper_f; P.tw = (const float2 *)pl.team_tw_f;
  P.in_row_stride = in_row_stride; P.flag_row_stride = flag_row_stride;
  P.nrows = nrows; P.total = nrows * nwin; P.delta_x = (float)delta_x; P.nwin = nwin;
  for (int i = 0; i < SDS_MAX_WINDOWS; ++i) P.win_start[i] = i < nwin ? win_start[i] : 0;
  switch (pl.nfreq) {
    case 64: return launch_k1_f<64>(P, st);
    case 128: return launch_k1_f<128>(P, st);
    case 256: return launch_k1_f<256>(P, st);
    case 512: return launch_k1_f<512>(P, st);
    case 1024: return launch_k1_f<1024>(P, st);
    case 2048: return launch_k1_f<2048>(P, st);
    default: return SDS_ENOTSUP;
  }
}

}  // namespace sds
