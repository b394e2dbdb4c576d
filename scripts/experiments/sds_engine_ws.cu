// EXPERIMENT kept as evidence (profiles/r2_engine_team_split_n2048_ncu_summary.txt): NOT part of libsds.so.
// Measured on SKA-512 (N = 2048): 7.77 ms per step against 7.30 ms for the lockstep kernel -- correct (all engine tests
// green), but 26 % more instructions, 22 % more shared-memory wavefronts, short-scoreboard + barrier stalls 43 %.
// Fused engine for LONG windows (N = 512 .. 2048), complex64, all three legs: the row team of
// sds_engine_f32.cu split into two teams of N / 8 threads that work on the SAME staged row --
//   team 0: flag mask * taper * delta_f -> transform of the data row -> S1 += X, S2 += |X|^2
//   team 1: noise draw * sigma(nsample, t_int) -> transform of the noise row -> sums; thermal sums
// Why: with N / 8 >= 64 threads per row the lockstep kernel needs ~240 registers per thread and
// one 8-warp CTA fills an SM at N = 2048 (profiles/r2_engine_n2048_ncu_summary.txt: issue slots 40 %
// busy, 0.59 eligible warps per cycle, barrier + MIO stalls 22 %): it is bound by latency at two
// warps per scheduler.  Each team keeps only its own leg's accumulators (<= 128 registers), so twice
// the warps are resident, and the two transforms of a row overlap across warps instead of inside
// one instruction stream.  The row is staged ONCE (1-D bulk TMA, mbarrier ring) and read by both
// teams; a CTA-wide barrier per row hands the ring slot back to the producer.
// Same item schedule, noise stream (moment-matched draw, sds_engine.cuh) and outputs as the lockstep
// kernel; sds_engine_run picks this one for legs = data | noise | thermal with noise_mode 1 or 2.
#include "sds_engine.cuh"
#include "sds_packed.cuh"

namespace sds {

constexpr int WS_STAGES = 3;

template <int N> struct WsCfg {
  static constexpr int TPR = N / 8;            // threads per team (one row)
  static_assert(TPR >= 64, "team-split engine: windows of 512 channels and more");
  static constexpr int THREADS = 2 * TPR;
  static constexpr int MINB = 512 / THREADS < 1 ? 1 : 512 / THREADS;  // 16 warps per SM: 128 registers per thread
  static constexpr int ROWB = 13 * N;          // vis 8N | nsample 4N | flags N
  static constexpr int BUF = N * 8;
  static constexpr int XREGION = 3 * BUF;      // team 0: one exchange buffer; team 1: two (thermal hand-over)
  static constexpr int TW_BYTES = ((fft_tw_total(N) * 8 + 127) / 128) * 128;
  // 64 B barriers | ring | 2 x N floats (per-team sums over times of |S1|^2)
  static constexpr int SMEM = BUF + XREGION + TW_BYTES + 64 + WS_STAGES * ROWB + 8 * N;
};

template <int N>
__global__ void __launch_bounds__(WsCfg<N>::THREADS, WsCfg<N>::MINB)
engine_ws_kernel(const __grid_constant__ EngineParams P) {
  using C = WsCfg<N>;
  using FFT = TeamFftX<N>;
  constexpr int TPR = C::TPR;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t s0 = smem_u32(smem_raw);
  unsigned char *smem = smem_raw + (((s0 + C::BUF - 1u) & ~(uint32_t)(C::BUF - 1)) - s0);
  f2 *tw_s = reinterpret_cast<f2 *>(smem + C::XREGION);
  for (int i = threadIdx.x; i < fft_tw_total(N); i += C::THREADS) tw_s[i] = reinterpret_cast<const f2 *>(P.tw_f)[i];
  const int role = threadIdx.x / TPR, tau = threadIdx.x % TPR, bar_id = 1 + role;
  const uint32_t xa = smem_u32(smem) + (uint32_t)role * C::BUF;  // team 1 owns [BUF, 3 BUF)
  const typename FFT::Addr fa = FFT::make_addr(xa, tau);
  unsigned char *wbase = smem + C::XREGION + C::TW_BYTES;
  uint64_t *full = reinterpret_cast<uint64_t *>(wbase);
  unsigned char *stages = wbase + 64;
  const uint32_t full_a = smem_u32(full), stages_a = smem_u32(stages);
  float *pw_s = reinterpret_cast<float *>(stages + WS_STAGES * C::ROWB) + role * N;
  for (int i = tau; i < N; i += TPR) pw_s[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < WS_STAGES; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // per-thread channel constants: taper * delta_f (team 0) / sigma coefficient * taper * delta_f (team 1, per item)
  float tapdf[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) tapdf[r] = __ldg(P.taper_f + tau + TPR * r) * (float)P.delta_f;
  const int total_items = P.item_start[P.ngroup];
  const int64_t row_stride = (int64_t)P.ntime * P.nchan;
  const bool issuer_warp = threadIdx.x < 32;
  uint32_t slot_c = 0, par_c = 0, slot_p = 0;
  __shared__ int item_s;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) item_s = atomicAdd(P.counter, 1);
    __syncthreads();
    const int item = item_s;
    if (item >= total_items) break;
    int lo = 0, hi = P.ngroup;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(P.item_start + mid) <= item) lo = mid; else hi = mid;
    }
    const int g = lo, b0 = __ldg(P.goff + g), ng = __ldg(P.goff + g + 1) - b0;
    const int gout = P.group_id ? __ldg(P.group_id + g) : g;
    const int64_t blg = P.group_bl0 ? __ldg(P.group_bl0 + g) : P.bl_offset + b0;
    const int chunk = eng_chunk(ng, P.ntime, P.target_rows);
    const int nchunks = (P.ntime + chunk - 1) / chunk;
    const int local = item - __ldg(P.item_start + g);
    const int tc = local % nchunks, sp = local / nchunks, p = sp % P.npol, s = sp / P.npol;
    const int t0 = tc * chunk, t1 = min(P.ntime, t0 + chunk);
    const int nsteps = (t1 - t0) * ng;
    const int64_t e00 = (((int64_t)p * P.nbl + b0) * P.ntime + t0) * P.nchan + P.win_start[s];
    const int64_t obase = (((int64_t)gout * P.nspw + s) * P.npol + p) * N;

    int pi0 = 0, pleft = nsteps;
    int64_t pe = e00, pe_t = e00;
    auto issue = [&]() {
      if (issuer_warp && elect_one()) {
        const uint32_t bar = full_a + 8u * slot_p;
        mbar_expect_tx_a(bar, (uint32_t)C::ROWB);
        const uint32_t dst = stages_a + slot_p * (uint32_t)C::ROWB;
        tma_load_1d_a(dst, P.vis + pe, 8 * N, bar);
        tma_load_1d_a(dst + 8 * N, P.nsample + pe, 4 * N, bar);
        tma_load_1d_a(dst + 12 * N, P.flags + pe, N, bar);
      }
      slot_p = slot_p + 1 == WS_STAGES ? 0 : slot_p + 1;
      pe += row_stride;
      if (++pi0 >= ng) {
        pi0 = 0;
        pe_t += P.nchan;
        pe = pe_t;
      }
      --pleft;
    };

    float coef[8];  // team 0: taper * delta_f; team 1: the noise multiplier per channel (ds.py:3575-3582: non-finite -> 0)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      coef[r] = tapdf[r];
      if (role == 1 && P.noise_mode == 1) {
        const float c = __ldg(P.sigma_coef + ((int64_t)s * P.npol + p) * N + tau + TPR * r);
        coef[r] = (fabsf(c) <= 3.0e38f) ? c * tapdf[r] : 0.f;
      }
    }
    f2 S1[8];     // sum over the rows of a time of the team's transform (data / noise)
    float S2[8];  // sum |X|^2 over the item
    f2 tA[4], tB[4], tC[4];  // team 1: thermal sums (sds_engine_f32.cu)
    float dfc[48];
    bool dirty_time = false;
    double th_all_acc = 0.0, th_auto_acc = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      S1[r] = 0ull;
      S2[r] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) tA[k] = tB[k] = tC[k] = 0ull;

    int ci = 0, ct = t0;
    int64_t it_off = (int64_t)b0 * P.ntime + t0;
    u32x4 nwords = {0u, 0u, 0u, 0u};
    const uint64_t nblk = (uint64_t)((P.nbl_total + 3) >> 2);
    const uint64_t blk_sp = (uint64_t)((int64_t)s * P.npol + p) * nblk;
    for (int k = 0; k < WS_STAGES - 1 && pleft > 0; ++k) issue();
    // (lane * P.zero = 0 keeps the value off the uniform datapath, see sds_engine_f32.cu)
    const int lane_zero = (int)(threadIdx.x & 31u) * P.zero;
    float tint_next = 1.f;
    if (role == 1) tint_next = __ldg(P.int_time + it_off + lane_zero);
    for (int step = 0; step < nsteps; ++step) {
      __syncthreads();  // both teams are done with the slot about to be refilled
      if (pleft > 0) issue();
      const float tint = tint_next;
      if (role == 1 && step + 1 < nsteps) {
        const bool wrap = ci + 1 >= ng;
        tint_next = __ldg(P.int_time + it_off + (wrap ? 1 : 0) + lane_zero + (int64_t)(wrap ? 0 : ci + 1) * P.ntime);
      }
      mbar_wait_a(full_a + 8u * slot_c, par_c);
      const unsigned char *row = stages + slot_c * C::ROWB;
      const unsigned char *fl_s = row + 12 * N + tau;
      f2 v[8];
      if (role == 0) {
        const f2 *vis_s = reinterpret_cast<const f2 *>(row) + tau;
#pragma unroll
        for (int r = 0; r < 8; ++r)  // (1 - flag) * taper * delta_f   (ds.py:3503-3505, utils.py:552-560)
          v[r] = f2_scale(vis_s[TPR * r], fl_s[TPR * r] == 0 ? coef[r] : 0.f);
      } else {
        const float *ns_s = reinterpret_cast<const float *>(row + 8 * N) + tau;
        const float rt = (tint > 0.f) ? fast_rsqrt(tint) : 0.f, itv = rt * rt;
        float av[8];
        bool all_ok = true;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float nv = ns_s[TPR * r];
          av[r] = nv > 0.f ? fast_rsqrt(nv) : 0.f;
          all_ok = all_ok && nv > 0.f;
        }
        if (P.noise_mode == 1) {  // moment-matched discrete draw (sds_engine.cuh)
          const uint32_t bsel = (uint32_t)(blg + ci) & 3u;
          if (ci == 0 || bsel == 0u) {
            const uint64_t bglob = (uint64_t)(blg + ci);
            nwords = engine_call_mm<ENG_PHILOX_ROUNDS>(
                (blk_sp + (bglob >> 2)) * (uint64_t)P.ntime_total + (uint64_t)(P.time_offset + ct), tau, P.pkey);
          }
          const uint32_t word = engine_pick_mm(nwords, bsel);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const bool kn = fl_s[TPR * r] == 0 && ((word >> r) & 1u);
            v[r] = engine_signs_mm(((kn ? coef[r] : 0.f) * rt) * av[r], word, r);
          }
        } else {  // injected frequency-domain noise (parity mode)
          const int64_t e = e00 + (int64_t)(ct - t0) * P.nchan + (int64_t)ci * row_stride;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float2 c = P.noise_in[e + tau + TPR * r];
            const float m = fl_s[TPR * r] == 0 ? coef[r] : 0.f;
            v[r] = f2_make(c.x * m, c.y * m);
          }
        }
        // thermal sums on the transform's channel map (sds_engine_f32.cu: A = sum a, B = sum a / t, C = sum a^2 / t)
        const f2 it2 = f2_make(itv, itv);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const f2 aa = f2_make(av[2 * m], av[2 * m + 1]);
          const f2 ab = f2_mul(aa, it2);
          tA[m] = f2_add(tA[m], aa);
          tB[m] = f2_add(tB[m], ab);
          tC[m] = f2_fma(aa, ab, tC[m]);
        }
        if (team_any<TPR>(!all_ok, bar_id)) {  // a row with an invalid sample: book what the panel mask removes
          if (!dirty_time) {
#pragma unroll 1
            for (int k = 0; k < 48; ++k) dfc[k] = 0.f;
            dirty_time = true;
          }
#pragma unroll 1
          for (int q = 0; q < 8; ++q) {
            if (tau + TPR * q + 1 >= N) continue;
            const float nl = ns_s[TPR * q], nr = ns_s[TPR * q + 1];
            if ((nl > 0.f) == (nr > 0.f)) continue;
            const float a1 = fast_rsqrt(nl > 0.f ? nl : nr);
            float *d = dfc + 6 * q + (nl > 0.f ? 0 : 3);
            d[0] += a1;
            d[1] += a1 * itv;
            d[2] += a1 * a1 * itv;
          }
        }
      }
      FFT::run(v, fa, tw_s, tau, bar_id);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        S1[r] = f2_add(S1[r], v[r]);
        S2[r] = fmaf(f2_lo(v[r]), f2_lo(v[r]), fmaf(f2_hi(v[r]), f2_hi(v[r]), S2[r]));
      }

      slot_c = slot_c + 1 == WS_STAGES ? 0 : slot_c + 1;
      if (slot_c == 0) par_c ^= 1u;
      if (++ci >= ng) {  // close the time
        ci = 0;
        ++ct;
        ++it_off;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          pw_s[tau + TPR * r] += f2_hsum(f2_mul(S1[r], S1[r]));
          S1[r] = 0ull;
        }
        if (role == 1) {
          const double *wlp = P.wl + ((int64_t)s * P.npol + p) * N + tau;
          const double *wrp = P.wr + ((int64_t)s * P.npol + p) * N + tau;
          team_sync<TPR>(bar_id);  // the step's last transform reads are done
          float Ar[8], Br[8], Cr[8], Ap[8], Bp[8], Cp[8];
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            Ar[2 * m] = f2_lo(tA[m]); Ar[2 * m + 1] = f2_hi(tA[m]);
            Br[2 * m] = f2_lo(tB[m]); Br[2 * m + 1] = f2_hi(tB[m]);
            Cr[2 * m] = f2_lo(tC[m]); Cr[2 * m + 1] = f2_hi(tC[m]);
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            sts_f2(xa + 8u * (uint32_t)(tau + TPR * r), f2_make(Ar[r], Br[r]));
            sts_f32(xa + C::BUF + 4u * (uint32_t)(tau + TPR * r), Cr[r]);
          }
          team_sync<TPR>(bar_id);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            if (r < 7 || tau + 1 < TPR) {  // no panel right of channel N - 1
              const f2 nx = lds_f2(xa + 8u * (uint32_t)(tau + TPR * r + 1));
              Ap[r] = f2_lo(nx); Bp[r] = f2_hi(nx);
              Cp[r] = lds_f32(xa + C::BUF + 4u * (uint32_t)(tau + TPR * r + 1));
            } else {
              Ap[r] = Bp[r] = Cp[r] = 0.f;
            }
          }
          if (dirty_time) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              Ar[r] -= dfc[6 * r + 0]; Br[r] -= dfc[6 * r + 1]; Cr[r] -= dfc[6 * r + 2];
              Ap[r] -= dfc[6 * r + 3]; Bp[r] -= dfc[6 * r + 4]; Cp[r] -= dfc[6 * r + 5];
            }
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const double l = __ldg(wlp + TPR * r), rr = __ldg(wrp + TPR * r);
            th_all_acc += l * (double)(Ar[r] * Br[r]) + rr * (double)(Ap[r] * Bp[r]);
            th_auto_acc += l * (double)Cr[r] + rr * (double)Cp[r];
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) tA[k] = tB[k] = tC[k] = 0ull;
          dirty_time = false;
        }
      }
    }

    // flush the item: fp64 atomics, fftshift as an index rotation
    double *o_all = role == 0 ? P.pow_all : P.noise_all, *o_auto = role == 0 ? P.pow_auto : P.noise_auto;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int kshift = (tau + TPR * r + N / 2) % N;
      atomicAdd(o_all + obase + kshift, (double)pw_s[tau + TPR * r]);
      atomicAdd(o_auto + obase + kshift, (double)S2[r]);
      pw_s[tau + TPR * r] = 0.f;
    }
    if (role == 1) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        th_all_acc += __shfl_xor_sync(0xffffffffu, th_all_acc, off);
        th_auto_acc += __shfl_xor_sync(0xffffffffu, th_auto_acc, off);
      }
      if ((tau & 31) == 0) {
        const double f = 1e-6 * __ldg(P.pol_factor + p);
        const int64_t o = ((int64_t)gout * P.nspw + s) * P.npol + p;
        atomicAdd(P.th_all + o, f * th_all_acc);
        atomicAdd(P.th_auto + o, f * th_auto_acc);
      }
    }
  }
}

template <int N>
static int launch_ws(const EngineParams &P, cudaStream_t st, int *launches) {
  using C = WsCfg<N>;
  static DeviceLatch configured;
  int cfg_dev = 0;
  if (configured.need(&cfg_dev)) {
    SDS_CUDA(cudaFuncSetAttribute(engine_ws_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    SDS_CUDA(cudaFuncSetAttribute(engine_ws_kernel<N>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    configured.set(cfg_dev);
  }
  int dev = 0, nsm = 148, per_sm = 1;
  SDS_CUDA(cudaGetDevice(&dev));
  SDS_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  SDS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, engine_ws_kernel<N>, C::THREADS, C::SMEM));
  if (per_sm < 1) return SDS_ENOTSUP;
  engine_ws_kernel<N><<<nsm * per_sm, C::THREADS, C::SMEM, st>>>(P);
  SDS_LAUNCH_CHECK("engine_ws_kernel");
  ++*launches;
  return SDS_OK;
}

// legs = data | noise | thermal, noise_mode 1 (moment-matched draw) or 2 (injected); SDS_ENOTSUP otherwise
int run_engine_ws(int N, const EngineParams &P, cudaStream_t st, int *launches) {
  if (P.noise_gauss || (P.noise_mode != 1 && P.noise_mode != 2)) return SDS_ENOTSUP;
  switch (N) {
    case 512: return launch_ws<512>(P, st, launches);
    case 1024: return launch_ws<1024>(P, st, launches);
    case 2048: return launch_ws<2048>(P, st, launches);
    default: return SDS_ENOTSUP;
  }
}

}  // namespace sds
