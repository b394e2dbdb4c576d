// Fused engine, complex64 / fp32 arithmetic (the bandwidth path of BASELINE.json).
//
// Same work decomposition as sds_engine.cu (item = group x window x pol x chunk of
// times, rows staged by 1-D bulk TMA through an mbarrier ring), re-organised around
// what the profile of the first version showed -- the kernel is bound by instruction
// ISSUE (65 % of slots at 8 warps / SM, 254 registers), not by DRAM or the FMA pipe:
//
//   * packed fp32x2 arithmetic (sds_packed.cuh): one FADD2 per complex add, two
//     FFMA2-class instructions per complex multiply, "* -i" as an operand modifier;
//   * a worker is TWO teams that consume the SAME staged rows: the data team
//     (flag mask * taper -> transform -> S1, S2 sums, thermal sums) and the noise
//     team (Philox4x32-10 + Box-Muller -> sigma(nsample, t_int) -> transform ->
//     sums).  Each team keeps only its own accumulators, which halves the
//     registers per thread and doubles the resident warps; the inputs are still
//     read from HBM exactly once (13 B / sample);
//   * slots are handed back through an `empty` mbarrier (one arrival per team), the
//     data team's lane 0 is the TMA producer.
#include "sds_engine.cuh"
#include "sds_packed.cuh"

namespace sds {

constexpr int E3_STAGES = 4;
#ifndef E3_MINBLOCKS
#define E3_MINBLOCKS 1
#endif
#ifndef E3_PHILOX_ROUNDS
#define E3_PHILOX_ROUNDS 10
#endif

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int ROUNDS>
__device__ __forceinline__ u32x4 philox4x32(u32x4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
    u32x4 n;
    n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
    n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
    n.w = (uint32_t)p0;
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}

template <int N, bool NOISE> struct E3Cfg {
  static constexpr int TPR = N / 8;                     // threads per row
  static constexpr int TEAM = TPR < 32 ? 32 : TPR;      // threads per team
  static constexpr int ROWS = TEAM / TPR;               // rows per step
  static constexpr int NTEAM = NOISE ? 2 : 1;
  static constexpr int WORKER = NTEAM * TEAM;
  static constexpr int NWORKER = WORKER >= 128 ? 1 : 128 / WORKER;
  static constexpr int THREADS = NWORKER * WORKER;
  static constexpr int ROWB = 13 * N;                   // vis 8N | nsample 4N | flags N
  static constexpr int SLOT = ROWS * ROWB;
  static constexpr int BUF = N * 8;                     // one exchange buffer of one row
  static constexpr int TW_BYTES = ((fft_tw_total(N) * 8 + 127) / 128) * 128;
  // per worker: 128 B (barriers, item mailbox) | E3_STAGES slots | NTEAM * ROWS * 2 buffers
  static constexpr int WORKER_BYTES = 128 + E3_STAGES * SLOT + NTEAM * ROWS * 2 * BUF;
  static constexpr int SMEM = TW_BYTES + NWORKER * WORKER_BYTES;
};

template <int N, int LEGS>
__global__ void __launch_bounds__(E3Cfg<N, (LEGS & LEG_NOISE) != 0>::THREADS, E3_MINBLOCKS)
engine_f32_kernel(const __grid_constant__ EngineParams P) {
  constexpr bool HAS_N = (LEGS & LEG_NOISE) != 0, HAS_T = (LEGS & LEG_THERMAL) != 0;
  using C = E3Cfg<N, HAS_N>;
  constexpr int TPR = C::TPR, TEAM = C::TEAM, ROWS = C::ROWS, WORKER = C::WORKER;
  extern __shared__ __align__(128) unsigned char smem[];

  f2 *tw_s = reinterpret_cast<f2 *>(smem);
  for (int i = threadIdx.x; i < fft_tw_total(N); i += C::THREADS)
    tw_s[i] = reinterpret_cast<const f2 *>(P.tw_f)[i];
  const int wid = threadIdx.x / WORKER, wt = threadIdx.x % WORKER;
  const int role = wt / TEAM;  // 0: data (+ thermal), 1: noise
  const int tl = wt % TEAM, sub = tl / TPR, tau = tl % TPR;
  const int bar_worker = 1 + 3 * wid, bar_team = bar_worker + 1 + role;
  unsigned char *wbase = smem + C::TW_BYTES + (size_t)wid * C::WORKER_BYTES;
  uint64_t *full = reinterpret_cast<uint64_t *>(wbase);
  uint64_t *empty = full + E3_STAGES;
  int *mailbox = reinterpret_cast<int *>(wbase + 16 * E3_STAGES);  // [2]
  unsigned char *stages = wbase + 128;
  f2 *bufA = reinterpret_cast<f2 *>(stages + E3_STAGES * C::SLOT +
                                    (size_t)((role * ROWS + sub) * 2) * C::BUF);
  f2 *bufB = bufA + N;
  if (wt == 0) {
    for (int i = 0; i < E3_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], C::NTEAM);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int total_items = P.item_start[P.ngroup];
  constexpr int ROW_BYTES = 13 * N;
  const int64_t row_stride = (int64_t)P.ntime * P.nchan;  // samples between baselines
  uint32_t slot_c = 0, par_c = 0;  // consumer ring position / phase parity
  uint32_t slot_p = 0, par_p = 1;  // producer ring position / parity of the `empty` wait

  for (int it = 0;; ++it) {
    int item;
    if constexpr (WORKER == 32) {
      item = 0;
      if (wt == 0) item = atomicAdd(P.counter, 1);
      item = __shfl_sync(0xffffffffu, item, 0);
    } else {
      if (wt == 0) mailbox[it & 1] = atomicAdd(P.counter, 1);
      asm volatile("bar.sync %0, %1;" ::"r"(bar_worker), "n"(WORKER) : "memory");
      item = mailbox[it & 1];
    }
    if (item >= total_items) break;
    int lo = 0, hi = P.ngroup;  // largest g with item_start[g] <= item
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(P.item_start + mid) <= item) lo = mid; else hi = mid;
    }
    const int g = lo, b0 = __ldg(P.goff + g), ng = __ldg(P.goff + g + 1) - b0;
    const int chunk = eng_chunk(ng, P.ntime, P.target_rows);
    const int nchunks = (P.ntime + chunk - 1) / chunk;
    const int local = item - __ldg(P.item_start + g);
    const int tc = local % nchunks, sp = local / nchunks, p = sp % P.npol, s = sp / P.npol;
    const int t0 = tc * chunk, t1 = min(P.ntime, t0 + chunk);
    const int steps_per_t = (ng + ROWS - 1) / ROWS, nsteps = (t1 - t0) * steps_per_t;
    // sample offset of (pol p, first baseline of the group, time t0, window start)
    const int64_t e00 = (((int64_t)p * P.nbl + b0) * P.ntime + t0) * P.nchan + P.win_start[s];
    const int64_t obase = (((int64_t)g * P.nspw + s) * P.npol + p) * N;
    int ci0 = 0;                                       // first row of the current step
    int64_t it_off = (int64_t)b0 * P.ntime + t0;       // int_time index of row 0 at this time

    f2 S1[8], S2[8];
    float Pw[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      S1[r] = 0ull; S2[r] = 0ull; Pw[r] = 0.f;
    }
    // close a time: P += |sum over the group's rows|^2
    auto close_time = [&]() {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        f2 sfull = S1[r];
#pragma unroll
        for (int off = TPR; off < 32; off <<= 1) sfull = f2_add(sfull, f2_shfl_xor(sfull, off));
        Pw[r] += f2_hsum(f2_mul(sfull, sfull));
        S1[r] = 0ull;
      }
    };
    // flush an item: fp64 atomics, fftshift as an index rotation
    auto flush = [&](double *out_all, double *out_auto) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int kshift = (tau + TPR * r + N / 2) % N;
        float s2 = f2_hsum(S2[r]);
#pragma unroll
        for (int off = TPR; off < 32; off <<= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        if (sub == 0) {
          atomicAdd(out_all + obase + kshift, (double)Pw[r]);
          atomicAdd(out_auto + obase + kshift, (double)s2);
        }
      }
    };

    if (role == 0) {
      // ======================= data team (+ thermal, + TMA producer) =================
      float tapdf[8];  // taper * delta_f at this thread's 8 transform channels
#pragma unroll
      for (int r = 0; r < 8; ++r) tapdf[r] = __ldg(P.taper_f + tau + TPR * r) * (float)P.delta_f;
      f2 tAA[8], tBB[8], tCC[8];  // thermal pair sums of the panel ends (left, right)
      double th_all_acc = 0.0, th_auto_acc = 0.0;
      if (HAS_T) {
#pragma unroll
        for (int r = 0; r < 8; ++r) tAA[r] = tBB[r] = tCC[r] = 0ull;
      }
      int pi0 = 0;          // first row of the next step to stage
      int64_t pe = e00;     // its sample offset
      int64_t pe_t = e00;   // offset of row 0 at the producer's current time
      int pleft = nsteps;   // steps not yet staged
      auto issue = [&]() {
        if (tl == 0) {
          mbar_wait(&empty[slot_p], par_p);  // both teams are done with the slot
          const int nact = min(ROWS, ng - pi0);
          mbar_expect_tx(&full[slot_p], (uint32_t)(nact * ROW_BYTES));
          unsigned char *dst = stages + slot_p * C::SLOT;
          int64_t e = pe;
          for (int rr = 0; rr < nact; ++rr, e += row_stride, dst += C::ROWB) {
            tma_load_1d(dst, P.vis + e, 8 * N, &full[slot_p]);
            tma_load_1d(dst + 8 * N, P.nsample + e, 4 * N, &full[slot_p]);
            tma_load_1d(dst + 12 * N, P.flags + e, N, &full[slot_p]);
          }
        }
        slot_p = slot_p + 1 == E3_STAGES ? 0 : slot_p + 1;
        if (slot_p == 0) par_p ^= 1u;
        pi0 += ROWS;
        pe += ROWS * row_stride;
        if (pi0 >= ng) {
          pi0 = 0;
          pe_t += P.nchan;
          pe = pe_t;
        }
        --pleft;
      };

      for (int k = 0; k < E3_STAGES - 1 && pleft > 0; ++k) issue();
      for (int step = 0; step < nsteps; ++step) {
        if (pleft > 0) issue();
        const int i = ci0 + sub;
        const bool active = ROWS == 1 || i < ng;
        const int ia = active ? i : 0;
        float tint = 1.f;
        if (HAS_T) tint = __ldg(P.int_time + it_off + (int64_t)ia * P.ntime);
        mbar_wait(&full[slot_c], par_c);
        const unsigned char *row = stages + slot_c * C::SLOT + sub * C::ROWB;
        const f2 *vis_s = reinterpret_cast<const f2 *>(row);
        const float *ns_s = reinterpret_cast<const float *>(row + 8 * N);
        const unsigned char *fl_s = row + 12 * N;

        f2 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          // (1 - flag) * taper * delta_f  (ds.py:3503-3505, utils.py:552-560)
          const float m = (active && fl_s[tau + TPR * r] == 0) ? tapdf[r] : 0.f;
          const f2 c = active ? vis_s[tau + TPR * r] : 0ull;
          v[r] = f2_scale(c, m);
        }
        TeamFftP<N>::run(v, bufA, bufB, tw_s, tau, bar_team);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          S1[r] = f2_add(S1[r], v[r]);
          S2[r] = f2_norm_acc(S2[r], v[r]);
        }

        if (HAS_T) {
          // thermal sums use their own channel map: two runs of 4 consecutive channels
          // per thread (c = 4 tau + 4 TPR h + j) -> two 128-bit loads + the right neighbours
          const float itv = (tint > 0.f) ? 1.f / tint : 0.f;
          const f2 it2 = f2_make(itv, itv);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c0 = 4 * tau + 4 * TPR * h;
            const float4 nv = *reinterpret_cast<const float4 *>(ns_s + c0);
            const float nx = (c0 + 4 < N) ? ns_s[c0 + 4] : 0.f;  // no panel right of channel N-1
            const float n5[5] = {nv.x, nv.y, nv.z, nv.w, nx};
            float a5[5];
            bool ok[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              ok[j] = active && n5[j] > 0.f;
              a5[j] = ok[j] ? fast_rsqrt(n5[j]) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // a panel counts only when both of its ends are valid (masked_invalid, ds.py:3697)
              const f2 aa = f2_make(ok[j + 1] ? a5[j] : 0.f, ok[j] ? a5[j + 1] : 0.f);
              const int q = 4 * h + j;
              tAA[q] = f2_add(tAA[q], aa);
              tBB[q] = f2_fma(aa, it2, tBB[q]);
              tCC[q] = f2_fma(f2_mul(aa, aa), it2, tCC[q]);
            }
          }
        }

        // every lane has read its part of the slot: hand it back to the producer
        team_sync<TEAM>(bar_team);
        if (tl == 0) mbar_arrive(&empty[slot_c]);
        slot_c = slot_c + 1 == E3_STAGES ? 0 : slot_c + 1;
        if (slot_c == 0) par_c ^= 1u;
        ci0 += ROWS;
        if (ci0 >= ng) {  // last rows of this time
          ci0 = 0;
          ++it_off;
          close_time();
          if (HAS_T) {
            const double *wlp = P.wl + ((int64_t)s * P.npol + p) * N;
            const double *wrp = P.wr + ((int64_t)s * P.npol + p) * N;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              f2 ab = tAA[q], bb = tBB[q];
#pragma unroll
              for (int off = TPR; off < 32; off <<= 1) {
                ab = f2_add(ab, f2_shfl_xor(ab, off));
                bb = f2_add(bb, f2_shfl_xor(bb, off));
              }
              const int c = 4 * tau + 4 * TPR * (q >> 2) + (q & 3);
              const double l = __ldg(wlp + c), rr = __ldg(wrp + c);
              if (sub == 0)
                th_all_acc += l * (double)f2_lo(ab) * (double)f2_lo(bb) +
                              rr * (double)f2_hi(ab) * (double)f2_hi(bb);
              th_auto_acc += l * (double)f2_lo(tCC[q]) + rr * (double)f2_hi(tCC[q]);
              tAA[q] = tBB[q] = tCC[q] = 0ull;
            }
          }
        }
      }
      flush(P.pow_all, P.pow_auto);
      if (HAS_T) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          th_all_acc += __shfl_xor_sync(0xffffffffu, th_all_acc, off);
          th_auto_acc += __shfl_xor_sync(0xffffffffu, th_auto_acc, off);
        }
        if ((tl & 31) == 0) {
          const double f = 1e-6 * __ldg(P.pol_factor + p);
          const int64_t o = ((int64_t)g * P.nspw + s) * P.npol + p;
          atomicAdd(P.th_all + o, f * th_all_acc);
          atomicAdd(P.th_auto + o, f * th_auto_acc);
        }
      }
    } else if constexpr (HAS_N) {
      // ================================ noise team ====================================
      float sigc[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float tap = __ldg(P.taper_f + tau + TPR * r) * (float)P.delta_f;
        if (P.noise_mode == 1) {
          const float c = __ldg(P.sigma_coef + ((int64_t)s * P.npol + p) * N + tau + TPR * r);
          // non-finite sigma -> 0 (ds.py:3575-3582); taper * delta_f folded in
          sigc[r] = (fabsf(c) <= 3.0e38f) ? c * tap : 0.f;
        } else {
          sigc[r] = tap;
        }
      }
      int ct = t0;
      // Philox counter base of (row 0, time t0): row_id * N/2 with
      // row_id = ((s npol + p) nbl_total + b) ntime_total + t   (include/sds.h)
      uint64_t rid_t = (((uint64_t)((int64_t)s * P.npol + p) * (uint64_t)P.nbl_total +
                         (uint64_t)(P.bl_offset + b0)) * (uint64_t)P.ntime_total +
                        (uint64_t)(P.time_offset + t0));

      for (int step = 0; step < nsteps; ++step) {
        const int i = ci0 + sub;
        const bool active = ROWS == 1 || i < ng;
        const int ia = active ? i : 0;
        const float tint = __ldg(P.int_time + it_off + (int64_t)ia * P.ntime);
        f2 v[8];
        if (P.noise_mode == 1) {
          // unit draws first: they do not depend on the staged row
          // counter = row_id * N/2 + channel mod N/2 ; words (x,y) -> channel f < N/2,
          // (z,w) -> f + N/2: both belong to this thread (slots q and q + 4)
          const uint64_t ctr0 = (rid_t + (uint64_t)ia * (uint64_t)P.ntime_total) * (uint64_t)(N / 2) +
                                (uint64_t)tau;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint64_t ctr = ctr0 + (uint64_t)(TPR * q);
            const u32x4 rn = philox4x32<E3_PHILOX_ROUNDS>(
                {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u}, P.seed_lo, P.seed_hi);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t a = h ? rn.z : rn.x, b = h ? rn.w : rn.y;
              const float u1 = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
              const float rad = fast_sqrt(-0.69314718055994531f * __log2f(u1));  // sqrt(-ln u1)
              float sn, cs;
              __sincosf(6.2831853071795865f * ((float)b * 2.3283064365386963e-10f), &sn, &cs);
              v[q + 4 * h] = f2_make(rad * cs, rad * sn);
            }
          }
        }
        mbar_wait(&full[slot_c], par_c);
        const unsigned char *row = stages + slot_c * C::SLOT + sub * C::ROWB;
        const float *ns_s = reinterpret_cast<const float *>(row + 8 * N);
        const unsigned char *fl_s = row + 12 * N;
        if (P.noise_mode == 1) {
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float tn = tint * ns_s[tau + TPR * r];
            const bool keep = active && fl_s[tau + TPR * r] == 0 && tn > 0.f;
            // sigma = coef / sqrt(t_int nsample), with mask, taper and delta_f folded into sigc
            const float amp = keep ? sigc[r] * fast_rsqrt(tn) : 0.f;
            v[r] = f2_scale(v[r], amp);
          }
        } else {  // injected freq-domain noise (parity mode): plain loads
          const int64_t e = e00 + (int64_t)(ct - t0) * P.nchan + (int64_t)ia * row_stride;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float m = (active && fl_s[tau + TPR * r] == 0) ? sigc[r] : 0.f;
            const float2 c = active ? P.noise_in[e + tau + TPR * r] : make_float2(0.f, 0.f);
            v[r] = f2_make(c.x * m, c.y * m);
          }
        }
        team_sync<TEAM>(bar_team);
        if (tl == 0) mbar_arrive(&empty[slot_c]);

        TeamFftP<N>::run(v, bufA, bufB, tw_s, tau, bar_team);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          S1[r] = f2_add(S1[r], v[r]);
          S2[r] = f2_norm_acc(S2[r], v[r]);
        }
        slot_c = slot_c + 1 == E3_STAGES ? 0 : slot_c + 1;
        if (slot_c == 0) par_c ^= 1u;
        ci0 += ROWS;
        if (ci0 >= ng) {
          ci0 = 0;
          ++ct;
          ++it_off;
          ++rid_t;
          close_time();
        }
      }
      flush(P.noise_all, P.noise_auto);
    }
  }
}

template <int N, int LEGS>
static int launch_f32(const EngineParams &P, cudaStream_t st, int *launches) {
  using C = E3Cfg<N, (LEGS & LEG_NOISE) != 0>;
  static bool configured = false;
  if (!configured) {
    SDS_CUDA(cudaFuncSetAttribute(engine_f32_kernel<N, LEGS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    // several CTAs per SM only fit when the L1/shared split favours shared memory
    SDS_CUDA(cudaFuncSetAttribute(engine_f32_kernel<N, LEGS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  int dev = 0, nsm = 148, per_sm = 1;
  SDS_CUDA(cudaGetDevice(&dev));
  SDS_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  SDS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, engine_f32_kernel<N, LEGS>,
                                                         C::THREADS, C::SMEM));
  if (per_sm < 1) {
    set_error("sds_engine_run: kernel does not fit an SM (smem %d B)", C::SMEM);
    return SDS_ENOTSUP;
  }
  engine_f32_kernel<N, LEGS><<<nsm * per_sm, C::THREADS, C::SMEM, st>>>(P);
  SDS_LAUNCH_CHECK("engine_f32_kernel");
  ++*launches;
  return SDS_OK;
}

template <int N>
static int run_f32_n(const EngineParams &P, int legs, cudaStream_t st, int *launches) {
  switch (legs) {
    case 1: return launch_f32<N, 1>(P, st, launches);
    case 3: return launch_f32<N, 3>(P, st, launches);
    case 5: return launch_f32<N, 5>(P, st, launches);
    case 7: return launch_f32<N, 7>(P, st, launches);
    default: break;
  }
  set_error("sds_engine_run: leg mask %d is not instantiated", legs);
  return SDS_ENOTSUP;
}

int run_engine_f32(int N, const EngineParams &P, int legs, cudaStream_t st, int *launches) {
  switch (N) {
    case 64: return run_f32_n<64>(P, legs, st, launches);
    case 128: return run_f32_n<128>(P, legs, st, launches);
    case 256: return run_f32_n<256>(P, legs, st, launches);
    case 512: return run_f32_n<512>(P, legs, st, launches);
    case 1024: return run_f32_n<1024>(P, legs, st, launches);
    case 2048: return run_f32_n<2048>(P, legs, st, launches);
    default:
      set_error("sds_engine_run: fused path needs nfreq in {64..2048, power of two}, got %d", N);
      return SDS_ENOTSUP;
  }
}

}  // namespace sds
