// Fused engine for spectral windows whose length N is NOT a power of two (21, 203 = 7 x 29
// channels of the reference's bundled files, BASELINE.json configs[0..1]) -- complex64 / fp32.
//
// Same work decomposition, staging and sums as sds_engine_f32.cu; the transform of a row is
// Bluestein's chirp-z form (sds_fft_bluestein.cu): with c[n] = exp(-i pi n^2 / N),
//     X[k] = c[k] * sum_n (x[n] c[n]) conj(c)[k - n],
// a circular convolution of power-of-two length M >= 2N - 1 done with two register-radix
// transforms (the inverse one as conj(FFT(conj(.))) / M).  A row team is M / 8 threads; thread
// tau owns positions tau + (M/8) r of the padded sequence, of which only r < 4 can be < N, so
// a thread carries at most FOUR channels / delays (accumulators, constants, draws).  The data
// and noise rows go through both transforms in lockstep (sds_packed.cuh: run2).
#include "sds_engine.cuh"
#include "sds_packed.cuh"

namespace sds {

#ifndef EB_STAGES
#define EB_STAGES 3
#endif

template <int M> struct EBCfg {
  static constexpr int TPR = M / 8;
  static constexpr int TEAM = TPR < 32 ? 32 : TPR;
  static constexpr int ROWS = TEAM / TPR;
  static constexpr int THREADS = TEAM < 128 ? 128 : TEAM;
  static constexpr int NWORKER = THREADS / TEAM;
  static constexpr int MINB = 384 / THREADS < 1 ? 1 : 384 / THREADS;  // aim at 12 warps / SM
  static constexpr int NPMAX = M / 2 + 16;               // padded window: ceil16(N) <= M/2 + 16
  static constexpr int ROWB = 13 * NPMAX;                // vis 8 NP | nsample 4 NP | flags NP
  static constexpr int SLOT = ROWS * ROWB;
  static constexpr int BUF = M * 8;
  static constexpr int XREGION = NWORKER * ROWS * 2 * BUF;  // all exchange buffers, aligned to BUF
  static constexpr int TW_BYTES = ((fft_tw_total(M) * 8 + 127) / 128) * 128;
  // per worker: 64 B barriers | EB_STAGES slots | 2 * 4 TPR floats
  static constexpr int WORKER_BYTES = 64 + EB_STAGES * SLOT + 32 * TPR;
  // BUF bytes of slack in front: the exchange region is aligned to BUF at run time (TeamFftX)
  static constexpr int SMEM = BUF + XREGION + TW_BYTES + NWORKER * WORKER_BYTES;
};

__device__ __forceinline__ f2 f2_conj(f2 a) { return f2_make(f2_lo(a), -f2_hi(a)); }

template <int M, int LEGS>
__global__ void __launch_bounds__(EBCfg<M>::THREADS, EBCfg<M>::MINB)
engine_blue_kernel(const __grid_constant__ EngineParams P) {
  constexpr bool HAS_N = (LEGS & LEG_NOISE) != 0, HAS_T = (LEGS & LEG_THERMAL) != 0;
  using C = EBCfg<M>;
  using FFT = TeamFftX<M>;
  constexpr int TPR = C::TPR, TEAM = C::TEAM, ROWS = C::ROWS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t s0 = smem_u32(smem_raw);
  unsigned char *smem = smem_raw + (((s0 + C::BUF - 1u) & ~(uint32_t)(C::BUF - 1)) - s0);

  f2 *tw_s = reinterpret_cast<f2 *>(smem + C::XREGION);
  for (int i = threadIdx.x; i < fft_tw_total(M); i += C::THREADS)
    tw_s[i] = reinterpret_cast<const f2 *>(P.blue_tw)[i];
  const int wid = threadIdx.x / TEAM, tl = threadIdx.x % TEAM;
  const int sub = tl / TPR, tau = tl % TPR, bar_id = 1 + wid;
  const uint32_t xa = smem_u32(smem) + (uint32_t)((wid * ROWS + sub) * 2) * C::BUF;
  const typename FFT::Addr fa = FFT::make_addr(xa, tau);
  unsigned char *wbase = smem + C::XREGION + C::TW_BYTES + (size_t)wid * C::WORKER_BYTES;
  uint64_t *full = reinterpret_cast<uint64_t *>(wbase);
  unsigned char *stages = wbase + 64;
  // per-worker sums over times of |S1|^2: [0, 4 TPR) data, [4 TPR, 8 TPR) noise
  float *pw_s = reinterpret_cast<float *>(stages + EB_STAGES * C::SLOT);
  for (int i = tl; i < 8 * TPR; i += TEAM) pw_s[i] = 0.f;
  if (tl == 0) {
    for (int i = 0; i < EB_STAGES; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int N = P.nfreq, NP = (N + 15) & ~15;  // window length and its padded row piece
  const int c0 = N / 2;
  // per-thread constants at its channels / delays k = tau + TPR r (r < 4; zero beyond N)
  f2 cin[4], cout[4], bh[8];
  float tapdf[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int n = tau + TPR * r;
    const float2 c = n < N ? __ldg(P.blue_chirp + n) : make_float2(0.f, 0.f);
    tapdf[r] = n < N ? __ldg(P.taper_f + n) * (float)P.delta_f : 0.f;
    cin[r] = f2_make(c.x, c.y);
    cout[r] = f2_make(c.x * (1.f / M), c.y * (1.f / M));  // output chirp and the inverse's 1/M
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const float2 b = __ldg(P.blue_bhat + tau + TPR * r);
    bh[r] = f2_make(b.x, b.y);
  }
  const int total_items = P.item_start[P.ngroup];
  const uint32_t row_bytes = 13u * (uint32_t)NP;
  const int64_t row_stride = (int64_t)P.ntime * P.nchan;
  const bool issuer_warp = TEAM == 32 || __shfl_sync(0xffffffffu, tl, 0) == 0;
  uint32_t slot_c = 0, par_c = 0, slot_p = 0;
  __shared__ int item_bcast[C::NWORKER];

  for (;;) {
    int item;
    if constexpr (TEAM == 32) {
      item = 0;
      if (tl == 0) item = atomicAdd(P.counter, 1);
      item = __shfl_sync(0xffffffffu, item, 0);
    } else {
      team_sync<TEAM>(bar_id);
      if (tl == 0) item_bcast[wid] = atomicAdd(P.counter, 1);
      team_sync<TEAM>(bar_id);
      item = __shfl_sync(0xffffffffu, item_bcast[wid], 0);
    }
    if (item >= total_items) break;
    int lo = 0, hi = P.ngroup;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(P.item_start + mid) <= item) lo = mid; else hi = mid;
    }
    const int g = lo, b0 = __ldg(P.goff + g), ng = __ldg(P.goff + g + 1) - b0;
    // a sub-list of groups (group sharding): where the group's sums go and where its baselines sit
    // in the whole array (the noise stream is keyed by the global baseline index)
    const int gout = P.group_id ? __ldg(P.group_id + g) : g;
    const int64_t blg = P.group_bl0 ? __ldg(P.group_bl0 + g) : P.bl_offset + b0;
    const int chunk = eng_chunk(ng, P.ntime, P.target_rows);
    const int nchunks = (P.ntime + chunk - 1) / chunk;
    const int local = item - __ldg(P.item_start + g);
    const int tc = local % nchunks, sp = local / nchunks, p = sp % P.npol, s = sp / P.npol;
    const int t0 = tc * chunk, t1 = min(P.ntime, t0 + chunk);
    const int steps_per_t = (ng + ROWS - 1) / ROWS, nsteps = (t1 - t0) * steps_per_t;
    const int64_t e00 = (((int64_t)p * P.nbl + b0) * P.ntime + t0) * P.nchan + P.win_start[s];
    const int64_t obase = (((int64_t)gout * P.nspw + s) * P.npol + p) * N;

    int pi0 = 0, pleft = nsteps;
    int64_t pe = e00, pe_t = e00;
    auto issue = [&]() {
      if (issuer_warp && elect_one()) {
        const int nact = min(ROWS, ng - pi0);
        mbar_expect_tx(&full[slot_p], (uint32_t)nact * row_bytes);
        unsigned char *dst = stages + slot_p * C::SLOT;
        int64_t e = pe;
        for (int rr = 0; rr < nact; ++rr, e += row_stride, dst += C::ROWB) {
          tma_load_1d(dst, P.vis + e, 8u * NP, &full[slot_p]);
          tma_load_1d(dst + 8 * C::NPMAX, P.nsample + e, 4u * NP, &full[slot_p]);
          tma_load_1d(dst + 12 * C::NPMAX, P.flags + e, (uint32_t)NP, &full[slot_p]);
        }
      }
      slot_p = slot_p + 1 == EB_STAGES ? 0 : slot_p + 1;
      pi0 += ROWS;
      pe += ROWS * row_stride;
      if (pi0 >= ng) {
        pi0 = 0;
        pe_t += P.nchan;
        pe = pe_t;
      }
      --pleft;
    };

    float sigc[4];
    if (HAS_N) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = tau + TPR * r;
        sigc[r] = 0.f;
        if (n < N) {
          if (P.noise_mode == 1) {
            const float c = __ldg(P.sigma_coef + ((int64_t)s * P.npol + p) * N + n);
            sigc[r] = (fabsf(c) <= 3.0e38f) ? c * ENG_SQRT_LN2 * tapdf[r] : 0.f;  // ds.py:3575-3582
          } else {
            sigc[r] = tapdf[r];
          }
        }
      }
    }
    f2 S1[4], Sn1[4];
    float S2[4], Sn2[4];
    f2 tAA[4], tBB[4], tCC[4];  // thermal pair sums (left, right end) of the panel at channel n
    double th_all_acc = 0.0, th_auto_acc = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      S1[r] = Sn1[r] = 0ull;
      S2[r] = Sn2[r] = 0.f;
      tAA[r] = tBB[r] = tCC[r] = 0ull;
    }
    int ci0 = 0, ct = t0;
    int64_t it_off = (int64_t)b0 * P.ntime + t0;
    uint64_t rid_t = (((uint64_t)((int64_t)s * P.npol + p) * (uint64_t)P.nbl_total +
                       (uint64_t)blg) * (uint64_t)P.ntime_total +
                      (uint64_t)(P.time_offset + t0));

    for (int k = 0; k < EB_STAGES - 1 && pleft > 0; ++k) issue();
    for (int step = 0; step < nsteps; ++step) {
      team_sync<TEAM>(bar_id);
      if (pleft > 0) issue();
      const int i = ci0 + sub;
      const bool active = ROWS == 1 || i < ng;
      const int ia = active ? i : 0;
      float tint = 1.f;
      if (HAS_N || HAS_T) tint = __ldg(P.int_time + it_off + (int64_t)ia * P.ntime);
      mbar_wait(&full[slot_c], par_c);
      const unsigned char *row = stages + slot_c * C::SLOT + sub * C::ROWB;
      const f2 *vis_s = reinterpret_cast<const f2 *>(row);
      const float *ns_s = reinterpret_cast<const float *>(row + 8 * C::NPMAX);
      const unsigned char *fl_s = row + 12 * C::NPMAX;

      f2 vd[8], vn[8];
#pragma unroll
      for (int r = 4; r < 8; ++r) vd[r] = vn[r] = 0ull;  // zero padding of the convolution
      float rad2[4], cs[4], sn[4];
      if (HAS_N && P.noise_mode == 1)
        engine_draw4<TPR>(rid_t + (uint64_t)ia * (uint64_t)P.ntime_total, tau, P.pkey, rad2, cs, sn);
      const int64_t e_in = e00 + (int64_t)(ct - t0) * P.nchan + (int64_t)ia * row_stride;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = tau + TPR * r;
        const bool in = active && n < N;
        const bool k = in && fl_s[in ? n : 0] == 0;
        // (1 - flag) * taper * delta_f * chirp  (ds.py:3503-3505, utils.py:552-560)
        vd[r] = f2_cmul(in ? vis_s[n] : 0ull, f2_scale(cin[r], k ? tapdf[r] : 0.f));
        if (HAS_N) {
          if (P.noise_mode == 1) {
            const float tn = tint * (in ? ns_s[n] : 0.f);
            const float amp = (k && tn > 0.f) ? sigc[r] * rad2[r] * fast_rsqrt(tn) : 0.f;
            vn[r] = f2_cmul(f2_make(cs[r], sn[r]), f2_scale(cin[r], amp));
          } else {  // injected frequency-domain noise (parity mode)
            const float2 c = in ? P.noise_in[e_in + n] : make_float2(0.f, 0.f);
            vn[r] = f2_cmul(f2_make(c.x, c.y), f2_scale(cin[r], k ? sigc[r] : 0.f));
          }
        }
      }

      if (HAS_T) {
        const float itv = (tint > 0.f) ? 1.f / tint : 0.f;
        const f2 it2 = f2_make(itv, itv);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int n = tau + TPR * r;
          if (active && n + 1 < N) {  // panel (n, n + 1): counts only when both ends are valid
            const float nl = ns_s[n], nr = ns_s[n + 1];
            if (nl > 0.f && nr > 0.f) {
              const f2 aa = f2_make(fast_rsqrt(nl), fast_rsqrt(nr));
              tAA[r] = f2_add(tAA[r], aa);
              tBB[r] = f2_fma(aa, it2, tBB[r]);
              tCC[r] = f2_fma(f2_mul(aa, aa), it2, tCC[r]);
            }
          }
        }
      }

      // chirp-z: FFT_M, multiply by the transform of the wrapped chirp, inverse FFT_M
      if (HAS_N) {
        FFT::run2(vd, vn, fa, tw_s, tau, bar_id);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          vd[r] = f2_conj(f2_cmul(vd[r], bh[r]));
          vn[r] = f2_conj(f2_cmul(vn[r], bh[r]));
        }
        team_sync<TEAM>(bar_id);
        FFT::run2(vd, vn, fa, tw_s, tau, bar_id);
      } else {
        FFT::run(vd, fa, tw_s, tau, bar_id);
#pragma unroll
        for (int r = 0; r < 8; ++r) vd[r] = f2_conj(f2_cmul(vd[r], bh[r]));
        team_sync<TEAM>(bar_id);
        FFT::run(vd, fa, tw_s, tau, bar_id);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {  // X[k] = c[k] / M * conj(.)   (zero beyond N: cout = 0)
        const f2 x = f2_cmul(f2_conj(vd[r]), cout[r]);
        S1[r] = f2_add(S1[r], x);
        S2[r] = fmaf(f2_lo(x), f2_lo(x), fmaf(f2_hi(x), f2_hi(x), S2[r]));
        if (HAS_N) {
          const f2 y = f2_cmul(f2_conj(vn[r]), cout[r]);
          Sn1[r] = f2_add(Sn1[r], y);
          Sn2[r] = fmaf(f2_lo(y), f2_lo(y), fmaf(f2_hi(y), f2_hi(y), Sn2[r]));
        }
      }

      slot_c = slot_c + 1 == EB_STAGES ? 0 : slot_c + 1;
      if (slot_c == 0) par_c ^= 1u;
      ci0 += ROWS;
      if (ci0 >= ng) {  // close the time
        ci0 = 0;
        ++ct;
        ++it_off;
        ++rid_t;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          f2 sd = S1[r], sq = Sn1[r];
#pragma unroll
          for (int off = TPR; off < 32; off <<= 1) {
            sd = f2_add(sd, f2_shfl_xor(sd, off));
            if (HAS_N) sq = f2_add(sq, f2_shfl_xor(sq, off));
          }
          if (sub == 0) pw_s[tau + TPR * r] += f2_hsum(f2_mul(sd, sd));
          S1[r] = 0ull;
          if (HAS_N) {
            if (sub == 0) pw_s[4 * TPR + tau + TPR * r] += f2_hsum(f2_mul(sq, sq));
            Sn1[r] = 0ull;
          }
        }
        if (HAS_T) {
          const double *wlp = P.wl + ((int64_t)s * P.npol + p) * N;
          const double *wrp = P.wr + ((int64_t)s * P.npol + p) * N;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            f2 ab = tAA[r], bb = tBB[r];
#pragma unroll
            for (int off = TPR; off < 32; off <<= 1) {
              ab = f2_add(ab, f2_shfl_xor(ab, off));
              bb = f2_add(bb, f2_shfl_xor(bb, off));
            }
            const int n = tau + TPR * r;
            if (n < N) {
              const double l = __ldg(wlp + n), rr = __ldg(wrp + n);
              if (sub == 0)
                th_all_acc += l * (double)f2_lo(ab) * (double)f2_lo(bb) +
                              rr * (double)f2_hi(ab) * (double)f2_hi(bb);
              th_auto_acc += l * (double)f2_lo(tCC[r]) + rr * (double)f2_hi(tCC[r]);
            }
            tAA[r] = tBB[r] = tCC[r] = 0ull;
          }
        }
      }
    }

    // flush the item: fp64 atomics, fftshift (bin k lands at (k + N // 2) mod N)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int k = tau + TPR * r;
      float s2 = S2[r], sn2 = Sn2[r];
#pragma unroll
      for (int off = TPR; off < 32; off <<= 1) {
        s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        if (HAS_N) sn2 += __shfl_xor_sync(0xffffffffu, sn2, off);
      }
      if (sub == 0 && k < N) {
        int kk = k + c0;
        if (kk >= N) kk -= N;
        atomicAdd(P.pow_all + obase + kk, (double)pw_s[k]);
        atomicAdd(P.pow_auto + obase + kk, (double)s2);
        if (HAS_N) {
          atomicAdd(P.noise_all + obase + kk, (double)pw_s[4 * TPR + k]);
          atomicAdd(P.noise_auto + obase + kk, (double)sn2);
        }
      }
      if (sub == 0) {
        pw_s[k] = 0.f;
        pw_s[4 * TPR + k] = 0.f;
      }
    }
    if (HAS_T) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        th_all_acc += __shfl_xor_sync(0xffffffffu, th_all_acc, off);
        th_auto_acc += __shfl_xor_sync(0xffffffffu, th_auto_acc, off);
      }
      if ((tl & 31) == 0) {
        const double f = 1e-6 * __ldg(P.pol_factor + p);
        const int64_t o = ((int64_t)gout * P.nspw + s) * P.npol + p;
        atomicAdd(P.th_all + o, f * th_all_acc);
        atomicAdd(P.th_auto + o, f * th_auto_acc);
      }
    }
  }
}

template <int M, int LEGS>
static int launch_blue(const EngineParams &P, cudaStream_t st, int *launches) {
  using C = EBCfg<M>;
  static DeviceLatch configured;
  int cfg_dev = 0;
  if (configured.need(&cfg_dev)) {
    SDS_CUDA(cudaFuncSetAttribute(engine_blue_kernel<M, LEGS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    SDS_CUDA(cudaFuncSetAttribute(engine_blue_kernel<M, LEGS>,
                                  cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    configured.set(cfg_dev);
  }
  int dev = 0, nsm = 148, per_sm = 1;
  SDS_CUDA(cudaGetDevice(&dev));
  SDS_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  SDS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, engine_blue_kernel<M, LEGS>,
                                                         C::THREADS, C::SMEM));
  if (per_sm < 1) {
    set_error("sds_engine_run: kernel does not fit an SM (smem %d B)", C::SMEM);
    return SDS_ENOTSUP;
  }
  engine_blue_kernel<M, LEGS><<<nsm * per_sm, C::THREADS, C::SMEM, st>>>(P);
  SDS_LAUNCH_CHECK("engine_blue_kernel");
  ++*launches;
  return SDS_OK;
}

template <int M>
static int run_blue_m(const EngineParams &P, int legs, cudaStream_t st, int *launches) {
  switch (legs) {
    case 1: return launch_blue<M, 1>(P, st, launches);
    case 3: return launch_blue<M, 3>(P, st, launches);
    case 5: return launch_blue<M, 5>(P, st, launches);
    case 7: return launch_blue<M, 7>(P, st, launches);
    default: break;
  }
  set_error("sds_engine_run: leg mask %d is not instantiated", legs);
  return SDS_ENOTSUP;
}

int run_engine_blue(int M, const EngineParams &P, int legs, cudaStream_t st, int *launches) {
  switch (M) {
    case 64: return run_blue_m<64>(P, legs, st, launches);
    case 128: return run_blue_m<128>(P, legs, st, launches);
    case 256: return run_blue_m<256>(P, legs, st, launches);
    case 512: return run_blue_m<512>(P, legs, st, launches);
    case 1024: return run_blue_m<1024>(P, legs, st, launches);
    case 2048: return run_blue_m<2048>(P, legs, st, launches);
    default:
      set_error("sds_engine_run: no chirp-z engine for convolution length %d", M);
      return SDS_ENOTSUP;
  }
}

}  // namespace sds
