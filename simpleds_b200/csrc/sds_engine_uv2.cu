// Fused engine for TWO data sets (Nuv = 2), complex64 / fp32 arithmetic.
//
// The reference cross-multiplies data_array[:, 0] with data_array[:, 1] (delay_spectrum.py:3620-3633,
// utils.py:387-389 with axis = 2):  P[i, j] = A2_i conj(A1_j),  A1 = transform of set 0, A2 = of set 1,
// and the same for the two noise realisations.  The sums over baseline pairs factorise like the
// one-set case, but stay complex:
//     sum_ij A2_i conj(A1_j) = (sum_i A2_i) conj(sum_j A1_j),      autos: sum_i A2_i conj(A1_i).
// Same work decomposition as sds_engine_f32.cu (item = group x window x pol x chunk of times, rows
// staged by 1-D bulk TMA through an mbarrier ring, one row team of max(32, N / 8) threads per worker,
// packed fp32x2 arithmetic, XOR-addressed exchange); the two transforms that run in lockstep are the
// two data sets' rows of one (baseline, time) -- or, in the noise launch, their two noise rows.
// One launch per leg (LEG_DATA: 18 B per sample pair; LEG_NOISE: 10 B): the accumulators of all four
// transforms do not fit the register budget of three CTAs per SM.  The thermal estimate of two data
// sets (delay_spectrum.py:3655-3707: N_ij = sqrt(n1_j n2_i), integration time of baseline j) rides in
// the noise launch, which stages both nsample arrays anyway: with a_u = nsample_u^-1/2 every factor
// separates,  sum_ij = sum_c w[c] (sum_i a2_i)(sum_j a1_j / t_j),  autos  sum_i a2_i a1_i / t_i,  and so
// does the panel mask (masked_invalid, 3697): a trapezoid panel counts for a pair when both of its ends
// are valid in set 1 for i and in set 0 for j (full sums + rare per-row deficits, as in the one-set
// engine).
#include "sds_engine.cuh"
#include "sds_packed.cuh"

namespace sds {

constexpr int UV_STAGES = 2;

template <int N, int LEG> struct UvCfg {
  static_assert(LEG == LEG_DATA || LEG == LEG_NOISE || LEG == (LEG_NOISE | LEG_THERMAL), "one leg per launch");
  static constexpr bool DATA = LEG == LEG_DATA;
  static constexpr int TPR = N / 8;
  static constexpr int TEAM = TPR < 32 ? 32 : TPR;
  static constexpr int ROWS = TEAM / TPR;
  static constexpr int THREADS = TEAM < 128 ? 128 : TEAM;
  static constexpr int NWORKER = THREADS / TEAM;
  static constexpr int MINB = 384 / THREADS < 1 ? 1 : 384 / THREADS;  // 12 warps per SM
  static constexpr int VALB = DATA ? 8 * N : 4 * N;   // vis (8 B) or nsample (4 B) per set
  static constexpr int ROWB = 2 * VALB + 2 * N;       // set 0 | set 1 | flags 0 | flags 1
  static constexpr int SLOT = ROWS * ROWB;
  static constexpr int BUF = N * 8;
  // the two lockstep transforms exchange through one padded buffer of 16-byte slots (TeamFftX::run2_il);
  // one-warp teams whose last pass has radix 4 (N = 256) make their last exchange through tensor memory
  static constexpr bool TM = TeamFftX<N>::TM_OK;
  static constexpr int XROW = TeamFftX<N>::IL_BYTES;
  static constexpr int XREGION = NWORKER * ROWS * XROW;
  static constexpr int TW_BYTES = ((fft_tw_total(N) * 8 + 127) / 128) * 128;
  // per worker: 64 B barriers | UV_STAGES slots | 2 N floats (complex sums over times)
  static constexpr int WORKER_BYTES = 64 + UV_STAGES * SLOT + 8 * N;
  static constexpr int SMEM = 128 + XREGION + TW_BYTES + NWORKER * WORKER_BYTES;
};

// acc + b conj(a)
__device__ __forceinline__ f2 f2_cmac_conj(f2 acc, f2 b, f2 a) {
  const float ax = f2_lo(a), ay = f2_hi(a);
  return f2_fma(f2_make(ay, ay), f2_mul_mi(b), f2_fma(f2_make(ax, ax), b, acc));
}

template <int N, int LEG>
__global__ void __launch_bounds__(UvCfg<N, LEG>::THREADS, UvCfg<N, LEG>::MINB)
engine_uv2_kernel(const __grid_constant__ EngineParams P) {
  using C = UvCfg<N, LEG>;
  using FFT = TeamFftX<N>;
  constexpr bool DATA = C::DATA, THERMAL = (LEG & LEG_THERMAL) != 0;
  constexpr int TPR = C::TPR, TEAM = C::TEAM, ROWS = C::ROWS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t s0 = smem_u32(smem_raw);
  unsigned char *smem = smem_raw + (((s0 + 127u) & ~127u) - s0);
  // shared-space addresses passed through an empty asm: not re-derived inside the step loop (sds_engine_f32.cu)
  uint32_t sb_a = smem_u32(smem);
  asm volatile("" : "+r"(sb_a));

  f2 *tw_s = reinterpret_cast<f2 *>(__cvta_shared_to_generic(sb_a + C::XREGION));
  for (int i = threadIdx.x; i < fft_tw_total(N); i += C::THREADS)
    tw_s[i] = reinterpret_cast<const f2 *>(P.tw_f)[C::TM ? FFT::tm_tw_src(i) : i];
  const int wid = threadIdx.x / TEAM, tl = threadIdx.x % TEAM;
  const int lane_zero = (int)(threadIdx.x & 31u) * P.zero;  // see sds_engine_f32.cu (int_time load)
  const int sub = tl / TPR, tau = tl % TPR, bar_id = 1 + wid;
  uint32_t xa = sb_a + (uint32_t)(wid * ROWS + sub) * C::XROW;
  uint32_t wb_a = sb_a + C::XREGION + C::TW_BYTES + (uint32_t)wid * C::WORKER_BYTES;
  asm volatile("" : "+r"(xa));
  asm volatile("" : "+r"(wb_a));
  const typename FFT::Addr fa = FFT::make_addr_il(xa, tau);
  unsigned char *wbase = reinterpret_cast<unsigned char *>(__cvta_shared_to_generic(wb_a));
  uint64_t *full = reinterpret_cast<uint64_t *>(wbase);
  unsigned char *stages = wbase + 64;
  const uint32_t full_a = wb_a, stages_a = wb_a + 64;
  // per-worker complex sums over times of (sum_i A2_i) conj(sum_j A1_j): (re, im) per delay
  f2 *pw_s = reinterpret_cast<f2 *>(stages + UV_STAGES * C::SLOT);
  for (int i = tl; i < N; i += TEAM) pw_s[i] = 0ull;
  if (tl == 0) {
    for (int i = 0; i < UV_STAGES; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __shared__ uint32_t tm_base_s;
  if (C::TM && threadIdx.x < 32) tm_alloc<32>(smem_u32(&tm_base_s));
  if (C::TM) tm_fence_before();
  __syncthreads();
  uint32_t tm_a = 0;
  if (C::TM) {
    tm_fence_after();
    tm_a = tm_base_s + (((threadIdx.x >> 5) & 3u) * 32u << 16);
  }
  // the spectrum bins a thread owns: tau_out + TPR r (the tensor-memory exchange renames the lanes)
  const int tau_out = C::TM ? FFT::tm_tau(tau) : tau;

  float tapdf[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) tapdf[r] = __ldg(P.taper_f + tau + TPR * r) * (float)P.delta_f;
  const int total_items = P.item_start[P.ngroup];
  const int64_t row_stride = (int64_t)P.ntime * P.nchan;            // samples between baselines
  const int64_t set_stride = (int64_t)P.npol * P.nbl * row_stride;  // samples between the two data sets
  const bool issuer_warp = TEAM == 32 || __shfl_sync(0xffffffffu, tl, 0) == 0;
  uint32_t slot_c = 0, par_c = 0, slot_p = 0;
  __shared__ int item_bcast[C::NWORKER];
  double *out_all = DATA ? P.pow_all : P.noise_all, *out_auto = DATA ? P.pow_auto : P.noise_auto;

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
        const int nact = ROWS == 1 ? 1 : min(ROWS, ng - pi0);
        const uint32_t bar = full_a + 8u * slot_p;
        mbar_expect_tx_a(bar, (uint32_t)nact * (uint32_t)C::ROWB);
        uint32_t dst = stages_a + slot_p * (uint32_t)C::SLOT;
        int64_t e = pe;
        for (int rr = 0; rr < nact; ++rr, e += row_stride, dst += C::ROWB) {
          if (DATA) {
            tma_load_1d_a(dst, P.vis + e, 8 * N, bar);
            tma_load_1d_a(dst + C::VALB, P.vis + set_stride + e, 8 * N, bar);
          } else {
            tma_load_1d_a(dst, P.nsample + e, 4 * N, bar);
            tma_load_1d_a(dst + C::VALB, P.nsample + set_stride + e, 4 * N, bar);
          }
          tma_load_1d_a(dst + 2 * C::VALB, P.flags + e, N, bar);
          tma_load_1d_a(dst + 2 * C::VALB + N, P.flags + set_stride + e, N, bar);
        }
      }
      slot_p = slot_p + 1 == UV_STAGES ? 0 : slot_p + 1;
      pi0 += ROWS;
      pe += ROWS * row_stride;
      if (pi0 >= ng) {
        pi0 = 0;
        pe_t += P.nchan;
        pe = pe_t;
      }
      --pleft;
    };

    // noise multiplier per channel (both sets share the radiometer tables): sigma coefficient * taper *
    // delta_f (non-finite sigma -> 0, ds.py:3575-3582); taper * delta_f alone for injected noise
    float sigc[8];
    if (!DATA) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (P.noise_mode == 1) {
          const float c = __ldg(P.sigma_coef + ((int64_t)s * P.npol + p) * N + tau + TPR * r);
          sigc[r] = (fabsf(c) <= 3.0e38f) ? c * tapdf[r] : 0.f;
        } else {
          sigc[r] = tapdf[r];
        }
      }
    }
    f2 Sa[8], Sb[8], Cauto[8];  // per time: sums of the two sets' transforms; per item: sum_i A2_i conj(A1_i)
#pragma unroll
    for (int r = 0; r < 8; ++r) Sa[r] = Sb[r] = Cauto[r] = 0ull;
    // thermal sums per channel, packed over channel pairs (tau + TPR 2m, tau + TPR (2m + 1)):
    // A = sum_i a2_i, B = sum_j a1_j / t_j, C = sum_i a2_i a1_i / t_i; dfc[]: what the panel mask removes
    // (per panel q: [0..2] left-end, [3..5] right-end deficits of A, B, C), local memory, rare path
    f2 tA[4], tB[4], tC[4];
    float dfc[48];
    bool dirty_time = false;
    double th_all_acc = 0.0, th_auto_acc = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) tA[k] = tB[k] = tC[k] = 0ull;

    int ci0 = 0, ct = t0;
    int64_t it_off = (int64_t)b0 * P.ntime + t0;
    // moment-matched draw (sds_engine.cuh): one call per block of four baselines and data set
    u32x4 nw0 = {0u, 0u, 0u, 0u}, nw1 = {0u, 0u, 0u, 0u};
    const uint64_t nblk = (uint64_t)((P.nbl_total + 3) >> 2);
    const uint64_t blk_sp0 = (uint64_t)(((int64_t)s * P.npol + p) * 2) * nblk, blk_sp1 = blk_sp0 + nblk;

    for (int k = 0; k < UV_STAGES - 1 && pleft > 0; ++k) issue();
    float tint_next = 1.f;
    if (!DATA) tint_next = __ldg(P.int_time + it_off + (int64_t)((ROWS == 1 || sub < ng) ? sub : 0) * P.ntime);
    for (int step = 0; step < nsteps; ++step) {
      team_sync<TEAM>(bar_id);  // everyone is done with the slot about to be refilled
      if (pleft > 0) issue();
      const int i = ci0 + sub;
      const bool active = ROWS == 1 || i < ng;
      const int ia = active ? i : 0;
      const float tint = tint_next;
      if (!DATA && step + 1 < nsteps) {
        const bool wrap = ci0 + ROWS >= ng;
        const int ni = (wrap ? 0 : ci0 + ROWS) + sub;
        tint_next = __ldg(P.int_time + it_off + (wrap ? 1 : 0) + lane_zero +
                          (int64_t)((ROWS == 1 || ni < ng) ? ni : 0) * P.ntime);
      }
      mbar_wait_a(full_a + 8u * slot_c, par_c);
      const unsigned char *row = stages + slot_c * C::SLOT + sub * C::ROWB;
      const unsigned char *fl0 = row + 2 * C::VALB + tau, *fl1 = fl0 + N;

      f2 va[8], vb[8];  // the two rows in lockstep: set 0, set 1
      if (DATA) {
        const f2 *v0 = reinterpret_cast<const f2 *>(row) + tau;
        const f2 *v1 = reinterpret_cast<const f2 *>(row + C::VALB) + tau;
#pragma unroll
        for (int r = 0; r < 8; ++r) {  // (1 - flag) * taper * delta_f   (ds.py:3503-3505, utils.py:552-560)
          va[r] = f2_scale(active ? v0[TPR * r] : 0ull, (active && fl0[TPR * r] == 0) ? tapdf[r] : 0.f);
          vb[r] = f2_scale(active ? v1[TPR * r] : 0ull, (active && fl1[TPR * r] == 0) ? tapdf[r] : 0.f);
        }
      } else {
        const float *n0 = reinterpret_cast<const float *>(row) + tau;
        const float *n1 = reinterpret_cast<const float *>(row + C::VALB) + tau;
        const float rt = (tint > 0.f) ? fast_rsqrt(tint) : 0.f;
        // a_u = nsample_u^-1/2 (0 where invalid): shared by the noise amplitude and the thermal sums
        float av0[8], av1[8];
        bool all_ok = true;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float s0v = n0[TPR * r], s1v = n1[TPR * r];
          const bool ok0 = active && s0v > 0.f, ok1 = active && s1v > 0.f;
          av0[r] = ok0 ? fast_rsqrt(s0v) : 0.f;
          av1[r] = ok1 ? fast_rsqrt(s1v) : 0.f;
          all_ok = all_ok && ok0 && ok1;
        }
        if (THERMAL) {
          const float itv = rt * rt;
          const f2 it2 = f2_make(itv, itv);
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const f2 a1 = f2_make(av1[2 * m], av1[2 * m + 1]);                 // set 1: baseline i
            const f2 b0 = f2_mul(f2_make(av0[2 * m], av0[2 * m + 1]), it2);    // set 0: baseline j, 1 / t_j
            tA[m] = f2_add(tA[m], a1);
            tB[m] = f2_add(tB[m], b0);
            tC[m] = f2_fma(a1, b0, tC[m]);
          }
          if (team_any<TEAM>(active && !all_ok, bar_id)) {
            if (!dirty_time) {
#pragma unroll 1
              for (int k = 0; k < 48; ++k) dfc[k] = 0.f;
              dirty_time = true;
            }
            if (active) {
#pragma unroll 1
              for (int q = 0; q < 8; ++q) {
                if (tau + TPR * q + 1 >= N) continue;
                const float l0 = n0[TPR * q], r0 = n0[TPR * q + 1], l1 = n1[TPR * q], r1 = n1[TPR * q + 1];
                const bool vl0 = l0 > 0.f, vr0 = r0 > 0.f, vl1 = l1 > 0.f, vr1 = r1 > 0.f;
                float *d = dfc + 6 * q;
                if (vl1 != vr1) d[vl1 ? 0 : 3] += fast_rsqrt(vl1 ? l1 : r1);
                if (vl0 != vr0) d[vl0 ? 1 : 4] += fast_rsqrt(vl0 ? l0 : r0) * itv;
                const bool wl = vl0 && vl1, wr = vr0 && vr1;
                if (wl != wr) d[wl ? 2 : 5] += wl ? fast_rsqrt(l0) * fast_rsqrt(l1) * itv : fast_rsqrt(r0) * fast_rsqrt(r1) * itv;
              }
            }
          }
        }
        if (P.noise_mode == 1) {
          const uint32_t bsel = (uint32_t)(blg + ia) & 3u;
          if (ci0 == 0 || bsel < (uint32_t)ROWS) {
            const uint64_t bq = (uint64_t)(blg + ia) >> 2, tg = (uint64_t)(P.time_offset + ct);
            nw0 = engine_call_mm<ENG_PHILOX_ROUNDS>((blk_sp0 + bq) * (uint64_t)P.ntime_total + tg, tau, P.pkey);
            nw1 = engine_call_mm<ENG_PHILOX_ROUNDS>((blk_sp1 + bq) * (uint64_t)P.ntime_total + tg, tau, P.pkey);
          }
          const uint32_t w0 = engine_pick_mm(nw0, bsel), w1 = engine_pick_mm(nw1, bsel);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const bool k0 = fl0[TPR * r] == 0 && ((w0 >> r) & 1u);
            const bool k1 = fl1[TPR * r] == 0 && ((w1 >> r) & 1u);
            const float a0 = ((k0 ? sigc[r] : 0.f) * rt) * av0[r];
            const float a1 = ((k1 ? sigc[r] : 0.f) * rt) * av1[r];
            va[r] = engine_signs_mm(a0, w0, r);
            vb[r] = engine_signs_mm(a1, w1, r);
          }
        } else {  // injected frequency-domain noise of both sets (parity mode)
          const int64_t e = e00 + (int64_t)(ct - t0) * P.nchan + (int64_t)ia * row_stride;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float2 c0 = active ? P.noise_in[e + tau + TPR * r] : make_float2(0.f, 0.f);
            const float2 c1 = active ? P.noise_in[set_stride + e + tau + TPR * r] : make_float2(0.f, 0.f);
            const float m0 = (active && fl0[TPR * r] == 0) ? sigc[r] : 0.f;
            const float m1 = (active && fl1[TPR * r] == 0) ? sigc[r] : 0.f;
            va[r] = f2_make(c0.x * m0, c0.y * m0);
            vb[r] = f2_make(c1.x * m1, c1.y * m1);
          }
        }
      }
      if (C::TM) FFT::run2_tm(va, vb, fa, tw_s, tau, bar_id, tm_a);
      else FFT::run2_il(va, vb, fa, tw_s, tau, bar_id);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        Sa[r] = f2_add(Sa[r], va[r]);
        Sb[r] = f2_add(Sb[r], vb[r]);
        Cauto[r] = f2_cmac_conj(Cauto[r], vb[r], va[r]);
      }

      slot_c = slot_c + 1 == UV_STAGES ? 0 : slot_c + 1;
      if (slot_c == 0) par_c ^= 1u;
      ci0 += ROWS;
      if (ci0 >= ng) {  // close the time: P += (sum_i A2_i) conj(sum_j A1_j)
        ci0 = 0;
        ++ct;
        ++it_off;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          f2 sa = Sa[r], sb = Sb[r];
#pragma unroll
          for (int off = TPR; off < 32; off <<= 1) {
            sa = f2_add(sa, f2_shfl_xor(sa, off));
            sb = f2_add(sb, f2_shfl_xor(sb, off));
          }
          if (sub == 0) pw_s[tau + TPR * r] = f2_cmac_conj(pw_s[tau + TPR * r], sb, sa);
          Sa[r] = Sb[r] = 0ull;
        }
        if (THERMAL) {
          const double *wlp = P.wl + ((int64_t)s * P.npol + p) * N + tau;
          const double *wrp = P.wr + ((int64_t)s * P.npol + p) * N + tau;
          // the right end of a panel is the next channel: another thread's sums, through this row's
          // exchange buffers ((A, B) pairs in the first, C in the second)
          team_sync<TEAM>(bar_id);
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
          team_sync<TEAM>(bar_id);
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
            f2 ab = f2_make(Ar[r], Ap[r]), bb = f2_make(Br[r], Bp[r]);
#pragma unroll
            for (int off = TPR; off < 32; off <<= 1) {
              ab = f2_add(ab, f2_shfl_xor(ab, off));
              bb = f2_add(bb, f2_shfl_xor(bb, off));
            }
            const double l = __ldg(wlp + TPR * r), rr = __ldg(wrp + TPR * r);
            if (sub == 0)
              th_all_acc += l * (double)(f2_lo(ab) * f2_lo(bb)) + rr * (double)(f2_hi(ab) * f2_hi(bb));
            th_auto_acc += l * (double)Cr[r] + rr * (double)Cp[r];
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) tA[k] = tB[k] = tC[k] = 0ull;
          dirty_time = false;
        }
      }
    }

    // flush the item: fp64 atomics on (re, im), fftshift as an index rotation
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int kshift = (tau_out + TPR * r + N / 2) % N;
      f2 ca = Cauto[r];
#pragma unroll
      for (int off = TPR; off < 32; off <<= 1) ca = f2_add(ca, f2_shfl_xor(ca, off));
      if (sub == 0) {
        const f2 pa = pw_s[tau + TPR * r];
        atomicAdd(out_all + 2 * (obase + kshift), (double)f2_lo(pa));
        atomicAdd(out_all + 2 * (obase + kshift) + 1, (double)f2_hi(pa));
        atomicAdd(out_auto + 2 * (obase + kshift), (double)f2_lo(ca));
        atomicAdd(out_auto + 2 * (obase + kshift) + 1, (double)f2_hi(ca));
        pw_s[tau + TPR * r] = 0ull;
      }
    }
    if (THERMAL) {
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
  if (C::TM) {
    tm_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tm_dealloc<32>(tm_base_s);
  }
}

template <int N, int LEG>
static int launch_uv2(const EngineParams &P, cudaStream_t st, int *launches) {
  using C = UvCfg<N, LEG>;
  static DeviceLatch configured;
  int cfg_dev = 0;
  if (configured.need(&cfg_dev)) {
    SDS_CUDA(cudaFuncSetAttribute(engine_uv2_kernel<N, LEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    SDS_CUDA(cudaFuncSetAttribute(engine_uv2_kernel<N, LEG>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    configured.set(cfg_dev);
  }
  int dev = 0, nsm = 148, per_sm = 1;
  SDS_CUDA(cudaGetDevice(&dev));
  SDS_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  SDS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, engine_uv2_kernel<N, LEG>, C::THREADS, C::SMEM));
  // (the occupancy query answers 1 for a kernel that allocates tensor memory: sds_engine_f32.cu)
  if (C::TM) per_sm = tmem_ctas_per_sm(engine_uv2_kernel<N, LEG>, C::THREADS, C::SMEM, 32);
  if (per_sm < 1) {
    set_error("sds_engine_run: two-set kernel does not fit an SM (smem %d B)", C::SMEM);
    return SDS_ENOTSUP;
  }
  engine_uv2_kernel<N, LEG><<<nsm * per_sm, C::THREADS, C::SMEM, st>>>(P);
  SDS_LAUNCH_CHECK("engine_uv2_kernel");
  ++*launches;
  return SDS_OK;
}

template <int N>
static int run_uv2_n(const EngineParams &P, int leg, cudaStream_t st, int *launches) {
  if (leg == LEG_DATA) return launch_uv2<N, LEG_DATA>(P, st, launches);
  if (leg == LEG_NOISE) return launch_uv2<N, LEG_NOISE>(P, st, launches);
  return launch_uv2<N, LEG_NOISE | LEG_THERMAL>(P, st, launches);
}

int run_engine_uv2(int N, const EngineParams &P, int leg, cudaStream_t st, int *launches) {
  switch (N) {
    case 64: return run_uv2_n<64>(P, leg, st, launches);
    case 128: return run_uv2_n<128>(P, leg, st, launches);
    case 256: return run_uv2_n<256>(P, leg, st, launches);
    case 512: return run_uv2_n<512>(P, leg, st, launches);
    case 1024: return run_uv2_n<1024>(P, leg, st, launches);
    default:
      set_error("sds_engine_run: the two-set fused path takes power-of-two windows of 64..1024 channels, got %d", N);
      return SDS_ENOTSUP;
  }
}

}  // namespace sds
