// Fused engine, complex64 / fp32 arithmetic (the bandwidth path of BASELINE.json).
//
// Same work decomposition as sds_engine.cu (item = group x window x pol x chunk of
// times; rows staged by 1-D bulk TMA through an mbarrier ring; one row team of
// max(32, N/8) threads per worker), re-written around what the profiles of the first
// version showed (profiles/r1_*): the kernel is bound by instruction issue and by
// the FMA pipe, not by DRAM -- so the arithmetic is packed fp32x2 (sds_packed.cuh:
// one FADD2 per complex add, two FFMA2-class instructions per complex multiply,
// "* -i" as an operand modifier) and the accumulators are slimmed to fit three
// CTAs per SM.
#include "sds_engine.cuh"
#include "sds_packed.cuh"

namespace sds {

#ifndef E3_STAGES
#define E3_STAGES 3
#endif
#ifndef E3_MIN_THREADS
#define E3_MIN_THREADS 128
#endif
#ifndef E3_WARPS_PER_SM
#define E3_WARPS_PER_SM 12  // register budget: 65536 / (12 * 32) = 170 per thread
#endif
template <int N> struct E3Cfg {
  static constexpr int TPR = N / 8;                     // threads per row
  static constexpr int TEAM = TPR < 32 ? 32 : TPR;      // threads per worker
  static constexpr int ROWS = TEAM / TPR;               // rows per step
  static constexpr int THREADS = TEAM < E3_MIN_THREADS ? E3_MIN_THREADS : TEAM;
  static constexpr int NWORKER = THREADS / TEAM;
  // resident CTAs per SM the register allocation aims at (12 warps: measured best at N = 256)
  static constexpr int MINB = (E3_WARPS_PER_SM * 32) / THREADS < 1 ? 1 : (E3_WARPS_PER_SM * 32) / THREADS;
  static constexpr int ROWB = 13 * N;                   // vis 8N | nsample 4N | flags N
  static constexpr int SLOT = ROWS * ROWB;
  static constexpr int BUF = N * 8;                     // one exchange buffer of one row
  static constexpr int TW_BYTES = ((fft_tw_total(N) * 8 + 127) / 128) * 128;
  // per worker: 64 B barriers | E3_STAGES slots | ROWS * 2 exchange buffers | 2 N floats (P sums)
  //             | 6 floats per thread (thermal neighbour exchange at the close of a time)
  static constexpr int WORKER_BYTES = 64 + E3_STAGES * SLOT + ROWS * 2 * BUF + 8 * N + 24 * TEAM;
  static constexpr int SMEM = TW_BYTES + NWORKER * WORKER_BYTES;
};

template <int N, int LEGS>
__global__ void __launch_bounds__(E3Cfg<N>::THREADS, E3Cfg<N>::MINB)
engine_f32_kernel(const __grid_constant__ EngineParams P) {
  // LEGS without LEG_DATA (the noise leg alone) serves the complex128 mode, whose data and
  // thermal legs run in fp64 (sds_engine.cu) while the noise realisation stays fp32
  constexpr bool HAS_D = (LEGS & LEG_DATA) != 0, HAS_N = (LEGS & LEG_NOISE) != 0,
                 HAS_T = (LEGS & LEG_THERMAL) != 0;
  using C = E3Cfg<N>;
  constexpr int TPR = C::TPR, TEAM = C::TEAM, ROWS = C::ROWS;
  extern __shared__ __align__(128) unsigned char smem[];

  f2 *tw_s = reinterpret_cast<f2 *>(smem);
  for (int i = threadIdx.x; i < fft_tw_total(N); i += C::THREADS)
    tw_s[i] = reinterpret_cast<const f2 *>(P.tw_f)[i];
  const int wid = threadIdx.x / TEAM, tl = threadIdx.x % TEAM;
  const int sub = tl / TPR, tau = tl % TPR, bar_id = 1 + wid;
  unsigned char *wbase = smem + C::TW_BYTES + (size_t)wid * C::WORKER_BYTES;
  uint64_t *full = reinterpret_cast<uint64_t *>(wbase);
  unsigned char *stages = wbase + 64;
  f2 *bufA = reinterpret_cast<f2 *>(stages + E3_STAGES * C::SLOT + (size_t)(sub * 2) * C::BUF);
  f2 *bufB = bufA + N;
  // per-worker sums over times of |S1|^2: [0, N) data, [N, 2N) noise (touched once per time)
  float *pw_s = reinterpret_cast<float *>(stages + E3_STAGES * C::SLOT + (size_t)(ROWS * 2) * C::BUF);
  float *nb_s = pw_s + 2 * N;  // [sub][run h][tau][3]: first-channel sums of each thermal run
  for (int i = tl; i < 2 * N; i += TEAM) pw_s[i] = 0.f;
  if (tl == 0) {
    for (int i = 0; i < E3_STAGES; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // per-thread constants: taper * delta_f at this thread's 8 transform channels
  float tapdf[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) tapdf[r] = __ldg(P.taper_f + tau + TPR * r) * (float)P.delta_f;
  const int total_items = P.item_start[P.ngroup];
  constexpr uint32_t ROW_BYTES = 13 * N;
  const int64_t row_stride = (int64_t)P.ntime * P.nchan;  // samples between baselines
  // the team's first warp stages its rows (warp-uniform test: shfl from lane 0)
  const bool issuer_warp = TEAM == 32 || __shfl_sync(0xffffffffu, tl, 0) == 0;
  uint32_t slot_c = 0, par_c = 0;  // consumer ring position / phase parity
  uint32_t slot_p = 0;             // producer ring position
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
      item = __shfl_sync(0xffffffffu, item_bcast[wid], 0);  // warp-uniform for the compiler too
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

    // ---- producer state (lane 0 stages rows i0 .. i0+ROWS-1 of time t) ----------
    int pi0 = 0;          // first row of the next step to stage
    int64_t pe = e00;     // its sample offset
    int64_t pe_t = e00;   // offset of row 0 at the producer's current time
    int pleft = nsteps;   // steps not yet staged
    auto issue = [&]() {
      if (issuer_warp && elect_one()) {
        const int nact = min(ROWS, ng - pi0);
        mbar_expect_tx(&full[slot_p], (uint32_t)nact * (HAS_D ? ROW_BYTES : 5u * N));
        unsigned char *dst = stages + slot_p * C::SLOT;
        int64_t e = pe;
        for (int rr = 0; rr < nact; ++rr, e += row_stride, dst += C::ROWB) {
          if (HAS_D) tma_load_1d(dst, P.vis + e, 8 * N, &full[slot_p]);
          tma_load_1d(dst + 8 * N, P.nsample + e, 4 * N, &full[slot_p]);
          tma_load_1d(dst + 12 * N, P.flags + e, N, &full[slot_p]);
        }
      }
      slot_p = slot_p + 1 == E3_STAGES ? 0 : slot_p + 1;
      pi0 += ROWS;
      pe += ROWS * row_stride;
      if (pi0 >= ng) {
        pi0 = 0;
        pe_t += P.nchan;
        pe = pe_t;
      }
      --pleft;
    };

    float sigc[8];
    if (HAS_N) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (P.noise_mode == 1) {
          const float c = __ldg(P.sigma_coef + ((int64_t)s * P.npol + p) * N + tau + TPR * r);
          // non-finite sigma -> 0 (ds.py:3575-3582); taper * delta_f folded in
          sigc[r] = (fabsf(c) <= 3.0e38f) ? c * ENG_SQRT_LN2 * tapdf[r] : 0.f;
        } else {
          sigc[r] = tapdf[r];
        }
      }
    }
    f2 S1[8], Sn1[8];                    // sum over the rows of a time: data and noise
    float S2[8], Sn2[8];                 // sum |x|^2
    // thermal sums per channel, packed over adjacent channels (index 2 h + m: channels
    // c0 + 2 m, c0 + 2 m + 1 of run h): A = sum a, B = sum a / t_int, C = sum a^2 / t_int with
    // a = nsample^-1/2 (0 where invalid).  Rows with a valid sample next to an invalid one
    // ("dirty" panels) additionally book what the panel mask removes into dfc[] (local memory,
    // touched only on that rare path): per panel q, [0..2] left-end and [3..5] right-end deficits.
    f2 tA[4], tB[4], tC[4];
    float dfc[48];
    bool dirty_time = false;
    double th_all_acc = 0.0, th_auto_acc = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      S1[r] = Sn1[r] = 0ull;
      S2[r] = Sn2[r] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) tA[k] = tB[k] = tC[k] = 0ull;

    // ---- consumer state ------------------------------------------------------------
    int ci0 = 0, ct = t0;                                   // first row / time of this step
    int64_t it_off = (int64_t)b0 * P.ntime + t0;            // int_time index of row 0
    // Philox counter base of (row 0, time t0): row_id * N/2 with
    // row_id = ((s npol + p) nbl_total + b) ntime_total + t   (include/sds.h)
    uint64_t rid_t = (((uint64_t)((int64_t)s * P.npol + p) * (uint64_t)P.nbl_total +
                       (uint64_t)(P.bl_offset + b0)) * (uint64_t)P.ntime_total +
                      (uint64_t)(P.time_offset + t0));

    for (int k = 0; k < E3_STAGES - 1 && pleft > 0; ++k) issue();
    for (int step = 0; step < nsteps; ++step) {
      team_sync<TEAM>(bar_id);  // everyone is done with the slot about to be refilled
      if (pleft > 0) issue();
      const int i = ci0 + sub;
      const bool active = ROWS == 1 || i < ng;
      const int ia = active ? i : 0;
      float tint = 1.f;
      if (HAS_N || HAS_T) tint = __ldg(P.int_time + it_off + (int64_t)ia * P.ntime);
      mbar_wait(&full[slot_c], par_c);
      const unsigned char *row = stages + slot_c * C::SLOT + sub * C::ROWB;
      const f2 *vis_s = reinterpret_cast<const f2 *>(row);
      const float *ns_s = reinterpret_cast<const float *>(row + 8 * N);
      const unsigned char *fl_s = row + 12 * N;

      // ---- data leg: (1 - flag) * taper * delta_f ; noise leg: draw * sigma(nsample, t_int) --
      f2 vd[8], vn[8];
      // mask = (1 - flag) (ds.py:3503-3505); taper and delta_f ride on the same multiplier
      // (utils.py:552-560): tapdf for the data, sigc (= sigma coefficient * tapdf) for the noise
      if (HAS_N && P.noise_mode == 1) {
        float rad[8], cs[8], sn[8];
        engine_draw8<TPR>(rid_t + (uint64_t)ia * (uint64_t)P.ntime_total, tau, P.seed_lo, P.seed_hi,
                          rad, cs, sn);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const bool k = active && fl_s[tau + TPR * r] == 0;
          if (HAS_D) vd[r] = f2_scale(active ? vis_s[tau + TPR * r] : 0ull, k ? tapdf[r] : 0.f);
          // sigma sqrt(-ln u1) = coef sqrt(-ln u1) / sqrt(t_int nsample)   (ds.py:3546-3581)
          const float tn = tint * ns_s[tau + TPR * r];
          const float amp = (k && tn > 0.f) ? sigc[r] * rad[r] * fast_rsqrt(tn) : 0.f;
          vn[r] = f2_scale(f2_make(cs[r], sn[r]), amp);
        }
      } else {
        const int64_t e = e00 + (int64_t)(ct - t0) * P.nchan + (int64_t)ia * row_stride;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const bool k = active && fl_s[tau + TPR * r] == 0;
          if (HAS_D) vd[r] = f2_scale(active ? vis_s[tau + TPR * r] : 0ull, k ? tapdf[r] : 0.f);
          if (HAS_N) {  // injected freq-domain noise (parity mode): plain loads
            const float2 c = active ? P.noise_in[e + tau + TPR * r] : make_float2(0.f, 0.f);
            const float m = k ? sigc[r] : 0.f;
            vn[r] = f2_make(c.x * m, c.y * m);
          }
        }
      }
      if (HAS_N) {
        // both transforms in lockstep: two independent dependency chains per team sync
        if (HAS_D) TeamFftP<N>::run2(vd, vn, bufA, bufB, tw_s, tau, bar_id);
        else TeamFftP<N>::run(vn, bufA, bufB, tw_s, tau, bar_id);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          Sn1[r] = f2_add(Sn1[r], vn[r]);
          Sn2[r] = fmaf(f2_lo(vn[r]), f2_lo(vn[r]), fmaf(f2_hi(vn[r]), f2_hi(vn[r]), Sn2[r]));
        }
      } else {
        TeamFftP<N>::run(vd, bufA, bufB, tw_s, tau, bar_id);
      }
      if (HAS_D) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          S1[r] = f2_add(S1[r], vd[r]);
          S2[r] = fmaf(f2_lo(vd[r]), f2_lo(vd[r]), fmaf(f2_hi(vd[r]), f2_hi(vd[r]), S2[r]));
        }
      }

      if (HAS_T) {
        // thermal sums use their own channel map: two runs of 4 consecutive channels per
        // thread (c = 4 tau + 4 TPR h + j) -> two 128-bit loads + the right neighbours
        const float itv = (tint > 0.f) ? 1.f / tint : 0.f;
        const f2 it2 = f2_make(itv, itv);
        bool dirty = false;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c0 = 4 * tau + 4 * TPR * h;
          const float4 nv = *reinterpret_cast<const float4 *>(ns_s + c0);
          const float n4[4] = {nv.x, nv.y, nv.z, nv.w};
          float a4[4];
          bool ok[5];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ok[j] = active && n4[j] > 0.f;
            a4[j] = ok[j] ? fast_rsqrt(n4[j]) : 0.f;
          }
          // right neighbour of the run (no panel right of channel N - 1)
          ok[4] = (c0 + 4 < N) ? (active && ns_s[c0 + 4] > 0.f) : ok[3];
          dirty |= (ok[0] != ok[1]) | (ok[1] != ok[2]) | (ok[2] != ok[3]) | (ok[3] != ok[4]);
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            const f2 aa = f2_make(a4[2 * m], a4[2 * m + 1]);
            tA[2 * h + m] = f2_add(tA[2 * h + m], aa);
            tB[2 * h + m] = f2_fma(aa, it2, tB[2 * h + m]);
            tC[2 * h + m] = f2_fma(f2_mul(aa, aa), it2, tC[2 * h + m]);
          }
        }
        if (dirty) {
          // a panel counts only when both of its ends are valid (masked_invalid, ds.py:3697):
          // book what the full sums above contain but the panel must not
          if (!dirty_time) {
#pragma unroll 1
            for (int k = 0; k < 48; ++k) dfc[k] = 0.f;
            dirty_time = true;
          }
#pragma unroll 1
          for (int q = 0; q < 8; ++q) {
            const int c = 4 * tau + 4 * TPR * (q >> 2) + (q & 3);
            if (c + 1 >= N) continue;
            const float nl = ns_s[c], nr = ns_s[c + 1];
            if ((nl > 0.f) == (nr > 0.f)) continue;
            const float av = fast_rsqrt(nl > 0.f ? nl : nr);
            float *d = dfc + 6 * q + (nl > 0.f ? 0 : 3);
            d[0] += av;
            d[1] += av * itv;
            d[2] += av * av * itv;
          }
        }
      }

      // ---- advance the consumer; close the time when its last row is done -----------
      slot_c = slot_c + 1 == E3_STAGES ? 0 : slot_c + 1;
      if (slot_c == 0) par_c ^= 1u;
      ci0 += ROWS;
      if (ci0 >= ng) {
        ci0 = 0;
        ++ct;
        ++it_off;
        ++rid_t;
#pragma unroll
        for (int r = 0; r < 8; ++r) {  // P += |sum over the group's rows|^2
          f2 sd = S1[r], sn = Sn1[r];
#pragma unroll
          for (int off = TPR; off < 32; off <<= 1) {
            sd = f2_add(sd, f2_shfl_xor(sd, off));
            if (HAS_N) sn = f2_add(sn, f2_shfl_xor(sn, off));
          }
          if (HAS_D && sub == 0) pw_s[tau + TPR * r] += f2_hsum(f2_mul(sd, sd));
          S1[r] = 0ull;
          if (HAS_N) {
            if (sub == 0) pw_s[N + tau + TPR * r] += f2_hsum(f2_mul(sn, sn));
            Sn1[r] = 0ull;
          }
        }
        if (HAS_T) {
          const double *wlp = P.wl + ((int64_t)s * P.npol + p) * N;
          const double *wrp = P.wr + ((int64_t)s * P.npol + p) * N;
          // the right end of a run's last panel is the first channel of the next thread's run
          float *mine = nb_s + ((sub * 2) * TPR + tau) * 3;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mine[h * TPR * 3 + 0] = f2_lo(tA[2 * h]);
            mine[h * TPR * 3 + 1] = f2_lo(tB[2 * h]);
            mine[h * TPR * 3 + 2] = f2_lo(tC[2 * h]);
          }
          team_sync<TEAM>(bar_id);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float nxt[3] = {0.f, 0.f, 0.f};
            if (tau + 1 < TPR || h == 0) {
              const float *src = (tau + 1 < TPR) ? mine + h * TPR * 3 + 3
                                                 : nb_s + ((sub * 2 + 1) * TPR) * 3;
              nxt[0] = src[0]; nxt[1] = src[1]; nxt[2] = src[2];
            }
            const float Ac[5] = {f2_lo(tA[2 * h]), f2_hi(tA[2 * h]), f2_lo(tA[2 * h + 1]),
                                 f2_hi(tA[2 * h + 1]), nxt[0]};
            const float Bc[5] = {f2_lo(tB[2 * h]), f2_hi(tB[2 * h]), f2_lo(tB[2 * h + 1]),
                                 f2_hi(tB[2 * h + 1]), nxt[1]};
            const float Cc[5] = {f2_lo(tC[2 * h]), f2_hi(tC[2 * h]), f2_lo(tC[2 * h + 1]),
                                 f2_hi(tC[2 * h + 1]), nxt[2]};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = 4 * h + j;
              float a = Ac[j], ap = Ac[j + 1], b = Bc[j], bp = Bc[j + 1], cc = Cc[j], cp = Cc[j + 1];
              if (dirty_time) {
                a -= dfc[6 * q + 0]; b -= dfc[6 * q + 1]; cc -= dfc[6 * q + 2];
                ap -= dfc[6 * q + 3]; bp -= dfc[6 * q + 4]; cp -= dfc[6 * q + 5];
              }
              f2 ab = f2_make(a, ap), bb = f2_make(b, bp);
#pragma unroll
              for (int off = TPR; off < 32; off <<= 1) {
                ab = f2_add(ab, f2_shfl_xor(ab, off));
                bb = f2_add(bb, f2_shfl_xor(bb, off));
              }
              const int c = 4 * tau + 4 * TPR * h + j;
              const double l = __ldg(wlp + c), rr = __ldg(wrp + c);
              if (sub == 0)
                th_all_acc += l * (double)f2_lo(ab) * (double)f2_lo(bb) +
                              rr * (double)f2_hi(ab) * (double)f2_hi(bb);
              th_auto_acc += l * (double)cc + rr * (double)cp;
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) tA[k] = tB[k] = tC[k] = 0ull;
          dirty_time = false;
        }
      }
    }

    // ---- flush the item: fp64 atomics, fftshift as an index rotation -------------
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int kshift = (tau + TPR * r + N / 2) % N;
      float s2 = S2[r], sn2 = Sn2[r];
#pragma unroll
      for (int off = TPR; off < 32; off <<= 1) {
        s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        if (HAS_N) sn2 += __shfl_xor_sync(0xffffffffu, sn2, off);
      }
      if (sub == 0) {
        if (HAS_D) {
          atomicAdd(P.pow_all + obase + kshift, (double)pw_s[tau + TPR * r]);
          atomicAdd(P.pow_auto + obase + kshift, (double)s2);
          pw_s[tau + TPR * r] = 0.f;
        }
        if (HAS_N) {
          atomicAdd(P.noise_all + obase + kshift, (double)pw_s[N + tau + TPR * r]);
          atomicAdd(P.noise_auto + obase + kshift, (double)sn2);
          pw_s[N + tau + TPR * r] = 0.f;
        }
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
        const int64_t o = ((int64_t)g * P.nspw + s) * P.npol + p;
        atomicAdd(P.th_all + o, f * th_all_acc);
        atomicAdd(P.th_auto + o, f * th_auto_acc);
      }
    }
  }
}

template <int N, int LEGS>
static int launch_f32(const EngineParams &P, cudaStream_t st, int *launches) {
  using C = E3Cfg<N>;
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
    case 2: return launch_f32<N, 2>(P, st, launches);
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
