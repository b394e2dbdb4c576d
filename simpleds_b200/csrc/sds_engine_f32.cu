// Fused engine, complex64 / fp32 arithmetic (the bandwidth path of BASELINE.json).
//
// Same work decomposition as sds_engine.cu (item = group x window x pol x chunk of
// times; rows staged by 1-D bulk TMA through an mbarrier ring; one row team of
// max(32, N/8) threads per worker), re-written around what the profiles of the first
// version showed (profiles/r1_*): the kernel is bound by instruction issue and by
// the FMA pipe, not by DRAM -- so the arithmetic is packed fp32x2 (sds_packed.cuh:
// one FADD2 per complex add, two FFMA2-class instructions per complex multiply,
// "* -i" as an operand modifier), the accumulators are slimmed to fit three CTAs per SM,
// the exchange addresses cost one LOP3 per access (TeamFftX), and whatever the noise and thermal
// legs both need (nsample^-1/2, the flag select) is computed once.  DESIGN.md sections 3.1 / 4.
#include "sds_engine.cuh"
#include "sds_packed.cuh"

namespace sds {

#ifndef E3_STAGES
#define E3_STAGES 3
#endif
#ifndef E3_MIN_THREADS
#define E3_MIN_THREADS 128
#endif
#ifndef E3_IL
#define E3_IL 1  // interleaved 128-bit exchange of the two lockstep transforms
#endif
#ifndef E3_TM
#define E3_TM 1  // last exchange of the lockstep transforms through tensor memory (N = 256)
#endif
#ifndef E3_TACC
#define E3_TACC 1  // accumulators in tensor memory: 128 registers, 16 warps per SM (N = 256, three legs)
#endif
#ifndef E3_F4
#define E3_F4 1  // N = 1024, 2048: one team-wide exchange, then one 256-point transform per warp (TeamFftRx256);
                 // at N = 512 (two-warp teams) the generic schedule measured 3 % faster (0.486 vs 0.473)
#endif
#ifndef E3_KEEP
#define E3_KEEP 1
#endif
#ifndef E3_TV
#define E3_TV 1
#endif
#ifndef E3_WARPS_PER_SM
#define E3_WARPS_PER_SM 12  // register budget: 65536 / (12 * 32) = 170 per thread
#endif
// GAUSS: the in-kernel draw is Box-Muller per channel (noise_mode 3) instead of the moment-matched
// discrete law (noise_mode 1); injected noise (mode 2) runs in the GAUSS = false instantiation
template <int N, int LEGS, bool GAUSS = false> struct E3Cfg {
  using F4T = TeamFftRx256<(N >= 512 ? N / 256 : 8)>;  // the four-step transform of this length (N >= 512)
  // the noise leg alone (complex128 mode) keeps neither visibilities nor the data accumulators:
  // 122 registers and 5 N staged bytes per row let 16 warps per SM run it
  static constexpr bool NOISE_ONLY = LEGS == LEG_NOISE;
  static constexpr int TPR = N / 8;                     // threads per row
  static constexpr int TEAM = TPR < 32 ? 32 : TPR;      // threads per worker
  static constexpr int ROWS = TEAM / TPR;               // rows per step
  static constexpr int THREADS = TEAM < E3_MIN_THREADS ? E3_MIN_THREADS : TEAM;
  static constexpr int NWORKER = THREADS / TEAM;
  // resident CTAs per SM the register allocation aims at (12 warps: measured best at N = 256)
  // (N = 2048 runs 256-thread CTAs: a second one would halve their registers -- measured slower)
  // TACC: every per-item accumulator (pair sums, |x|^2 sums, thermal sums, the per-time |S1|^2 sums) lives
  // in thread-private tensor-memory columns and is read-modified-written once per row with one
  // instruction per 16 / 32 words: 72 registers fewer, 128 registers per thread, four CTAs per SM
  static constexpr bool TACC = E3_TACC && E3_TM && E3_IL && (LEGS & ~LEG_THERMAL) == (LEG_DATA | LEG_NOISE) &&
                               (N == 256 || (E3_F4 && N >= 1024));
  // TM1: the instantiations with ONE transform per row (data without noise; the noise + fp64-thermal leg
  // of the complex128 mode) make their last exchange through tensor memory as well; their accumulators
  // stay in registers (in tensor memory they measured
  // slower: 7.44 vs 5.96 ms, 80 words of read-modify-write per row and spills at 128 registers)
  static constexpr bool SINGLE = !((LEGS & LEG_DATA) && (LEGS & LEG_NOISE)) && (LEGS & (LEG_DATA | LEG_NOISE));
  static constexpr bool F4S = E3_F4 && E3_TM && SINGLE && N >= 1024;  // four-step, one transform per row
  static constexpr bool TM1 = E3_TM && SINGLE && (TeamFftX<N>::TM_OK || F4S);
  static constexpr int STAGES = TACC ? 2 : E3_STAGES;
  // tensor-memory columns of one warp (exchange 32 + accumulators 88), and of the CTA: warps w and w + 4
  // share TMEM lanes, so every set of four warps has its own column range
  static constexpr int TM_WCOLS = TACC ? 128 : 32;
  static constexpr int TM_COLS = TM_WCOLS * (THREADS / 128);
  static constexpr int WARPS_PER_SM = (TACC || (NOISE_ONLY && THREADS <= 128)) ? 16 : E3_WARPS_PER_SM;
  static constexpr int MINB = (WARPS_PER_SM * 32) / THREADS < 1 ? 1 : (WARPS_PER_SM * 32) / THREADS;
  static constexpr int VISB = (LEGS & LEG_DATA) ? 8 * N : 0;
  static constexpr int ROWB = VISB + 5 * N;             // vis 8N (data leg only) | nsample 4N | flags N
  static constexpr int SLOT = ROWS * ROWB;
  static constexpr int BUF = N * 8;                     // one exchange buffer of one row
  // data and noise transforms in lockstep exchange through ONE padded buffer of 16-byte (data, noise)
  // slots with 128-bit accesses (TeamFftX::run2_il): 18 N bytes per row, no alignment requirement;
  // otherwise a pair of XOR-addressed buffers of 8 N bytes, aligned to 8 N
  static constexpr bool IL = E3_IL && (LEGS & LEG_DATA) && (LEGS & LEG_NOISE);
  static constexpr bool F4 = E3_F4 && E3_TM && IL && N >= 1024;
  static constexpr bool TM = E3_TM && IL && (TeamFftX<N>::TM_OK || F4);
  static constexpr int XROW = IL ? TeamFftX<N>::IL_BYTES : 2 * BUF;  // exchange bytes of one row
  static constexpr int XALIGN = IL ? 128 : BUF;
  static constexpr int XREGION = NWORKER * ROWS * XROW;  // all exchange buffers
  static constexpr int TW_ENTRIES = (E3_F4 && E3_TM && N >= 1024 && (SINGLE || (E3_IL && (LEGS & LEG_DATA) && (LEGS & LEG_NOISE))))
                                        ? F4T::TW_TOTAL : fft_tw_total(N);
  static constexpr int TW_BYTES = ((TW_ENTRIES * 8 + 127) / 128) * 128;
  // per worker: 64 B barriers | 128 B integration times of 32 steps | E3_STAGES slots | 2 N floats (P sums)
  static constexpr int HEAD = 448;  // 64 B barriers | 128 B t_int^-1/2 of 32 steps | 256 B 1 / t_int in fp64 (T64)
  static constexpr int WORKER_BYTES = HEAD + STAGES * SLOT + (TACC ? 0 : 8 * N);
  // XALIGN bytes of slack in front: the exchange region is aligned at run time
  static constexpr int SMEM = XALIGN + XREGION + TW_BYTES + NWORKER * WORKER_BYTES;
};

template <int N, int LEGS, bool GAUSS>
__global__ void __launch_bounds__(E3Cfg<N, LEGS, GAUSS>::THREADS, E3Cfg<N, LEGS, GAUSS>::MINB)
engine_f32_kernel(const __grid_constant__ EngineParams P) {
  // LEGS without LEG_DATA (the noise leg alone) serves the complex128 mode, whose data and
  // thermal legs run in fp64 (sds_engine.cu) while the noise realisation stays fp32
  // LEG_THERMAL64: the thermal sums in fp64 (complex128 mode, where the thermal estimate must hold
  // 1e-10: its sums ride in the noise launch, whose fp64 pipe is idle and which stages nsample anyway)
  constexpr bool HAS_D = (LEGS & LEG_DATA) != 0, HAS_N = (LEGS & LEG_NOISE) != 0,
                 HAS_T = (LEGS & LEG_THERMAL) != 0, T64 = (LEGS & LEG_THERMAL64) != 0;
  static_assert(!(HAS_T && T64), "thermal sums are fp32 or fp64, not both");
  using C = E3Cfg<N, LEGS, GAUSS>;
  using FFT = TeamFftX<N>;
  using F4T = typename C::F4T;
  constexpr int TPR = C::TPR, TEAM = C::TEAM, ROWS = C::ROWS;
  constexpr bool TV = E3_TV && ROWS == 1 && (HAS_N || HAS_T || T64);  // integration times as a lane vector
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // exchange region first, aligned to the size of one buffer (TeamFftX addressing)
  const uint32_t s0 = smem_u32(smem_raw);
  unsigned char *smem = smem_raw + (((s0 + C::XALIGN - 1u) & ~(uint32_t)(C::XALIGN - 1)) - s0);

  const int wid = threadIdx.x / TEAM, tl = threadIdx.x % TEAM;
  // 0 at run time (P.zero is 0), lane-dependent as far as ptxas can tell: see the int_time load
  const int lane_zero = (int)(threadIdx.x & 31u) * P.zero;
  const int sub = tl / TPR, tau = tl % TPR, bar_id = 1 + wid;
  // Shared-space addresses, each passed through an empty asm: ptxas otherwise re-derives them from
  // the CTA's shared window (SR_CgaCtaId, the run-time alignment, the warp index) at every use inside
  // the step loop -- some forty instructions per row
  uint32_t sb_a = smem_u32(smem);
#if E3_KEEP
  asm volatile("" : "+r"(sb_a));
#endif
  f2 *tw_s = reinterpret_cast<f2 *>(__cvta_shared_to_generic(sb_a + C::XREGION));
  if constexpr (C::F4 || C::F4S) {
    F4T::fill_tables(tw_s, threadIdx.x, C::THREADS);
  } else {
    for (int i = threadIdx.x; i < fft_tw_total(N); i += C::THREADS)
      tw_s[i] = reinterpret_cast<const f2 *>(P.tw_f)[(C::TM || C::TM1) ? FFT::tm_tw_src(i) : i];  // (not F4 / F4S: above)
  }
  // this row's exchange buffers: [xa, xa + BUF) and [xa + BUF, xa + 2 BUF)
  uint32_t xa = sb_a + (uint32_t)(wid * ROWS + sub) * C::XROW;
  uint32_t wb_a = sb_a + C::XREGION + C::TW_BYTES + (uint32_t)wid * C::WORKER_BYTES;
#if E3_KEEP
  asm volatile("" : "+r"(xa));
  asm volatile("" : "+r"(wb_a));
#endif
  const typename FFT::Addr fa = C::IL ? FFT::make_addr_il(xa, tau) : (C::TM1 && !C::F4S) ? FFT::make_addr_p1(xa, tau) : FFT::make_addr(xa, tau);
  const typename TeamFftX<256>::Addr fa4s =
      TeamFftX<256>::make_addr_p1(xa + (uint32_t)(tau >> 5) * F4T::REGION1, tau & 31);
  // four-step: the addresses of the warp-local 256-point transform in the warp's region of the row's buffer
  const typename TeamFftX<256>::Addr fa4 =
      TeamFftX<256>::make_addr_il(xa + (uint32_t)(tau >> 5) * F4T::REGION, tau & 31);
  unsigned char *wbase = reinterpret_cast<unsigned char *>(__cvta_shared_to_generic(wb_a));
  uint64_t *full = reinterpret_cast<uint64_t *>(wbase);
  unsigned char *stages = wbase + C::HEAD;
  const uint32_t tv_a = wb_a + 64;  // [32] floats t_int^-1/2 of the next steps, then [32] doubles 1 / t_int (T64)
  const uint32_t full_a = wb_a, stages_a = wb_a + C::HEAD;  // shared-space addresses
  const uint32_t free_a = wb_a + 32;  // four-step: the buffer-release barrier
  uint32_t par_f = 1u;                // its phase parity (a fresh barrier passes a wait on parity 1)
  const uint32_t stored_a = wb_a + 40, flag_a = wb_a + 48;
  uint32_t par_s = 0u, seq_f4 = 0u;
  // per-worker sums over times of |S1|^2: [0, N) data, [N, 2N) noise (touched once per time)
  float *pw_s = reinterpret_cast<float *>(stages + C::STAGES * C::SLOT);
  if (!C::TACC)
    for (int i = tl; i < 2 * N; i += TEAM) pw_s[i] = 0.f;
  if (tl == 0) {
    for (int i = 0; i < C::STAGES; ++i) mbar_init(&full[i], 1);
    if (C::F4) {
      mbar_init(&full[4], TEAM / 32);  // four-step: "the team's exchange buffer is free" (one arrival per warp)
      mbar_init(&full[5], TEAM / 32);  // four-step: "the row's transposed stores are done"
      full[6] = 0;                     // the two flag words of the invalid-sample vote
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // tensor memory for the last exchange: 32 columns per CTA, each warp its own 32 lanes
  __shared__ uint32_t tm_base_s;
  constexpr bool TMANY = C::TM || C::TM1;
  if (TMANY && threadIdx.x < 32) tm_alloc<C::TM_COLS>(smem_u32(&tm_base_s));
  if (TMANY) tm_fence_before();
  __syncthreads();
  uint32_t tm_a = 0;
  if (TMANY) {
    tm_fence_after();
    const uint32_t warp = threadIdx.x >> 5;
    tm_a = tm_base_s + (((warp & 3u) * 32u) << 16) + (warp >> 2) * (uint32_t)C::TM_WCOLS;
    tm_a = __shfl_sync(0xffffffffu, tm_a, 0);  // warp-uniform for the compiler too: lives in a uniform register
  }
  // tensor-memory columns of the accumulators (TACC), after the 32 columns of the exchange:
  // [32, 64) S1, Sn1 | [64, 80) S2, Sn2 | [80, 104) tA, tB, tC | [104, 120) per-time |S1|^2 sums
  constexpr uint32_t TC_S1 = 32, TC_S2 = 64, TC_T = 80, TC_PW = 104;
  if (C::TACC) {
    const uint32_t z16[16] = {};
    tm_st_w16(tm_a + TC_PW, z16);
  }
  // the spectrum bins a thread owns: tau_out + TPR r (the tensor-memory exchange renames the lanes)
  const int tau_out = (C::F4 || C::F4S) ? (tau >> 5) + (N / 256) * TeamFftX<256>::tm_tau(tau & 31) : (C::TM || C::TM1) ? FFT::tm_tau(tau) : tau;

  // per-thread constants: taper * delta_f at this thread's 8 transform channels
  float tapdf[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) tapdf[r] = __ldg(P.taper_f + tau + TPR * r) * (float)P.delta_f;
  const int total_items = P.item_start[P.ngroup];
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
    // sample offset of (pol p, first baseline of the group, time t0, window start)
    const int64_t e00 = (((int64_t)p * P.nbl + b0) * P.ntime + t0) * P.nchan + P.win_start[s];
    const int64_t obase = (((int64_t)gout * P.nspw + s) * P.npol + p) * N;

    // ---- producer state (one elected lane stages rows i0 .. i0+ROWS-1 of time t) ----------
    int pi0 = 0;          // first row of the next step to stage
    int64_t pe = e00;     // its sample offset
    int64_t pe_t = e00;   // offset of row 0 at the producer's current time
    int pleft = nsteps;   // steps not yet staged
    auto issue = [&]() {
      if (issuer_warp && elect_one()) {
        const int nact = ROWS == 1 ? 1 : min(ROWS, ng - pi0);
        const uint32_t bar = full_a + 8u * slot_p;
        mbar_expect_tx_a(bar, (uint32_t)nact * (uint32_t)C::ROWB);
        uint32_t dst = stages_a + slot_p * (uint32_t)C::SLOT;
        int64_t e = pe;
        for (int rr = 0; rr < nact; ++rr, e += row_stride, dst += C::ROWB) {
          if (HAS_D) tma_load_1d_a(dst, P.vis + e, 8 * N, bar);
          tma_load_1d_a(dst + C::VISB, P.nsample + e, 4 * N, bar);
          tma_load_1d_a(dst + C::VISB + 4 * N, P.flags + e, N, bar);
        }
      }
      slot_p = slot_p + 1 == C::STAGES ? 0 : slot_p + 1;
      pi0 += ROWS;
      pe += ROWS * row_stride;
      if (pi0 >= ng) {
        pi0 = 0;
        pe_t += P.nchan;
        pe = pe_t;
      }
      --pleft;
    };

    // noise multiplier per channel: sigma coefficient * sqrt(ln 2) * taper * delta_f (non-finite
    // sigma -> 0, ds.py:3575-3582); taper * delta_f alone for injected noise
    float sigc[8];
    if (HAS_N) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (P.noise_mode == 1) {
          const float c = __ldg(P.sigma_coef + ((int64_t)s * P.npol + p) * N + tau + TPR * r);
          sigc[r] = (fabsf(c) <= 3.0e38f) ? c * (GAUSS ? ENG_SQRT_LN2 : 1.f) * tapdf[r] : 0.f;
        } else {
          sigc[r] = tapdf[r];
        }
      }
    }
    f2 S1[8], Sn1[8];                    // sum over the rows of a time: data and noise
    float S2[8], Sn2[8];                 // sum |x|^2
    // thermal sums per channel, packed over the thread's channel pairs (tau + TPR 2m, tau + TPR (2m+1)):
    // A = sum a, B = sum a / t_int, C = sum a^2 / t_int with a = nsample^-1/2 (0 where invalid).
    // Rows with an invalid sample additionally book what the panel mask removes into dfc[]
    // (local memory, touched only on that rare path): per panel q = channel tau + TPR q and its
    // right neighbour, [0..2] left-end and [3..5] right-end deficits.
    f2 tA[4], tB[4], tC[4];
    float dfc[48];
    double dA[8], dB[8], dC[8];  // T64: the same sums per channel tau + TPR r in fp64
    double ddfc[T64 ? 48 : 1];
    bool dirty_time = false, dirty_row = false;
    double th_all_acc = 0.0, th_auto_acc = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      S1[r] = Sn1[r] = 0ull;
      S2[r] = Sn2[r] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) tA[k] = tB[k] = tC[k] = 0ull;
#pragma unroll
    for (int r = 0; r < 8; ++r) dA[r] = dB[r] = dC[r] = 0.0;
    if constexpr (C::TACC) {  // the item's accumulators start at zero in tensor memory
      const uint32_t z32[32] = {};
      const uint32_t z16[16] = {};
      const uint32_t z8[8] = {};
      tm_st_w32(tm_a + TC_S1, z32);
      tm_st_w16(tm_a + TC_S2, z16);
      tm_st_w16(tm_a + TC_T, z16);
      tm_st_w8(tm_a + TC_T + 16, z8);
      tm_wait_st();
    }

    // ---- consumer state ------------------------------------------------------------
    int ci0 = 0, ct = t0;                                   // first row / time of this step
    int64_t it_off = (int64_t)b0 * P.ntime + t0;            // int_time index of row 0
    const int64_t it0 = it_off;
    // Philox counter words 0, 1 of (row 0, time t0):
    // row_id = ((s npol + p) nbl_total + b) ntime_total + t   (include/sds.h)
    uint64_t rid_t = (((uint64_t)((int64_t)s * P.npol + p) * (uint64_t)P.nbl_total +
                       (uint64_t)blg) * (uint64_t)P.ntime_total +
                      (uint64_t)(P.time_offset + t0));

    // moment-matched draw: the Philox call of the current block of four baselines
    u32x4 nwords = {0u, 0u, 0u, 0u};
    const uint64_t nblk = (uint64_t)((P.nbl_total + 3) >> 2);
    const uint64_t blk_sp = (uint64_t)((int64_t)s * P.npol + p) * nblk;
    // (four-step: a row's slot is refilled from the MIDDLE of the step that consumed it, so every slot is in flight)
    for (int k = 0; k < C::STAGES - (C::F4 ? 0 : 1) && pleft > 0; ++k) issue();
    // integration time of the step's row, loaded one step ahead (an L2 round trip otherwise sits
    // on the critical path of every row)
    float tint_next = 1.f;
    if (!TV && (HAS_N || HAS_T || T64))
      tint_next = __ldg(P.int_time + it_off + (int64_t)((ROWS == 1 || sub < ng) ? sub : 0) * P.ntime);
    // one row per step: entry (f & 31) of a small shared-memory ring holds the integration time of
    // step f of the item (t_int^-1/2; the fp64 thermal sums also keep 1 / t_int in fp64 there); the half of
    // the ring that is not being consumed is refilled every 16 steps, 16 to 31 steps ahead of its use,
    // and a row takes its value with one broadcast load instead of an address computation and an L2 load
    auto tv_fill = [&](int f) {
      float v = 0.f;
      if (f < nsteps) {
        // f / ng without a hoisted reciprocal (registers): approximate quotient, exact fix-up
        float rcp;
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"((float)ng));
        int tt = (int)((float)f * rcp), ii = f - tt * ng;
        if (ii < 0) { ii += ng; --tt; }
        if (ii >= ng) { ii -= ng; ++tt; }
        v = __ldg(P.int_time + it0 + (int64_t)ii * P.ntime + tt);
        if (T64) {
          const double id = v > 0.f ? 1.0 / (double)v : 0.0;
          asm volatile("st.shared.f64 [%0], %1;" ::"r"(tv_a + 128u + 8u * (uint32_t)(f & 31)), "d"(id) : "memory");
        }
        v = v > 0.f ? fast_rsqrt(v) : 0.f;
      } else if (T64) {
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(tv_a + 128u + 8u * (uint32_t)(f & 31)), "d"(0.0) : "memory");
      }
      sts_f32(tv_a + 4u * (uint32_t)(f & 31), v);
    };
    if (TV) {
      team_sync<TEAM>(bar_id);  // the previous item's last reads of the ring are done
      if ((threadIdx.x & 16u) == 0) tv_fill((int)(threadIdx.x & 15u));
      if (C::F4) team_sync<TEAM>(bar_id);  // (the other schedules sync at the top of every step)
    }
    for (int step = 0; step < nsteps; ++step) {
      if (!C::F4) {
        team_sync<TEAM>(bar_id);  // everyone is done with the slot about to be refilled
        if (pleft > 0) issue();
      }
      const int i = ci0 + sub;
      const bool active = ROWS == 1 || i < ng;
      const int ia = active ? i : 0;
      float tint = tint_next;
      double itd_ring = 0.0;
      if constexpr (TV) {
        // (every warp of a team writes the same values; the step's team sync orders them before the reads)
        if ((step & 15) == 0 && (((int)threadIdx.x ^ step) & 16) != 0)
          tv_fill(step + 16 + (int)(threadIdx.x & 15u));
        tint = lds_f32(tv_a + 4u * (uint32_t)(step & 31));
        if (T64) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(itd_ring) : "r"(tv_a + 128u + 8u * (uint32_t)(step & 31)) : "memory");
      }
      if (!TV && (HAS_N || HAS_T || T64) && step + 1 < nsteps) {
        const bool wrap = ci0 + ROWS >= ng;
        const int ni = (wrap ? 0 : ci0 + ROWS) + sub;
        // lane_zero (= 0 at run time, but lane-dependent as far as ptxas can tell) keeps the address, and with it
        // the loaded value, out of the uniform datapath: for warp-sized teams the value is
        // warp-uniform, and a uniform register is waited on right behind the load (R2UR) instead of
        // one step later (10 % of the issue stalls, profiles/r2_engine_mm3)
        tint_next = __ldg(P.int_time + it_off + (wrap ? 1 : 0) + lane_zero +
                          (int64_t)((ROWS == 1 || ni < ng) ? ni : 0) * P.ntime);
      }
      mbar_wait_a(full_a + 8u * slot_c, par_c);
      const unsigned char *row = stages + slot_c * C::SLOT + sub * C::ROWB;
      const f2 *vis_s = reinterpret_cast<const f2 *>(row) + tau;
      const float *ns_s = reinterpret_cast<const float *>(row + C::VISB) + tau;
      const unsigned char *fl_s = row + C::VISB + 4 * N + tau;

      // rt = t_int^-1/2, itv = 1 / t_int (0 where the integration time is not positive)
      const float rt = TV ? tint : (tint > 0.f) ? fast_rsqrt(tint) : 0.f;
      const float itv = rt * rt;
      // a = nsample^-1/2 (0 where invalid): shared by the noise amplitude and the thermal sums
      float av[8];
      bool all_ok = true;
      if (HAS_N || HAS_T || T64) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float nv = ns_s[TPR * r];
          const bool ok = active && nv > 0.f;
          av[r] = ok ? fast_rsqrt(nv) : 0.f;
        }
        // every sample valid <=> the smallest nsample^-1/2 is positive (three-input minima: four instructions)
        const float m0 = fminf(fminf(av[0], av[1]), av[2]), m1 = fminf(fminf(av[3], av[4]), av[5]);
        all_ok = fminf(fminf(m0, m1), fminf(av[6], av[7])) > 0.f;
      }

      // ---- data leg: (1 - flag) * taper * delta_f ; noise leg: draw * sigma(nsample, t_int) --
      f2 vd[8], vn[8];
      // mask = (1 - flag) (ds.py:3503-3505); taper and delta_f ride on the same multiplier
      // (utils.py:552-560): tapdf for the data, sigc (= sigma coefficient * tapdf) for the noise
      if (HAS_N && !GAUSS && P.noise_mode == 1) {
        // a new block of four baselines starts at the first step of a time and whenever the
        // (global) baseline index crosses a multiple of four; uniform over the row team
        const uint32_t bsel = (uint32_t)(blg + ia) & 3u;
        if (ci0 == 0 || bsel < (uint32_t)ROWS) {
          const uint64_t bglob = (uint64_t)(blg + ia);
          const uint64_t blk = (blk_sp + (bglob >> 2)) * (uint64_t)P.ntime_total + (uint64_t)(P.time_offset + ct);
          nwords = engine_call_mm<ENG_PHILOX_ROUNDS>(blk, tau, P.pkey);
        }
        const uint32_t word = engine_pick_mm(nwords, bsel);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          // sigma = coef / sqrt(t_int nsample) (ds.py:3546-3581); the on/off bit rides on the flag select
          float dsel, nsel;
          engine_selects_mm(active ? (uint32_t)fl_s[TPR * r] : 1u, word, 1u << r, tapdf[r], sigc[r], dsel, nsel);
          if (HAS_D) vd[r] = f2_scale(active ? vis_s[TPR * r] : 0ull, dsel);
          const float amp = (nsel * rt) * av[r];
          vn[r] = engine_signs_mm(amp, word, r);
        }
      } else if (HAS_N && GAUSS && P.noise_mode == 1) {
        float rad[8], cs[8], sn[8];
        engine_draw8<TPR>(rid_t + (uint64_t)ia * (uint64_t)P.ntime_total, tau, P.pkey, rad, cs, sn);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const bool k = active && fl_s[TPR * r] == 0;
          if (HAS_D) vd[r] = f2_scale(active ? vis_s[TPR * r] : 0ull, k ? tapdf[r] : 0.f);
          // sigma sqrt(-ln u1) = coef sqrt(-ln u1) / sqrt(t_int nsample)   (ds.py:3546-3581); the two
          // selects run on the ALU pipe, the three products on the (binding) FP32 pipe
          const float amp = ((k ? sigc[r] : 0.f) * rt) * (rad[r] * av[r]);
          vn[r] = f2_scale(f2_make(cs[r], sn[r]), amp);
        }
      } else {
        const int64_t e = e00 + (int64_t)(ct - t0) * P.nchan + (int64_t)ia * row_stride;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const bool k = active && fl_s[TPR * r] == 0;
          if (HAS_D) vd[r] = f2_scale(active ? vis_s[TPR * r] : 0ull, k ? tapdf[r] : 0.f);
          if (HAS_N) {  // injected freq-domain noise (parity mode): plain loads
            const float2 c = active ? P.noise_in[e + tau + TPR * r] : make_float2(0.f, 0.f);
            const float m = k ? sigc[r] : 0.f;
            vn[r] = f2_make(c.x * m, c.y * m);
          }
        }
      }
      // thermal sums ahead of the transforms: nsample^-1/2 is dead before the butterflies start (four-step:
      // in the shadow of the team-wide exchange, between a warp's arrival and its wait)
      auto thermal_sums = [&]() {
        const f2 it2 = f2_make(itv, itv);
        if constexpr (C::TACC) {
          // (the stores of the previous row were fenced by its exchange)
          uint32_t w16[16], w8[8];
          tm_ld_w16(tm_a + TC_T, w16);
          tm_ld_w8(tm_a + TC_T + 16, w8);
          tm_wait_ld();
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            tA[m] = f2_from_words(w16[2 * m], w16[2 * m + 1]);
            tB[m] = f2_from_words(w16[8 + 2 * m], w16[8 + 2 * m + 1]);
            tC[m] = f2_from_words(w8[2 * m], w8[2 * m + 1]);
          }
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const f2 aa = f2_make(av[2 * m], av[2 * m + 1]);
          const f2 ab = f2_mul(aa, it2);
          tA[m] = f2_add(tA[m], aa);
          tB[m] = f2_add(tB[m], ab);
          tC[m] = f2_fma(aa, ab, tC[m]);
        }
        if constexpr (C::TACC) {
          uint32_t w16[16], w8[8];
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            w16[2 * m] = (uint32_t)tA[m]; w16[2 * m + 1] = (uint32_t)(tA[m] >> 32);
            w16[8 + 2 * m] = (uint32_t)tB[m]; w16[8 + 2 * m + 1] = (uint32_t)(tB[m] >> 32);
            w8[2 * m] = (uint32_t)tC[m]; w8[2 * m + 1] = (uint32_t)(tC[m] >> 32);
          }
          tm_st_w16(tm_a + TC_T, w16);
          tm_st_w8(tm_a + TC_T + 16, w8);
        }
      };
      if (HAS_T && !C::F4) thermal_sums();
      // a panel counts only when both of its ends are valid (masked_invalid, ds.py:3697); a row
      // with any invalid sample books what the full sums above contain but the panel must not
      auto book_deficits = [&]() {
        if (!dirty_time) {
#pragma unroll 1
          for (int k = 0; k < 48; ++k) dfc[k] = 0.f;
          dirty_time = true;
        }
        if (active) {
#pragma unroll 1
          for (int q = 0; q < 8; ++q) {
            const int c = tau + TPR * q;
            if (c + 1 >= N) continue;
            const float nl = ns_s[TPR * q], nr = ns_s[TPR * q + 1];
            if ((nl > 0.f) == (nr > 0.f)) continue;
            const float a1 = fast_rsqrt(nl > 0.f ? nl : nr);
            float *d = dfc + 6 * q + (nl > 0.f ? 0 : 3);
            d[0] += a1;
            d[1] += a1 * itv;
            d[2] += a1 * a1 * itv;
          }
        }
      };
      if (HAS_N) {
        // both transforms in lockstep: two independent dependency chains per team sync
        if constexpr (C::F4) {
          F4T::front(vd, vn);
          mbar_wait_a(free_a, par_f);  // every warp has made its last read of the previous row's buffer
          par_f ^= 1u;
          F4T::store(vd, vn, xa, tau);
          // the row's ONE team-wide synchronisation as an arrive / wait pair on an mbarrier (one arrival
          // per warp), with the thermal sums in between; the "some sample of the row is invalid" vote
          // travels as the row's sequence number in one of two flag words (a word is rewritten two rows
          // later, when every warp has long read it)
          ++seq_f4;
          {
            const bool wd = __any_sync(0xffffffffu, active && !all_ok);
            __syncwarp();
            if ((threadIdx.x & 31u) == 0) {
              if (wd) sts_u32(flag_a + 4u * (seq_f4 & 1u), seq_f4);
              asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(stored_a) : "memory");
            }
          }
          if (HAS_T) thermal_sums();
          f2 tw4[8];
          F4T::twiddles(tw4, tw_s, tau);
          mbar_wait_a(stored_a, par_s);
          par_s ^= 1u;
          dirty_row = lds_u32(flag_a + 4u * (seq_f4 & 1u)) == seq_f4;
          if (HAS_T && dirty_row) {  // rare: the booking re-reads nsample of the staged row
            book_deficits();
            team_sync<TEAM>(bar_id);
          }
          // all warps have consumed this row's slot: refill it (two rows ahead)
          if (pleft > 0) issue();
          F4T::back(vd, vn, tw4, xa, fa4, tw_s, tau, bar_id, tm_a, free_a);
        } else if (C::TM) FFT::run2_tm(vd, vn, fa, tw_s, tau, bar_id, tm_a);
        else if (C::IL) FFT::run2_il(vd, vn, fa, tw_s, tau, bar_id);
        else if (HAS_D) FFT::run2(vd, vn, fa, tw_s, tau, bar_id);
        else if constexpr (C::F4S) F4T::run1(vn, xa, fa4s, tw_s, tau, bar_id, tm_a);
        else if (C::TM1) FFT::run_tm1(vn, fa, tw_s, tau, bar_id, tm_a);
        else FFT::run(vn, fa, tw_s, tau, bar_id);
        if constexpr (C::TACC) {
          // (the exchange inside the transforms fenced the stores of the previous row)
          uint32_t w32[32], w16[16];
          tm_ld_w32(tm_a + TC_S1, w32);
          tm_ld_w16(tm_a + TC_S2, w16);
          tm_wait_ld();
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            S1[r] = f2_from_words(w32[2 * r], w32[2 * r + 1]);
            Sn1[r] = f2_from_words(w32[16 + 2 * r], w32[16 + 2 * r + 1]);
            S2[r] = __uint_as_float(w16[r]);
            Sn2[r] = __uint_as_float(w16[8 + r]);
          }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          Sn1[r] = f2_add(Sn1[r], vn[r]);
          Sn2[r] = fmaf(f2_lo(vn[r]), f2_lo(vn[r]), fmaf(f2_hi(vn[r]), f2_hi(vn[r]), Sn2[r]));
        }
      } else {
        if constexpr (C::F4S) F4T::run1(vd, xa, fa4s, tw_s, tau, bar_id, tm_a);
        else if (C::TM1) FFT::run_tm1(vd, fa, tw_s, tau, bar_id, tm_a);
        else FFT::run(vd, fa, tw_s, tau, bar_id);
      }
      if (HAS_D) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          S1[r] = f2_add(S1[r], vd[r]);
          S2[r] = fmaf(f2_lo(vd[r]), f2_lo(vd[r]), fmaf(f2_hi(vd[r]), f2_hi(vd[r]), S2[r]));
        }
      }
      if constexpr (C::TACC) {
        uint32_t w16[16];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          w16[r] = __float_as_uint(S2[r]);
          w16[8 + r] = __float_as_uint(Sn2[r]);
        }
        tm_st_w16(tm_a + TC_S2, w16);
        // (two self-contained branches: merging them costs a register move per accumulator on every row)
        if (ci0 + ROWS >= ng) {  // the time closes with this row: P += |sum over the group's rows|^2
          uint32_t pw[16];
          tm_ld_w16(tm_a + TC_PW, pw);
          tm_wait_ld();
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            pw[r] = __float_as_uint(__uint_as_float(pw[r]) + f2_hsum(f2_mul(S1[r], S1[r])));
            pw[8 + r] = __float_as_uint(__uint_as_float(pw[8 + r]) + f2_hsum(f2_mul(Sn1[r], Sn1[r])));
          }
          tm_st_w16(tm_a + TC_PW, pw);
          const uint32_t z32[32] = {};
          tm_st_w32(tm_a + TC_S1, z32);
        } else {
          uint32_t w32[32];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            w32[2 * r] = (uint32_t)S1[r]; w32[2 * r + 1] = (uint32_t)(S1[r] >> 32);
            w32[16 + 2 * r] = (uint32_t)Sn1[r]; w32[16 + 2 * r + 1] = (uint32_t)(Sn1[r] >> 32);
          }
          tm_st_w32(tm_a + TC_S1, w32);
        }
      }

      if (HAS_T && !C::F4) {
        if (team_any<TEAM>(active && !all_ok, bar_id)) book_deficits();
      }

      if (T64) {
        // a = nsample^-1/2 in fp64: the fp32 approximation (2 ulp) and one Newton step (-> 1e-13)
        const double itd = TV ? itd_ring : (tint > 0.f) ? 1.0 / (double)tint : 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const double nvd = (double)ns_s[TPR * r], y0 = (double)av[r];
          // invalid sample or inactive row (whose slot holds stale bytes): a = 0
          const double a = av[r] > 0.f ? y0 * fma(-0.5 * nvd * y0, y0, 1.5) : 0.0;
          const double ab = a * itd;
          dA[r] += a;
          dB[r] += ab;
          dC[r] = fma(a, ab, dC[r]);
        }
        if (team_any<TEAM>(active && !all_ok, bar_id)) {
          if (!dirty_time) {
#pragma unroll 1
            for (int k = 0; k < 48; ++k) ddfc[k] = 0.0;
            dirty_time = true;
          }
          if (active) {
#pragma unroll 1
            for (int q = 0; q < 8; ++q) {
              const int c = tau + TPR * q;
              if (c + 1 >= N) continue;
              const float nl = ns_s[TPR * q], nr = ns_s[TPR * q + 1];
              if ((nl > 0.f) == (nr > 0.f)) continue;
              const double a1 = rsqrt_t<double>(nl > 0.f ? nl : nr);
              double *d = ddfc + 6 * q + (nl > 0.f ? 0 : 3);
              d[0] += a1;
              d[1] += a1 * itd;
              d[2] += a1 * a1 * itd;
            }
          }
        }
      }

      // ---- advance the consumer; close the time when its last row is done -----------
      slot_c = slot_c + 1 == C::STAGES ? 0 : slot_c + 1;
      if (slot_c == 0) par_c ^= 1u;
      ci0 += ROWS;
      if (ci0 >= ng) {
        ci0 = 0;
        ++ct;
        ++it_off;
        ++rid_t;
#pragma unroll
        for (int r = 0; r < 8 && !C::TACC; ++r) {  // P += |sum over the group's rows|^2
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
          const double *wlp = P.wl + ((int64_t)s * P.npol + p) * N + tau;
          const double *wrp = P.wr + ((int64_t)s * P.npol + p) * N + tau;
          float Ar[8], Br[8], Cr[8], Ap[8], Bp[8], Cp[8];
          if constexpr (C::TACC) {  // (stored before this row's exchange, which fenced them)
            uint32_t w16[16], w8[8];
            tm_ld_w16(tm_a + TC_T, w16);
            tm_ld_w8(tm_a + TC_T + 16, w8);
            tm_wait_ld();
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              tA[m] = f2_from_words(w16[2 * m], w16[2 * m + 1]);
              tB[m] = f2_from_words(w16[8 + 2 * m], w16[8 + 2 * m + 1]);
              tC[m] = f2_from_words(w8[2 * m], w8[2 * m + 1]);
            }
            const uint32_t z16[16] = {};
            const uint32_t z8[8] = {};
            tm_st_w16(tm_a + TC_T, z16);
            tm_st_w8(tm_a + TC_T + 16, z8);
          }
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            Ar[2 * m] = f2_lo(tA[m]); Ar[2 * m + 1] = f2_hi(tA[m]);
            Br[2 * m] = f2_lo(tB[m]); Br[2 * m + 1] = f2_hi(tB[m]);
            Cr[2 * m] = f2_lo(tC[m]); Cr[2 * m + 1] = f2_hi(tC[m]);
          }
          if (ROWS == 1 && !dirty_time) {
            // no invalid sample in this time: every panel counts, and the left-end weight of panel c and the
            // right-end weight of panel c - 1 multiply the same sums -- one weight per channel, no neighbour
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const double w = __ldg(wlp + TPR * r) + ((r > 0 || tau > 0) ? __ldg(wrp + TPR * r - 1) : 0.0);
              th_all_acc += w * (double)(Ar[r] * Br[r]);
              th_auto_acc += w * (double)Cr[r];
            }
          } else {
          // the right end of a panel is the next channel: another thread's sums, through this
          // row's exchange buffers ((A, B) pairs in the first, C in the second)
          team_sync<TEAM>(bar_id);  // the step's last transform reads are done
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
          // four-step: the next row's transposed stores are gated by the buffer-release barrier, which does
          // not know about these reads
          if (C::F4) team_sync<TEAM>(bar_id);
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
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) tA[k] = tB[k] = tC[k] = 0ull;
          dirty_time = false;
        }
        if (T64) {
          const double *wlp = P.wl + ((int64_t)s * P.npol + p) * N + tau;
          const double *wrp = P.wr + ((int64_t)s * P.npol + p) * N + tau;
          // right neighbours' sums (channel + 1) through this row's first exchange buffer, one
          // quantity per round (N doubles fill the buffer)
          double Ap[8], Bp[8], Cp[8];
#pragma unroll
          for (int rnd = 0; rnd < 3; ++rnd) {
            team_sync<TEAM>(bar_id);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const double v = rnd == 0 ? dA[r] : rnd == 1 ? dB[r] : dC[r];
              asm volatile("st.shared.f64 [%0], %1;" ::"r"(xa + 8u * (uint32_t)(tau + TPR * r)), "d"(v) : "memory");
            }
            team_sync<TEAM>(bar_id);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              double v = 0.0;
              if (r < 7 || tau + 1 < TPR)
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(xa + 8u * (uint32_t)(tau + TPR * r + 1)) : "memory");
              if (rnd == 0) Ap[r] = v; else if (rnd == 1) Bp[r] = v; else Cp[r] = v;
            }
          }
          if (dirty_time) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              dA[r] -= ddfc[6 * r + 0]; dB[r] -= ddfc[6 * r + 1]; dC[r] -= ddfc[6 * r + 2];
              Ap[r] -= ddfc[6 * r + 3]; Bp[r] -= ddfc[6 * r + 4]; Cp[r] -= ddfc[6 * r + 5];
            }
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            double a = dA[r], b = dB[r], ap = Ap[r], bp = Bp[r];
#pragma unroll
            for (int off = TPR; off < 32; off <<= 1) {
              a += __shfl_xor_sync(0xffffffffu, a, off);
              b += __shfl_xor_sync(0xffffffffu, b, off);
              ap += __shfl_xor_sync(0xffffffffu, ap, off);
              bp += __shfl_xor_sync(0xffffffffu, bp, off);
            }
            const double l = __ldg(wlp + TPR * r), rr = __ldg(wrp + TPR * r);
            if (sub == 0) th_all_acc += l * a * b + rr * ap * bp;
            th_auto_acc += l * dC[r] + rr * Cp[r];
            dA[r] = dB[r] = dC[r] = 0.0;
          }
          dirty_time = false;
        }
      }
    }

    // ---- flush the item: fp64 atomics, fftshift as an index rotation -------------
    uint32_t pwt[16];
    if constexpr (C::TACC) {
      uint32_t w16[16];
      tm_wait_st();
      tm_ld_w16(tm_a + TC_S2, w16);
      tm_ld_w16(tm_a + TC_PW, pwt);
      tm_wait_ld();
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        S2[r] = __uint_as_float(w16[r]);
        Sn2[r] = __uint_as_float(w16[8 + r]);
      }
      const uint32_t z16[16] = {};
      tm_st_w16(tm_a + TC_PW, z16);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int kshift = (tau_out + TPR * r + N / 2) % N;
      float s2 = S2[r], sn2 = Sn2[r];
#pragma unroll
      for (int off = TPR; off < 32; off <<= 1) {
        s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        if (HAS_N) sn2 += __shfl_xor_sync(0xffffffffu, sn2, off);
      }
      if (sub == 0) {
        if (HAS_D) {
          atomicAdd(P.pow_all + obase + kshift, (double)(C::TACC ? __uint_as_float(pwt[r]) : pw_s[tau + TPR * r]));
          atomicAdd(P.pow_auto + obase + kshift, (double)s2);
          if (!C::TACC) pw_s[tau + TPR * r] = 0.f;
        }
        if (HAS_N) {
          atomicAdd(P.noise_all + obase + kshift,
                    (double)(C::TACC ? __uint_as_float(pwt[8 + r]) : pw_s[N + tau + TPR * r]));
          atomicAdd(P.noise_auto + obase + kshift, (double)sn2);
          if (!C::TACC) pw_s[N + tau + TPR * r] = 0.f;
        }
      }
    }
    if (HAS_T || T64) {
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
  if (TMANY) {
    tm_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tm_dealloc<C::TM_COLS>(tm_base_s);
  }
}

template <int N, int LEGS, bool GAUSS = false>
static int launch_f32(const EngineParams &P, cudaStream_t st, int *launches) {
  using C = E3Cfg<N, LEGS, GAUSS>;
  static DeviceLatch configured;
  int cfg_dev = 0;
  if (configured.need(&cfg_dev)) {
    SDS_CUDA(cudaFuncSetAttribute(engine_f32_kernel<N, LEGS, GAUSS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    // several CTAs per SM only fit when the L1/shared split favours shared memory
    SDS_CUDA(cudaFuncSetAttribute(engine_f32_kernel<N, LEGS, GAUSS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    configured.set(cfg_dev);
  }
  int dev = 0, nsm = 148, per_sm = 1;
  SDS_CUDA(cudaGetDevice(&dev));
  SDS_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  SDS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, engine_f32_kernel<N, LEGS, GAUSS>,
                                                         C::THREADS, C::SMEM));
  if (per_sm < 1) {
    set_error("sds_engine_run: kernel does not fit an SM (smem %d B)", C::SMEM);
    return SDS_ENOTSUP;
  }
  // the occupancy query answers 1 for a kernel that allocates tensor memory, whatever it allocates: the
  // grid is sized from the kernel's own register / shared-memory / column budget (the work queue is
  // dynamic: a CTA that had to wait for an SM would still be correct)
  if (C::TM || C::TM1) per_sm = tmem_ctas_per_sm(engine_f32_kernel<N, LEGS, GAUSS>, C::THREADS, C::SMEM, C::TM_COLS);
  engine_f32_kernel<N, LEGS, GAUSS><<<nsm * per_sm, C::THREADS, C::SMEM, st>>>(P);
  SDS_LAUNCH_CHECK("engine_f32_kernel");
  ++*launches;
  return SDS_OK;
}

template <int N>
static int run_f32_n(const EngineParams &P, int legs, cudaStream_t st, int *launches) {
#ifdef E3_DEV_ONLY  // development builds: one instantiation (SASS inspection, quick compiles)
  return launch_f32<N, 7>(P, st, launches);
#else
  if (P.noise_mode == 1 && P.noise_gauss) {
    switch (legs) {
      case LEG_NOISE | LEG_THERMAL64: return launch_f32<N, LEG_NOISE | LEG_THERMAL64, true>(P, st, launches);
      case 2: return launch_f32<N, 2, true>(P, st, launches);
      case 3: return launch_f32<N, 3, true>(P, st, launches);
      case 7: return launch_f32<N, 7, true>(P, st, launches);
      default: break;
    }
  }
  switch (legs) {
    case 1: return launch_f32<N, 1>(P, st, launches);
    case LEG_NOISE | LEG_THERMAL64: return launch_f32<N, LEG_NOISE | LEG_THERMAL64>(P, st, launches);
    case 2: return launch_f32<N, 2>(P, st, launches);
    case 3: return launch_f32<N, 3>(P, st, launches);
    case 5: return launch_f32<N, 5>(P, st, launches);
    case 7: return launch_f32<N, 7>(P, st, launches);
    default: break;
  }
  set_error("sds_engine_run: leg mask %d is not instantiated", legs);
  return SDS_ENOTSUP;
#endif
}

int run_engine_f32(int N, const EngineParams &P, int legs, cudaStream_t st, int *launches) {
#ifdef E3_DEV_ONLY
  return run_f32_n<E3_DEV_ONLY>(P, legs, st, launches);
#endif
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
