// Fused engine: one pass over the 13 B/sample inputs (complex64 visibility, u8
// flag, f32 nsample) produces, per redundant group / spectral window /
// polarisation / delay, the pair sums of the data power, of a matching noise
// realisation, and the thermal-noise sums -- the reference's
//   select_spectral_windows -> generate_noise -> delay_transform ->
//   calculate_delay_spectrum -> mean over baseline pairs and time
// (ds.py:3196-3338, 3340-3343, 3487-3583, 3585-3707; utils.py:387-389, 486-566)
// without any (Nbls x Nbls) tensor.
//
// Layout of the work
//   item   = (group g, window s, pol p, chunk of times); items are claimed from
//            an atomic counter by "workers" (a row team of max(32, N/8) threads).
//   step   = one row (N/8 >= 32) or 32/(N/8) rows of the item, staged in shared
//            memory by 1-D bulk TMA copies (cp.async.bulk, mbarrier complete_tx)
//            through a ring of STAGES slots per worker.
//   row    = flag mask * taper * delta_f -> length-N transform in registers and
//            shared memory (sds_radix.cuh) -> S1 += X, S2 += |X|^2; the same for
//            the noise row generated in registers (Philox4x32-10 + Box-Muller);
//            thermal partial sums from nsample.
//   end of a time: P += |S1|^2 (the sum over ALL ordered pairs factorises),
//   end of an item: fp64 atomics into the outputs (fftshift = index rotation).
#include "sds_engine.cuh"
#include "sds_packed.cuh"

namespace sds {

// ---- setup: item table, thermal weights, counter -------------------------------
__global__ void engine_setup_kernel(const int32_t *__restrict__ goff, int ngroup, int ntime,
                                    int nspw, int npol, int N, int target_rows,
                                    const double *__restrict__ thermal_coef,
                                    const double *__restrict__ freq, int32_t *item_start,
                                    double *wl, double *wr, int *counter) {
  // thermal weights: wl[k] = dnu_k/2 * c[k], wr[k] = dnu_k/2 * c[k+1]; a panel whose
  // channel factor is not finite is excluded (masked_invalid, ds.py:3697)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nspw * npol * N;
       i += gridDim.x * blockDim.x) {
    const int k = i % N, s = i / (N * npol);
    double a = 0.0, b = 0.0;
    if (thermal_coef && k < N - 1) {
      const double c0 = thermal_coef[i], c1 = thermal_coef[i + 1];
      if (isfinite(c0) && isfinite(c1)) {
        const double hd = 0.5 * (freq[s * N + k + 1] - freq[s * N + k]);
        a = hd * c0;
        b = hd * c1;
      }
    }
    wl[i] = a;
    wr[i] = b;
  }
  if (blockIdx.x == 0) {  // exclusive scan of items per group (ngroup <= 65535: one block)
    __shared__ int carry, part[1024];
    if (threadIdx.x == 0) {
      carry = 0;
      *counter = 0;
    }
    __syncthreads();
    for (int g0 = 0; g0 < ngroup; g0 += blockDim.x) {
      const int g = g0 + threadIdx.x;
      int cnt = 0;
      if (g < ngroup) {
        const int ng = goff[g + 1] - goff[g];
        if (ng > 0) {
          const int c = eng_chunk(ng, ntime, target_rows);
          cnt = ((ntime + c - 1) / c) * nspw * npol;
        }
      }
      part[threadIdx.x] = cnt;
      __syncthreads();
      for (int off = 1; off < blockDim.x; off <<= 1) {
        int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
      }
      if (g < ngroup) item_start[g] = carry + part[threadIdx.x] - cnt;
      __syncthreads();
      if (threadIdx.x == blockDim.x - 1) carry += part[threadIdx.x];
      __syncthreads();
    }
    if (threadIdx.x == 0) item_start[ngroup] = carry;
  }
}

// ---- complex128 transform of length 256, last exchange through tensor memory ------------------------
// The fp64 data leg is bound by the shared-memory data pipe (89 % busy: 16-byte values through two
// exchanges and 13 twiddle loads per row).  Its last exchange goes through tensor memory exactly like the
// packed fp32 engine's (sds_packed.cuh, TeamFftX::pass_il): a double is a 64-bit column pair, which the
// 16x256b fragment keeps whole, so the real parts of the thread's eight values travel in columns [0, 16)
// and the imaginary parts in [16, 32) -- the two planes play the role of the fp32 engine's data and noise
// rows.  Afterwards lane l acts as team member tm_tau(l) and owns bins tm_tau(l) + 32 r.
struct TeamFftD256 {
  using F = TeamFftX<256>;
  static __device__ __forceinline__ unsigned long long bits(double v) { return (unsigned long long)__double_as_longlong(v); }
  static __device__ __forceinline__ double dbl(unsigned long long v) { return __longlong_as_double((long long)v); }
  // first exchange: a complex128 value is one 16-byte slot of the padded layout of the fp32 engine's
  // interleaved exchange (slot j at j + (j >> 3): conflict-free 128-bit accesses at [register + immediate])
  template <int p> static __device__ __forceinline__ void store0(uint32_t st, const cpx<double> (&a)[8]) {
    if constexpr (p < 8) {
      sts_f2x2_off<16 * p>(st, bits(a[brev(p, 3)].x), bits(a[brev(p, 3)].y));
      store0<p + 1>(st, a);
    }
  }
  template <int r> static __device__ __forceinline__ void load0(uint32_t ld, cpx<double> (&v)[8]) {
    if constexpr (r < 8) {
      unsigned long long re, im;
      lds_f2x2_off<16 * F::il_pad(32 * r)>(ld, re, im);
      v[r] = {dbl(re), dbl(im)};
      load0<r + 1>(ld, v);
    }
  }
  // xa: shared-space address of the row's F::IL_BYTES exchange region
  static __device__ __forceinline__ void run_tm(cpx<double> (&v)[8], uint32_t xa, const cpx<double> *__restrict__ tw,
                                                int tau, uint32_t tm) {
    constexpr int TPR = 32;
    {  // pass 0: radix 8, Stockham store (NS = 1)
      cpx<double> a[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = v[q];
      Radix<double, 8>::run(a);
      store0<0>(xa + 16u * (uint32_t)F::il_pad(8 * tau), a);
      __syncwarp();
      load0<0>(xa + 16u * (uint32_t)F::il_pad(tau), v);
    }
    {  // pass 1: twiddles, radix 8, tensor-memory exchange
      constexpr int OFF = fft_tw_offset(256, 1);
#pragma unroll
      for (int q = 1; q < 8; ++q) v[q] = cmul(v[q], tw[OFF + (q - 1) * TPR + tau]);
      Radix<double, 8>::run(v);
      tm_st8(tm, bits(v[brev(0, 3)].x), bits(v[brev(1, 3)].x), bits(v[brev(2, 3)].x), bits(v[brev(3, 3)].x),
             bits(v[brev(4, 3)].x), bits(v[brev(5, 3)].x), bits(v[brev(6, 3)].x), bits(v[brev(7, 3)].x));
      tm_st8(tm + 16u, bits(v[brev(0, 3)].y), bits(v[brev(1, 3)].y), bits(v[brev(2, 3)].y), bits(v[brev(3, 3)].y),
             bits(v[brev(4, 3)].y), bits(v[brev(5, 3)].y), bits(v[brev(6, 3)].y), bits(v[brev(7, 3)].y));
    }
    // (this lane acts as team member tm_tau(tau) from here on; the last pass's twiddles are stored permuted)
    constexpr int OFF2 = fft_tw_offset(256, 2);
    tm_wait_st();
    unsigned long long re[8], im[8];
    // loaded slot (tau4 t2 tau3) is logical slot (tau4 tau3 t2): 1 <-> 2, 5 <-> 6
    tm_ld4(tm, re[0], re[2], re[1], re[3]);
    tm_ld4(tm + (16u << 16), re[4], re[6], re[5], re[7]);
    tm_ld4(tm + 16u, im[0], im[2], im[1], im[3]);
    tm_ld4(tm + (16u << 16) + 16u, im[4], im[6], im[5], im[7]);
    tm_wait_ld();
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = {dbl(re[r]), dbl(im[r])};
    // pass 2: radix 4, two butterflies (m = 0, 1) on slots m + 2 q
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      cpx<double> a[4];
      a[0] = v[m];
#pragma unroll
      for (int q = 1; q < 4; ++q) a[q] = cmul(v[m + 2 * q], tw[OFF2 + (m * 3 + (q - 1)) * TPR + tau]);
      Radix<double, 4>::run(a);
#pragma unroll
      for (int p = 0; p < 4; ++p) v[m + 2 * p] = a[brev(p, 2)];
    }
  }
};

// ---- complex128 transform of length 2048 = 8 x 256, four-step (as TeamFft8x256 of the fp32 engine) ----
// One radix-8 pass across the 256-thread team, ONE team-wide exchange (16-byte slots: region k1 of 4608
// bytes, position n2), then a 256-point transform per warp (TeamFftD256) in the warp's own region, with no
// further team barrier.  On return lane l of warp w owns bins w + 8 (tm_tau(l) + 32 r).
struct TeamFftD8x256 {
  using D = TeamFftD256;
  static constexpr int REGION = TeamFftX<256>::IL_BYTES;
  static constexpr int XBYTES = 8 * REGION;
  static constexpr int TW_A = 2048;                           // [k1][n2] W_2048^(n2 k1)
  static constexpr int TW_TOTAL = TW_A + fft_tw_total(256);   // + the team table of length 256 (last pass permuted)
  static __device__ __forceinline__ void fill_tables(cpx<double> *tab, int tid, int nthreads) {
    for (int i = tid; i < TW_A; i += nthreads) {
      const int k1 = i >> 8, n2 = i & 255;
      double sn, cs;
      sincospi(-2.0 * (double)((n2 * k1) & 2047) / 2048.0, &sn, &cs);
      tab[i] = {cs, sn};
    }
    constexpr int NP = fft_npass(256);
#pragma unroll
    for (int p = 1; p < NP; ++p) {
      const int R = fft_radix(256, p), NB = 8 / R, NS = fft_ns(256, p), off = fft_tw_offset(256, p);
      for (int i = tid; i < NB * (R - 1) * 32; i += nthreads) {
        const int tau = p == NP - 1 ? TeamFftX<256>::tm_tau(i & 31) : (i & 31);
        const int mq = i >> 5, m = mq / (R - 1), q = mq % (R - 1) + 1;
        const int k = (tau + 32 * m) % NS;
        double sn, cs;
        sincospi(-2.0 * (double)(q * k) / (double)(NS * R), &sn, &cs);
        tab[TW_A + off + i] = {cs, sn};
      }
    }
  }
  template <int k1> static __device__ __forceinline__ void store_t(uint32_t st, const cpx<double> (&a)[8]) {
    if constexpr (k1 < 8) {
      sts_f2x2_off<k1 * REGION>(st, D::bits(a[brev(k1, 3)].x), D::bits(a[brev(k1, 3)].y));
      store_t<k1 + 1>(st, a);
    }
  }
  template <int r> static __device__ __forceinline__ void load_t(uint32_t ld, cpx<double> (&v)[8]) {
    if constexpr (r < 8) {
      unsigned long long re, im;
      lds_f2x2_off<512 * r>(ld, re, im);
      v[r] = {D::dbl(re), D::dbl(im)};
      load_t<r + 1>(ld, v);
    }
  }
  // (the caller team-syncs between two calls: the step loop does, at the top of every step)
  static __device__ __forceinline__ void run(cpx<double> (&v)[8], uint32_t xT, const cpx<double> *__restrict__ tw,
                                             int tau, int bar, uint32_t tm) {
    Radix<double, 8>::run(v);
    store_t<0>(xT + 16u * (uint32_t)tau, v);
    team_sync<256>(bar);
    const int w = tau >> 5, lane = tau & 31;
    load_t<0>(xT + (uint32_t)w * REGION + 16u * (uint32_t)lane, v);
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = cmul(v[r], tw[w * 256 + lane + 32 * r]);
    __syncwarp();  // the warp has read its region; its own exchange may overwrite it
    D::run_tm(v, xT + (uint32_t)w * REGION, tw + TW_A, lane, tm);
  }
};

// ---- the fused kernel ---------------------------------------------------------
template <typename T, int N, int LEGS> struct EngCfg {
  static constexpr int TPR = N / 8;
  static constexpr int WORKER = TPR < 32 ? 32 : TPR;   // threads per worker
  static constexpr int ROWS = WORKER / TPR;            // rows per step
  static constexpr int THREADS = WORKER < 128 ? 128 : WORKER;
  static constexpr int NWORKER = THREADS / WORKER;
  // (the fp64 four-step data leg stages only what it reads: vis 8N | flags N -- two CTAs per SM)
  static constexpr bool F4D = sizeof(T) == 8 && LEGS == LEG_DATA && N == 2048;
  static constexpr int FL_OFF = F4D ? 8 * N : 12 * N;  // flags inside a staged row
  static constexpr int ROWB = F4D ? 9 * N : 13 * N;    // vis 8N | nsample 4N | flags N
  static constexpr int SLOT = ROWS * ROWB;
  static constexpr int BUF = fft_buf_elems(N) * (int)sizeof(cpx<T>);
  // twiddle tables only for the legs of this instantiation (the data leg's in T, the noise
  // leg's in float): at N = 2048 in fp64 both tables, the ring and the exchange buffers
  // would not fit the 227 KB of an SM
  static constexpr bool SAME = sizeof(T) == sizeof(float);
  static constexpr int TW_BYTES = F4D ? TeamFftD8x256::TW_TOTAL * (int)sizeof(cpx<T>)
                                  : (LEGS & LEG_DATA) ? fft_tw_total(N) * (int)sizeof(cpx<T>) : 0;
  static constexpr int TWF_BYTES =
      ((LEGS & LEG_NOISE) && !(SAME && (LEGS & LEG_DATA))) ? fft_tw_total(N) * (int)sizeof(cpx<float>) : 0;
  static constexpr int TAB_BYTES = ((TW_BYTES + TWF_BYTES + 127) / 128) * 128;
  // per worker: barriers (64 B) | STAGES slots | 2 exchange buffers per row |
  // thermal leg: 6 values per thread (first-channel sums of its two runs, for the left neighbour)
  static constexpr int NB_BYTES = (LEGS & LEG_THERMAL) ? 6 * WORKER * (int)sizeof(T) : 0;
  // the fp64 data leg: a 2-slot ring leaves shared memory for three CTAs (12 warps) per SM at
  // N = 256, which beats the deeper prefetch at 8 warps (measured 7.23 -> 6.69 ms per HERA-350 batch)
  static constexpr bool D64 = sizeof(T) == 8 && LEGS == LEG_DATA;
  static constexpr bool TM = D64 && (N == 256 || F4D);  // last exchange through tensor memory (TeamFftD256)
#ifndef ENG_D64_TACC
#define ENG_D64_TACC 1
#endif
  // the fp64 accumulators (S1: 32 words, S2: 16, the per-time sums: 16) in thread-private tensor-memory
  // columns [32, 96), as in the fp32 engine: 128 registers, four CTAs per SM
  static constexpr bool TACC = TM && ENG_D64_TACC;
  static constexpr int TM_WCOLS = TACC ? 128 : 32;             // per set of four warps (they share TMEM lanes)
  static constexpr int TM_COLS = TM_WCOLS * (THREADS / 128);
  // the thermal leg transforms nothing: no exchange buffers, and a register budget for 16 warps / SM
  static constexpr int XB_BYTES = F4D ? TeamFftD8x256::XBYTES : TM ? ROWS * TeamFftX<256>::IL_BYTES
                                  : (LEGS & (LEG_DATA | LEG_NOISE)) ? 2 * ROWS * BUF : 0;
  static constexpr int STAGES = D64 ? 2 : ENG_STAGES;
  static constexpr int MINB = (LEGS == LEG_THERMAL) ? (512 / THREADS < 1 ? 1 : 512 / THREADS)
                              : TACC ? 512 / THREADS : D64 ? (384 / THREADS < 1 ? 1 : 384 / THREADS) : ENG_MINBLOCKS;
  static constexpr int WORKER_BYTES = 64 + STAGES * SLOT + XB_BYTES + NB_BYTES;
  static constexpr int SMEM = TAB_BYTES + NWORKER * WORKER_BYTES;
};

template <typename T, int N, int LEGS>
__global__ void __launch_bounds__(EngCfg<T, N, LEGS>::THREADS, EngCfg<T, N, LEGS>::MINB)
engine_kernel(const __grid_constant__ EngineParams P) {
  using C = EngCfg<T, N, LEGS>;
  constexpr int TPR = C::TPR, WORKER = C::WORKER, ROWS = C::ROWS;
  constexpr bool SAME = sizeof(T) == sizeof(float);
  constexpr bool HAS_D = (LEGS & LEG_DATA) != 0, HAS_N = (LEGS & LEG_NOISE) != 0,
                 HAS_T = (LEGS & LEG_THERMAL) != 0;
  extern __shared__ __align__(128) unsigned char smem[];

  // twiddle tables -> shared memory (conflict-free [m][q][tau] layout)
  cpx<T> *tw_s = reinterpret_cast<cpx<T> *>(smem);
  cpx<float> *twf_s = SAME ? reinterpret_cast<cpx<float> *>(smem)
                           : reinterpret_cast<cpx<float> *>(smem + C::TW_BYTES);
  if constexpr (C::F4D) {
    TeamFftD8x256::fill_tables(reinterpret_cast<cpx<double> *>(smem), threadIdx.x, C::THREADS);
  } else {
    for (int i = threadIdx.x; i < fft_tw_total(N); i += C::THREADS) {
      if (C::TW_BYTES) tw_s[i] = reinterpret_cast<const cpx<T> *>(P.tw)[C::TM ? TeamFftX<N>::tm_tw_src(i) : i];
      if (C::TWF_BYTES) twf_s[i] = {P.tw_f[i].x, P.tw_f[i].y};
    }
  }
  const int wid = threadIdx.x / WORKER, wl = threadIdx.x % WORKER;
  const int sub = wl / TPR, tau = wl % TPR, bar_id = 1 + wid;
  unsigned char *wbase = smem + C::TAB_BYTES + (size_t)wid * C::WORKER_BYTES;
  uint64_t *full = reinterpret_cast<uint64_t *>(wbase);
  unsigned char *stages = wbase + 64;
  cpx<T> *bufA = reinterpret_cast<cpx<T> *>(stages + C::STAGES * C::SLOT + (size_t)sub * 2 * C::BUF);
  cpx<T> *bufB = reinterpret_cast<cpx<T> *>(reinterpret_cast<unsigned char *>(bufA) + C::BUF);
  T *nb_s = reinterpret_cast<T *>(stages + C::STAGES * C::SLOT + C::XB_BYTES);
  if (wl == 0) {
    for (int i = 0; i < C::STAGES; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __shared__ uint32_t tm_base_s;
  if (C::TM && threadIdx.x < 32) tm_alloc<C::TM_COLS>(smem_u32(&tm_base_s));
  if (C::TM) tm_fence_before();
  __syncthreads();
  uint32_t tm_a = 0;
  if (C::TM) {
    tm_fence_after();
    const uint32_t warp = (uint32_t)threadIdx.x >> 5;
    tm_a = tm_base_s + (((warp & 3u) * 32u) << 16) + (warp >> 2) * (uint32_t)C::TM_WCOLS;
  }
  // the spectrum bins a thread owns: tau_out + TPR r (the tensor-memory exchange renames the lanes)
  const int tau_out = C::F4D ? (tau >> 5) + 8 * TeamFftX<256>::tm_tau(tau & 31)
                      : C::TM ? TeamFftX<256>::tm_tau(tau & 31) : tau;
  uint32_t xa_tm = smem_u32(stages + C::STAGES * C::SLOT);  // TM: one exchange region per worker (ROWS = 1)
  asm volatile("" : "+r"(xa_tm));
  constexpr uint32_t TC_S1 = 32, TC_S2 = 64, TC_PW = 80;
  if (C::TACC) {
    const uint32_t z16[16] = {};
    tm_st_w16(tm_a + TC_PW, z16);
  }

  // per-thread constants: taper * delta_f at this thread's 8 transform channels
  T tapdf[8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
    tapdf[r] = reinterpret_cast<const T *>(P.taper)[tau + TPR * r] * (T)P.delta_f;
  const int total_items = P.item_start[P.ngroup];
  constexpr int ROW_BYTES = (HAS_D ? 8 * N : 0) + ((HAS_N || HAS_T) ? 4 * N : 0) +
                            ((HAS_D || HAS_N) ? N : 0);
  const int64_t row_stride = (int64_t)P.ntime * P.nchan;  // samples between baselines
  // the team's first warp stages its rows (warp-uniform test: shfl from lane 0)
  const bool issuer_warp = WORKER == 32 || __shfl_sync(0xffffffffu, wl, 0) == 0;
  uint32_t slot_c = 0, par_c = 0;  // consumer ring position / phase parity
  uint32_t slot_p = 0;             // producer ring position
  __shared__ int item_bcast[C::NWORKER];

  for (;;) {
    int item;
    if constexpr (WORKER == 32) {
      item = 0;
      if (wl == 0) item = atomicAdd(P.counter, 1);
      item = __shfl_sync(0xffffffffu, item, 0);
    } else {
      team_sync<WORKER>(bar_id);
      if (wl == 0) item_bcast[wid] = atomicAdd(P.counter, 1);
      team_sync<WORKER>(bar_id);
      item = item_bcast[wid];
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

    // ---- producer state (lane 0 stages rows i0 .. i0+ROWS-1 of time t) ----------
    int pi0 = 0;          // first row of the next step to stage
    int64_t pe = e00;     // its sample offset
    int64_t pe_t = e00;   // offset of row 0 at the producer's current time
    int pleft = nsteps;   // steps not yet staged
    auto issue = [&]() {
      if (issuer_warp && elect_one()) {
        const int nact = min(ROWS, ng - pi0);
        mbar_expect_tx(&full[slot_p], (uint32_t)(nact * ROW_BYTES));
        unsigned char *dst = stages + slot_p * C::SLOT;
        int64_t e = pe;
        for (int rr = 0; rr < nact; ++rr, e += row_stride, dst += C::ROWB) {
          if (HAS_D) tma_load_1d(dst, P.vis + e, 8 * N, &full[slot_p]);
          if (HAS_N || HAS_T) tma_load_1d(dst + 8 * N, P.nsample + e, 4 * N, &full[slot_p]);
          if (HAS_D || HAS_N) tma_load_1d(dst + C::FL_OFF, P.flags + e, N, &full[slot_p]);
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

    float sigc[8];
    if (HAS_N) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float c = __ldg(P.sigma_coef + ((int64_t)s * P.npol + p) * N + tau + TPR * r);
        // non-finite sigma -> 0 (ds.py:3575-3582); taper * delta_f folded in
        sigc[r] = (fabsf(c) <= 3.0e38f) ? c * ENG_SQRT_LN2 * (float)tapdf[r] : 0.f;
      }
    }
    cpx<T> S1[8];
    T S2[8], Pw[8];
    cpx<float> Sn1[8];
    float Sn2[8], Pn[8];
    // thermal sums per channel (c = 4 tau + 4 TPR h + j, q = 4 h + j): A = sum a, B = sum a / t_int,
    // C = sum a^2 / t_int, a = nsample^-1/2 (0 where invalid).  Rows with a valid sample next to an
    // invalid one additionally book what the panel mask (masked_invalid, ds.py:3697) removes into
    // dfc[] (local memory, rare): per panel q, [0..2] left-end and [3..5] right-end deficits.
    T tA[8], tB[8], tC[8];
    T dfc[48];
    bool dirty_time = false;
    double th_all_acc = 0.0, th_auto_acc = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      S1[r] = {(T)0, (T)0}; S2[r] = (T)0; Pw[r] = (T)0;
      Sn1[r] = {0.f, 0.f}; Sn2[r] = 0.f; Pn[r] = 0.f;
      tA[r] = tB[r] = tC[r] = (T)0;
    }

    if constexpr (C::TACC) {
      const uint32_t z32[32] = {};
      const uint32_t z16[16] = {};
      tm_st_w32(tm_a + TC_S1, z32);
      tm_st_w16(tm_a + TC_S2, z16);
      tm_wait_st();
    }
    // ---- consumer state ------------------------------------------------------------
    int ci0 = 0, ct = t0;                                   // first row / time of this step
    int64_t it_off = (int64_t)b0 * P.ntime + t0;            // int_time index of row 0
    // Philox counter base of (row 0, time t0): row_id * N/2, row_id = ((s npol + p) nbl_total + b) ntime_total + t (include/sds.h)
    uint64_t rid_t = (((uint64_t)((int64_t)s * P.npol + p) * (uint64_t)P.nbl_total +
                       (uint64_t)blg) * (uint64_t)P.ntime_total +
                      (uint64_t)(P.time_offset + t0));

    for (int k = 0; k < C::STAGES - 1 && pleft > 0; ++k) issue();
    for (int step = 0; step < nsteps; ++step) {
      team_sync<WORKER>(bar_id);  // everyone is done with the slot about to be refilled
      if (pleft > 0) issue();
      const int i = ci0 + sub;
      const bool active = i < ng;
      const int ia = active ? i : 0;
      float tint = 1.f;
      if (HAS_N || HAS_T) tint = __ldg(P.int_time + it_off + (int64_t)ia * P.ntime);
      mbar_wait(&full[slot_c], par_c);
      const unsigned char *row = stages + slot_c * C::SLOT + sub * C::ROWB;
      const float2 *vis_s = reinterpret_cast<const float2 *>(row);
      const float *ns_s = reinterpret_cast<const float *>(row + 8 * N);
      const unsigned char *fl_s = row + C::FL_OFF;

      uint32_t keep = 0;  // bit r set: channel tau + TPR r of an active row is unflagged
      if (HAS_D || HAS_N) {
#pragma unroll
        for (int r = 0; r < 8; ++r) keep |= (fl_s[tau + TPR * r] == 0 ? 1u : 0u) << r;
        if (!active) keep = 0;
      }

      if (HAS_D) {
        cpx<T> v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float2 c = active ? vis_s[tau + TPR * r] : make_float2(0.f, 0.f);
          const T m = (keep >> r) & 1u ? tapdf[r] : (T)0;  // (1 - flag) * taper * df
          v[r] = {(T)c.x * m, (T)c.y * m};
        }
        if constexpr (C::F4D) TeamFftD8x256::run(v, xa_tm, reinterpret_cast<const cpx<double> *>(tw_s), tau, bar_id, tm_a);
        else if constexpr (C::TM) TeamFftD256::run_tm(v, xa_tm, tw_s, tau, tm_a);
        else TeamFft<T, N>::run(v, bufA, bufB, tw_s, tau, bar_id);
        if constexpr (C::TACC) {  // (the exchange inside the transform fenced the stores of the previous row)
          uint32_t w32[32], w16[16];
          tm_ld_w32(tm_a + TC_S1, w32);
          tm_ld_w16(tm_a + TC_S2, w16);
          tm_wait_ld();
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            S1[r] = {(T)TeamFftD256::dbl(f2_from_words(w32[4 * r], w32[4 * r + 1])),
                     (T)TeamFftD256::dbl(f2_from_words(w32[4 * r + 2], w32[4 * r + 3]))};
            S2[r] = (T)TeamFftD256::dbl(f2_from_words(w16[2 * r], w16[2 * r + 1]));
          }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          S1[r] = cadd(S1[r], v[r]);
          S2[r] += v[r].x * v[r].x + v[r].y * v[r].y;
        }
        if constexpr (C::TACC) {
          uint32_t w16[16];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const unsigned long long b2 = TeamFftD256::bits((double)S2[r]);
            w16[2 * r] = (uint32_t)b2; w16[2 * r + 1] = (uint32_t)(b2 >> 32);
          }
          tm_st_w16(tm_a + TC_S2, w16);
          // (two self-contained branches: merging them costs register moves on every row)
          if (ci0 + ROWS >= ng) {  // the time closes with this row
            uint32_t pw[16];
            tm_ld_w16(tm_a + TC_PW, pw);
            tm_wait_ld();
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const double pv = TeamFftD256::dbl(f2_from_words(pw[2 * r], pw[2 * r + 1])) +
                                (double)(S1[r].x * S1[r].x + S1[r].y * S1[r].y);
              const unsigned long long b = TeamFftD256::bits(pv);
              pw[2 * r] = (uint32_t)b; pw[2 * r + 1] = (uint32_t)(b >> 32);
            }
            tm_st_w16(tm_a + TC_PW, pw);
            const uint32_t z32[32] = {};
            tm_st_w32(tm_a + TC_S1, z32);
          } else {
            uint32_t w32[32];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const unsigned long long bx = TeamFftD256::bits((double)S1[r].x), by = TeamFftD256::bits((double)S1[r].y);
              w32[4 * r] = (uint32_t)bx; w32[4 * r + 1] = (uint32_t)(bx >> 32);
              w32[4 * r + 2] = (uint32_t)by; w32[4 * r + 3] = (uint32_t)(by >> 32);
            }
            tm_st_w32(tm_a + TC_S1, w32);
          }
        }
      }

      if (HAS_N) {
        cpx<float> v[8];
        if (P.noise_mode == 1) {
          float rad[8], cs[8], sn[8];
          engine_draw8<TPR>(rid_t + (uint64_t)ia * (uint64_t)P.ntime_total, tau, P.pkey, rad, cs, sn);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float tn = tint * ns_s[tau + TPR * r];
            // sigma sqrt(-ln u1), with mask, taper and delta_f folded into sigc / keep
            float amp = sigc[r] * rad[r] * fast_rsqrt(tn);
            if (!(tn > 0.f) || !((keep >> r) & 1u)) amp = 0.f;
            v[r] = {amp * cs[r], amp * sn[r]};
          }
        } else {  // injected freq-domain noise (parity mode): plain loads
          const int64_t e = e00 + (int64_t)(ct - t0) * P.nchan + (int64_t)ia * row_stride;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float2 c = active ? P.noise_in[e + tau + TPR * r] : make_float2(0.f, 0.f);
            const float m = (keep >> r) & 1u ? (float)tapdf[r] : 0.f;
            v[r] = {c.x * m, c.y * m};
          }
        }
        TeamFft<float, N>::run(v, reinterpret_cast<cpx<float> *>(bufB),
                               reinterpret_cast<cpx<float> *>(bufA), twf_s, tau, bar_id);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          Sn1[r] = cadd(Sn1[r], v[r]);
          Sn2[r] += v[r].x * v[r].x + v[r].y * v[r].y;
        }
      }

      if (HAS_T) {
        // thermal sums use their own channel map: two runs of 4 consecutive channels
        // per thread (c = 4 tau + 4 TPR h + j) -> two 128-bit loads + the right neighbours
        const T it = (tint > 0.f) ? (T)1 / (T)tint : (T)0;
        bool dirty = false;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c0 = 4 * tau + 4 * TPR * h;
          const float4 nv = *reinterpret_cast<const float4 *>(ns_s + c0);
          const float n4[4] = {nv.x, nv.y, nv.z, nv.w};
          bool ok[5];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ok[j] = active && n4[j] > 0.f;
            const T a = ok[j] ? rsqrt_t<T>(n4[j]) : (T)0;
            const T ai = a * it;
            tA[4 * h + j] += a;
            tB[4 * h + j] += ai;
            tC[4 * h + j] += a * ai;
          }
          ok[4] = (c0 + 4 < N) ? (active && ns_s[c0 + 4] > 0.f) : ok[3];  // no panel right of N - 1
          dirty |= (ok[0] != ok[1]) | (ok[1] != ok[2]) | (ok[2] != ok[3]) | (ok[3] != ok[4]);
        }
        if (dirty) {
          if (!dirty_time) {
#pragma unroll 1
            for (int k = 0; k < 48; ++k) dfc[k] = (T)0;
            dirty_time = true;
          }
#pragma unroll 1
          for (int q = 0; q < 8; ++q) {
            const int c = 4 * tau + 4 * TPR * (q >> 2) + (q & 3);
            if (c + 1 >= N) continue;
            const float nl = ns_s[c], nr = ns_s[c + 1];
            if ((nl > 0.f) == (nr > 0.f)) continue;
            const T a1 = rsqrt_t<T>(nl > 0.f ? nl : nr);
            T *d = dfc + 6 * q + (nl > 0.f ? 0 : 3);
            d[0] += a1;
            d[1] += a1 * it;
            d[2] += a1 * a1 * it;
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
        if (HAS_D && !C::TACC) {
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            cpx<T> sfull = S1[r];
#pragma unroll
            for (int off = TPR; off < 32; off <<= 1) {
              sfull.x += __shfl_xor_sync(0xffffffffu, sfull.x, off);
              sfull.y += __shfl_xor_sync(0xffffffffu, sfull.y, off);
            }
            Pw[r] += sfull.x * sfull.x + sfull.y * sfull.y;
            S1[r] = {(T)0, (T)0};
          }
        }
        if (HAS_N) {
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            cpx<float> sfull = Sn1[r];
#pragma unroll
            for (int off = TPR; off < 32; off <<= 1) {
              sfull.x += __shfl_xor_sync(0xffffffffu, sfull.x, off);
              sfull.y += __shfl_xor_sync(0xffffffffu, sfull.y, off);
            }
            Pn[r] += sfull.x * sfull.x + sfull.y * sfull.y;
            Sn1[r] = {0.f, 0.f};
          }
        }
        if (HAS_T) {
          const double *wlp = P.wl + ((int64_t)s * P.npol + p) * N;
          const double *wrp = P.wr + ((int64_t)s * P.npol + p) * N;
          // the right end of a run's last panel is the first channel of the next thread's run
          T *mine = nb_s + ((sub * 2) * TPR + tau) * 3;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mine[h * TPR * 3 + 0] = tA[4 * h];
            mine[h * TPR * 3 + 1] = tB[4 * h];
            mine[h * TPR * 3 + 2] = tC[4 * h];
          }
          team_sync<WORKER>(bar_id);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            T nxt[3] = {(T)0, (T)0, (T)0};
            if (tau + 1 < TPR || h == 0) {
              const T *src = (tau + 1 < TPR) ? mine + h * TPR * 3 + 3 : nb_s + ((sub * 2 + 1) * TPR) * 3;
              nxt[0] = src[0]; nxt[1] = src[1]; nxt[2] = src[2];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = 4 * h + j;
              T a = tA[q], b = tB[q], cc = tC[q];
              T ap = j < 3 ? tA[q + 1] : nxt[0], bp = j < 3 ? tB[q + 1] : nxt[1], cp = j < 3 ? tC[q + 1] : nxt[2];
              if (dirty_time) {
                a -= dfc[6 * q + 0]; b -= dfc[6 * q + 1]; cc -= dfc[6 * q + 2];
                ap -= dfc[6 * q + 3]; bp -= dfc[6 * q + 4]; cp -= dfc[6 * q + 5];
              }
#pragma unroll
              for (int off = TPR; off < 32; off <<= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, off);
                ap += __shfl_xor_sync(0xffffffffu, ap, off);
                b += __shfl_xor_sync(0xffffffffu, b, off);
                bp += __shfl_xor_sync(0xffffffffu, bp, off);
              }
              const int c = 4 * tau + 4 * TPR * h + j;
              const double l = __ldg(wlp + c), rr = __ldg(wrp + c);
              if (sub == 0) th_all_acc += l * (double)a * (double)b + rr * (double)ap * (double)bp;
              th_auto_acc += l * (double)cc + rr * (double)cp;
            }
          }
          team_sync<WORKER>(bar_id);  // nb_s is read before the next time can overwrite it
#pragma unroll
          for (int q = 0; q < 8; ++q) tA[q] = tB[q] = tC[q] = (T)0;
          dirty_time = false;
        }
      }
    }

    // ---- flush the item: fp64 atomics, fftshift as an index rotation -------------
    const int64_t obase = (((int64_t)gout * P.nspw + s) * P.npol + p) * N;
    if constexpr (C::TACC) {
      uint32_t w16[16], pw[16];
      tm_wait_st();
      tm_ld_w16(tm_a + TC_S2, w16);
      tm_ld_w16(tm_a + TC_PW, pw);
      tm_wait_ld();
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        S2[r] = (T)TeamFftD256::dbl(f2_from_words(w16[2 * r], w16[2 * r + 1]));
        Pw[r] = (T)TeamFftD256::dbl(f2_from_words(pw[2 * r], pw[2 * r + 1]));
      }
      const uint32_t z16[16] = {};
      tm_st_w16(tm_a + TC_PW, z16);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int kshift = (tau_out + TPR * r + N / 2) % N;
      if (HAS_D) {
        T s2 = S2[r];
#pragma unroll
        for (int off = TPR; off < 32; off <<= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        if (sub == 0) {
          atomicAdd(P.pow_all + obase + kshift, (double)Pw[r]);
          atomicAdd(P.pow_auto + obase + kshift, (double)s2);
        }
      }
      if (HAS_N) {
        float s2 = Sn2[r];
#pragma unroll
        for (int off = TPR; off < 32; off <<= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        if (sub == 0) {
          atomicAdd(P.noise_all + obase + kshift, (double)Pn[r]);
          atomicAdd(P.noise_auto + obase + kshift, (double)s2);
        }
      }
    }
    if (HAS_T) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        th_all_acc += __shfl_xor_sync(0xffffffffu, th_all_acc, off);
        th_auto_acc += __shfl_xor_sync(0xffffffffu, th_auto_acc, off);
      }
      if ((wl & 31) == 0) {
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
    if (threadIdx.x < 32) tm_dealloc<C::TM_COLS>(tm_base_s);
  }
}

// workspace layout: [counter 256 B][item_start (ngroup+1) int32, padded][wl][wr]
struct EngWorkspace {
  int *counter;
  int32_t *item_start;
  double *wl, *wr;
  int64_t bytes;
};
static EngWorkspace carve(void *ws, int ngroup, int nspw, int npol, int N) {
  EngWorkspace w;
  unsigned char *b = (unsigned char *)ws;
  int64_t off = 0;
  w.counter = (int *)(b + off); off += 256;
  w.item_start = (int32_t *)(b + off); off += (((int64_t)(ngroup + 1) * 4 + 255) / 256) * 256;
  w.wl = (double *)(b + off); off += (((int64_t)nspw * npol * N * 8 + 255) / 256) * 256;
  w.wr = (double *)(b + off); off += (((int64_t)nspw * npol * N * 8 + 255) / 256) * 256;
  w.bytes = off;
  return w;
}

// launches enqueued by the sds_engine_run in progress on this host thread (reported through its
// launches_out argument; thread-local so that concurrent calls from several host threads do not mix)
static thread_local int g_last_launches = 0;

template <typename T, int N, int LEGS>
static int launch_engine(const EngineParams &P, cudaStream_t st) {
  using C = EngCfg<T, N, LEGS>;
  static DeviceLatch configured;
  int cfg_dev = 0;
  if (configured.need(&cfg_dev)) {
    SDS_CUDA(cudaFuncSetAttribute(engine_kernel<T, N, LEGS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    // several CTAs per SM only fit when the L1/shared split favours shared memory
    SDS_CUDA(cudaFuncSetAttribute(engine_kernel<T, N, LEGS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    configured.set(cfg_dev);
  }
  int dev = 0, nsm = 148, per_sm = 1;
  SDS_CUDA(cudaGetDevice(&dev));
  SDS_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  SDS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, engine_kernel<T, N, LEGS>,
                                                         C::THREADS, C::SMEM));
  if (per_sm < 1) {
    set_error("sds_engine_run: kernel does not fit an SM (smem %d B)", C::SMEM);
    return SDS_ENOTSUP;
  }
  // (the occupancy query answers 1 for a kernel that allocates tensor memory: sds_engine_f32.cu)
  if (C::TM) per_sm = tmem_ctas_per_sm(engine_kernel<T, N, LEGS>, C::THREADS, C::SMEM, C::TM_COLS);
  engine_kernel<T, N, LEGS><<<nsm * per_sm, C::THREADS, C::SMEM, st>>>(P);
  SDS_LAUNCH_CHECK("engine_kernel");
  ++g_last_launches;
  return SDS_OK;
}

// fp64 (register budget): one leg per launch; the noise leg (fp32 in every mode) is served
// by sds_engine_f32.cu
template <typename T, int N>
static int run_engine_n(EngineParams &P, int legs, cudaStream_t st) {
  switch (legs) {
    case 1: return launch_engine<T, N, 1>(P, st);
    case 4: return launch_engine<T, N, 4>(P, st);
    default: break;
  }
  set_error("sds_engine_run: leg mask %d is not instantiated", legs);
  return SDS_ENOTSUP;
}

template <typename T>
static int run_engine(int N, EngineParams &P, int legs, cudaStream_t st) {
  switch (N) {
    case 64: return run_engine_n<T, 64>(P, legs, st);
    case 128: return run_engine_n<T, 128>(P, legs, st);
    case 256: return run_engine_n<T, 256>(P, legs, st);
    case 512: return run_engine_n<T, 512>(P, legs, st);
    case 1024: return run_engine_n<T, 1024>(P, legs, st);
    case 2048: return run_engine_n<T, 2048>(P, legs, st);
    default:
      set_error("sds_engine_run: fused path needs nfreq in {64..2048, power of two}, got %d", N);
      return SDS_ENOTSUP;
  }
}

}  // namespace sds

extern "C" int64_t sds_engine_workspace_bytes(const sds_plan *plan, const sds_engine_desc *d) {
  if (!plan || !d) return SDS_EINVAL;
  return sds::carve(nullptr, d->ngroup, d->nspw, d->npol, plan->p.nfreq).bytes;
}

extern "C" int sds_engine_run(const sds_plan *plan, const sds_engine_desc *d, const void *vis,
                              const uint8_t *flags, const float *nsample, const float *int_time,
                              const int32_t *group_offset, const float *sigma_coef,
                              const double *thermal_coef, const double *freq,
                              const double *pol_factor, const void *noise_in, double *pow_all,
                              double *pow_auto, double *noise_all, double *noise_auto,
                              double *thermal_all, double *thermal_auto, void *workspace,
                              sds_stream stream, int32_t *launches_out) {
  using namespace sds;
  g_last_launches = 0;
  struct Report {  // whatever way the call ends, the caller learns what was enqueued
    int32_t *out;
    ~Report() { if (out) *out = g_last_launches; }
  } report{launches_out};
  SDS_CHECK_ARG(plan && d, "sds_engine_run: NULL plan/desc");
  const Plan &pl = plan->p;
  const int N = pl.nfreq;
  SDS_CHECK_ARG(vis && flags && nsample && group_offset && workspace && pow_all && pow_auto,
                "sds_engine_run: NULL vis/flags/nsample/group_offset/workspace/pow outputs");
  SDS_CHECK_ARG(d->npol >= 1 && d->nbl >= 1 && d->ntime >= 1 && d->ngroup >= 1 && d->nspw >= 1 &&
                    d->nspw <= SDS_MAX_WINDOWS && d->ngroup <= 65535,
                "sds_engine_run: bad dimensions");
  SDS_CHECK_ARG(d->noise_mode >= 0 && d->noise_mode <= 3, "sds_engine_run: noise_mode not in 0..3");
  SDS_CHECK_ARG(d->noise_mode == 0 || (int_time && sigma_coef && noise_all && noise_auto),
                "sds_engine_run: noise needs int_time, sigma_coef and the noise outputs");
  SDS_CHECK_ARG(d->noise_mode != 2 || noise_in, "sds_engine_run: noise_mode 2 needs noise_in");
  const bool thermal = thermal_coef != nullptr;
  SDS_CHECK_ARG(!thermal || (freq && pol_factor && int_time && thermal_all && thermal_auto),
                "sds_engine_run: thermal needs freq, pol_factor, int_time and the thermal outputs");
  SDS_CHECK_ARG(d->nuv == 1 || d->nuv == 2, "sds_engine_run: nuv must be 1 or 2 (got %d)", d->nuv);
  if (d->nuv == 2 && (pl.math != SDS_F32 || !pl.is_pow2 || N < 64 || N > 1024 || d->noise_mode == 3 ||
                      (thermal && d->noise_mode == 0))) {
    set_error("sds_engine_run: two data sets (nuv = 2) take the fused path in complex64 with power-of-two "
              "windows of 64..1024 channels and noise_mode 1 or 2 (0 without the thermal leg, which rides "
              "in the noise launch)");
    return SDS_ENOTSUP;
  }
  const bool pow2_path = pl.is_pow2 && N >= 64 && N <= 2048 && pl.team_tw_f;
  // other window lengths: chirp-z engine (complex64 only)
  const bool blue_path = !pl.is_pow2 && pl.blue_m != 0 && pl.math == SDS_F32;
  if (!pow2_path && !blue_path) {
    set_error("sds_engine_run: fused path needs a power-of-two nfreq in 64..2048, or another "
              "length in 8..1024 with complex64 arithmetic; got nfreq %d", N);
    return SDS_ENOTSUP;
  }
  const int npad = blue_path ? (N + 15) & ~15 : N;  // channels one staged row piece covers
  // 1-D bulk copies need 16-byte aligned sources for every array (flags are 1 B/sample)
  if (d->nchan % 16 != 0 || ((uintptr_t)vis | (uintptr_t)flags | (uintptr_t)nsample) % 16 != 0) {
    set_error("sds_engine_run: fused path needs nchan %% 16 == 0 and 16-byte aligned inputs");
    return SDS_ENOTSUP;
  }
  for (int w = 0; w < d->nspw; ++w) {
    SDS_CHECK_ARG(d->win_start[w] >= 0 && d->win_start[w] + N <= d->nchan,
                  "sds_engine_run: window %d [%d, %d) outside the %d channels", w, d->win_start[w],
                  d->win_start[w] + N, d->nchan);
    if (d->win_start[w] + npad > d->nchan) {
      set_error("sds_engine_run: fused path stages %d channels per window; window %d does not "
                "leave that many before the end of the row", npad, w);
      return SDS_ENOTSUP;
    }
    if (d->win_start[w] % 16 != 0) {
      set_error("sds_engine_run: fused path needs window starts that are multiples of 16");
      return SDS_ENOTSUP;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  EngWorkspace ws = carve(workspace, d->ngroup, d->nspw, d->npol, N);
  EngineParams P;
  P.vis = (const float2 *)vis; P.flags = flags; P.nsample = nsample; P.int_time = int_time;
  P.goff = group_offset; P.group_id = d->group_index; P.group_bl0 = d->group_bl_start; P.sigma_coef = sigma_coef; P.noise_in = (const float2 *)noise_in;
  P.pol_factor = pol_factor;
  P.taper_f = pl.taper_f; P.tw_f = (const float2 *)pl.team_tw_f;
  if (pl.math == SDS_F64) { P.taper = pl.taper_d; P.tw = pl.team_tw_d; }
  else { P.taper = pl.taper_f; P.tw = pl.team_tw_f; }
  P.item_start = ws.item_start; P.wl = ws.wl; P.wr = ws.wr; P.counter = ws.counter;
  P.pow_all = pow_all; P.pow_auto = pow_auto; P.noise_all = noise_all; P.noise_auto = noise_auto;
  P.th_all = thermal_all; P.th_auto = thermal_auto;
  P.time_offset = d->time_offset; P.ntime_total = d->ntime_total > 0 ? d->ntime_total : d->ntime;
  P.bl_offset = d->bl_offset; P.nbl_total = d->nbl_total > 0 ? d->nbl_total : d->nbl;
  P.delta_f = d->delta_f; P.seed_lo = (uint32_t)d->seed; P.seed_hi = (uint32_t)(d->seed >> 32);
  philox_fill_keys(P.seed_lo, P.seed_hi, P.pkey);
  P.npol = d->npol; P.nbl = d->nbl; P.ntime = d->ntime; P.nchan = d->nchan; P.nspw = d->nspw;
  P.ngroup = d->ngroup;
  // in-kernel draws: mode 1 = moment-matched discrete law (power-of-two engine; the chirp-z engine
  // draws Box-Muller in both modes), mode 3 = Box-Muller per channel
  P.noise_mode = d->noise_mode == 3 ? 1 : d->noise_mode;
  P.noise_gauss = d->noise_mode == 3;
  P.zero = 0;
  P.target_rows = 64;
  for (int i = 0; i < SDS_MAX_WINDOWS; ++i) P.win_start[i] = i < d->nspw ? d->win_start[i] : 0;
  P.nfreq = N;
  P.blue_chirp = (const float2 *)pl.blue_chirp_f;
  P.blue_bhat = (const float2 *)pl.blue_bhat_f;
  P.blue_tw = (const float2 *)pl.blue_tw_f;

  const int legs = LEG_DATA | (d->noise_mode ? LEG_NOISE : 0) | (thermal ? LEG_THERMAL : 0);
  int rc = SDS_OK;
  // one setup launch per engine launch (it also rewinds the item counter)
  auto setup = [&]() -> int {
    engine_setup_kernel<<<8, 1024, 0, st>>>(group_offset, d->ngroup, d->ntime, d->nspw, d->npol, N,
                                            P.target_rows, thermal_coef, freq, ws.item_start,
                                            ws.wl, ws.wr, ws.counter);
    SDS_LAUNCH_CHECK("engine_setup_kernel");
    ++g_last_launches;
    return SDS_OK;
  };
  if (d->nuv == 2) {  // one launch per leg, the two data sets' rows in lockstep (sds_engine_uv2.cu)
    if ((rc = setup()) != SDS_OK) return rc;
    if ((rc = run_engine_uv2(N, P, LEG_DATA, st, &g_last_launches)) != SDS_OK) return rc;
    if (legs & LEG_NOISE) {
      if ((rc = setup()) != SDS_OK) return rc;
      rc = run_engine_uv2(N, P, LEG_NOISE | (legs & LEG_THERMAL), st, &g_last_launches);
    }
    return rc;
  }
  if (pl.math == SDS_F32) {  // one pass over the inputs for every leg
    if ((rc = setup()) != SDS_OK) return rc;
    if (blue_path) return run_engine_blue(pl.blue_m, P, legs, st, &g_last_launches);
    return run_engine_f32(N, P, legs, st, &g_last_launches);
  }
  // complex128: the data leg in fp64 (sds_engine.cu), then ONE launch of the packed engine for the
  // noise realisation (fp32 in every mode) and the thermal sums in fp64 -- it stages nsample and the
  // flags anyway and its fp64 pipe is idle: 9 + 5 = 14 bytes per sample in two launches (the first
  // version took three: 9 + 5 + 4).  Thermal without noise keeps the fp64 thermal-only kernel.
  if (legs & LEG_DATA) {
    if ((rc = setup()) != SDS_OK) return rc;
    if ((rc = run_engine<double>(N, P, LEG_DATA, st)) != SDS_OK) return rc;
  }
  if (legs & LEG_NOISE) {
    if ((rc = setup()) != SDS_OK) return rc;
    EngineParams Pn = P;
    Pn.taper = pl.taper_f;
    Pn.tw = pl.team_tw_f;
    rc = run_engine_f32(N, Pn, LEG_NOISE | ((legs & LEG_THERMAL) ? LEG_THERMAL64 : 0), st, &g_last_launches);
  } else if (legs & LEG_THERMAL) {
    if ((rc = setup()) != SDS_OK) return rc;
    rc = run_engine<double>(N, P, LEG_THERMAL, st);
  }
  return rc;
}


// ---- means over pairs and times from the accumulated sums (sds.h: sds_engine_means) ---------------
namespace sds {
__global__ void engine_means_kernel(int rows_per_group, int row_len, const int32_t *__restrict__ group_size,
                                    const double *__restrict__ times_done, double ntimes,
                                    const double *__restrict__ pow_all, const double *__restrict__ pow_auto,
                                    const double *__restrict__ noise_all, const double *__restrict__ noise_auto,
                                    const double *__restrict__ th_all, const double *__restrict__ th_auto,
                                    double *__restrict__ o_pa, double *__restrict__ o_pn, double *__restrict__ o_na,
                                    double *__restrict__ o_nn, double *__restrict__ o_ta, double *__restrict__ o_tn) {
  const int64_t row = blockIdx.x;  // (group, spw, pol)
  const int g = (int)(row / rows_per_group);
  const double n = (double)group_size[g];
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  double T = ntimes > 0.0 ? ntimes : times_done[g];
  if (!(T > 0.0)) T = nan;
  const double f_all = 1.0 / (n * n * T);
  const double f_no = n > 1.0 ? 1.0 / (n * (n - 1.0) * T) : nan;
  for (int k = threadIdx.x; k < row_len; k += blockDim.x) {
    const int64_t e = row * row_len + k;
    if (pow_all) {
      const double a = pow_all[e];
      o_pa[e] = a * f_all;
      o_pn[e] = n > 1.0 ? (a - pow_auto[e]) * f_no : nan;
    }
    if (noise_all) {
      const double a = noise_all[e];
      o_na[e] = a * f_all;
      o_nn[e] = n > 1.0 ? (a - noise_auto[e]) * f_no : nan;
    }
  }
  if (th_all && threadIdx.x == 0) {
    const double a = th_all[row];
    o_ta[row] = a * f_all;
    o_tn[row] = n > 1.0 ? (a - th_auto[row]) * f_no : nan;
  }
}
}  // namespace sds

extern "C" int sds_engine_means(int32_t ngroup, int32_t rows_per_group, int32_t row_len, const int32_t *group_size,
                                const double *times_done, double ntimes, const double *pow_all,
                                const double *pow_auto, const double *noise_all, const double *noise_auto,
                                const double *thermal_all, const double *thermal_auto, double *power_all,
                                double *power_noautos, double *noise_power_all, double *noise_power_noautos,
                                double *thermal_mean_all, double *thermal_mean_noautos, sds_stream stream) {
  using namespace sds;
  if (ngroup < 0 || rows_per_group <= 0 || row_len <= 0 || !group_size || (!times_done && !(ntimes > 0.0))) {
    set_error("sds_engine_means: bad arguments");
    return SDS_EINVAL;
  }
  if ((pow_all && (!pow_auto || !power_all || !power_noautos)) ||
      (noise_all && (!noise_auto || !noise_power_all || !noise_power_noautos)) ||
      (thermal_all && (!thermal_auto || !thermal_mean_all || !thermal_mean_noautos))) {
    set_error("sds_engine_means: a sum needs its auto sum and both outputs");
    return SDS_EINVAL;
  }
  if (ngroup == 0) return SDS_OK;
  const int64_t rows = (int64_t)ngroup * rows_per_group;
  engine_means_kernel<<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(
      rows_per_group, row_len, group_size, times_done, ntimes, pow_all, pow_auto, noise_all, noise_auto, thermal_all,
      thermal_auto, power_all, power_noautos, noise_power_all, noise_power_noautos, thermal_mean_all,
      thermal_mean_noautos);
  SDS_LAUNCH_CHECK("engine_means_kernel");
  return SDS_OK;
}
