// Plan management and the error channel of libsds.so.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <utility>
#include <vector>

#include "sds_radix.cuh"

namespace sds {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *where) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
  return (int)e;
}

}  // namespace sds

extern "C" const char *sds_last_error(void) { return sds::g_err; }
extern "C" int sds_version(void) { return SDS_VERSION; }

extern "C" int sds_plan_create(sds_plan **plan, int nfreq, int math, const double *taper_host) {
  SDS_CHECK_ARG(plan != nullptr, "sds_plan_create: plan is NULL");
  *plan = nullptr;
  SDS_CHECK_ARG(nfreq >= 1 && nfreq <= SDS_MAX_NFREQ,
                "sds_plan_create: nfreq=%d outside [1, %d]", nfreq, SDS_MAX_NFREQ);
  SDS_CHECK_ARG(math == SDS_F32 || math == SDS_F64,
                "sds_plan_create: math must be SDS_F32 or SDS_F64, got %d", math);
  SDS_CHECK_ARG(taper_host != nullptr, "sds_plan_create: taper is NULL");

  sds_plan *pl = (sds_plan *)calloc(1, sizeof(sds_plan));
  if (!pl) {
    sds::set_error("sds_plan_create: out of host memory");
    return SDS_ENOMEM;
  }
  sds::Plan &p = pl->p;
  p.nfreq = nfreq;
  p.math = math;
  p.is_pow2 = (nfreq & (nfreq - 1)) == 0;

  // factorise: 4s first, then 2, then odd primes (a large prime factor becomes
  // one O(r^2) stage -- correct for any length).
  int n = nfreq, nr = 0;
  while (n % 4 == 0 && nr < 15) { p.radix[nr++] = 4; n /= 4; }
  while (n % 2 == 0 && nr < 15) { p.radix[nr++] = 2; n /= 2; }
  for (int f = 3; (int64_t)f * f <= n && nr < 15; f += 2)
    while (n % f == 0 && nr < 15) { p.radix[nr++] = f; n /= f; }
  if (n > 1) p.radix[nr++] = n;
  p.nradix = nr;

  std::vector<double> tw(2 * (size_t)nfreq);
  std::vector<float> twf(2 * (size_t)nfreq), tapf(nfreq);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int k = 0; k < nfreq; ++k) {
    long double a = -two_pi * (long double)k / (long double)nfreq;
    tw[2 * k] = (double)cosl(a);
    tw[2 * k + 1] = (double)sinl(a);
    twf[2 * k] = (float)tw[2 * k];
    twf[2 * k + 1] = (float)tw[2 * k + 1];
    tapf[k] = (float)taper_host[k];
  }
  cudaError_t e;
#define PLAN_TRY(call)                      \
  if ((e = (call)) != cudaSuccess) {        \
    sds_plan_destroy(pl);                   \
    return sds::cuda_fail(e, #call);        \
  }
  PLAN_TRY(cudaMalloc(&p.taper_d, sizeof(double) * nfreq));
  PLAN_TRY(cudaMalloc(&p.taper_f, sizeof(float) * nfreq));
  PLAN_TRY(cudaMalloc(&p.twiddle_d, sizeof(double2) * nfreq));
  PLAN_TRY(cudaMalloc(&p.twiddle_f, sizeof(float2) * nfreq));
  PLAN_TRY(cudaMemcpy(p.taper_d, taper_host, sizeof(double) * nfreq, cudaMemcpyHostToDevice));
  PLAN_TRY(cudaMemcpy(p.taper_f, tapf.data(), sizeof(float) * nfreq, cudaMemcpyHostToDevice));
  PLAN_TRY(cudaMemcpy(p.twiddle_d, tw.data(), sizeof(double2) * nfreq, cudaMemcpyHostToDevice));
  PLAN_TRY(cudaMemcpy(p.twiddle_f, twf.data(), sizeof(float2) * nfreq, cudaMemcpyHostToDevice));
  if (p.is_pow2 && nfreq >= 64) {
    const int nt = sds::fft_tw_total(nfreq);
    std::vector<sds::cpx<float>> tf(nt);
    std::vector<sds::cpx<double>> td(nt);
    sds::fill_team_twiddles<float>(nfreq, tf.data());
    sds::fill_team_twiddles<double>(nfreq, td.data());
    PLAN_TRY(cudaMalloc(&p.team_tw_f, sizeof(sds::cpx<float>) * nt));
    PLAN_TRY(cudaMalloc(&p.team_tw_d, sizeof(sds::cpx<double>) * nt));
    PLAN_TRY(cudaMemcpy(p.team_tw_f, tf.data(), sizeof(sds::cpx<float>) * nt, cudaMemcpyHostToDevice));
    PLAN_TRY(cudaMemcpy(p.team_tw_d, td.data(), sizeof(sds::cpx<double>) * nt, cudaMemcpyHostToDevice));
  }
  if (!p.is_pow2 && nfreq >= 8 && nfreq <= 1024) {
    // Bluestein: X[k] = c[k] sum_n (x[n] c[n]) b[k - n], c[n] = exp(-i pi n^2 / N), b = conj(c)
    int M = 64;
    while (M < 2 * nfreq - 1) M <<= 1;
    const long double pi = 3.141592653589793238462643383279502884L;
    std::vector<sds::cpx<double>> chirp(nfreq), bhat(M);
    std::vector<std::pair<long double, long double>> b(M, {0.0L, 0.0L}), wm(M);
    for (int n = 0; n < nfreq; ++n) {
      const long long q = ((long long)n * n) % (2LL * nfreq);  // n^2 mod 2N keeps the angle exact
      const long double a = pi * (long double)q / (long double)nfreq;
      chirp[n] = {(double)cosl(a), (double)-sinl(a)};
      b[n] = {cosl(a), sinl(a)};
      if (n) b[M - n] = b[n];
    }
    for (int k = 0; k < M; ++k) {
      const long double a = -2.0L * pi * (long double)k / (long double)M;
      wm[k] = {cosl(a), sinl(a)};
    }
    for (int k = 0; k < M; ++k) {  // plain O(M^2) transform in long double, once per plan
      long double sr = 0.0L, si = 0.0L;
      int idx = 0;
      for (int m = 0; m < M; ++m) {
        sr += b[m].first * wm[idx].first - b[m].second * wm[idx].second;
        si += b[m].first * wm[idx].second + b[m].second * wm[idx].first;
        idx += k;
        if (idx >= M) idx -= M;
      }
      bhat[k] = {(double)sr, (double)si};
    }
    std::vector<sds::cpx<float>> chirpf(nfreq), bhatf(M);
    for (int n = 0; n < nfreq; ++n) chirpf[n] = {(float)chirp[n].x, (float)chirp[n].y};
    for (int k = 0; k < M; ++k) bhatf[k] = {(float)bhat[k].x, (float)bhat[k].y};
    const int nt = sds::fft_tw_total(M);
    std::vector<sds::cpx<float>> tf(nt);
    std::vector<sds::cpx<double>> td(nt);
    sds::fill_team_twiddles<float>(M, tf.data());
    sds::fill_team_twiddles<double>(M, td.data());
    PLAN_TRY(cudaMalloc(&p.blue_chirp_f, sizeof(sds::cpx<float>) * nfreq));
    PLAN_TRY(cudaMalloc(&p.blue_chirp_d, sizeof(sds::cpx<double>) * nfreq));
    PLAN_TRY(cudaMalloc(&p.blue_bhat_f, sizeof(sds::cpx<float>) * M));
    PLAN_TRY(cudaMalloc(&p.blue_bhat_d, sizeof(sds::cpx<double>) * M));
    PLAN_TRY(cudaMalloc(&p.blue_tw_f, sizeof(sds::cpx<float>) * nt));
    PLAN_TRY(cudaMalloc(&p.blue_tw_d, sizeof(sds::cpx<double>) * nt));
    PLAN_TRY(cudaMemcpy(p.blue_chirp_f, chirpf.data(), sizeof(sds::cpx<float>) * nfreq, cudaMemcpyHostToDevice));
    PLAN_TRY(cudaMemcpy(p.blue_chirp_d, chirp.data(), sizeof(sds::cpx<double>) * nfreq, cudaMemcpyHostToDevice));
    PLAN_TRY(cudaMemcpy(p.blue_bhat_f, bhatf.data(), sizeof(sds::cpx<float>) * M, cudaMemcpyHostToDevice));
    PLAN_TRY(cudaMemcpy(p.blue_bhat_d, bhat.data(), sizeof(sds::cpx<double>) * M, cudaMemcpyHostToDevice));
    PLAN_TRY(cudaMemcpy(p.blue_tw_f, tf.data(), sizeof(sds::cpx<float>) * nt, cudaMemcpyHostToDevice));
    PLAN_TRY(cudaMemcpy(p.blue_tw_d, td.data(), sizeof(sds::cpx<double>) * nt, cudaMemcpyHostToDevice));
    p.blue_m = M;
  }
#undef PLAN_TRY
  *plan = pl;
  return SDS_OK;
}

extern "C" int sds_plan_destroy(sds_plan *plan) {
  if (!plan) return SDS_OK;
  cudaFree(plan->p.taper_d);
  cudaFree(plan->p.taper_f);
  cudaFree(plan->p.twiddle_d);
  cudaFree(plan->p.twiddle_f);
  cudaFree(plan->p.team_tw_f);
  cudaFree(plan->p.team_tw_d);
  cudaFree(plan->p.blue_chirp_f);
  cudaFree(plan->p.blue_chirp_d);
  cudaFree(plan->p.blue_bhat_f);
  cudaFree(plan->p.blue_bhat_d);
  cudaFree(plan->p.blue_tw_f);
  cudaFree(plan->p.blue_tw_d);
  free(plan);
  return SDS_OK;
}

extern "C" int sds_plan_nfreq(const sds_plan *plan) { return plan ? plan->p.nfreq : SDS_EINVAL; }
extern "C" int sds_plan_math(const sds_plan *plan) { return plan ? plan->p.math : SDS_EINVAL; }

// ---- noise leg of the unfused API in one call ---------------------------
// sigma (ds.py:3544-3583) -> realisation (utils.py:486-506) -> delay transform with the flag
// mask (ds.py:3503-3517, utils.py:509-566): the three kernels above enqueued back to back on
// `stream`; nothing is allocated (the caller passes the sigma scratch and receives the
// frequency-domain realisation, which the reference keeps as noise_array until delay_transform).
extern "C" int sds_noise_transform(const sds_plan *plan, const double *coef, const double *int_time,
                                   const void *nsample, int nsample_dtype, int S, int U, int P, int B,
                                   int T, int F, const uint8_t *flags, double delta_f, uint64_t seed,
                                   uint64_t offset, const double *g_re, const double *g_im,
                                   double *sigma_scratch, void *noise_freq, void *noise_delay,
                                   int dtype, sds_stream stream) {
  SDS_CHECK_ARG(plan && sigma_scratch && noise_freq && noise_delay, "sds_noise_transform: NULL pointer");
  SDS_CHECK_ARG(F == plan->p.nfreq, "sds_noise_transform: nfreq %d does not match the plan (%d)", F,
                plan->p.nfreq);
  const int64_t rows = (int64_t)S * U * P * B * T;
  int rc = sds_noise_sigma(coef, int_time, nsample, nsample_dtype, S, U, P, B, T, F, sigma_scratch, stream);
  if (rc != SDS_OK) return rc;
  rc = sds_generate_noise(sigma_scratch, rows * F, g_re, g_im, seed, offset, noise_freq, dtype, stream);
  if (rc != SDS_OK) return rc;
  const int32_t start = 0;
  return sds_delay_transform(plan, noise_freq, dtype, F, flags, F, rows, 1, &start, delta_f, 0,
                             noise_delay, dtype, stream);
}
