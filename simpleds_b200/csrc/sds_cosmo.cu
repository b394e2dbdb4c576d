// Cosmological normalisation of the delay power spectrum: the flat-LambdaCDM
// background the reference takes from astropy (cosmo.py:20 sets Planck15) -- E(z) with photons
// and massive neutrinos, the comoving distance integral -- and the per-(window, polarisation)
// conversion integrals of DelaySpectrum.update_cosmology (delay_spectrum.py:1246-1334).
// Everything is fp64; the inputs are the (Nspws, Nfreqs) frequency table and the
// (Nspws, Npols, Nfreqs) beam table, a few kilobytes, so the kernels are latency-bound by design:
// one thread per redshift for the distance (fixed 48-panel x 16-point Gauss-Legendre, error
// < 1e-14 for z <= 30), one warp per (window, polarisation) for the trapezoid integrals.
#include "sds_common.cuh"

namespace sds {

struct CosmoDev {
  double om0, ode0, ogamma0, neff_per_nu, nmassless, nu_y[SDS_COSMO_MAX_NU];
  int nmassive, has_nu;
  double hubble_distance_mpc;
};

// astropy FLRW.nu_relative_density (Komatsu et al. 2011, eq. 26 fit: p = 1.83, k = 0.3173)
__device__ __forceinline__ double nu_relative_density(const CosmoDev &c, double z) {
  const double prefac = 0.22710731766;  // 7/8 (4/11)^(4/3)
  if (!c.has_nu) return 0.0;
  if (c.nmassive == 0) return prefac * c.neff_per_nu * c.nmassless;
  const double p = 1.83, invp = 0.54644808743, k = 0.3173;
  double rel_mass = c.nmassless;
  for (int i = 0; i < c.nmassive; ++i)
    rel_mass += pow(1.0 + pow(k * c.nu_y[i] / (1.0 + z), p), invp);
  return prefac * c.neff_per_nu * rel_mass;
}
// astropy FlatLambdaCDM.efunc: sqrt((1+z)^3 (Or (1+z) + Om0) + Ode0), Or = Ogamma0 (1 + nu density)
__device__ __forceinline__ double efunc(const CosmoDev &c, double z) {
  const double zp1 = 1.0 + z;
  const double orad = c.ogamma0 * (1.0 + nu_relative_density(c, z));
  return sqrt(zp1 * zp1 * zp1 * (orad * zp1 + c.om0) + c.ode0);
}

__constant__ double GL16_X[8] = {0.0950125098376374401853193, 0.2816035507792589132304605,
                                 0.4580167776572273863424194, 0.6178762444026437484466718,
                                 0.7554044083550030338951012, 0.8656312023878317438804679,
                                 0.9445750230732325760779884, 0.9894009349916499325961542};
__constant__ double GL16_W[8] = {0.1894506104550684962853967, 0.1826034150449235888667637,
                                 0.1691565193950025381893121, 0.1495959888165767320815017,
                                 0.1246289712555338720524763, 0.0951585116824927848099251,
                                 0.0622535239386478928628438, 0.0271524594117540948517806};

__global__ void cosmo_kernel(CosmoDev c, const double *__restrict__ z, int64_t n,
                             double *__restrict__ e_out, double *__restrict__ dc_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double zi = z[i];
    if (e_out) e_out[i] = efunc(c, zi);
    if (dc_out) {
      const int NP = 48;
      const double h = zi / NP;
      double acc = 0.0;
      for (int p = 0; p < NP; ++p) {
        const double mid = (p + 0.5) * h, half = 0.5 * h;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          s += GL16_W[k] * (1.0 / efunc(c, mid - half * GL16_X[k]) + 1.0 / efunc(c, mid + half * GL16_X[k]));
        acc += s * half;
      }
      dc_out[i] = c.hubble_distance_mpc * acc;
    }
  }
}

// out[s,p] = trapz_f( freq^pw / x2y * taper^2 * beam_sq )   (delay_spectrum.py:1260-1268, 1286-1297)
__global__ void conversion_integral_kernel(const double *__restrict__ freq, const double *__restrict__ x2y,
                                           const double *__restrict__ taper, const double *__restrict__ beam_sq,
                                           int nspw, int npol, int nfreq, int freq_power,
                                           double *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= nspw * npol) return;
  const int s = warp / npol;
  const double *f = freq + (int64_t)s * nfreq, *x = x2y + (int64_t)s * nfreq;
  const double *b = beam_sq + (int64_t)warp * nfreq;
  auto g = [&](int k) {
    const double fk = f[k], f2 = fk * fk;
    const double fp = freq_power == 4 ? f2 * f2 : 1.0;
    return fp / x[k] * (taper[k] * taper[k]) * b[k];
  };
  double acc = 0.0;
  for (int k = lane; k + 1 < nfreq; k += 32) acc += 0.5 * (f[k + 1] - f[k]) * (g(k) + g(k + 1));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) out[warp] = acc;
}

static CosmoDev to_dev(const sds_cosmo_params *p) {
  CosmoDev c;
  c.om0 = p->om0; c.ode0 = p->ode0; c.ogamma0 = p->ogamma0; c.neff_per_nu = p->neff_per_nu;
  c.nmassless = p->nmassless_nu; c.nmassive = p->nmassive_nu; c.has_nu = p->has_massive_nu_or_radiation;
  for (int i = 0; i < SDS_COSMO_MAX_NU; ++i) c.nu_y[i] = p->nu_y[i];
  c.hubble_distance_mpc = 299792.458 / p->h0_km_s_mpc;
  return c;
}

}  // namespace sds

using namespace sds;

extern "C" int sds_cosmo_evaluate(const sds_cosmo_params *params, const double *z, int64_t n,
                                  double *efunc_out, double *comoving_distance_mpc_out,
                                  sds_stream stream) {
  SDS_CHECK_ARG(params && z, "sds_cosmo_evaluate: null pointer");
  SDS_CHECK_ARG(n >= 0 && params->nmassive_nu >= 0 && params->nmassive_nu <= SDS_COSMO_MAX_NU,
                "sds_cosmo_evaluate: bad size");
  SDS_CHECK_ARG(params->h0_km_s_mpc > 0.0, "sds_cosmo_evaluate: H0 must be positive");
  if (n == 0) return SDS_OK;
  const int block = 128;
  int64_t grid = (n + block - 1) / block;
  if (grid > 148 * 8) grid = 148 * 8;
  cosmo_kernel<<<(int)grid, block, 0, (cudaStream_t)stream>>>(to_dev(params), z, n, efunc_out,
                                                              comoving_distance_mpc_out);
  SDS_LAUNCH_CHECK("sds_cosmo_evaluate");
  return SDS_OK;
}

extern "C" int sds_conversion_integral(const double *freq, const double *x2y, const double *taper,
                                       const double *beam_sq_area, int nspw, int npol, int nfreq,
                                       int freq_power, double *out, sds_stream stream) {
  SDS_CHECK_ARG(freq && x2y && taper && beam_sq_area && out, "sds_conversion_integral: null pointer");
  SDS_CHECK_ARG(nspw >= 1 && npol >= 1 && nfreq >= 2, "sds_conversion_integral: bad dims");
  SDS_CHECK_ARG(freq_power == 0 || freq_power == 4, "sds_conversion_integral: freq_power must be 0 or 4");
  const int warps = nspw * npol, block = 128;
  conversion_integral_kernel<<<(warps * 32 + block - 1) / block, block, 0, (cudaStream_t)stream>>>(
      freq, x2y, taper, beam_sq_area, nspw, npol, nfreq, freq_power, out);
  SDS_LAUNCH_CHECK("sds_conversion_integral");
  return SDS_OK;
}
