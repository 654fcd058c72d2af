// Run-time-size entry points (include/dynare_b200.h, "models without a generated device function"): a handle built
// from index sets alone, the solve / likelihood from host-evaluated Jacobians, the batched doubling Lyapunov
// solver and the generic dense Kalman filter.
#include <cstdlib>
#include <algorithm>

#include <atomic>
#include "internal.h"
#include "kalman_univariate.cuh"
#include "kalman_generic.cuh"
#include "kalman_dmma.cuh"
#include "kalman_dmma_fold.cuh"
#include "kalman_dmma_large.cuh"
#include "kalman_lanes.cuh"
#include "solve_generic.cuh"
#include "solve_block.cuh"
#include "kalman_smoother.cuh"
#include "kalman_diffuse.cuh"

namespace dyb {

struct GenericHost {
  GenericModel gm{};           // sizes + device pointers into d_idx
  DevBuf d_idx;
  std::vector<int> obs_state;  // ny: Kalman state observed by each row of Y
  size_t smem = 0;             // bytes of shared memory per draw
  bool warp_groups = false;    // one warp per draw (small models) instead of one CTA per draw
  bool block = false;          // medium models: register-blocked CTA-per-draw kernel (solve_block.cuh)
  SbLayout sbl{};              // its shared-memory layout
  bool block_small = false;    // ... and the narrower instantiation fits (m + k + nf <= 80, m + nx <= 80, k + nf <= 48)
  DevBuf wscr;                 // its global scratch: the working Jacobian after the static elimination, SB_GRID CTAs
  DevBuf bT, bQ, bP, byo, bH, bst, bz;   // persistent work buffers of dyb_loglik_from_jacobian_batch
  DevBuf bJ, byb, bse, bsm, bsi;         // persistent staging of its host inputs
};

void generic_release(dyb_handle* h) {
  if (h && h->generic) {
    h->generic->d_idx.release();
    h->generic->wscr.release();
    DevBuf* all[] = {&h->generic->bT, &h->generic->bQ, &h->generic->bP, &h->generic->byo, &h->generic->bH, &h->generic->bst,
                     &h->generic->bz, &h->generic->bJ, &h->generic->byb, &h->generic->bse, &h->generic->bsm, &h->generic->bsi};
    for (DevBuf* b : all) b->release();
    delete h->generic;
    h->generic = nullptr;
  }
}

// launch the dense Kalman filter on device-resident state spaces (shared by dyb_kalman_batch and the
// from-Jacobian likelihood); scratch (2 ns^2 doubles per CTA) only when P and T P do not fit in shared memory
// PROCESS-WIDE A/B switches set by dyb_set_kalman_variant (every handle, device and thread reads them; atomics so that
// a concurrent reader sees a consistent value -- a test that toggles them changes the kernel other handles use)
static std::atomic<int> g_kalman_variant{0};      // 0 auto, 1 scalar (kalman_generic.cuh), 2 tensor-core (kalman_dmma*.cuh)
static std::atomic<int> g_kalman_fold{1};         // ns <= 88: correction folded behind the products + scalar warp (kalman_dmma_fold.cuh)
static std::atomic<int> g_kalman_lanes_max{16};   // auto choice: sub-warp kernel up to this many states
static std::atomic<int> g_kalman_bulk{1};         // large-ns kernel: stage the GEMM panels with cp.async.bulk when alignment allows

static cudaError_t launch_kalman_generic(KalmanGenericArgs g, int device, cudaStream_t st, Stager& sg) {
  int dev_sms = 148, max_smem = 0;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  // small state spaces: four (ns <= 12) or eight lanes per draw (kalman_lanes.cuh); variant 4 forces it, variants 1-3 bypass it
  if (g.ns <= 16 && g.ny <= KLN_NY && (g_kalman_variant == 4 || (g_kalman_variant == 0 && g.ns <= g_kalman_lanes_max))) {
    g.scratch = nullptr;
    auto launch = [&](auto kern, size_t smem, int draws_per_cta) -> cudaError_t {
      if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
      }
      long long blocks = (g.n_draws + draws_per_cta - 1) / draws_per_cta;
      const long long cap = (long long)dev_sms * 64;
      if (blocks > cap) blocks = cap;
      kern<<<(unsigned)blocks, KLN_THREADS, smem, st>>>(g);
      return cudaGetLastError();
    };
    if (g.ns <= 8) return launch(kalman_lanes_kernel<8, 4>, KalmanLanesCfg<8, 4>::SMEM_BYTES, KalmanLanesCfg<8, 4>::DRAWS);
    if (g.ns <= 12) return launch(kalman_lanes_kernel<12, 4>, KalmanLanesCfg<12, 4>::SMEM_BYTES, KalmanLanesCfg<12, 4>::DRAWS);
    return launch(kalman_lanes_kernel<16, 8>, KalmanLanesCfg<16, 8>::SMEM_BYTES, KalmanLanesCfg<16, 8>::DRAWS);
  }
  {
    const KalmanDmmaGeom q = kalman_dmma_geom(g.ns, g.ny);
    const size_t smem = sizeof(double) * kalman_dmma_smem_doubles(g.ns, g.ny);
    const bool fits = q.nsp / 8 <= DMMA_MAX_TILES && smem + 1024 <= (size_t)max_smem;
    const size_t smem_f = sizeof(double) * kalman_fold_smem_doubles(g.ns, g.ny);
    if (fits && g_kalman_variant != 1 && g_kalman_fold && g.ny <= 32 && smem_f + 1024 <= (size_t)max_smem) {
      long long grid = g.n_draws < (long long)dev_sms * 32 ? g.n_draws : (long long)dev_sms * 32;
      g.scratch = nullptr;
      const int ntile = q.nsp / 8;
      auto launch = [&](auto kern) -> cudaError_t {
        if (smem_f > 48 * 1024) {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
          if (e != cudaSuccess) return e;
        }
        // up to four tiles (five CTAs per SM by registers): the whole L1 / shared array as shared memory, the default carve-out
        // stops at four CTAs (ns = 30: 12.9 -> 12.6 ms per 8 192 draws; larger CTAs keep the L1: ns = 40 measured 2 % slower without)
        if (ntile <= 4) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        kern<<<(unsigned)grid, 32 * (ntile + 1), smem_f, st>>>(g, q);
        return cudaGetLastError();
      };
      switch (ntile) {
        case 1: return launch(kalman_dmma_fold_kernel<1>);
        case 2: return launch(kalman_dmma_fold_kernel<2>);
        case 3: return launch(kalman_dmma_fold_kernel<3>);
        case 4: return launch(kalman_dmma_fold_kernel<4>);
        case 5: return launch(kalman_dmma_fold_kernel<5>);
        case 6: return launch(kalman_dmma_fold_kernel<6>);
        case 7: return launch(kalman_dmma_fold_kernel<7>);
        case 8: return launch(kalman_dmma_fold_kernel<8>);
        case 9: return launch(kalman_dmma_fold_kernel<9>);
        case 10: return launch(kalman_dmma_fold_kernel<10>);
        default: return launch(kalman_dmma_fold_kernel<DMMA_MAX_TILES>);
      }
    }
    if (fits && g_kalman_variant != 1) {
      long long grid = g.n_draws < (long long)dev_sms * 32 ? g.n_draws : (long long)dev_sms * 32;
      g.scratch = nullptr;
      const int ntile = q.nsp / 8;
      auto launch = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return e;
        }
        kern<<<(unsigned)grid, 32 * ntile, smem, st>>>(g, q);
        return cudaGetLastError();
      };
      if (ntile <= 3) return launch(kalman_dmma_kernel<3>);
      if (ntile <= 5) return launch(kalman_dmma_kernel<5>);
      if (ntile <= 8) return launch(kalman_dmma_kernel<8>);
      return launch(kalman_dmma_kernel<DMMA_MAX_TILES>);
    }
  }
  {
    // large state spaces: persistent CTAs, P and T P in an L2-resident scratch, blocked tensor-core GEMMs
    const KalmanLargeGeom q = kalman_large_geom(g.ns, g.ny);
    // bulk-copy staging needs every operand column to be a 16-byte aligned run: ns a multiple of 8, aligned bases
    const bool bulk = g_kalman_bulk && g.ns % 8 == 0 && ((uintptr_t)g.T % 16 == 0) && ((uintptr_t)g.RQR % 16 == 0) &&
                      sizeof(double) * kalman_large_smem_doubles(g.ns, g.ny, true) + 1024 <= (size_t)max_smem;
    const size_t smem = sizeof(double) * kalman_large_smem_doubles(g.ns, g.ny, bulk);
    const int threads = 32 * q.nwarps;
    const bool fits = q.nwarps <= 16 && smem + 1024 <= (size_t)max_smem &&
                      (size_t)8 * threads >= (size_t)q.nsp * DL_BK + (size_t)DL_BK * 8 * DL_NTP;
    if (fits && g_kalman_variant != 1 && g.ns > 8 * DMMA_MAX_TILES) {
      long long grid = g.n_draws < dev_sms ? g.n_draws : dev_sms;
      if (const char* e = getenv("DYB_KL_GRID")) { const long long c = atoll(e); if (c > 0 && c < grid) grid = c; }   // experiments only
      g.scratch = sg.temp<double>(kalman_large_scratch_doubles(g.ns, g.ny) * (size_t)grid);
      if (sg.err != cudaSuccess) return sg.err;
      auto launch = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return e;
        }
        kern<<<(unsigned)grid, threads, smem, st>>>(g, q);
        return cudaGetLastError();
      };
      return bulk ? launch(kalman_dmma_large_kernel<true>) : launch(kalman_dmma_large_kernel<false>);
    }
  }
  const size_t ns2 = (size_t)g.ns * g.ns;
  const size_t small = sizeof(double) * ((size_t)g.ny * g.ns + (size_t)g.ny * g.ny + 2 * (size_t)g.ns + 2 * (size_t)g.ny + 32);
  const size_t full = small + sizeof(double) * 2 * ns2;
  size_t smem;
  long long grid;
  if (full + 1024 <= (size_t)max_smem) {
    smem = full; g.scratch = nullptr;
    grid = g.n_draws;
    if (grid > 65535LL * 16) grid = 65535LL * 16;
  } else {
    smem = small;
    grid = (long long)dev_sms * 2 < g.n_draws ? (long long)dev_sms * 2 : g.n_draws;
    g.scratch = sg.temp<double>(2 * ns2 * (size_t)grid);
    if (sg.err != cudaSuccess) return sg.err;
  }
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kalman_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kalman_generic_kernel<<<(unsigned)grid, 256, smem, st>>>(g);
  return cudaGetLastError();
}


// ---- prior-predictive quantities per draw (src/estimation/priorprediction.jl:64-98): stdev of every endogenous
// variable from the unconditional variance (compute_variance!, robustsqrt: src/perturbations.jl:195, 205) and impulse
// responses to a one-standard-deviation shock (irfs!, src/perturbations.jl:670-694).  One CTA per draw. ------------
struct PriorPredArgs {
  int n, k, nx, periods;
  int bk[64];
  const double* g1;          // n x (k+nx) per draw
  const double* sigma_e;     // nx x nx per draw
  const int* status_in;
  long long n_draws;
  double lyap_tol; int lyap_maxit;
  double* stdev;             // n per draw (may be null)
  double* irfs;              // n x periods x nx per draw (may be null)
  int* status;
};

// Sigma_e of every draw: the shock part of set_estimated_parameters! (src/estimation/estimation.jl:517-601)
struct SigmaEArgs { int np, nx, ny; int est_type[MAX_EST], est_i[MAX_EST], est_j[MAX_EST]; };
// sigma_m0 / out_m may be null (only Sigma_e wanted)
__global__ void sigma_e_kernel(SigmaEArgs a, const double* __restrict__ sigma_e0, const double* __restrict__ sigma_m0,
                               const double* __restrict__ theta, long long n_draws, double* __restrict__ out,
                               double* __restrict__ out_m) {
  const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_draws) return;
  double* S = out + d * a.nx * a.nx;
  double* Sm = out_m ? out_m + d * a.ny * a.ny : nullptr;
  for (int e = 0; e < a.nx * a.nx; ++e) S[e] = sigma_e0[e];
  if (Sm) for (int e = 0; e < a.ny * a.ny; ++e) Sm[e] = sigma_m0[e];
  for (int q = 0; q < a.np; ++q) {
    const double v = theta[d * a.np + q];
    const int i = a.est_i[q], j = a.est_j[q], ty = a.est_type[q];
    if (ty == EST_SD_SHOCK) S[i + a.nx * j] = v * v;
    else if (ty == EST_VAR_SHOCK) S[i + a.nx * j] = v;
    else if (ty == EST_CORR_SHOCK) { const double c = v * sqrt(S[i + a.nx * i] * S[j + a.nx * j]); S[i + a.nx * j] = c; S[j + a.nx * i] = c; }
    else if (Sm && ty == EST_SD_MEAS) Sm[i + a.ny * j] = v * v;
    else if (Sm && ty == EST_VAR_MEAS) Sm[i + a.ny * j] = v;
    else if (Sm && ty == EST_CORR_MEAS) { const double c = v * sqrt(Sm[i + a.ny * i] * Sm[j + a.ny * j]); Sm[i + a.ny * j] = c; Sm[j + a.ny * i] = c; }
  }
}

// state space with EVERY endogenous variable as a state, as calibsmoother! builds it (src/filters/kalman/kalman.jl:78-125):
// T[:, i_bkwrd_b] = g1_1, R = g1_2, P0 = endogenous_variance.  One CTA per draw.
struct FullStateArgs {
  int n, k, nx, ny;
  int bk[64], obs[64];
  const double* g1; const double* sigma_e; const double* ybar; const int* status_in; long long n_draws;
  double lyap_tol; int lyap_maxit;
  double* T; double* R; double* P0; double* ybar_obs; int* status;
};
__global__ void __launch_bounds__(128) full_state_space_kernel(FullStateArgs a) {
  extern __shared__ double sm[];
  const int n = a.n, k = a.k, nx = a.nx, tid = threadIdx.x, nt = blockDim.x;
  double* Ak = sm; double* Sg = Ak + k * k; double* Tm = Sg + k * k; double* BQ = Tm + k * k; double* red = BQ + k * nx;
  for (long long d = blockIdx.x; d < a.n_draws; d += gridDim.x) {
    __syncthreads();
    int status = a.status_in[d];
    const double* g1 = a.g1 + (size_t)d * n * (k + nx);
    const double* g2 = g1 + (size_t)n * k;
    const double* Se = a.sigma_e + (size_t)d * nx * nx;
    if (status == ST_OK) {
      for (int e = tid; e < k * k; e += nt) Ak[e] = g1[a.bk[e % k] + n * (e / k)];
      for (int e = tid; e < k * nx; e += nt) {
        const int i = e % k, q = e / k;
        double s = 0.0;
        for (int f = 0; f < nx; ++f) s = fma(g2[a.bk[i] + n * f], Se[f + nx * q], s);
        BQ[e] = s;
      }
      __syncthreads();
      for (int e = tid; e < k * k; e += nt) {
        const int i = e % k, j = e / k;
        double s = 0.0;
        for (int q = 0; q < nx; ++q) s = fma(BQ[i + k * q], g2[a.bk[j] + n * q], s);
        Sg[e] = s;
      }
      __syncthreads();
      int it = 0;
      if (!blk_lyapunov<128>(Ak, Sg, Tm, k, a.lyap_tol, a.lyap_maxit, red, &it)) status = ST_NONSTATIONARY;
      __syncthreads();
    }
    if (status == ST_OK) {
      double* To = a.T + (size_t)d * n * n;
      double* Ro = a.R + (size_t)d * n * nx;
      double* Po = a.P0 + (size_t)d * n * n;
      for (int e = tid; e < n * n; e += nt) To[e] = 0.0;
      __syncthreads();
      for (int e = tid; e < n * k; e += nt) To[(e % n) + (size_t)n * a.bk[e / n]] = g1[e];
      for (int e = tid; e < n * nx; e += nt) Ro[e] = g2[e];
      for (int e = tid; e < n * n; e += nt) {
        const int v = e % n, w = e / n;
        if (v < w) continue;
        double s = 0.0;
        for (int i = 0; i < k; ++i) {
          double rr = 0.0;
          for (int j = 0; j < k; ++j) rr = fma(Sg[i + k * j], g1[w + n * j], rr);
          s = fma(g1[v + n * i], rr, s);
        }
        for (int q = 0; q < nx; ++q) {
          double rr = 0.0;
          for (int f = 0; f < nx; ++f) rr = fma(Se[q + nx * f], g2[w + n * f], rr);
          s = fma(g2[v + n * q], rr, s);
        }
        Po[v + (size_t)n * w] = s; Po[w + (size_t)n * v] = s;
      }
      for (int i = tid; i < a.ny; i += nt) a.ybar_obs[(size_t)d * a.ny + i] = a.ybar[(size_t)d * n + a.obs[i]];
    }
    if (tid == 0) a.status[d] = status;
  }
}

__global__ void __launch_bounds__(128) prior_predictive_kernel(PriorPredArgs a) {
  extern __shared__ double sm[];
  const int n = a.n, k = a.k, nx = a.nx, tid = threadIdx.x, nt = blockDim.x;
  double* Ak = sm; double* Sg = Ak + k * k; double* Tm = Sg + k * k; double* BQ = Tm + k * k;   // BQ: k x nx
  double* red = BQ + k * nx; double* ycur = red + 8; double* ynext = ycur + n;
  for (long long d = blockIdx.x; d < a.n_draws; d += gridDim.x) {
    __syncthreads();
    int status = a.status_in[d];
    const double* g1 = a.g1 + (size_t)d * n * (k + nx);
    const double* g2 = g1 + (size_t)n * k;
    const double* Se = a.sigma_e + (size_t)d * nx * nx;
    if (status == ST_OK && a.stdev) {
      for (int e = tid; e < k * k; e += nt) Ak[e] = g1[a.bk[e % k] + n * (e / k)];
      for (int e = tid; e < k * nx; e += nt) {
        const int i = e % k, q = e / k;
        double s = 0.0;
        for (int f = 0; f < nx; ++f) s = fma(g2[a.bk[i] + n * f], Se[f + nx * q], s);
        BQ[e] = s;
      }
      __syncthreads();
      for (int e = tid; e < k * k; e += nt) {
        const int i = e % k, j = e / k;
        double s = 0.0;
        for (int q = 0; q < nx; ++q) s = fma(BQ[i + k * q], g2[a.bk[j] + n * q], s);
        Sg[e] = s;
      }
      __syncthreads();
      int it = 0;
      if (!blk_lyapunov<128>(Ak, Sg, Tm, k, a.lyap_tol, a.lyap_maxit, red, &it)) status = ST_NONSTATIONARY;
      __syncthreads();
      // diag of Sigma_y = g1_1 Sigma_s g1_1' + g1_2 Sigma_e g1_2'
      for (int v = tid; v < n; v += nt) {
        double s = 0.0;
        for (int i = 0; i < k; ++i) {
          double r = 0.0;
          for (int j = 0; j < k; ++j) r = fma(Sg[i + k * j], g1[v + n * j], r);
          s = fma(g1[v + n * i], r, s);
        }
        for (int q = 0; q < nx; ++q) {
          double r = 0.0;
          for (int f = 0; f < nx; ++f) r = fma(Se[q + nx * f], g2[v + n * f], r);
          s = fma(g2[v + n * q], r, s);
        }
        a.stdev[(size_t)d * n + v] = status == ST_OK ? sqrt(s + 2.220446049250313e-16) : NAN;
      }
    } else if (a.stdev) {
      for (int v = tid; v < n; v += nt) a.stdev[(size_t)d * n + v] = NAN;
    }
    if (a.irfs) {
      double* out = a.irfs + (size_t)d * n * a.periods * nx;
      for (int q = 0; q < nx; ++q) {
        __syncthreads();
        const double x = sqrt(Se[q + nx * q] + 2.220446049250313e-16);
        for (int v = tid; v < n; v += nt) {
          const double y = status == ST_OK || status == ST_NONSTATIONARY ? g2[v + n * q] * x : NAN;
          ycur[v] = y;
          out[v + (size_t)n * (0 + (size_t)a.periods * q)] = y;
        }
        for (int p = 1; p < a.periods; ++p) {
          __syncthreads();
          for (int v = tid; v < n; v += nt) {
            double s = 0.0;
            for (int j = 0; j < k; ++j) s = fma(g1[v + n * j], ycur[a.bk[j]], s);
            ynext[v] = s;
            out[v + (size_t)n * (p + (size_t)a.periods * q)] = s;
          }
          __syncthreads();
          double* t_ = ycur; ycur = ynext; ynext = t_;
        }
      }
    }
    if (tid == 0) a.status[d] = status;
  }
}


// unconditional forecast from a given initial point (forecast_!, src/forecast.jl:154-159 -> simul_first_order! with zero
// shocks): Y[:, 0] = y0, Y[:, h] = ybar + g1_1 (Y[:, h-1] - ybar)[i_bkwrd_b].  One warp per draw.
struct ForecastArgs {
  int n, k, periods;
  int bk[64];
  const double* g1; long long g1_stride; const double* ybar; const double* y0; const int* status_in; long long n_draws;
  double* Y;
};
__global__ void __launch_bounds__(128) forecast_kernel(ForecastArgs a) {
  const int lane = threadIdx.x & 31;
  const long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (d >= a.n_draws) return;
  const int n = a.n, k = a.k;
  double* Y = a.Y + (size_t)d * n * (a.periods + 1);
  const double* g1 = a.g1 + (size_t)d * a.g1_stride;
  const double* yb = a.ybar + (size_t)d * n;
  const bool ok = a.status_in[d] == ST_OK;
  for (int v = lane; v < n; v += 32) Y[v] = ok ? a.y0[(size_t)d * n + v] : NAN;
  __syncwarp();
  for (int h = 1; h <= a.periods; ++h) {
    const double* prev = Y + (size_t)n * (h - 1);
    for (int v = lane; v < n; v += 32) {
      double s = yb[v];
      for (int j = 0; j < k; ++j) s = fma(g1[v + n * j], prev[a.bk[j]] - yb[a.bk[j]], s);
      Y[v + (size_t)n * h] = ok ? s : NAN;
    }
    __syncwarp();
  }
}

}  // namespace dyb

using namespace dyb;

// the univariate fallback on the draws the production filter flagged (mode 1) / on every draw (mode 2)
static cudaError_t launch_kalman_univariate(KalmanGenericArgs g, int mode, const int* status_in, int device, cudaStream_t st, DevBuf& scratch) {
  int dev_sms = 148;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, device);
  long long blocks = g.n_draws < (long long)dev_sms * 8 ? g.n_draws : (long long)dev_sms * 8;
  if (blocks < 1) return cudaSuccess;
  cudaError_t e = scratch.reserve(sizeof(double) * 2 * (size_t)g.ns * g.ns * (size_t)blocks);
  if (e != cudaSuccess) return e;
  g.scratch = scratch.as<double>();
  const size_t smem = kalman_univariate_smem_bytes(g.ns, g.ny);
  if (smem > 48 * 1024) {
    e = cudaFuncSetAttribute(kalman_univariate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kalman_univariate_kernel<<<(unsigned)blocks, 64, smem, st>>>(g, KalmanUnivariateSel{mode, status_in});
  return cudaGetLastError();
}

namespace dyb {
int univariate_fallback_from_handoff(dyb_handle* h, long long n, const int* d_status_solve, double* d_loglik, int* d_status) {
  const dyb_model_dims& d = h->dims;
  const size_t N = (size_t)n, ns2 = (size_t)d.ns * d.ns;
  DYB_CUDA(h->d_uT.reserve(8 * ns2 * N)); DYB_CUDA(h->d_uQ.reserve(8 * ns2 * N)); DYB_CUDA(h->d_uP.reserve(8 * ns2 * N));
  DYB_CUDA(h->d_uyb.reserve(8 * (size_t)d.ny * N)); DYB_CUDA(h->d_uH.reserve(8 * (size_t)d.ny * d.ny * N));
  DYB_CUDA(h->d_uz.reserve(4 * (size_t)d.ny));
  std::vector<int> z((size_t)d.ny);
  h->ops->obs_state_positions(z.data());
  DYB_CUDA(cudaMemcpyAsync(h->d_uz.p, z.data(), 4 * (size_t)d.ny, cudaMemcpyHostToDevice, h->stream));
  DYB_CUDA(h->ops->unpack(h->d_ss.as<double>(), n, h->d_uT.as<double>(), h->d_uQ.as<double>(), h->d_uP.as<double>(),
                          h->d_uyb.as<double>(), h->d_uH.as<double>(), h->stream));
  KalmanGenericArgs g{};
  g.ns = d.ns; g.ny = d.ny; g.nobs = h->nobs; g.presample = h->hs.opt.presample;
  g.z_idx = h->d_uz.as<int>(); g.Y = h->d_Y.as<double>();
  g.T = h->d_uT.as<double>(); g.RQR = h->d_uQ.as<double>(); g.P0 = h->d_uP.as<double>(); g.a0 = nullptr;
  g.H = h->d_uH.as<double>(); g.ybar_obs = h->d_uyb.as<double>(); g.status_in = d_status_solve;
  g.n_draws = n; g.loglik = d_loglik; g.status = d_status;
  DYB_CUDA(launch_kalman_univariate(g, h->hs.opt.univariate_fallback, d_status_solve, h->device, h->stream, h->d_uscr));
  h->launches += 2;
  return DYB_OK;
}
}  // namespace dyb

extern "C" {

int dyb_kalman_batch(int device, int32_t ns, int32_t ny, int32_t nobs, int32_t presample,
                     const int32_t* z_idx, const double* Y, const double* T, const double* RQR,
                     const double* P0, const double* a0, const double* H, int64_t n,
                     double* loglik, int32_t* status) {
  if (ns <= 0 || ny <= 0 || ny > 64 || ny > ns || nobs < 0 || presample < 0 || n < 0)
    return fail(DYB_ERR_ARG, "bad sizes (need 0 < ny <= min(ns, 64))");
  if (n == 0) return DYB_OK;
  if (!z_idx || !Y || !T || !RQR || !P0 || !loglik || !status) return fail(DYB_ERR_ARG, "null argument");
  DYB_CUDA(cudaSetDevice(device));
  cudaStream_t st = device_stream(device);
  Stager sg(st);
  const size_t N = (size_t)n, ns2 = (size_t)ns * ns;
  KalmanGenericArgs g{};
  g.ns = ns; g.ny = ny; g.nobs = nobs; g.presample = presample;
  g.z_idx = sg.in(z_idx, (size_t)ny);
  g.Y = sg.in(Y, (size_t)ny * (nobs > 0 ? nobs : 1));
  g.T = sg.in(T, ns2 * N);
  g.RQR = sg.in(RQR, ns2 * N);
  g.P0 = sg.in(P0, ns2 * N);
  g.a0 = sg.in(a0, (size_t)ns * N);
  g.H = sg.in(H, (size_t)ny * ny * N);
  g.n_draws = n;
  g.loglik = sg.out(loglik, N);
  g.status = sg.out(status, N);
  if (!is_device_ptr(z_idx)) for (int i = 0; i < ny; ++i) if (z_idx[i] < 0 || z_idx[i] >= ns) return fail(DYB_ERR_ARG, "z_idx out of range");
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  g.scratch = nullptr;
  DYB_CUDA(launch_kalman_generic(g, device, st, sg));
  DYB_CUDA(sg.finish());
  return DYB_OK;
}


int dyb_kalman_batch_fallback(int device, int32_t ns, int32_t ny, int32_t nobs, int32_t presample,
                              const int32_t* z_idx, const double* Y, const double* T, const double* RQR,
                              const double* P0, const double* a0, const double* H, int64_t n, int32_t mode,
                              double* loglik, int32_t* status) {
  if (ns <= 0 || ny <= 0 || ny > 64 || ny > ns || nobs < 0 || presample < 0 || n < 0 || mode < 1 || mode > 2)
    return fail(DYB_ERR_ARG, "bad sizes (need 0 < ny <= min(ns, 64)) or mode (1, 2)");
  if (n == 0) return DYB_OK;
  if (!z_idx || !Y || !T || !RQR || !P0 || !loglik || !status) return fail(DYB_ERR_ARG, "null argument");
  DYB_CUDA(cudaSetDevice(device));
  cudaStream_t st = device_stream(device);
  Stager sg(st);
  const size_t N = (size_t)n, ns2 = (size_t)ns * ns;
  KalmanGenericArgs g{};
  g.ns = ns; g.ny = ny; g.nobs = nobs; g.presample = presample;
  g.z_idx = sg.in(z_idx, (size_t)ny);
  g.Y = sg.in(Y, (size_t)ny * (nobs > 0 ? nobs : 1));
  g.T = sg.in(T, ns2 * N); g.RQR = sg.in(RQR, ns2 * N); g.P0 = sg.in(P0, ns2 * N);
  g.a0 = sg.in(a0, (size_t)ns * N); g.H = sg.in(H, (size_t)ny * ny * N);
  g.n_draws = n; g.loglik = sg.out(loglik, N); g.status = sg.out(status, N);
  if (!is_device_ptr(z_idx)) for (int i = 0; i < ny; ++i) if (z_idx[i] < 0 || z_idx[i] >= ns) return fail(DYB_ERR_ARG, "z_idx out of range");
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  g.scratch = nullptr;
  DYB_CUDA(launch_kalman_generic(g, device, st, sg));
  DevBuf scratch;
  cudaError_t e = launch_kalman_univariate(g, mode, nullptr, device, st, scratch);
  cudaError_t e2 = sg.finish();
  scratch.release();
  DYB_CUDA(e); DYB_CUDA(e2);
  return DYB_OK;
}

int dyb_set_kalman_variant(int32_t variant) {
  if (variant < 0 || variant > 5)
    return fail(DYB_ERR_ARG, "variant must be 0 (auto), 1 (scalar), 2 (tensor-core), 3 (tensor-core, no bulk copies), 4 (sub-warp, ns <= 16) or 5 (tensor-core, unfolded ns <= 88 kernel)");
  g_kalman_bulk = variant != 3;
  g_kalman_fold = variant != 5;
  g_kalman_variant = (variant == 3 || variant == 5) ? 2 : variant;
  return DYB_OK;
}

// ---- the run-time description (index arrays on the device, shared-memory plan) on an existing handle -----------------------
static int generic_build(dyb_handle* h, const dyb_model_desc* md, dyb_model_dims* dims_out) {
  if (!md || md->n < 1 || md->nx < 1 || md->ny < 1 || md->n_static < 0 || md->n_bkwrd_b < 0 || md->n_fwrd_b < 0 ||
      md->n_static >= md->n || md->ny > 64 || (md->n_static > 0 && !md->i_static) || (md->n_bkwrd_b > 0 && !md->i_bkwrd_b) ||
      (md->n_fwrd_b > 0 && !md->i_fwrd_b) || !md->obs_idx)
    return fail(DYB_ERR_ARG, "bad model description");
  const int n = md->n, nx = md->nx, ny = md->ny, s = md->n_static, k = md->n_bkwrd_b, nf = md->n_fwrd_b, m = n - s;
  auto in_range_sorted = [&](const int32_t* v, int cnt) {
    for (int i = 0; i < cnt; ++i) if (v[i] < 0 || v[i] >= n || (i > 0 && v[i] <= v[i - 1])) return false;
    return true;
  };
  if (!in_range_sorted(md->i_static, s) || !in_range_sorted(md->i_bkwrd_b, k) || !in_range_sorted(md->i_fwrd_b, nf))
    return fail(DYB_ERR_ARG, "index sets must be strictly increasing 0-based variable indices");
  std::vector<int> var_static(n, 0), var_pos(n, 0);
  for (int i = 0; i < s; ++i) var_static[md->i_static[i]] = 1;
  for (int i = 0; i < k; ++i) if (var_static[md->i_bkwrd_b[i]]) return fail(DYB_ERR_ARG, "a static variable cannot have a lag");
  for (int i = 0; i < nf; ++i) if (var_static[md->i_fwrd_b[i]]) return fail(DYB_ERR_ARG, "a static variable cannot have a lead");
  { int ps = 0, pd = 0; for (int v = 0; v < n; ++v) var_pos[v] = var_static[v] ? ps++ : pd++; }
  std::vector<int> bk(k), fw(nf);
  for (int i = 0; i < k; ++i) bk[i] = var_pos[md->i_bkwrd_b[i]];
  for (int i = 0; i < nf; ++i) fw[i] = var_pos[md->i_fwrd_b[i]];
  // compressed Jacobian columns [lags | current | leads | shocks] (src/perturbations.jl:616-635) -> working columns
  const int ncol = k + n + nf + nx;
  std::vector<int> jperm(ncol);
  for (int j = 0; j < k; ++j) jperm[j] = s + j;
  for (int v = 0; v < n; ++v) jperm[k + v] = var_static[v] ? var_pos[v] : s + k + var_pos[v];
  for (int j = 0; j < nf; ++j) jperm[k + n + j] = s + k + m + j;
  for (int e = 0; e < nx; ++e) jperm[k + n + nf + e] = s + k + m + nf + e;
  // Kalman states: sort(union(obs_idx, i_bkwrd_b)) (src/estimation/estimation.jl:384-390)
  std::vector<int> state_ids(md->i_bkwrd_b, md->i_bkwrd_b + k);
  for (int i = 0; i < ny; ++i) {
    if (md->obs_idx[i] < 0 || md->obs_idx[i] >= n) return fail(DYB_ERR_ARG, "obs_idx out of range");
    state_ids.push_back(md->obs_idx[i]);
  }
  std::sort(state_ids.begin(), state_ids.end());
  state_ids.erase(std::unique(state_ids.begin(), state_ids.end()), state_ids.end());
  const int ns = (int)state_ids.size();
  auto pos_in_states = [&](int v) { return (int)(std::lower_bound(state_ids.begin(), state_ids.end(), v) - state_ids.begin()); };
  std::vector<int> obs_state(ny), lagged(k), obs_var(md->obs_idx, md->obs_idx + ny);
  for (int i = 0; i < ny; ++i) obs_state[i] = pos_in_states(md->obs_idx[i]);
  for (int j = 0; j < k; ++j) lagged[j] = pos_in_states(md->i_bkwrd_b[j]);

  DYB_CUDA(cudaSetDevice(h->device));
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
  const size_t smem = sizeof(double) * generic_solve_smem_doubles(n, nx, ny, s, k, nf, ns);
  if (smem + 1024 > (size_t)max_smem)
    return fail(DYB_ERR_UNSUPPORTED, "model too large for the shared-memory generic solver (" + std::to_string(smem / 1024) + " KB per draw)");
  GenericHost* gh = new GenericHost();
  // pack the index arrays into one device buffer
  std::vector<int> pack;
  auto push = [&](const std::vector<int>& v) { const size_t off = pack.size(); pack.insert(pack.end(), v.begin(), v.end()); return off; };
  const size_t o_jperm = push(jperm), o_bk = push(bk), o_fw = push(fw), o_vs = push(var_static), o_vp = push(var_pos),
               o_sid = push(state_ids), o_ov = push(obs_var), o_lag = push(lagged);
  cudaError_t e = gh->d_idx.reserve(sizeof(int) * pack.size());
  if (e == cudaSuccess) e = cudaMemcpy(gh->d_idx.p, pack.data(), sizeof(int) * pack.size(), cudaMemcpyHostToDevice);
  const int* base = gh->d_idx.as<int>();
  gh->gm = GenericModel{n, nx, ny, s, k, nf, m, ncol, ns, base + o_jperm, base + o_bk, base + o_fw, base + o_vs, base + o_vp,
                        base + o_sid, base + o_ov, base + o_lag};
  gh->obs_state = obs_state;
  gh->smem = smem;
  gh->warp_groups = smem <= 24 * 1024;
  // medium models (more than 12 dynamic variables, up to 32): register-blocked kernel, instantiation <4, 6>
  if (!gh->warp_groups && m > 12 && m <= 32 && m + k + nf <= 96 && m + nx <= 96 && k + nf <= 64) {
    gh->sbl = sb_layout(n, nx, s, k, nf, ns, 32, 96);
    const size_t sb = sizeof(double) * (size_t)gh->sbl.total;
    if (sb + 1024 <= (size_t)max_smem) {
      gh->block = true;
      e = gh->wscr.reserve(sizeof(double) * (size_t)gh->sbl.wdoubles * SB_GRID);
      gh->block_small = (m + k + nf <= 80) && (m + nx <= 80) && (k + nf <= 48);
      // test hook: DYB_SB_WIDE=1 in the environment when the handle is built forces the wide instantiation <4, 6, 4>
      if (const char* wide = std::getenv("DYB_SB_WIDE")) { if (wide[0] == '1') gh->block_small = false; }
      if (e == cudaSuccess && sb > 48 * 1024) {
        e = cudaFuncSetAttribute(solve_block_kernel<4, 6, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(solve_block_kernel<4, 5, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb);
      }
    }
  }
  if (e == cudaSuccess) {
    if (gh->warp_groups) {
      if (4 * smem > 48 * 1024) e = cudaFuncSetAttribute(solve_generic_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * smem));
    } else if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(solve_generic_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
  }
  if (e != cudaSuccess) { gh->d_idx.release(); gh->wscr.release(); delete gh; return fail(DYB_ERR_CUDA, cudaGetErrorString(e)); }
  h->generic = gh;
  if (dims_out) *dims_out = dyb_model_dims{n, nx, 0, k, nf, s, ncol, ns, ny, 0, 0, 0};
  return DYB_OK;
}

// ---- handle from index sets alone --------------------------------------------------------------------------------
int dyb_create_generic(const dyb_model_desc* md, int device, dyb_handle** out) {
  if (!out) return fail(DYB_ERR_ARG, "out is null");
  *out = nullptr;
  int ndev = 0;
  DYB_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(DYB_ERR_ARG, "no such CUDA device");
  DYB_CUDA(cudaSetDevice(device));
  dyb_handle* h = new dyb_handle();
  h->device = device;
  h->hs.np = 0;
  h->hs.opt.cr_tol = 1e-8; h->hs.opt.cr_maxit = 100; h->hs.opt.steady_tol = std::cbrt(2.220446049250313e-16);
  h->hs.opt.lyap_tol = 1e-16; h->hs.opt.lyap_maxit = 60; h->hs.opt.presample = 0; h->hs.opt.solver = 0; h->hs.opt.filter_lanes = 0;
  h->hs.opt.univariate_fallback = 0;
  cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  h->stream = h->own_stream;
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&h->ev[i]);
  if (e != cudaSuccess) { dyb_destroy(h); return fail(DYB_ERR_CUDA, cudaGetErrorString(e)); }
  const int rc = generic_build(h, md, &h->dims);
  if (rc) { dyb_destroy(h); return rc; }
  *out = h;
  return DYB_OK;
}

static int generic_solve(dyb_handle* h, GenericSolveArgs a) {
  GenericHost* gh = h->generic;
  a.cr_tol = h->hs.opt.cr_tol; a.cr_maxit = h->hs.opt.cr_maxit; a.lyap_tol = h->hs.opt.lyap_tol; a.lyap_maxit = h->hs.opt.lyap_maxit;
  const int per_draw = (int)(gh->smem / sizeof(double));
  if (gh->block && h->hs.opt.solver != 1) {      // medium model: one CTA per draw, algebra in registers (solve_block.cuh)
    long long grid = a.n_draws;
    if (grid > SB_GRID) grid = SB_GRID;
    a.wscratch = gh->wscr.as<double>();
    const size_t sb = sizeof(double) * (size_t)gh->sbl.total;
    if (gh->block_small) solve_block_kernel<4, 5, 3><<<(unsigned)grid, SB_THREADS, sb, h->stream>>>(gh->gm, a, gh->sbl);
    else solve_block_kernel<4, 6, 4><<<(unsigned)grid, SB_THREADS, sb, h->stream>>>(gh->gm, a, gh->sbl);
  } else if (gh->warp_groups) {                 // small model: one warp per draw, four draws per CTA
    long long grid = (a.n_draws + 3) / 4;
    if (grid > 148LL * 64) grid = 148LL * 64;
    solve_generic_kernel<32><<<(unsigned)grid, 128, 4 * gh->smem, h->stream>>>(gh->gm, a, per_draw);
  } else {
    long long grid = a.n_draws;
    if (grid > 148LL * 64) grid = 148LL * 64;
    solve_generic_kernel<128><<<(unsigned)grid, 128, gh->smem, h->stream>>>(gh->gm, a, per_draw);
  }
  DYB_CUDA(cudaGetLastError());
  h->launches += 1;
  return DYB_OK;
}

int dyb_first_order_from_jacobian_batch(dyb_handle* h, const double* jac, const double* sigma_e, int32_t sigma_e_per_draw,
                                        int64_t n, double* g1, double* T, double* RQR, double* P0,
                                        int32_t* cr_iters, int32_t* lyap_iters, int32_t* status) {
  if (h && h->ops && h->ops->solve_from_jacobian)       // a compiled sizes-only model: the register-resident kernels
    return jacfed_first_order(h, jac, sigma_e, sigma_e_per_draw, n, g1, T, RQR, P0, cr_iters, lyap_iters, status);
  if (!h || !h->generic) return fail(DYB_ERR_ARG, "needs a handle made by dyb_create_generic");
  if (n < 0 || (n > 0 && (!jac || !sigma_e || !status))) return fail(DYB_ERR_ARG, "null argument");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  const dyb_model_dims& d = h->dims;
  Stager sg(h->stream);
  const size_t N = (size_t)n, ns2 = (size_t)d.ns * d.ns;
  GenericSolveArgs a{};
  a.jac = sg.in(jac, (size_t)d.n * d.ncol * N);
  a.sigma_e = sg.in(sigma_e, (size_t)d.nx * d.nx * (sigma_e_per_draw ? N : 1));
  a.stride_se = sigma_e_per_draw ? (long long)d.nx * d.nx : 0;
  a.n_draws = n;
  a.g1 = sg.out(g1, (size_t)d.n * (d.k + d.nx) * N);
  a.T = sg.out(T, ns2 * N); a.RQR = sg.out(RQR, ns2 * N); a.P0 = sg.out(P0, ns2 * N);
  a.status = sg.out(status, N); a.cr_iters = sg.out(cr_iters, N); a.lyap_iters = sg.out(lyap_iters, N);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  if (a.g1) DYB_CUDA(cudaMemsetAsync(a.g1, 0, sizeof(double) * (size_t)d.n * (d.k + d.nx) * N, h->stream));
  int rc = generic_solve(h, a);
  if (rc) return rc;
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

// solve + dense filter from device-resident Jacobians (shared by dyb_loglik_from_jacobian_batch and the generated-model path
// for models beyond the register-resident solver), in chunks that bound the working set of dense state spaces
static int loglik_from_device_jacobians(dyb_handle* h, const double* djac, const double* dyb_, const double* dse, bool se_per_draw,
                                        const double* dsm, bool sm_per_draw, const int* dsi, long long n, double* dll, int* dst,
                                        Stager& sg) {
  const dyb_model_dims& d = h->dims;
  GenericHost* gh = h->generic;
  const size_t ns2 = (size_t)d.ns * d.ns;
  const long long chunk = std::max<long long>(1, std::min<long long>(n, (long long)((1ULL << 30) / (sizeof(double) * (3 * ns2 + d.ny * (d.ny + 1))))));
  DYB_CUDA(gh->bT.reserve(sizeof(double) * ns2 * chunk)); DYB_CUDA(gh->bQ.reserve(sizeof(double) * ns2 * chunk));
  DYB_CUDA(gh->bP.reserve(sizeof(double) * ns2 * chunk)); DYB_CUDA(gh->byo.reserve(sizeof(double) * d.ny * chunk));
  DYB_CUDA(gh->bH.reserve(sizeof(double) * d.ny * d.ny * chunk)); DYB_CUDA(gh->bst.reserve(sizeof(int) * chunk));
  DYB_CUDA(gh->bz.reserve(sizeof(int) * d.ny));
  double* dT = gh->bT.as<double>(); double* dQ = gh->bQ.as<double>(); double* dP = gh->bP.as<double>();
  double* dyo = gh->byo.as<double>(); double* dH = gh->bH.as<double>();
  int* dst1 = gh->bst.as<int>();
  int* dz = gh->bz.as<int>();
  DYB_CUDA(cudaMemcpyAsync(dz, gh->obs_state.data(), sizeof(int) * d.ny, cudaMemcpyHostToDevice, h->stream));
  for (long long lo = 0; lo < n; lo += chunk) {
    const long long c = std::min<long long>(chunk, n - lo);
    GenericSolveArgs a{};
    a.jac = djac + (size_t)lo * d.n * d.ncol;
    a.ybar = dyb_ ? dyb_ + (size_t)lo * d.n : nullptr;
    a.sigma_e = dse + (se_per_draw ? (size_t)lo * d.nx * d.nx : 0);
    a.stride_se = se_per_draw ? (long long)d.nx * d.nx : 0;
    a.sigma_m = dsm ? dsm + (sm_per_draw ? (size_t)lo * d.ny * d.ny : 0) : nullptr;
    a.stride_sm = sm_per_draw ? (long long)d.ny * d.ny : 0;
    a.status_in = dsi ? dsi + lo : nullptr;
    a.n_draws = c;
    a.T = dT; a.RQR = dQ; a.P0 = dP; a.ybar_obs = dyo; a.H = dH; a.status = dst1;
    int rc = generic_solve(h, a);
    if (rc) return rc;
    KalmanGenericArgs g{};
    g.ns = d.ns; g.ny = d.ny; g.nobs = h->nobs; g.presample = h->hs.opt.presample;
    g.z_idx = dz; g.Y = h->d_Y.as<double>(); g.T = dT; g.RQR = dQ; g.P0 = dP; g.a0 = nullptr; g.H = dH;
    g.ybar_obs = dyo; g.status_in = dst1;
    g.n_draws = c; g.loglik = dll + lo; g.status = dst + lo; g.scratch = nullptr;
    DYB_CUDA(launch_kalman_generic(g, h->device, h->stream, sg));
    if (h->hs.opt.univariate_fallback)
      DYB_CUDA(launch_kalman_univariate(g, h->hs.opt.univariate_fallback, dst1, h->device, h->stream, h->d_uscr));
    h->launches += 1;
  }
  return DYB_OK;
}

int dyb_loglik_from_jacobian_batch(dyb_handle* h, const double* jac, const double* ybar, const double* sigma_e,
                                   int32_t sigma_e_per_draw, const double* sigma_m, int32_t sigma_m_per_draw,
                                   const int32_t* status_in, int64_t n, double* loglik, int32_t* status) {
  if (h && h->ops && h->ops->solve_from_jacobian)       // a compiled sizes-only model: the register-resident kernels
    return jacfed_loglik(h, jac, ybar, sigma_e, sigma_e_per_draw, sigma_m, sigma_m_per_draw, status_in, n, loglik, status);
  if (!h || !h->generic || h->ops) return fail(DYB_ERR_ARG, "needs a handle made by dyb_create_generic");
  if (n < 0 || (n > 0 && (!jac || !sigma_e || !loglik || !status))) return fail(DYB_ERR_ARG, "null argument");
  if (h->d_Y.p == nullptr) return fail(DYB_ERR_STATE, "dyb_set_observations has not been called");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  const dyb_model_dims& d = h->dims;
  GenericHost* gh = h->generic;
  Stager sg(h->stream);
  const size_t N = (size_t)n;
  // inputs: device pointers are used in place, host arrays go through persistent device buffers (pinned host memory from
  // dyb_host_alloc is copied at full PCIe speed)
  cudaError_t se_ = cudaSuccess;
  auto stage = [&](DevBuf& b, const void* p, size_t bytes) -> const void* {
    if (!p || is_device_ptr(p)) return p;
    if (se_ == cudaSuccess) se_ = b.reserve(bytes);
    if (se_ == cudaSuccess) se_ = cudaMemcpyAsync(b.p, p, bytes, cudaMemcpyHostToDevice, h->stream);
    return b.p;
  };
  const double* djac = static_cast<const double*>(stage(gh->bJ, jac, sizeof(double) * d.n * d.ncol * N));
  const double* dyb_ = static_cast<const double*>(stage(gh->byb, ybar, sizeof(double) * d.n * N));
  const double* dse = static_cast<const double*>(stage(gh->bse, sigma_e, sizeof(double) * d.nx * d.nx * (sigma_e_per_draw ? N : 1)));
  const double* dsm = static_cast<const double*>(stage(gh->bsm, sigma_m, sizeof(double) * d.ny * d.ny * (sigma_m_per_draw ? N : 1)));
  const int* dsi = static_cast<const int*>(stage(gh->bsi, status_in, sizeof(int) * N));
  if (se_ != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(se_));
  double* dll = sg.out(loglik, N);
  int* dst = sg.out(status, N);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  int rc = loglik_from_device_jacobians(h, djac, dyb_, dse, sigma_e_per_draw != 0, dsm, sigma_m_per_draw != 0, dsi, n, dll, dst, sg);
  if (rc) return rc;
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

// ---- generated models beyond the register-resident solver: model function -> Jacobians on the device -> the path above ----
static int hybrid_prepare(dyb_handle* h, const double* d_theta, long long n) {
  const dyb_model_dims& d = h->dims;
  if (!h->generic) {
    std::vector<int> is((size_t)d.s), ib((size_t)d.k), ifw((size_t)d.nf), io((size_t)d.ny);
    h->ops->static_indices(is.data()); h->ops->bkwrd_indices(ib.data()); h->ops->fwrd_indices(ifw.data()); h->ops->obs_indices(io.data());
    dyb_model_desc md{d.n, d.nx, d.ny, d.s, d.k, d.nf, is.data(), ib.data(), ifw.data(), io.data()};
    int rc = generic_build(h, &md, nullptr);
    if (rc) return rc;
  }
  GenericHost* gh = h->generic;
  const size_t N = (size_t)n;
  DYB_CUDA(gh->bJ.reserve(sizeof(double) * d.n * d.ncol * N)); DYB_CUDA(gh->byb.reserve(sizeof(double) * d.n * N));
  DYB_CUDA(gh->bse.reserve(sizeof(double) * d.nx * d.nx * N)); DYB_CUDA(gh->bsm.reserve(sizeof(double) * d.ny * d.ny * N));
  DYB_CUDA(gh->bsi.reserve(sizeof(int) * N));
  DYB_CUDA(h->ops->jacobian_batch(h->hs, d_theta, n, gh->bJ.as<double>(), gh->byb.as<double>(), gh->bse.as<double>(),
                                  gh->bsm.as<double>(), gh->bsi.as<int>(), h->stream));
  h->launches += 1;
  return DYB_OK;
}

}  // extern "C"

namespace dyb {
int hybrid_loglik(dyb_handle* h, const double* d_theta, long long n, double* d_loglik, int* d_status) {
  int rc = hybrid_prepare(h, d_theta, n);
  if (rc) return rc;
  GenericHost* gh = h->generic;
  Stager sg(h->stream);
  rc = loglik_from_device_jacobians(h, gh->bJ.as<double>(), gh->byb.as<double>(), gh->bse.as<double>(), true,
                                    gh->bsm.as<double>(), true, gh->bsi.as<int>(), n, d_loglik, d_status, sg);
  if (rc) return rc;
  if (!sg.temps.empty()) DYB_CUDA(sg.finish());      // scratch of the large-state filter: released after the stream has drained
  return DYB_OK;
}

// LRE.first_order_solver! for such a model: g1, ybar (may be null), iteration counts, status; device pointers
int hybrid_first_order(dyb_handle* h, const double* d_theta, long long n, double* d_g1, double* d_ybar, int* d_cr_iters,
                       int* d_lyap_iters, int* d_status) {
  int rc = hybrid_prepare(h, d_theta, n);
  if (rc) return rc;
  const dyb_model_dims& d = h->dims;
  GenericHost* gh = h->generic;
  GenericSolveArgs a{};
  a.jac = gh->bJ.as<double>(); a.ybar = gh->byb.as<double>();
  a.sigma_e = gh->bse.as<double>(); a.stride_se = (long long)d.nx * d.nx;
  a.sigma_m = gh->bsm.as<double>(); a.stride_sm = (long long)d.ny * d.ny;
  a.status_in = gh->bsi.as<int>(); a.n_draws = n;
  a.g1 = d_g1; a.status = d_status; a.cr_iters = d_cr_iters; a.lyap_iters = d_lyap_iters;
  if (d_g1) DYB_CUDA(cudaMemsetAsync(d_g1, 0, sizeof(double) * (size_t)d.n * (d.k + d.nx) * (size_t)n, h->stream));
  rc = generic_solve(h, a);
  if (rc) return rc;
  if (d_ybar) DYB_CUDA(cudaMemcpyAsync(d_ybar, gh->byb.p, sizeof(double) * (size_t)d.n * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
  return DYB_OK;
}
}  // namespace dyb

extern "C" {

int dyb_prior_predictive_batch(dyb_handle* h, const double* theta, int64_t n, int32_t periods, double* ybar, double* stdev,
                                double* irfs, int32_t* status) {
  if (h && !h->ops) return fail(DYB_ERR_UNSUPPORTED, "needs a handle with a device model function");
  if (!h || n < 0 || periods < 0 || (n > 0 && (!theta || !status))) return fail(DYB_ERR_ARG, "bad argument");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  const dyb_model_dims& d = h->dims;
  if (d.k > 64) return fail(DYB_ERR_UNSUPPORTED, "more than 64 state variables");
  const size_t N = (size_t)n;
  Stager sg(h->stream);
  double* dg1 = sg.temp<double>((size_t)d.n * (d.k + d.nx) * N);
  double* dse = sg.temp<double>((size_t)d.nx * d.nx * N);
  int* dst1 = sg.temp<int>(N);
  double* dse0 = sg.temp<double>((size_t)d.nx * d.nx);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  DYB_CUDA(cudaMemcpyAsync(dse0, h->hs.sigma_e.data(), sizeof(double) * d.nx * d.nx, cudaMemcpyHostToDevice, h->stream));
  int rc = dyb_first_order_batch(h, theta, n, dg1, ybar, nullptr, nullptr, reinterpret_cast<int32_t*>(dst1));
  if (rc) return rc;
  const double* dth = sg.in(theta, (size_t)h->hs.np * N);
  int* dst = sg.out(status, N);
  double* dsd = sg.out(stdev, (size_t)d.n * N);
  double* dirf = (irfs && periods > 0) ? sg.out(irfs, (size_t)d.n * periods * d.nx * N) : nullptr;
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  SigmaEArgs sa{};
  sa.np = h->hs.np; sa.nx = d.nx; sa.ny = d.ny;
  for (int q = 0; q < h->hs.np; ++q) { sa.est_type[q] = h->hs.est_type[q]; sa.est_i[q] = h->hs.est_i[q]; sa.est_j[q] = h->hs.est_j[q]; }
  sigma_e_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(sa, dse0, nullptr, dth, n, dse, nullptr);
  PriorPredArgs pa{};
  pa.n = d.n; pa.k = d.k; pa.nx = d.nx; pa.periods = periods;
  h->ops->bkwrd_indices(pa.bk);
  pa.g1 = dg1; pa.sigma_e = dse; pa.status_in = dst1; pa.n_draws = n;
  pa.lyap_tol = h->hs.opt.lyap_tol; pa.lyap_maxit = h->hs.opt.lyap_maxit;
  pa.stdev = dsd; pa.irfs = dirf; pa.status = dst;
  const size_t smem = sizeof(double) * (3 * (size_t)d.k * d.k + (size_t)d.k * d.nx + 8 + 2 * (size_t)d.n);
  long long grid = n < 148LL * 16 ? n : 148LL * 16;
  prior_predictive_kernel<<<(unsigned)grid, 128, smem, h->stream>>>(pa);
  DYB_CUDA(cudaGetLastError());
  h->launches += 2;
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

// persistent CTAs of the smoothers keep one period-by-period workspace each: fewer CTAs rather than more than
// SMOOTHER_WORKSPACE_BYTES of scratch (ns = 96, T = 500 is 37 MB per CTA)
constexpr size_t SMOOTHER_WORKSPACE_BYTES = (size_t)4 << 30;
static long long cap_grid_by_workspace(long long grid, size_t bytes_per_cta) {
  const long long fit = (long long)(SMOOTHER_WORKSPACE_BYTES / (bytes_per_cta ? bytes_per_cta : 1));
  if (fit < 1) return 1;
  return grid < fit ? grid : fit;
}

static int launch_smoother(KalmanSmootherArgs g, int device, cudaStream_t st, Stager& sg) {
  int dev_sms = 148, max_smem = 0;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  const size_t smem = sizeof(double) * smoother_smem_doubles(g.ns, g.ny, g.nx);
  if (smem + 1024 > (size_t)max_smem) return fail(DYB_ERR_UNSUPPORTED, "state space too large for the shared-memory smoother");
  long long grid = g.n_draws < (long long)dev_sms * 4 ? g.n_draws : (long long)dev_sms * 4;
  grid = cap_grid_by_workspace(grid, sizeof(double) * smoother_work_doubles(g.ns, g.ny, g.nobs));
  g.work = sg.temp<double>(smoother_work_doubles(g.ns, g.ny, g.nobs) * (size_t)grid);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  if (smem > 48 * 1024) DYB_CUDA(cudaFuncSetAttribute(kalman_smoother_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kalman_smoother_kernel<<<(unsigned)grid, 256, smem, st>>>(g);
  DYB_CUDA(cudaGetLastError());
  return DYB_OK;
}

int dyb_kalman_smoother_batch(int device, int32_t ns, int32_t ny, int32_t nx, int32_t nobs, const int32_t* z_idx,
                              const double* Y, const double* T, const double* R, const double* Q, const double* P0,
                              const double* a0, const double* H, int64_t n, double* alphah, double* etah, double* epsilonh,
                              double* att, double* a_pred, double* loglik, int32_t* status) {
  if (ns <= 0 || ny <= 0 || ny > 64 || ny > ns || nx <= 0 || nobs <= 0 || n < 0) return fail(DYB_ERR_ARG, "bad sizes");
  if (n == 0) return DYB_OK;
  if (!z_idx || !Y || !T || !R || !Q || !P0 || !status) return fail(DYB_ERR_ARG, "null argument");
  DYB_CUDA(cudaSetDevice(device));
  cudaStream_t st = device_stream(device);
  Stager sg(st);
  const size_t N = (size_t)n, ns2 = (size_t)ns * ns;
  KalmanSmootherArgs g{};
  g.ns = ns; g.ny = ny; g.nx = nx; g.nobs = nobs;
  g.z_idx = sg.in(z_idx, (size_t)ny);
  g.Y = sg.in(Y, (size_t)ny * nobs);
  g.T = sg.in(T, ns2 * N); g.R = sg.in(R, (size_t)ns * nx * N); g.Q = sg.in(Q, (size_t)nx * nx * N); g.P0 = sg.in(P0, ns2 * N);
  g.a0 = sg.in(a0, (size_t)ns * N); g.H = sg.in(H, (size_t)ny * ny * N);
  g.n_draws = n;
  g.alphah = sg.out(alphah, (size_t)ns * nobs * N); g.etah = sg.out(etah, (size_t)nx * nobs * N);
  g.epsilonh = sg.out(epsilonh, (size_t)ny * nobs * N); g.att = sg.out(att, (size_t)ns * nobs * N);
  g.a_pred = sg.out(a_pred, (size_t)ns * (nobs + 1) * N);
  g.loglik = sg.out(loglik, N); g.status = sg.out(status, N);
  if (!is_device_ptr(z_idx)) for (int i = 0; i < ny; ++i) if (z_idx[i] < 0 || z_idx[i] >= ns) return fail(DYB_ERR_ARG, "z_idx out of range");
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  int rc = launch_smoother(g, device, st, sg);
  if (rc) return rc;
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

int dyb_kalman_diffuse_smoother_batch(int device, int32_t ns, int32_t ny, int32_t nx, int32_t nobs, const int32_t* z_idx,
                                      const double* Y, const double* T, const double* R, const double* Q,
                                      const double* Pinf0, const double* Pstar0, const double* a0, const double* Hdiag,
                                      const double* ybar_obs, double tol, int64_t n, double* alphah, double* etah, double* epsilonh, double* att,
                                      double* a_pred, double* loglik, int32_t* n_diffuse, int32_t* status) {
  if (ns <= 0 || ny <= 0 || ny > 64 || ny > ns || nx <= 0 || nobs <= 0 || n < 0 || !(tol > 0.0)) return fail(DYB_ERR_ARG, "bad sizes");
  if (n == 0) return DYB_OK;
  if (!z_idx || !Y || !T || !R || !Q || !Pinf0 || !Pstar0 || !status) return fail(DYB_ERR_ARG, "null argument");
  DYB_CUDA(cudaSetDevice(device));
  cudaStream_t st = device_stream(device);
  Stager sg(st);
  const size_t N = (size_t)n, ns2 = (size_t)ns * ns;
  KalmanDiffuseArgs g{};
  g.ns = ns; g.ny = ny; g.nx = nx; g.nobs = nobs; g.tol = tol;
  g.z_idx = sg.in(z_idx, (size_t)ny);
  g.Y = sg.in(Y, (size_t)ny * nobs);
  g.T = sg.in(T, ns2 * N); g.R = sg.in(R, (size_t)ns * nx * N); g.Q = sg.in(Q, (size_t)nx * nx * N);
  g.Pinf0 = sg.in(Pinf0, ns2 * N); g.Pstar0 = sg.in(Pstar0, ns2 * N);
  g.a0 = sg.in(a0, (size_t)ns * N); g.Hdiag = sg.in(Hdiag, (size_t)ny * N); g.ybar_obs = sg.in(ybar_obs, (size_t)ny * N);
  g.n_draws = n;
  g.alphah = sg.out(alphah, (size_t)ns * nobs * N); g.etah = sg.out(etah, (size_t)nx * nobs * N);
  g.epsilonh = sg.out(epsilonh, (size_t)ny * nobs * N); g.att = sg.out(att, (size_t)ns * nobs * N);
  g.a_pred = sg.out(a_pred, (size_t)ns * (nobs + 1) * N);
  g.loglik = sg.out(loglik, N); g.n_diffuse = sg.out(n_diffuse, N); g.status = sg.out(status, N);
  if (!is_device_ptr(z_idx)) for (int i = 0; i < ny; ++i) if (z_idx[i] < 0 || z_idx[i] >= ns) return fail(DYB_ERR_ARG, "z_idx out of range");
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  int dev_sms = 148, max_smem = 0;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  const size_t smem = sizeof(double) * diffuse_smem_doubles(ns, nx);
  if (smem + 1024 > (size_t)max_smem) return fail(DYB_ERR_UNSUPPORTED, "state space too large for the shared-memory diffuse smoother");
  long long grid = n < (long long)dev_sms * 2 ? n : (long long)dev_sms * 2;
  grid = cap_grid_by_workspace(grid, sizeof(double) * diffuse_work_doubles(ns, ny, nobs));
  g.work = sg.temp<double>(diffuse_work_doubles(ns, ny, nobs) * (size_t)grid);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  if (smem > 48 * 1024) DYB_CUDA(cudaFuncSetAttribute(kalman_diffuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kalman_diffuse_kernel<<<(unsigned)grid, 256, smem, st>>>(g);
  DYB_CUDA(cudaGetLastError());
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

int dyb_shock_covariances_batch(dyb_handle* h, const double* theta, int64_t n, double* sigma_e, double* sigma_m) {
  if (!h || n < 0 || (n > 0 && (!theta || (!sigma_e && !sigma_m)))) return fail(DYB_ERR_ARG, "bad argument");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  const dyb_model_dims& d = h->dims;
  const size_t N = (size_t)n;
  Stager sg(h->stream);
  const double* dth = sg.in(theta, (size_t)h->hs.np * N);
  double* dse = sg.out(sigma_e, (size_t)d.nx * d.nx * N);
  double* dsm = sg.out(sigma_m, (size_t)d.ny * d.ny * N);
  if (!dse) dse = sg.temp<double>((size_t)d.nx * d.nx * N);
  double* dse0 = sg.temp<double>((size_t)d.nx * d.nx);
  double* dsm0 = sg.temp<double>((size_t)d.ny * d.ny);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  DYB_CUDA(cudaMemcpyAsync(dse0, h->hs.sigma_e.data(), sizeof(double) * d.nx * d.nx, cudaMemcpyHostToDevice, h->stream));
  DYB_CUDA(cudaMemcpyAsync(dsm0, h->hs.sigma_m.data(), sizeof(double) * d.ny * d.ny, cudaMemcpyHostToDevice, h->stream));
  SigmaEArgs sa{};
  sa.np = h->hs.np; sa.nx = d.nx; sa.ny = d.ny;
  for (int q = 0; q < h->hs.np; ++q) { sa.est_type[q] = h->hs.est_type[q]; sa.est_i[q] = h->hs.est_i[q]; sa.est_j[q] = h->hs.est_j[q]; }
  sigma_e_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(sa, dse0, dsm0, dth, n, dse, dsm);
  DYB_CUDA(cudaGetLastError());
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

int dyb_smoother_batch(dyb_handle* h, const double* theta, int64_t n, double* alphah, double* etah, double* epsilonh,
                       double* loglik, int32_t* status) {
  if (h && !h->ops) return fail(DYB_ERR_UNSUPPORTED, "needs a handle with a device model function");
  if (!h || n < 0 || (n > 0 && (!theta || !status))) return fail(DYB_ERR_ARG, "bad argument");
  if (h->d_Y.p == nullptr) return fail(DYB_ERR_STATE, "dyb_set_observations has not been called");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  const dyb_model_dims& d = h->dims;
  if (d.k > 64) return fail(DYB_ERR_UNSUPPORTED, "more than 64 state variables");
  const size_t N = (size_t)n, n2 = (size_t)d.n * d.n;
  Stager sg(h->stream);
  double* dg1 = sg.temp<double>((size_t)d.n * (d.k + d.nx) * N);
  double* dyb_ = sg.temp<double>((size_t)d.n * N);
  double* dse = sg.temp<double>((size_t)d.nx * d.nx * N);
  double* dsm = sg.temp<double>((size_t)d.ny * d.ny * N);
  double* dse0 = sg.temp<double>((size_t)d.nx * d.nx);
  double* dsm0 = sg.temp<double>((size_t)d.ny * d.ny);
  int* dst1 = sg.temp<int>(N);
  int* dst2 = sg.temp<int>(N);
  double* dT = sg.temp<double>(n2 * N); double* dR = sg.temp<double>((size_t)d.n * d.nx * N); double* dP = sg.temp<double>(n2 * N);
  double* dyo = sg.temp<double>((size_t)d.ny * N);
  int* dz = sg.temp<int>((size_t)d.ny);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  DYB_CUDA(cudaMemcpyAsync(dse0, h->hs.sigma_e.data(), sizeof(double) * d.nx * d.nx, cudaMemcpyHostToDevice, h->stream));
  DYB_CUDA(cudaMemcpyAsync(dsm0, h->hs.sigma_m.data(), sizeof(double) * d.ny * d.ny, cudaMemcpyHostToDevice, h->stream));
  int rc = dyb_first_order_batch(h, theta, n, dg1, dyb_, nullptr, nullptr, reinterpret_cast<int32_t*>(dst1));
  if (rc) return rc;
  const double* dth = sg.in(theta, (size_t)h->hs.np * N);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  SigmaEArgs sa{};
  sa.np = h->hs.np; sa.nx = d.nx; sa.ny = d.ny;
  for (int q = 0; q < h->hs.np; ++q) { sa.est_type[q] = h->hs.est_type[q]; sa.est_i[q] = h->hs.est_i[q]; sa.est_j[q] = h->hs.est_j[q]; }
  sigma_e_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(sa, dse0, dsm0, dth, n, dse, dsm);
  FullStateArgs fa{};
  fa.n = d.n; fa.k = d.k; fa.nx = d.nx; fa.ny = d.ny;
  h->ops->bkwrd_indices(fa.bk);
  h->ops->obs_indices(fa.obs);
  fa.g1 = dg1; fa.sigma_e = dse; fa.ybar = dyb_; fa.status_in = dst1; fa.n_draws = n;
  fa.lyap_tol = h->hs.opt.lyap_tol; fa.lyap_maxit = h->hs.opt.lyap_maxit;
  fa.T = dT; fa.R = dR; fa.P0 = dP; fa.ybar_obs = dyo; fa.status = dst2;
  const size_t smem_fs = sizeof(double) * (3 * (size_t)d.k * d.k + (size_t)d.k * d.nx + 8);
  long long grid = n < 148LL * 16 ? n : 148LL * 16;
  full_state_space_kernel<<<(unsigned)grid, 128, smem_fs, h->stream>>>(fa);
  DYB_CUDA(cudaGetLastError());
  DYB_CUDA(cudaMemcpyAsync(dz, fa.obs, sizeof(int) * d.ny, cudaMemcpyHostToDevice, h->stream));
  DYB_CUDA(cudaStreamSynchronize(h->stream));                 // fa.obs lives on this stack frame
  KalmanSmootherArgs g{};
  g.ns = d.n; g.ny = d.ny; g.nx = d.nx; g.nobs = h->nobs;
  g.z_idx = dz; g.Y = h->d_Y.as<double>(); g.T = dT; g.R = dR; g.Q = dse; g.P0 = dP; g.a0 = nullptr; g.H = dsm;
  g.ybar_obs = dyo; g.level = dyb_; g.status_in = dst2; g.n_draws = n;
  g.alphah = sg.out(alphah, (size_t)d.n * h->nobs * N); g.etah = sg.out(etah, (size_t)d.nx * h->nobs * N);
  g.epsilonh = sg.out(epsilonh, (size_t)d.ny * h->nobs * N);
  g.loglik = sg.out(loglik, N); g.status = sg.out(status, N);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  rc = launch_smoother(g, h->device, h->stream, sg);
  if (rc) return rc;
  h->launches += 3;
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

int dyb_forecast_batch(dyb_handle* h, const double* theta, int64_t n, int32_t periods, const double* y0, double* Y, int32_t* status) {
  if (h && !h->ops) return fail(DYB_ERR_UNSUPPORTED, "needs a handle with a device model function");
  if (!h || n < 0 || periods < 0 || (n > 0 && (!theta || !y0 || !Y || !status))) return fail(DYB_ERR_ARG, "bad argument");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  const dyb_model_dims& d = h->dims;
  if (d.k > 64) return fail(DYB_ERR_UNSUPPORTED, "more than 64 state variables");
  const size_t N = (size_t)n;
  Stager sg(h->stream);
  double* dg1 = sg.temp<double>((size_t)d.n * (d.k + d.nx) * N);
  double* dyb_ = sg.temp<double>((size_t)d.n * N);
  int* dst = sg.out(status, N);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  int rc = dyb_first_order_batch(h, theta, n, dg1, dyb_, nullptr, nullptr, reinterpret_cast<int32_t*>(dst));
  if (rc) return rc;
  ForecastArgs fa{};
  fa.n = d.n; fa.k = d.k; fa.periods = periods;
  h->ops->bkwrd_indices(fa.bk);
  fa.g1 = dg1; fa.g1_stride = (long long)d.n * (d.k + d.nx); fa.ybar = dyb_; fa.y0 = sg.in(y0, (size_t)d.n * N); fa.status_in = dst; fa.n_draws = n;
  fa.Y = sg.out(Y, (size_t)d.n * (periods + 1) * N);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  forecast_kernel<<<(unsigned)((n * 32 + 127) / 128), 128, 0, h->stream>>>(fa);
  DYB_CUDA(cudaGetLastError());
  h->launches += 1;
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

int dyb_lyapunov_batch(int device, int32_t k, const double* A, const double* B, int64_t n, double tol, int32_t maxit,
                       double* X, int32_t* iters, int32_t* status) {
  if (k < 1 || n < 0 || maxit < 1 || !(tol >= 0) || (n > 0 && (!A || !B || !X || !status))) return fail(DYB_ERR_ARG, "bad argument");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(device));
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  const size_t smem = sizeof(double) * (3 * (size_t)k * k + 8);
  if (smem + 1024 > (size_t)max_smem) return fail(DYB_ERR_UNSUPPORTED, "k too large for the shared-memory Lyapunov solver");
  cudaStream_t st = device_stream(device);
  Stager sg(st);
  const size_t N = (size_t)n, k2 = (size_t)k * k;
  LyapunovArgs a{k, sg.in(A, k2 * N), sg.in(B, k2 * N), sg.out(X, k2 * N), sg.out(iters, N), sg.out(status, N), n, tol, maxit};
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  if (smem > 48 * 1024) DYB_CUDA(cudaFuncSetAttribute(lyapunov_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = n < 148LL * 64 ? n : 148LL * 64;
  lyapunov_generic_kernel<<<(unsigned)grid, 128, smem, st>>>(a);
  DYB_CUDA(cudaGetLastError());
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

}  // extern "C"
