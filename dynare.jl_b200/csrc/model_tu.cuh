// Included by each generated model translation unit after its model header:
//   #include "generated/model_<name>.cuh"
//   #define DYB_MODEL dyb::Model_<name>
//   #include "model_tu.cuh"
// Instantiates the solve / filter kernels for the model and registers its launchers.
#include <cstdio>
#include <cstdlib>
#include "registry.h"
#include "solve_kernel.cuh"
#include "solve_coop.cuh"
#include "kalman_kernel.cuh"
#include "kalman_sublanes.cuh"
#include "kalman_coop.cuh"

namespace dyb {
namespace {

using MT = DYB_MODEL;

void mt_defaults(HostState& hs) {
  hs.params.resize(MT::NPAR);
  for (int i = 0; i < MT::NPAR; ++i) hs.params[i] = MT::param0(i);
  hs.sigma_e.resize(MT::NX * MT::NX);
  for (int i = 0; i < MT::NX * MT::NX; ++i) hs.sigma_e[i] = MT::sigma_e0(i);
  hs.sigma_m.resize(MT::NY * MT::NY);
  for (int i = 0; i < MT::NY * MT::NY; ++i) hs.sigma_m[i] = MT::sigma_m0(i);
  hs.np = MT::NP_DEFAULT;
  for (int i = 0; i < MT::NP_DEFAULT && i < MAX_EST; ++i) {
    hs.est_type[i] = MT::est_type0(i); hs.est_i[i] = MT::est_i0(i); hs.est_j[i] = MT::est_j0(i);
  }
}

constexpr bool mt_coop_fits = CrCfg<MT>::FITS;
// Models beyond the per-thread / register-resident kernels (dynamic block > 12 or more than 16 Kalman states): none of
// the model-templated solve / filter kernels is instantiated (their unrolled code would take tens of minutes to compile
// and could not launch); the generated model function feeds the run-time-size solver and the dense tensor-core filter
// instead (ModelOps::jacobian_batch, generic.cu).
constexpr bool mt_big = (MT::M > 12) || (MT::NS > 16);
constexpr bool mt_jac_fed = is_jac_fed<MT>::value;      // sizes-only model: Jacobians come from the host (modelgen.emit_cuda_jacobian_fed)
constexpr bool mt_fast_fits = CrCfg<MT>::FITS_LU && !mt_jac_fed;
int mt_solve_launches(const HostState& hs) {
  return mt_big ? 2 : ((mt_coop_fits && hs.opt.solver != 1) ? ((mt_fast_fits && hs.opt.solver == 0) ? 4 : 3) : 1);
}

// FNV-1a over what the calibration draw of the pivot order depends on
unsigned long long mt_perm_key(const HostState& hs) {
  unsigned long long k = 1469598103934665603ULL;
  auto mix = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { k ^= b[i]; k *= 1099511628211ULL; } };
  mix(hs.params.data(), 8 * hs.params.size()); mix(hs.sigma_e.data(), 8 * hs.sigma_e.size()); mix(hs.sigma_m.data(), 8 * hs.sigma_m.size());
  mix(&hs.opt.steady_tol, 8);
  return k | 1ULL;
}

cudaError_t mt_solve(const HostState& hs, const double* d_theta, long long n, SolveOut out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if constexpr (mt_big || mt_jac_fed) { return cudaErrorNotSupported; } else {
  SolveParams<MT> prm;
  for (int i = 0; i < MT::NPAR; ++i) prm.params0[i] = hs.params[i];
  for (int i = 0; i < MT::NX * MT::NX; ++i) prm.sigma_e0[i] = hs.sigma_e[i];
  for (int i = 0; i < MT::NY * MT::NY; ++i) prm.sigma_m0[i] = hs.sigma_m[i];
  prm.np = hs.np;
  for (int i = 0; i < MAX_EST; ++i) { prm.est_type[i] = hs.est_type[i]; prm.est_i[i] = hs.est_i[i]; prm.est_j[i] = hs.est_j[i]; }
  prm.cr_tol = hs.opt.cr_tol; prm.cr_maxit = hs.opt.cr_maxit; prm.steady_tol = hs.opt.steady_tol;
  prm.lyap_tol = hs.opt.lyap_tol; prm.lyap_maxit = hs.opt.lyap_maxit;
  const int block = 128;
  const long long grid = (n + block - 1) / block;
  constexpr int G = CrCfg<MT>::G;
  const bool coop = mt_coop_fits && hs.opt.solver != 1;
  if constexpr (mt_coop_fits) { if (coop) {
    model_kernel<MT><<<(unsigned)((n + DYB_MODEL_BLOCK - 1) / DYB_MODEL_BLOCK), DYB_MODEL_BLOCK, 0, st>>>(prm, d_theta, n, out.mid, out.status, out.ybar);
    if (out.ev_after_model) cudaEventRecord(out.ev_after_model, st);
    const long long gridc = (n * G + CrCfg<MT>::BLOCK - 1) / CrCfg<MT>::BLOCK;
    CrOptions co{hs.opt.cr_tol, hs.opt.cr_maxit};
    bool fast = false;
    if constexpr (mt_fast_fits) {
      // fast pass (no pivoting, verified) + Householder pass on what it could not verify; needs the pivot order of the
      // calibration point, derived once per configuration (not inside a stream capture: it allocates)
      CrPermCache& c = hs.crp;
      fast = hs.opt.solver == 0;
      const unsigned long long key = mt_perm_key(hs);
      if (fast && !(c.valid && c.key == key)) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        if (cap != cudaStreamCaptureStatusNone && !c.d_perm) fast = false;
        else {
          if (!c.d_perm) {
            if (cudaMalloc(&c.d_perm, sizeof(int) * (MT::M + 1)) != cudaSuccess || cudaMalloc(&c.d_mid1, sizeof(double) * MidLayout<MT>::SIZE) != cudaSuccess ||
                cudaMalloc(&c.d_st1, sizeof(int)) != cudaSuccess || cudaMalloc(&c.d_redo, 2 * sizeof(unsigned long long)) != cudaSuccess) {
              c.release(); cudaGetLastError(); fast = false;
            } else cudaMemsetAsync(c.d_redo, 0, 2 * sizeof(unsigned long long), st);
          }
          if (fast) {
            SolveParams<MT> prm0 = prm;
            prm0.np = 0;                              // the calibration itself: no estimated parameter overrides it
            model_kernel<MT><<<1, 32, 0, st>>>(prm0, nullptr, 1, c.d_mid1, c.d_st1, nullptr);
            cr_perm_kernel<MT><<<1, 32, 0, st>>>(c.d_mid1, c.d_st1, c.d_perm);
            if (cap == cudaStreamCaptureStatusNone) { c.key = key; c.valid = true; }    // a captured launch has not run yet
            if (cap == cudaStreamCaptureStatusNone && getenv("DYB_DEBUG_PERM")) {
              int hp[MT::M], hst = -1;
              cudaStreamSynchronize(st);
              cudaMemcpy(hp, c.d_perm, sizeof(hp), cudaMemcpyDeviceToHost); cudaMemcpy(&hst, c.d_st1, sizeof(int), cudaMemcpyDeviceToHost);
              fprintf(stderr, "[dyb] %s: calibration status %d, equation order", MT::name, hst);
              for (int r = 0; r < MT::M; ++r) fprintf(stderr, " %d", hp[r]);
              fprintf(stderr, "\n");
            }
          }
        }
      }
      if (fast && c.list_cap < n) {                   // the list of draws for the second pass grows with the widest batch seen
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        if (cap != cudaStreamCaptureStatusNone) fast = false;
        else {
          cudaStreamSynchronize(st);                  // a previous call may still be using the old list
          if (c.d_list) cudaFree(c.d_list);
          c.d_list = nullptr; c.list_cap = 0;
          if (cudaMalloc(&c.d_list, sizeof(long long) * (size_t)n) == cudaSuccess) c.list_cap = n;
          else { cudaGetLastError(); fast = false; }
        }
      }
      if (fast) {
        cudaMemsetAsync(c.d_redo + 1, 0, sizeof(unsigned long long), st);
        cr_kernel<MT, G, CR_LU, false><<<(unsigned)gridc, CrCfg<MT>::BLOCK, 0, st>>>(out.mid, n, co, out.status, out.cr_iters, c.d_perm, c.d_redo, c.d_list);
        cr_kernel<MT, G, CR_QR, true><<<(unsigned)gridc, CrCfg<MT>::BLOCK, 0, st>>>(out.mid, n, co, out.status, out.cr_iters, nullptr, c.d_redo, c.d_list);
      }
    }
    if (!fast)
      cr_kernel<MT, G, (DYB_CR_SOLVER == 1 ? CR_GJ : CR_QR), false><<<(unsigned)gridc, CrCfg<MT>::BLOCK, 0, st>>>(out.mid, n, co, out.status, out.cr_iters, nullptr, nullptr, nullptr);
    if (out.ev_after_cr) cudaEventRecord(out.ev_after_cr, st);
    assemble_kernel<MT><<<(unsigned)((n + DYB_ASM_BLOCK - 1) / DYB_ASM_BLOCK), DYB_ASM_BLOCK, 0, st>>>(out.mid, n, hs.opt.lyap_tol, hs.opt.lyap_maxit, out.ss,
                                                          out.status, out.g1, out.lyap_iters);
    return cudaGetLastError();
  } }
  (void)coop;
  SolveExtra ex{out.g1, out.ybar, out.cr_iters, out.lyap_iters};
  solve_kernel<MT><<<(unsigned)grid, block, 0, st>>>(prm, d_theta, n, out.ss, out.status, ex);
  return cudaGetLastError();
  }
}

// sizes-only models: host-evaluated Jacobians -> static elimination -> the Householder cyclic reduction -> assembly
cudaError_t mt_solve_from_jacobian(const HostState& hs, const double* d_jac, const double* d_ybar, const double* d_se,
                                   int se_per_draw, const double* d_sm, int sm_per_draw, const int* d_status_in,
                                   long long n, SolveOut out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if constexpr (!(mt_jac_fed && mt_coop_fits)) { return cudaErrorNotSupported; } else {
  constexpr int G = CrCfg<MT>::G;
  jac_model_kernel<MT><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_jac, d_ybar, d_se, se_per_draw, d_sm, sm_per_draw,
                                                                     d_status_in, n, out.mid, out.status);
  if (out.ev_after_model) cudaEventRecord(out.ev_after_model, st);
  const long long gridc = (n * G + CrCfg<MT>::BLOCK - 1) / CrCfg<MT>::BLOCK;
  CrOptions co{hs.opt.cr_tol, hs.opt.cr_maxit};
  cr_kernel<MT, G, CR_QR, false><<<(unsigned)gridc, CrCfg<MT>::BLOCK, 0, st>>>(out.mid, n, co, out.status, out.cr_iters, nullptr, nullptr, nullptr);
  if (out.ev_after_cr) cudaEventRecord(out.ev_after_cr, st);
  assemble_kernel<MT><<<(unsigned)((n + DYB_ASM_BLOCK - 1) / DYB_ASM_BLOCK), DYB_ASM_BLOCK, 0, st>>>(out.mid, n, hs.opt.lyap_tol, hs.opt.lyap_maxit, out.ss,
                                                                                                   out.status, out.g1, out.lyap_iters);
  return cudaGetLastError();
  }
}

// G lanes per draw for small batches (kalman_sublanes.cuh)
template <int G>
cudaError_t mt_kalman_sub(const double* d_ss, long long n, const double* d_Y, int nobs, int presample,
                          const int* d_status_in, double* d_loglik, int* d_status_out, cudaStream_t st) {
  if constexpr (!KalmanSubCfg<MT, G>::FITS) { return cudaErrorNotSupported; } else {
  using C = KalmanSubCfg<MT, G>;
  const size_t smem = C::smem_bytes(nobs);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kalman_sublanes_kernel<MT, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const long long grid = (n * G + C::BLOCK - 1) / C::BLOCK;
  kalman_sublanes_kernel<MT, G><<<(unsigned)grid, C::BLOCK, smem, st>>>(d_ss, n, d_Y, nobs, presample, d_status_in, d_loglik, d_status_out);
  return cudaGetLastError();
  }
}

// G = 16 / 32 lanes per draw for the smallest batches (kalman_coop.cuh)
template <int G>
cudaError_t mt_kalman_coop(const double* d_ss, long long n, const double* d_Y, int nobs, int presample,
                           const int* d_status_in, double* d_loglik, int* d_status_out, cudaStream_t st) {
  if constexpr (!KalmanCoopCfg<MT, G>::FITS) { return cudaErrorNotSupported; } else {
  using C = KalmanCoopCfg<MT, G>;
  const size_t smem = C::smem_bytes(nobs);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kalman_coop_kernel<MT, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const long long grid = (n + C::DRAWS - 1) / C::DRAWS;
  kalman_coop_kernel<MT, G><<<(unsigned)grid, C::BLOCK, smem, st>>>(d_ss, n, d_Y, nobs, presample, d_status_in, d_loglik, d_status_out);
  return cudaGetLastError();
  }
}

// lanes per draw of the filter: 1 = one thread per draw (throughput kernel), 4 / 8 = shared-tile kernel with the time update
// dealt over the lanes, 16 / 32 = every stage of a period dealt over the lanes (kalman_coop.cuh), for batches that leave
// schedulers idle; `want` = dyb_options.filter_lanes (0 = choose by batch width).  Measured with scripts/small_batch_latency.py
// (profiles/r02_n_small_batch_coop.txt), ls2003 filter at 2 048 draws without missing observations: 0.107 ms with 16 lanes,
// 0.129 with 32, 0.146 with 8, 0.154 with one thread per draw; the 16-lane kernel keeps 2 368 draws resident (242 registers),
// beyond that a second wave starts (4 096 draws: 0.203 ms) and one thread per draw is the fastest (0.154 ms up to 16 384)
#ifndef DYB_COOP32_MAX
#define DYB_COOP32_MAX 0
#endif
#ifndef DYB_COOP16_MAX
#define DYB_COOP16_MAX 2304
#endif
int mt_kalman_lanes(long long n, int nobs, int want) {
  const bool fits32 = KalmanCoopCfg<MT, 32>::FITS && KalmanCoopCfg<MT, 32>::smem_bytes(nobs) <= 200 * 1024;
  const bool fits16 = KalmanCoopCfg<MT, 16>::FITS && KalmanCoopCfg<MT, 16>::smem_bytes(nobs) <= 200 * 1024;
  int g = want;
  if (g != 1 && g != 4 && g != 8 && g != 16 && g != 32) {
    if (n <= DYB_COOP32_MAX && fits32) g = 32;
    else if (n <= DYB_COOP16_MAX && fits16) g = 16;
    else g = 1;      // without missing observations one thread per draw (0.154 ms up to 16 384 ls2003 draws) beats 4 / 8 lanes
  }
  if ((g == 32 && !fits32) || (g == 16 && !fits16)) g = 1;
  if (g == 4 && (!KalmanSubCfg<MT, 4>::FITS || KalmanSubCfg<MT, 4>::smem_bytes(nobs) > 200 * 1024)) g = 8;
  if (g == 8 && (!KalmanSubCfg<MT, 8>::FITS || KalmanSubCfg<MT, 8>::smem_bytes(nobs) > 200 * 1024)) g = 1;
  return g;
}

cudaError_t mt_kalman(const double* d_ss, long long n, const double* d_Y, int nobs, int presample, int want_lanes,
                      const int* d_status_in, double* d_loglik, int* d_status_out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if constexpr (mt_big) { return cudaErrorNotSupported; } else {
  const int lanes = mt_kalman_lanes(n, nobs, want_lanes);
  if (lanes == 32) return mt_kalman_coop<32>(d_ss, n, d_Y, nobs, presample, d_status_in, d_loglik, d_status_out, st);
  if (lanes == 16) return mt_kalman_coop<16>(d_ss, n, d_Y, nobs, presample, d_status_in, d_loglik, d_status_out, st);
  if (lanes == 8) return mt_kalman_sub<8>(d_ss, n, d_Y, nobs, presample, d_status_in, d_loglik, d_status_out, st);
  if (lanes == 4) return mt_kalman_sub<4>(d_ss, n, d_Y, nobs, presample, d_status_in, d_loglik, d_status_out, st);
  using Cfg = KalmanCfg<MT>;
  const int block = Cfg::BLOCK;
  const long long grid = (n + block - 1) / block;
  const size_t smem = sizeof(double) * ((size_t)MT::NY * (size_t)nobs + (size_t)Cfg::SMEM_DOUBLES_PER_THREAD * block);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kalman_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kalman_kernel<MT><<<(unsigned)grid, block, smem, st>>>(d_ss, n, d_Y, nobs, presample, d_status_in, d_loglik, d_status_out);
  return cudaGetLastError();
  }
}

__global__ void mt_unpack_kernel(const double* __restrict__ ss, long long n, double* T, double* RQR, double* P0,
                                 double* yb, double* H) {
  using L = SSLayout<MT>;
  constexpr int NS = MT::NS, K = MT::K, NY = MT::NY;
  const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n) return;
  const double* s = ss + d;
  if (T) {
    double* t = T + d * NS * NS;
    for (int i = 0; i < NS * NS; ++i) t[i] = 0.0;
    for (int r = 0; r < NS; ++r) for (int j = 0; j < K; ++j) t[r + NS * MT::lagged_state_ids(j)] = s[(L::OFF_T + r * K + j) * n];
  }
  if (RQR) for (int r = 0; r < NS; ++r) for (int c = 0; c < NS; ++c) RQR[d * NS * NS + r + NS * c] = s[(L::OFF_QQ + sym(r, c)) * n];
  if (P0) for (int r = 0; r < NS; ++r) for (int c = 0; c < NS; ++c) P0[d * NS * NS + r + NS * c] = s[(L::OFF_P0 + sym(r, c)) * n];
  if (yb) for (int i = 0; i < NY; ++i) yb[d * NY + i] = s[(L::OFF_YB + i) * n];
  if (H) for (int r = 0; r < NY; ++r) for (int c = 0; c < NY; ++c) H[d * NY * NY + r + NY * c] = s[(L::OFF_H + sym(r, c)) * n];
}

cudaError_t mt_unpack(const double* d_ss, long long n, double* T, double* RQR, double* P0, double* yb, double* H, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (mt_big) return cudaErrorNotSupported;
  const int block = 128;
  mt_unpack_kernel<<<(unsigned)((n + block - 1) / block), block, 0, st>>>(d_ss, n, T, RQR, P0, yb, H);
  return cudaGetLastError();
}

void mt_bkwrd_indices(int* out) { for (int i = 0; i < MT::K; ++i) out[i] = MT::i_bkwrd_b(i); }

void mt_obs_indices(int* out) { for (int i = 0; i < MT::NY; ++i) out[i] = MT::state_ids(MT::obs_idx_state(i)); }

void mt_static_indices(int* out) { for (int i = 0; i < MT::S; ++i) out[i] = MT::i_static(i); }
void mt_fwrd_indices(int* out) { for (int i = 0; i < MT::NF; ++i) out[i] = MT::i_fwrd_b(i); }

// theta -> dense compressed Jacobian, steady state, Sigma_e, Sigma_m, status: one thread per draw (K1-K2 of solve_draw with
// the Jacobian written out instead of consumed); feeds solve_generic_kernel for models beyond the register-resident solver
struct JacDense {
  double* w;
  DYB_D double& operator[](int idx) const { return w[idx]; }
};
__global__ void __launch_bounds__(128)
mt_jacobian_kernel(const SolveParams<MT> prm, const double* __restrict__ theta, long long n_draws, double* __restrict__ jac,
                   double* __restrict__ ybar, double* __restrict__ sigma_e, double* __restrict__ sigma_m, int* __restrict__ status) {
  constexpr int N = MT::N, NX = MT::NX, NY = MT::NY, NCOL = MT::NCOL;
  const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_draws) return;
  const double* th = theta + d * (long long)prm.np;
  double p[MT::NPAR > 0 ? MT::NPAR : 1];
  double* Se = sigma_e + d * NX * NX;
  double* Sm = sigma_m + d * NY * NY;
  for (int i = 0; i < MT::NPAR; ++i) p[i] = prm.params0[i];
  for (int i = 0; i < NX * NX; ++i) Se[i] = prm.sigma_e0[i];
  for (int i = 0; i < NY * NY; ++i) Sm[i] = prm.sigma_m0[i];
  for (int q = 0; q < prm.np; ++q) {
    const double v = th[q];
    const int ty = prm.est_type[q], i = prm.est_i[q], j = prm.est_j[q];
    switch (ty) {
      case EST_PARAM: p[i] = v; break;
      case EST_SD_SHOCK: Se[i + NX * j] = v * v; break;
      case EST_VAR_SHOCK: Se[i + NX * j] = v; break;
      case EST_CORR_SHOCK: { const double c = v * sqrt(Se[i + NX * i] * Se[j + NX * j]); Se[i + NX * j] = c; Se[j + NX * i] = c; } break;
      case EST_SD_MEAS: Sm[i + NY * j] = v * v; break;
      case EST_VAR_MEAS: Sm[i + NY * j] = v; break;
      case EST_CORR_MEAS: { const double c = v * sqrt(Sm[i + NY * i] * Sm[j + NY * j]); Sm[i + NY * j] = c; Sm[j + NY * i] = c; } break;
      default: break;
    }
  }
  double* y = ybar + d * N;
  MT::steady_state(p, y);
  bool ok = true;
  for (int i = 0; i < N; ++i) ok = ok && isfinite(y[i]);
  if (ok) ok = MT::static_resid_ss(p, y) <= prm.steady_tol * prm.steady_tol;      // SteadyState.jl:476-482
  double* J = jac + d * (long long)(N * NCOL);
  for (int e = 0; e < N * NCOL; ++e) J[e] = 0.0;
  int st = ok ? ST_OK : ST_STEADY;
  if (ok) {
    JacDense acc{J};
    MT::jacobian(p, y, acc);
    double chk = 0.0;
    for (int e = 0; e < N * NCOL; ++e) chk = fma(J[e], 0.0, chk);
    if (!(chk == 0.0)) st = ST_STEADY;                  // NaN / Inf in the Jacobian
  }
  status[d] = st;
}

cudaError_t mt_jacobian_batch(const HostState& hs, const double* d_theta, long long n, double* d_jac, double* d_ybar,
                              double* d_se, double* d_sm, int* d_status, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  SolveParams<MT> prm;
  for (int i = 0; i < MT::NPAR; ++i) prm.params0[i] = hs.params[i];
  for (int i = 0; i < MT::NX * MT::NX; ++i) prm.sigma_e0[i] = hs.sigma_e[i];
  for (int i = 0; i < MT::NY * MT::NY; ++i) prm.sigma_m0[i] = hs.sigma_m[i];
  prm.np = hs.np;
  for (int i = 0; i < MAX_EST; ++i) { prm.est_type[i] = hs.est_type[i]; prm.est_i[i] = hs.est_i[i]; prm.est_j[i] = hs.est_j[i]; }
  prm.cr_tol = hs.opt.cr_tol; prm.cr_maxit = hs.opt.cr_maxit; prm.steady_tol = hs.opt.steady_tol;
  prm.lyap_tol = hs.opt.lyap_tol; prm.lyap_maxit = hs.opt.lyap_maxit;
  mt_jacobian_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(prm, d_theta, n, d_jac, d_ybar, d_se, d_sm, d_status);
  return cudaGetLastError();
}

void mt_obs_state_positions(int* out) { for (int i = 0; i < MT::NY; ++i) out[i] = MT::obs_idx_state(i); }

const ModelOps mt_ops = {
  MT::name,
  {MT::N, MT::NX, MT::NPAR, MT::K, MT::NF, MT::S, MT::NCOL, MT::NS, MT::NY, MT::NP_DEFAULT, SSLayout<MT>::SIZE, MidLayout<MT>::SIZE},
  MidLayout<MT>::SIZE,
  mt_defaults, mt_solve_launches, mt_solve, mt_kalman, mt_unpack, mt_bkwrd_indices, mt_obs_indices,
  mt_obs_state_positions, !mt_big, mt_static_indices, mt_fwrd_indices, mt_jacobian_batch,
  (mt_jac_fed && mt_coop_fits) ? mt_solve_from_jacobian : nullptr
};

struct Registrar { Registrar() { register_model(&mt_ops); } } mt_registrar;

}  // namespace
}  // namespace dyb
