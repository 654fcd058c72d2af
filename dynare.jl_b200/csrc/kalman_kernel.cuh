// Fused Kalman-filter log-likelihood, model-specialised: one thread per draw, the whole
// T-step recursion in one launch, state (a, P) and the draw's transition data in registers.
//
// Replaces the reference's `kalman_likelihood(Y,Z,H,T,R,Q,a0,P,start,last,presample,ws[,data_pattern])`
// call (src/estimation/estimation.jl:665-701; KalmanFilterTools 0.1.x, not vendored):
//     v = y - Z a ; F = Z P Z' + H ; lik_t = log det F + v' F^-1 v ; K = P Z' F^-1
//     a <- T (a + K v) ; P <- T (P - K Z P) T' + R Q R'
//     loglik = -1/2 ( n_observed * log 2pi + sum_{t >= presample} lik_t )
// Structure exploited (none of it changes the result beyond rounding):
//   * Z is a selection (estimation.jl:644-646): Z a and Z P Z' are gathers;
//   * T has non-zero columns only at the K lagged states (estimation.jl:648-649): the time update
//     only needs the updated mean / covariance on those K states:  a <- Tl a_l,  P <- Tl P_ll Tl' + RQR';
//   * P is symmetric: packed lower storage, half of the products.
//   * the observations (and therefore the missing-data pattern) are shared by every draw, so the
//     `data_pattern` variant costs no divergence: a missing row i gets v_i = 0, F_ii = 1, F_ij = 0 and
//     a zero column in P Z', which removes it from the likelihood and from the update exactly.
#pragma once
#include "model_traits.cuh"
#include "solve_kernel.cuh"

namespace dyb {

// minimum CTAs per SM the register allocation aims at.  Measured on the headline (bench.py, two batches in flight) after
// the no-missing-observation instantiation: 8 -> 104.4, 7 -> 103.7-104.4, 6 -> 107.0, 5 -> 107.7 M evals/s; at 7 and 8
// the kernel spills (200+ bytes per thread, re-read every period) and holds the whole register file while it runs;
// hs_nk filter per 65 536 draws: 0.324 ms at 7, 0.195 ms at 5; ls2003 (P in shared memory, one-warp CTAs) unchanged
#ifndef DYB_KF_MIN_CTAS
#define DYB_KF_MIN_CTAS 5
#endif

template <class M>
struct KalmanCfg {
  // packed R Q R' goes to shared memory ([element][thread], conflict-free) to keep the register count low.  Models whose
  // packed covariance would not fit the register file beside T_l (ls2003: 55 + 55 + 60 doubles) keep P in shared memory as
  // well and use one-warp CTAs.
  static constexpr int QQ_N = M::NS * (M::NS + 1) / 2;
  static constexpr bool P_SMEM = (QQ_N + M::NS * M::K > 96);
  static constexpr int BLOCK = P_SMEM ? 32 : 64;
  static constexpr bool QQ_SMEM = P_SMEM || (QQ_N * BLOCK * 8 <= 16 * 1024);
  static constexpr int SMEM_DOUBLES_PER_THREAD = (QQ_SMEM ? QQ_N : 0) + (P_SMEM ? QQ_N : 0);
  static constexpr int MIN_CTAS = DYB_KF_MIN_CTAS;
};

// the recursion of one draw; MISS = false is the instantiation for data without a single missing observation (the pattern
// is shared by every draw, so the choice is uniform): the neutralising selects of the measurement update are compiled out
template <class M, bool MISS>
DYB_D void kalman_draw(const double* __restrict__ ss, long long n_draws, const double* ysh, int nobs, int presample,
                       const int* __restrict__ status_in, double* __restrict__ loglik, int* __restrict__ status_out) {
  constexpr int NS = M::NS, K = M::K, NY = M::NY;
  using L = SSLayout<M>;
  using Cfg = KalmanCfg<M>;
  double* qsh = const_cast<double*>(ysh) + NY * nobs + threadIdx.x;   // this thread's column of the R Q R' tile
  double* psh = qsh + Cfg::QQ_N * Cfg::BLOCK;           // ... and of the P tile (P_SMEM only)

  const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_draws) return;
  if (status_in[d] != ST_OK) {
    loglik[d] = -INFINITY;
    status_out[d] = status_in[d];
    return;
  }
  const double* s_ = ss + d;
  const long long st = n_draws;

  double Tl[NS * K];                      // row-major
  double QQ[Cfg::QQ_SMEM ? 1 : NS * (NS + 1) / 2];
  double Preg[Cfg::P_SMEM ? 1 : NS * (NS + 1) / 2];
  auto Pget = [&](int e) -> double { return Cfg::P_SMEM ? psh[e * Cfg::BLOCK] : Preg[Cfg::P_SMEM ? 0 : e]; };
  auto Pset = [&](int e, double v) { if (Cfg::P_SMEM) psh[e * Cfg::BLOCK] = v; else Preg[Cfg::P_SMEM ? 0 : e] = v; };
  double a[NS];
  double yb[NY];
  double H[NY * (NY + 1) / 2];
#pragma unroll
  for (int i = 0; i < NS * K; ++i) Tl[i] = s_[(L::OFF_T + i) * st];
#pragma unroll
  for (int i = 0; i < NS * (NS + 1) / 2; ++i) {
    const double q = s_[(L::OFF_QQ + i) * st];
    if (Cfg::QQ_SMEM) qsh[i * Cfg::BLOCK] = q; else QQ[i] = q;
    Pset(i, s_[(L::OFF_P0 + i) * st]);
  }
#pragma unroll
  for (int i = 0; i < NY; ++i) yb[i] = s_[(L::OFF_YB + i) * st];
#pragma unroll
  for (int i = 0; i < NY * (NY + 1) / 2; ++i) H[i] = s_[(L::OFF_H + i) * st];
#pragma unroll
  for (int i = 0; i < NS; ++i) a[i] = 0.0;

  double acc = 0.0;
  LogDetAcc detp;
  int nseen = 0;
  bool bad = false;

  for (int t = 0; t < nobs; ++t) {
    // ---- measurement update ----------------------------------------------------------------
    double v[NY], F[NY * (NY + 1) / 2];
    bool miss[NY];
    int nobs_t = 0;
#pragma unroll
    for (int i = 0; i < NY; ++i) {
      const double yi = ysh[i + NY * t];
      miss[i] = MISS && isnan(yi);
      nobs_t += miss[i] ? 0 : 1;
      v[i] = miss[i] ? 0.0 : (yi - yb[i] - a[M::obs_idx_state(i)]);
    }
#pragma unroll
    for (int i = 0; i < NY; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        double f = Pget(sym(M::obs_idx_state(i), M::obs_idx_state(j))) + H[tri(i, j)];
        if (miss[i] || miss[j]) f = (i == j) ? 1.0 : 0.0;
        F[tri(i, j)] = f;
      }
    // Cholesky F = L L' in place (packed lower); rdiag = 1 / L_ii
    double rdiag[NY];
    double det = 1.0;
#pragma unroll
    for (int j = 0; j < NY; ++j) {
      double djj = F[tri(j, j)];
#pragma unroll
      for (int q = 0; q < j; ++q) djj = fma(-F[tri(j, q)], F[tri(j, q)], djj);
      if (!(djj > 0.0) || !(djj < 1e300)) bad = true;
      det *= djj;
      const double r = fast_rsqrt(djj);
      rdiag[j] = r;
      F[tri(j, j)] = djj * r;
#pragma unroll
      for (int i = j + 1; i < NY; ++i) {
        double x = F[tri(i, j)];
#pragma unroll
        for (int q = 0; q < j; ++q) x = fma(-F[tri(i, q)], F[tri(j, q)], x);
        F[tri(i, j)] = x * r;
      }
    }
    // w = L^-1 v
    double w[NY];
    double quad = 0.0;
#pragma unroll
    for (int i = 0; i < NY; ++i) {
      double x = v[i];
#pragma unroll
      for (int q = 0; q < i; ++q) x = fma(-F[tri(i, q)], w[q], x);
      w[i] = x * rdiag[i];
      quad = fma(w[i], w[i], quad);
    }
    if (t >= presample) {
      // sum_t log det F_t: running product with its exponent split off (LogDetAcc), one logarithm at the end
      detp.mul(det);
      acc += quad;
      nseen += nobs_t;
    }
    if (t == nobs - 1) break;

    // U = L^-1 (Z P)[:, lagged]   (NY x K)
    double U[NY * K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
#pragma unroll
      for (int i = 0; i < NY; ++i) {
        double x = miss[i] ? 0.0 : Pget(sym(M::obs_idx_state(i), M::lagged_state_ids(j)));
#pragma unroll
        for (int q = 0; q < i; ++q) x = fma(-F[tri(i, q)], U[q + NY * j], x);
        U[i + NY * j] = x * rdiag[i];
      }
    }
    // updated mean / covariance on the lagged states only
    double al[K], Pll[K * (K + 1) / 2];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      double x = a[M::lagged_state_ids(j)];
#pragma unroll
      for (int i = 0; i < NY; ++i) x = fma(U[i + NY * j], w[i], x);
      al[j] = x;
#pragma unroll
      for (int q = 0; q <= j; ++q) {
        double pjq = Pget(sym(M::lagged_state_ids(j), M::lagged_state_ids(q)));
#pragma unroll
        for (int i = 0; i < NY; ++i) pjq = fma(-U[i + NY * j], U[i + NY * q], pjq);
        Pll[tri(j, q)] = pjq;
      }
    }
    if (Cfg::P_SMEM) asm volatile("" ::: "memory");
    // ---- time update ---------------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      double x = 0.0;
      double mrow[K];                                    // row s of Tl * Pll
#pragma unroll
      for (int j = 0; j < K; ++j) x = fma(Tl[s * K + j], al[j], x);
      a[s] = x;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        double m = 0.0;
#pragma unroll
        for (int q = 0; q < K; ++q) m = fma(Tl[s * K + q], Pll[sym(q, j)], m);
        mrow[j] = m;
      }
#pragma unroll
      for (int t2 = 0; t2 <= s; ++t2) {
        double x2 = Cfg::QQ_SMEM ? qsh[tri(s, t2) * Cfg::BLOCK] : QQ[tri(s, t2)];
#pragma unroll
        for (int j = 0; j < K; ++j) x2 = fma(mrow[j], Tl[t2 * K + j], x2);
        Pset(tri(s, t2), x2);
      }
      if (Cfg::P_SMEM) asm volatile("" ::: "memory");   // keep the shared-memory loads of row s inside its iteration
    }
  }
  acc += detp.value();
  const double ll = -0.5 * (nseen * 1.8378770664093453 + acc);     // log(2 pi)
  if (bad || !isfinite(ll)) {
    loglik[d] = -INFINITY;
    status_out[d] = ST_F_NOT_PD;
  } else {
    loglik[d] = ll;
    status_out[d] = ST_OK;
  }
}

template <class M>
__global__ void __launch_bounds__(KalmanCfg<M>::BLOCK, KalmanCfg<M>::MIN_CTAS)
kalman_kernel(const double* __restrict__ ss, long long n_draws,
              const double* __restrict__ Y,            // NY x nobs, column-major, NaN = missing
              int nobs, int presample,
              const int* __restrict__ status_in,
              double* __restrict__ loglik, int* __restrict__ status_out) {
  extern __shared__ double ysh[];                      // NY * nobs, then (optionally) QQ_N * BLOCK
  int nan_seen = 0;
  for (int i = threadIdx.x; i < M::NY * nobs; i += blockDim.x) { const double y = Y[i]; ysh[i] = y; nan_seen |= isnan(y) ? 1 : 0; }
  const int any_missing = __syncthreads_or(nan_seen);
  if (any_missing) kalman_draw<M, true>(ss, n_draws, ysh, nobs, presample, status_in, loglik, status_out);
  else kalman_draw<M, false>(ss, n_draws, ysh, nobs, presample, status_in, loglik, status_out);
}

}  // namespace dyb
