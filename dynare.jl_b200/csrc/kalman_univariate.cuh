// Kalman log-likelihood with the UNIVARIATE FALLBACK of KalmanFilterTools' `kalman_likelihood` (the package behind the
// reference call src/estimation/estimation.jl:671-700; its source is not vendored, the behaviour is restated from the
// published package): when the Cholesky factorisation of F_t = Z P Z' + H fails, the period is not an error -- its
// observations are processed one at a time,
//     F_i = P[z_i, z_i] + H_ii ;  if F_i > kalman_tol:  v_i = y_i - a[z_i],  K = P[:, z_i] / F_i,
//     a += K v_i,  P -= K K' F_i,  lik += log F_i + v_i^2 / F_i        (an observation with F_i <= kalman_tol is skipped)
// followed by the usual time update.  Every observed entry still counts in the n log(2 pi) constant.  Needs a diagonal H
// (a draw with correlated measurement errors keeps status 6).
//
// This kernel is the slow path behind dyb_options.univariate_fallback: it re-evaluates only the SELECTED draws of a
// batch (those the production filter flagged "F not positive definite"; mode 2: every draw, every period univariate --
// a diagnostic that must reproduce the multivariate likelihood) from dense state spaces, one CTA per draw, P and T P in a
// global scratch.  Throughput is irrelevant here: such draws are rare.
#pragma once
#include "kalman_generic.cuh"

namespace dyb {

constexpr double KALMAN_TOL = 1e-10;

struct KalmanUnivariateSel {
  int mode;                 // 1: draws with status[d] == ST_F_NOT_PD, per-period fallback; 2: draws with status_in OK, all periods univariate
  const int* status_in;     // upstream status (null = all ok)
};

// launch <<<min(n_draws, cap), 64, smem>>>; g.scratch: 2 ns^2 doubles per CTA; g.status / g.loglik are read and updated
__global__ void __launch_bounds__(64) kalman_univariate_kernel(KalmanGenericArgs g, KalmanUnivariateSel sel) {
  extern __shared__ double sm_ku[];
  const int ns = g.ns, ny = g.ny, tid = threadIdx.x, nt = blockDim.x;
  double* P = g.scratch + (size_t)blockIdx.x * 2 * ns * ns;
  double* Mx = P + (size_t)ns * ns;
  double* F = sm_ku;                 // ny x ny
  double* U = F + (size_t)ny * ny;   // ny x ns
  double* a = U + (size_t)ny * ns;
  double* au = a + ns;
  double* v = au + ns;
  double* w = v + ny;
  double* K = w + ny;                // ns
  __shared__ int zi[64];
  __shared__ int chol_ok, bad;
  __shared__ double s_acc, s_fi, s_vi;
  __shared__ int s_nseen;
  for (long long d = blockIdx.x; d < g.n_draws; d += gridDim.x) {
    __syncthreads();
    const bool selected = sel.mode == 2 ? (!sel.status_in || sel.status_in[d] == ST_OK) : (g.status[d] == ST_F_NOT_PD);
    if (!selected) continue;
    const double* Tm = g.T + (size_t)d * ns * ns;
    const double* QQ = g.RQR + (size_t)d * ns * ns;
    const double* P0 = g.P0 + (size_t)d * ns * ns;
    const double* Hm = g.H ? g.H + (size_t)d * ny * ny : nullptr;
    const double* yb = g.ybar_obs ? g.ybar_obs + (size_t)d * ny : nullptr;
    for (int i = tid; i < ny; i += nt) zi[i] = g.z_idx[i];
    for (int i = tid; i < ns * ns; i += nt) P[i] = P0[i];
    for (int i = tid; i < ns; i += nt) a[i] = g.a0 ? g.a0[(size_t)d * ns + i] : 0.0;
    if (tid == 0) {
      s_acc = 0.0; s_nseen = 0; bad = 0;
      if (Hm) for (int i = 0; i < ny; ++i) for (int j = 0; j < ny; ++j) if (i != j && Hm[i + ny * j] != 0.0) bad = 2;   // H not diagonal
    }
    __syncthreads();
    for (int t = 0; t < g.nobs && bad != 2; ++t) {
      const double* yt = g.Y + (size_t)t * ny;
      const bool counted = t >= g.presample;
      for (int i = tid; i < ny; i += nt) {
        const double yi = yt[i];
        v[i] = isnan(yi) ? 0.0 : (yi - (yb ? yb[i] : 0.0)) - a[zi[i]];
      }
      for (int e = tid; e < ny * ny; e += nt) {
        const int i = e % ny, j = e / ny;
        double f = P[zi[i] + (size_t)ns * zi[j]] + (Hm ? Hm[e] : 0.0);
        if (isnan(yt[i]) || isnan(yt[j])) f = (i == j) ? 1.0 : 0.0;
        F[e] = f;
      }
      __syncthreads();
      if (tid == 0) {                    // Cholesky attempt (lower, in place), w = L^-1 v
        int ok = sel.mode != 2;
        double ldet = 0.0, quad = 0.0;
        for (int j = 0; j < ny && ok; ++j) {
          double djj = F[j + ny * j];
          for (int q = 0; q < j; ++q) djj -= F[j + ny * q] * F[j + ny * q];
          if (!(djj > 0.0) || !(djj < 1e300)) { ok = 0; break; }
          const double ljj = sqrt(djj);
          F[j + ny * j] = ljj;
          ldet += log(ljj);
          for (int i = j + 1; i < ny; ++i) {
            double x = F[i + ny * j];
            for (int q = 0; q < j; ++q) x -= F[i + ny * q] * F[j + ny * q];
            F[i + ny * j] = x / ljj;
          }
        }
        if (ok) {
          for (int i = 0; i < ny; ++i) {
            double x = v[i];
            for (int q = 0; q < i; ++q) x -= F[i + ny * q] * w[q];
            w[i] = x / F[i + ny * i];
            quad += w[i] * w[i];
          }
          if (counted) s_acc += 2.0 * ldet + quad;
        }
        if (counted) for (int i = 0; i < ny; ++i) s_nseen += isnan(yt[i]) ? 0 : 1;
        chol_ok = ok;
      }
      __syncthreads();
      if (chol_ok) {
        for (int s = tid; s < ns; s += nt) {       // U[:, s] = L^-1 (Z P)[:, s], au = a + U'w
          double x = a[s];
          for (int i = 0; i < ny; ++i) {
            double u = isnan(yt[i]) ? 0.0 : P[zi[i] + (size_t)ns * s];
            for (int q = 0; q < i; ++q) u -= F[i + ny * q] * U[q + ny * s];
            u /= F[i + ny * i];
            U[i + ny * s] = u;
            x += u * w[i];
          }
          au[s] = x;
        }
        __syncthreads();
        for (int e = tid; e < ns * ns; e += nt) {
          const int i = e % ns, j = e / ns;
          double x = P[e];
          for (int q = 0; q < ny; ++q) x -= U[q + ny * i] * U[q + ny * j];
          P[e] = x;
        }
        __syncthreads();
      } else {
        for (int s = tid; s < ns; s += nt) au[s] = a[s];
        __syncthreads();
        for (int i = 0; i < ny; ++i) {
          if (isnan(yt[i])) continue;
          if (tid == 0) {
            const double fi = P[zi[i] + (size_t)ns * zi[i]] + (Hm ? Hm[i + ny * i] : 0.0);
            const double vi = (yt[i] - (yb ? yb[i] : 0.0)) - au[zi[i]];
            s_fi = fi; s_vi = vi;
            if (fi > KALMAN_TOL && counted) s_acc += log(fi) + vi * vi / fi;
          }
          __syncthreads();
          const double fi = s_fi, vi = s_vi;
          if (fi > KALMAN_TOL) {
            for (int s = tid; s < ns; s += nt) K[s] = P[s + (size_t)ns * zi[i]] / fi;
            __syncthreads();
            for (int s = tid; s < ns; s += nt) au[s] += K[s] * vi;
            for (int e = tid; e < ns * ns; e += nt) P[e] -= K[e % ns] * K[e / ns] * fi;
          }
          __syncthreads();
        }
      }
      // time update: a = T au ; M = T P ; P = M T' + R Q R'
      for (int i = tid; i < ns; i += nt) {
        double x = 0.0;
        for (int q = 0; q < ns; ++q) x = fma(Tm[i + (size_t)ns * q], au[q], x);
        a[i] = x;
      }
      for (int e = tid; e < ns * ns; e += nt) {
        const int i = e % ns, j = e / ns;
        double x = 0.0;
        for (int q = 0; q < ns; ++q) x = fma(Tm[i + (size_t)ns * q], P[q + (size_t)ns * j], x);
        Mx[e] = x;
      }
      __syncthreads();
      for (int e = tid; e < ns * ns; e += nt) {
        const int i = e % ns, j = e / ns;
        double x = QQ[e];
        for (int q = 0; q < ns; ++q) x = fma(Mx[i + (size_t)ns * q], Tm[j + (size_t)ns * q], x);
        P[e] = x;
      }
      __syncthreads();
      for (int e = tid; e < ns * ns; e += nt) {      // keep P symmetric, as the update formula is
        const int i = e % ns, j = e / ns;
        if (i > j) { const double m = 0.5 * (P[e] + P[j + (size_t)ns * i]); P[e] = m; P[j + (size_t)ns * i] = m; }
      }
      __syncthreads();
    }
    if (tid == 0) {
      const double ll = -0.5 * (s_nseen * 1.8378770664093453 + s_acc);
      if (bad == 2 || !isfinite(ll)) { g.loglik[d] = -INFINITY; g.status[d] = ST_F_NOT_PD; }
      else { g.loglik[d] = ll; g.status[d] = ST_OK; }
    }
  }
}

inline size_t kalman_univariate_smem_bytes(int ns, int ny) {
  return sizeof(double) * ((size_t)ny * ny + (size_t)ny * ns + 3 * (size_t)ns + 2 * (size_t)ny);
}

}  // namespace dyb
