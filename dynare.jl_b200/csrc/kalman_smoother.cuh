// Batched fixed-interval Kalman smoother for dense state spaces of run-time size: one CTA per draw, forward filter
// with the per-period quantities the backward pass needs kept in a global workspace (a_t, P_t, u_t = F_t^-1 v_t,
// K_t), then the backward recursion
//     eta_hat_t = Q R' r_t ;  e_t = u_t - K_t' r_t ;  eps_hat_t = H e_t ;  r_{t-1} = T' r_t + Z' e_t ;
//     alpha_hat_t = a_t + P_t r_{t-1}                                   (Durbin & Koopman 2012, sections 4.3-4.5)
// Missing observations (NaN, pattern shared by all draws) are removed exactly as in the filters.
// Replaces, for N draws at once, the `kalman_smoother!` call of `calibsmoother!`
// (reference src/filters/kalman/kalman.jl:52-160; KalmanFilterTools 0.1.x, not vendored).  Correctness-first version
// of SURVEY.md section 8f rank 4: the workspace traffic (ns^2 doubles written and read per period) is its roofline.
#pragma once
#include "model_traits.cuh"

namespace dyb {

struct KalmanSmootherArgs {
  int ns, ny, nx, nobs;
  const int* z_idx;        // ny
  const double* Y;         // ny x nobs
  const double* T;         // ns x ns per draw
  const double* R;         // ns x nx per draw
  const double* Q;         // nx x nx per draw
  const double* P0;        // ns x ns per draw
  const double* a0;        // ns per draw or null
  const double* H;         // ny x ny per draw or null
  const double* ybar_obs;  // ny per draw or null
  const double* level;     // ns per draw or null: added to the state outputs (steady state of the state variables)
  const int* status_in;    // per draw or null
  long long n_draws;
  double* work;            // per CTA: ns*(nobs+1) + ns*ns*nobs + ny*nobs + ns*ny*nobs doubles
  // outputs per draw (any but status may be null)
  double* alphah;          // ns x nobs
  double* etah;            // nx x nobs
  double* epsilonh;        // ny x nobs
  double* att;             // ns x nobs
  double* a_pred;          // ns x (nobs + 1)
  double* loglik;
  int* status;
};

DYB_HD size_t smoother_work_doubles(int ns, int ny, int nobs) {
  return (size_t)ns * (nobs + 1) + (size_t)ns * ns * nobs + (size_t)ny * nobs + (size_t)ns * ny * nobs;
}
inline size_t smoother_smem_doubles(int ns, int ny, int nx) {
  return 2 * (size_t)ns * ns + 2 * (size_t)ny * ns + (size_t)ny * ny + 4 * (size_t)ns + 4 * (size_t)ny + (size_t)nx + 8;
}

__global__ void __launch_bounds__(256) kalman_smoother_kernel(KalmanSmootherArgs g) {
  extern __shared__ double sm[];
  const int ns = g.ns, ny = g.ny, nx = g.nx, nobs = g.nobs;
  const int tid = threadIdx.x, nt = blockDim.x;
  double* P = sm;                                // ns x ns
  double* Mx = P + (size_t)ns * ns;              // T P in the forward pass, T itself in the backward pass
  double* U = Mx + (size_t)ns * ns;              // ny x ns: U[i + ny s]
  double* G = U + (size_t)ny * ns;               // ns x ny: G[s + ns i] = (P Z' F^-1)[s, i]
  double* F = G + (size_t)ny * ns;               // ny x ny lower Cholesky
  double* a = F + (size_t)ny * ny;
  double* au = a + ns;
  double* r = au + ns;
  double* rn = r + ns;
  double* v = rn + ns;
  double* w = v + ny;
  double* u = w + ny;
  double* e = u + ny;
  double* tx = e + ny;                           // nx
  __shared__ int bad;
  __shared__ int zi[64];
  double* Astore = g.work + (size_t)blockIdx.x * smoother_work_doubles(ns, ny, nobs);
  double* Pstore = Astore + (size_t)ns * (nobs + 1);
  double* ustore = Pstore + (size_t)ns * ns * nobs;
  double* Kstore = ustore + (size_t)ny * nobs;

  for (long long d = blockIdx.x; d < g.n_draws; d += gridDim.x) {
    __syncthreads();
    if (g.status_in && g.status_in[d] != ST_OK) {
      if (tid == 0) { g.status[d] = g.status_in[d]; if (g.loglik) g.loglik[d] = -INFINITY; }
      continue;
    }
    const double* Tm = g.T + (size_t)d * ns * ns;
    const double* Rm = g.R + (size_t)d * ns * nx;
    const double* Qm = g.Q + (size_t)d * nx * nx;
    const double* P0 = g.P0 + (size_t)d * ns * ns;
    const double* Hm = g.H ? g.H + (size_t)d * ny * ny : nullptr;
    const double* yb = g.ybar_obs ? g.ybar_obs + (size_t)d * ny : nullptr;
    const double* lev = g.level ? g.level + (size_t)d * ns : nullptr;
    if (tid == 0) bad = 0;
    for (int i = tid; i < ny; i += nt) zi[i] = g.z_idx[i];
    for (int i = tid; i < ns * ns; i += nt) P[i] = P0[i];
    for (int i = tid; i < ns; i += nt) a[i] = g.a0 ? g.a0[(size_t)d * ns + i] : 0.0;
    __syncthreads();
    double acc = 0.0;
    int nseen = 0;
    // ---- forward pass -------------------------------------------------------------------------------------------
    for (int t = 0; t < nobs; ++t) {
      const double* yt = g.Y + (size_t)t * ny;
      for (int i = tid; i < ns; i += nt) Astore[i + (size_t)ns * t] = a[i];
      for (int i = tid; i < ns * ns; i += nt) Pstore[i + (size_t)ns * ns * t] = P[i];
      for (int i = tid; i < ny; i += nt) {
        const double yi = yt[i];
        v[i] = isnan(yi) ? 0.0 : (yi - (yb ? yb[i] : 0.0)) - a[zi[i]];
      }
      for (int q = tid; q < ny * ny; q += nt) {
        const int i = q % ny, j = q / ny;
        const bool mi = isnan(yt[i]), mj = isnan(yt[j]);
        double f = P[zi[i] + (size_t)ns * zi[j]] + (Hm ? Hm[q] : 0.0);
        if (mi || mj) f = (i == j) ? 1.0 : 0.0;
        F[q] = f;
      }
      __syncthreads();
      if (tid < 32) {
        for (int j = 0; j < ny; ++j) {
          const double djj = F[j + ny * j];
          __syncwarp();
          if (tid == 0) { if (!(djj > 0.0)) bad = 1; F[j + ny * j] = sqrt(djj); }
          __syncwarp();
          const double ljj = F[j + ny * j];
          for (int i = j + 1 + tid; i < ny; i += 32) F[i + ny * j] /= ljj;
          __syncwarp();
          for (int c = j + 1; c < ny; ++c) {
            const double lcj = F[c + ny * j];
            for (int i = c + tid; i < ny; i += 32) F[i + ny * c] -= F[i + ny * j] * lcj;
          }
          __syncwarp();
        }
        if (tid == 0) {
          double quad = 0.0, ldet = 0.0;
          int cnt = 0;
          for (int i = 0; i < ny; ++i) {
            double x = v[i];
            for (int p = 0; p < i; ++p) x -= F[i + ny * p] * w[p];
            x /= F[i + ny * i];
            w[i] = x;
            quad += x * x;
            ldet += log(F[i + ny * i]);
            cnt += isnan(yt[i]) ? 0 : 1;
          }
          for (int i = ny - 1; i >= 0; --i) {               // u = F^-1 v = L^-T w
            double x = w[i];
            for (int p = i + 1; p < ny; ++p) x -= F[p + ny * i] * u[p];
            u[i] = x / F[i + ny * i];
          }
          acc += 2.0 * ldet + quad; nseen += cnt;
        }
      }
      __syncthreads();
      for (int i = tid; i < ny; i += nt) ustore[i + (size_t)ny * t] = u[i];
      // U[:, s] = L^-1 (Z P)[:, s] ; G[s, :] = L^-T U[:, s] ; att = a + U'w
      for (int s = tid; s < ns; s += nt) {
        double x = a[s];
        for (int i = 0; i < ny; ++i) {
          double q = isnan(yt[i]) ? 0.0 : P[zi[i] + (size_t)ns * s];
          for (int p = 0; p < i; ++p) q -= F[i + ny * p] * U[p + ny * s];
          q /= F[i + ny * i];
          U[i + ny * s] = q;
          x += q * w[i];
        }
        for (int i = ny - 1; i >= 0; --i) {
          double q = U[i + ny * s];
          for (int p = i + 1; p < ny; ++p) q -= F[p + ny * i] * G[s + ns * p];
          G[s + ns * i] = q / F[i + ny * i];
        }
        au[s] = x;
        if (g.att) g.att[(size_t)d * ns * nobs + s + (size_t)ns * t] = x + (lev ? lev[s] : 0.0);
      }
      __syncthreads();
      // K = T G (stored) ; P <- P - U'U
      for (int q = tid; q < ns * ny; q += nt) {
        const int s = q % ns, i = q / ns;
        double x = 0.0;
        for (int p = 0; p < ns; ++p) x = fma(Tm[s + (size_t)ns * p], G[p + ns * i], x);
        Kstore[q + (size_t)ns * ny * t] = x;
      }
      for (int q = tid; q < ns * ns; q += nt) {
        const int i = q % ns, j = q / ns;
        double x = P[q];
        for (int p = 0; p < ny; ++p) x -= U[p + ny * i] * U[p + ny * j];
        P[q] = x;
      }
      __syncthreads();
      for (int i = tid; i < ns; i += nt) {
        double x = 0.0;
        for (int p = 0; p < ns; ++p) x += Tm[i + (size_t)ns * p] * au[p];
        a[i] = x;
      }
      for (int q = tid; q < ns * ns; q += nt) {
        const int i = q % ns, j = q / ns;
        double x = 0.0;
        for (int p = 0; p < ns; ++p) x += Tm[i + (size_t)ns * p] * P[p + (size_t)ns * j];
        Mx[q] = x;
      }
      __syncthreads();
      for (int q = tid; q < ns * ns; q += nt) {
        const int i = q % ns, j = q / ns;
        if (i < j) continue;
        double x = 0.0;
        for (int p = 0; p < ns; ++p) x += Mx[i + (size_t)ns * p] * Tm[j + (size_t)ns * p];
        for (int p = 0; p < nx; ++p) {                       // + (R Q R')[i, j]
          double rq = 0.0;
          for (int f = 0; f < nx; ++f) rq = fma(Rm[i + (size_t)ns * f], Qm[f + nx * p], rq);
          x = fma(rq, Rm[j + (size_t)ns * p], x);
        }
        P[q] = x;
        P[j + (size_t)ns * i] = x;
      }
      __syncthreads();
    }
    for (int i = tid; i < ns; i += nt) Astore[i + (size_t)ns * nobs] = a[i];
    // ---- backward pass ------------------------------------------------------------------------------------------
    for (int i = tid; i < ns * ns; i += nt) Mx[i] = Tm[i];
    for (int i = tid; i < ns; i += nt) r[i] = 0.0;
    __syncthreads();
    for (int t = nobs - 1; t >= 0; --t) {
      const double* Kt = Kstore + (size_t)ns * ny * t;
      const double* Pt = Pstore + (size_t)ns * ns * t;
      for (int p = tid; p < nx; p += nt) {                   // tx = R' r
        double x = 0.0;
        for (int s = 0; s < ns; ++s) x = fma(Rm[s + (size_t)ns * p], r[s], x);
        tx[p] = x;
      }
      for (int i = tid; i < ny; i += nt) {                   // e = u_t - K_t' r
        double x = ustore[i + (size_t)ny * t];
        for (int s = 0; s < ns; ++s) x = fma(-Kt[s + ns * i], r[s], x);
        e[i] = x;
      }
      __syncthreads();
      if (g.etah) for (int p = tid; p < nx; p += nt) {
        double x = 0.0;
        for (int f = 0; f < nx; ++f) x = fma(Qm[p + nx * f], tx[f], x);
        g.etah[(size_t)d * nx * nobs + p + (size_t)nx * t] = x;
      }
      if (g.epsilonh) for (int i = tid; i < ny; i += nt) {
        double x = 0.0;
        if (Hm) for (int j = 0; j < ny; ++j) x = fma(Hm[i + ny * j], e[j], x);
        g.epsilonh[(size_t)d * ny * nobs + i + (size_t)ny * t] = x;
      }
      for (int s = tid; s < ns; s += nt) {                   // rn = T' r + Z' e
        double x = 0.0;
        for (int p = 0; p < ns; ++p) x = fma(Mx[p + (size_t)ns * s], r[p], x);
        for (int i = 0; i < ny; ++i) if (zi[i] == s) x += e[i];
        rn[s] = x;
      }
      __syncthreads();
      for (int s = tid; s < ns; s += nt) {                   // alpha_hat_t = a_t + P_t r_{t-1}
        double x = Astore[s + (size_t)ns * t];
        for (int p = 0; p < ns; ++p) x = fma(Pt[s + (size_t)ns * p], rn[p], x);
        if (g.alphah) g.alphah[(size_t)d * ns * nobs + s + (size_t)ns * t] = x + (lev ? lev[s] : 0.0);
        r[s] = rn[s];
      }
      __syncthreads();
    }
    if (g.a_pred) for (int q = tid; q < ns * (nobs + 1); q += nt)
      g.a_pred[(size_t)d * ns * (nobs + 1) + q] = Astore[q] + (lev ? lev[q % ns] : 0.0);
    if (tid == 0) {
      const double ll = -0.5 * (nseen * 1.8378770664093453 + acc);
      const bool fail = bad || !isfinite(ll);
      if (g.loglik) g.loglik[d] = fail ? -INFINITY : ll;
      g.status[d] = fail ? ST_F_NOT_PD : ST_OK;
    }
  }
}

}  // namespace dyb
