// Batched exact-diffuse Kalman filter and smoother for state spaces with unit roots: one CTA per draw, univariate
// treatment of the observations (Koopman & Durbin 2000; Durbin & Koopman 2012, sections 5.2-5.4 and 6.4), so that a
// rank-deficient Finf -- several observables loading on one stochastic trend -- needs no special case.  H must be
// diagonal.  A period is diffuse while max diag(Pinf) > tol; afterwards the same recursion is the ordinary univariate
// filter.  Replaces, for N draws at once, the `diffuse_kalman_smoother!` call of calibsmoother!
// (reference src/filters/kalman/kalman.jl:163-223; KalmanFilterTools 0.1.x, not vendored) in the ORIGINAL basis:
// the caller passes Pinf0 = Z1 Z1', Pstar0 = Z2 P22 Z2' built from the ordered Schur vectors [Z1 Z2] of T.
// Forward pass keeps a_t, Pstar_t, Pinf_t and per observation (v, Finf, Fstar, Kinf, Kstar) in a global workspace;
// backward pass runs the (r0, r1) recursion:
//     Finf > tol :  r1 <- r1 + z (v/Finf + ((Kinf Fstar/Finf - Kstar)'r0 - Kinf'r1)/Finf) ;  r0 <- r0 - z Kinf'r0/Finf
//     else       :  r0 <- r0 + z (v - Kstar'r0)/Fstar ;  r1 <- r1 - z Kstar'r1/Fstar            (z = the observed state)
//     alpha_hat_t = a_t + Pstar_t r0 + Pinf_t r1 ;  eta_hat_t = Q R' r0_t
#pragma once
#include "model_traits.cuh"

namespace dyb {

struct KalmanDiffuseArgs {
  int ns, ny, nx, nobs;
  double tol;
  const int* z_idx;        // ny
  const double* Y;         // ny x nobs
  const double* T;         // ns x ns per draw
  const double* R;         // ns x nx per draw
  const double* Q;         // nx x nx per draw
  const double* Pinf0;     // ns x ns per draw
  const double* Pstar0;    // ns x ns per draw
  const double* a0;        // ns per draw or null
  const double* Hdiag;     // ny per draw or null
  const double* ybar_obs;  // ny per draw or null: subtracted from the observations (steady state of the observables)
  long long n_draws;
  double* work;            // per CTA: diffuse_work_doubles
  double* alphah;          // ns x nobs
  double* etah;            // nx x nobs
  double* epsilonh;        // ny x nobs
  double* att;             // ns x nobs
  double* a_pred;          // ns x (nobs + 1)
  double* loglik;
  int* n_diffuse;          // number of diffuse periods, per draw (may be null)
  int* status;
};

DYB_HD size_t diffuse_work_doubles(int ns, int ny, int nobs) {
  return (size_t)ns * (nobs + 1) + 2 * (size_t)ns * ns * nobs + 4 * (size_t)ny * nobs + 2 * (size_t)ns * ny * nobs +
         (size_t)ns * ns;
}
inline size_t diffuse_smem_doubles(int ns, int nx) { return 3 * (size_t)ns * ns + 5 * (size_t)ns + (size_t)nx + 8; }

__global__ void __launch_bounds__(256) kalman_diffuse_kernel(KalmanDiffuseArgs g) {
  extern __shared__ double sm[];
  const int ns = g.ns, ny = g.ny, nx = g.nx, nobs = g.nobs;
  const int tid = threadIdx.x, nt = blockDim.x;
  const double tol = g.tol;
  double* Ps = sm;
  double* Pi = Ps + (size_t)ns * ns;
  double* Mx = Pi + (size_t)ns * ns;
  double* a = Mx + (size_t)ns * ns;
  double* r0 = a + ns;
  double* r1 = r0 + ns;
  double* Ks = r1 + ns;
  double* Ki = Ks + ns;
  double* tx = Ki + ns;                         // nx
  __shared__ int zi[64];
  double* Astore = g.work + (size_t)blockIdx.x * diffuse_work_doubles(ns, ny, nobs);
  double* Pstore = Astore + (size_t)ns * (nobs + 1);
  double* Pistore = Pstore + (size_t)ns * ns * nobs;
  double* sstore = Pistore + (size_t)ns * ns * nobs;    // (kind, v, Finf, Fstar) per (t, i)
  double* Kstore = sstore + 4 * (size_t)ny * nobs;      // Kstar then Kinf, ns each, per (t, i)
  double* QQ = Kstore + 2 * (size_t)ns * ny * nobs;     // R Q R'

  for (long long d = blockIdx.x; d < g.n_draws; d += gridDim.x) {
    __syncthreads();
    const double* Tm = g.T + (size_t)d * ns * ns;
    const double* Rm = g.R + (size_t)d * ns * nx;
    const double* Qm = g.Q + (size_t)d * nx * nx;
    const double* Hd = g.Hdiag ? g.Hdiag + (size_t)d * ny : nullptr;
    const double* yb = g.ybar_obs ? g.ybar_obs + (size_t)d * ny : nullptr;
    for (int i = tid; i < ny; i += nt) zi[i] = g.z_idx[i];
    for (int i = tid; i < ns * ns; i += nt) {
      Ps[i] = g.Pstar0[(size_t)d * ns * ns + i];
      Pi[i] = g.Pinf0[(size_t)d * ns * ns + i];
    }
    for (int i = tid; i < ns; i += nt) a[i] = g.a0 ? g.a0[(size_t)d * ns + i] : 0.0;
    for (int q = tid; q < ns * nx; q += nt) {              // Mx[:, p] = (R Q)[:, p]
      const int i = q % ns, p = q / ns;
      double x = 0.0;
      for (int f = 0; f < nx; ++f) x = fma(Rm[i + (size_t)ns * f], Qm[f + nx * p], x);
      Mx[q] = x;
    }
    __syncthreads();
    for (int q = tid; q < ns * ns; q += nt) {
      const int i = q % ns, j = q / ns;
      double x = 0.0;
      for (int p = 0; p < nx; ++p) x = fma(Mx[i + (size_t)ns * p], Rm[j + (size_t)ns * p], x);
      QQ[q] = x;
    }
    __syncthreads();
    double acc = 0.0;
    int nseen = 0, ndiff = 0;
    bool diffuse;
    {
      double mx = 0.0;
      for (int s = 0; s < ns; ++s) mx = fmax(mx, Pi[s + (size_t)ns * s]);
      diffuse = mx > tol;
    }
    // ---- forward pass -------------------------------------------------------------------------------------------
    for (int t = 0; t < nobs; ++t) {
      const double* yt = g.Y + (size_t)t * ny;
      for (int i = tid; i < ns; i += nt) Astore[i + (size_t)ns * t] = a[i];
      for (int i = tid; i < ns * ns; i += nt) Pstore[i + (size_t)ns * ns * t] = Ps[i];
      if (diffuse) {
        for (int i = tid; i < ns * ns; i += nt) Pistore[i + (size_t)ns * ns * t] = Pi[i];
        ndiff = t + 1;
      }
      for (int i = 0; i < ny; ++i) {
        double* st = sstore + 4 * ((size_t)ny * t + i);
        const double yi = yt[i];
        if (isnan(yi)) { if (tid == 0) st[0] = -1.0; continue; }
        const int z = zi[i];
        const double v = (yi - (yb ? yb[i] : 0.0)) - a[z];
        const double Fs = Ps[z + (size_t)ns * z] + (Hd ? Hd[i] : 0.0);
        const double Fi = diffuse ? Pi[z + (size_t)ns * z] : 0.0;
        double* Kst = Kstore + 2 * (size_t)ns * ((size_t)ny * t + i);
        for (int s = tid; s < ns; s += nt) {
          const double ks = Ps[s + (size_t)ns * z];
          Ks[s] = ks; Kst[s] = ks;
          if (diffuse) { const double ki = Pi[s + (size_t)ns * z]; Ki[s] = ki; Kst[ns + s] = ki; }
        }
        const int kind = (diffuse && Fi > tol) ? 1 : (Fs > tol ? 0 : -1);
        if (tid == 0) { st[0] = (double)kind; st[1] = v; st[2] = Fi; st[3] = Fs; }
        __syncthreads();
        if (kind == 1) {
          const double c1 = Fs / (Fi * Fi), c2 = 1.0 / Fi;
          for (int s = tid; s < ns; s += nt) a[s] = fma(Ki[s], v * c2, a[s]);
          for (int q = tid; q < ns * ns; q += nt) {
            const int r = q % ns, c = q / ns;
            const double kir = Ki[r], kic = Ki[c];
            Ps[q] += kir * kic * c1 - (Ks[r] * kic + kir * Ks[c]) * c2;
            Pi[q] -= kir * kic * c2;
          }
          acc += log(Fi); ++nseen;
        } else if (kind == 0) {
          const double c2 = 1.0 / Fs;
          for (int s = tid; s < ns; s += nt) a[s] = fma(Ks[s], v * c2, a[s]);
          for (int q = tid; q < ns * ns; q += nt) {
            const int r = q % ns, c = q / ns;
            Ps[q] -= Ks[r] * Ks[c] * c2;
          }
          acc += log(Fs) + v * v * c2; ++nseen;
        }
        __syncthreads();
      }
      if (g.att) for (int s = tid; s < ns; s += nt) g.att[(size_t)d * ns * nobs + s + (size_t)ns * t] = a[s];
      // prediction: a = T a ; Pstar = T Pstar T' + R Q R' ; Pinf = T Pinf T'
      for (int i = tid; i < ns; i += nt) {
        double x = 0.0;
        for (int p = 0; p < ns; ++p) x = fma(Tm[i + (size_t)ns * p], a[p], x);
        Ks[i] = x;
      }
      for (int q = tid; q < ns * ns; q += nt) {
        const int i = q % ns, j = q / ns;
        double x = 0.0;
        for (int p = 0; p < ns; ++p) x = fma(Tm[i + (size_t)ns * p], Ps[p + (size_t)ns * j], x);
        Mx[q] = x;
      }
      __syncthreads();
      for (int i = tid; i < ns; i += nt) a[i] = Ks[i];
      for (int q = tid; q < ns * ns; q += nt) {
        const int i = q % ns, j = q / ns;
        if (i < j) continue;
        double x = QQ[q];
        for (int p = 0; p < ns; ++p) x = fma(Mx[i + (size_t)ns * p], Tm[j + (size_t)ns * p], x);
        Ps[q] = x; Ps[j + (size_t)ns * i] = x;
      }
      __syncthreads();
      if (diffuse) {
        for (int q = tid; q < ns * ns; q += nt) {
          const int i = q % ns, j = q / ns;
          double x = 0.0;
          for (int p = 0; p < ns; ++p) x = fma(Tm[i + (size_t)ns * p], Pi[p + (size_t)ns * j], x);
          Mx[q] = x;
        }
        __syncthreads();
        for (int q = tid; q < ns * ns; q += nt) {
          const int i = q % ns, j = q / ns;
          if (i < j) continue;
          double x = 0.0;
          for (int p = 0; p < ns; ++p) x = fma(Mx[i + (size_t)ns * p], Tm[j + (size_t)ns * p], x);
          Pi[q] = x; Pi[j + (size_t)ns * i] = x;
        }
        __syncthreads();
        double mx = 0.0;
        for (int s = 0; s < ns; ++s) mx = fmax(mx, Pi[s + (size_t)ns * s]);
        diffuse = mx > tol;
      }
    }
    for (int i = tid; i < ns; i += nt) Astore[i + (size_t)ns * nobs] = a[i];
    // ---- backward pass ------------------------------------------------------------------------------------------
    for (int i = tid; i < ns * ns; i += nt) Mx[i] = Tm[i];
    for (int i = tid; i < ns; i += nt) { r0[i] = 0.0; r1[i] = 0.0; }
    __syncthreads();
    for (int t = nobs - 1; t >= 0; --t) {
      const bool dper = t < ndiff;
      const double* yt = g.Y + (size_t)t * ny;
      for (int p = tid; p < nx; p += nt) {                   // tx = R' r0
        double x = 0.0;
        for (int s = 0; s < ns; ++s) x = fma(Rm[s + (size_t)ns * p], r0[s], x);
        tx[p] = x;
      }
      for (int s = tid; s < ns; s += nt) {                   // T' r0, T' r1
        double x0 = 0.0, x1 = 0.0;
        for (int p = 0; p < ns; ++p) {
          const double tp = Mx[p + (size_t)ns * s];
          x0 = fma(tp, r0[p], x0); x1 = fma(tp, r1[p], x1);
        }
        Ks[s] = x0; Ki[s] = x1;
      }
      __syncthreads();
      if (g.etah) for (int p = tid; p < nx; p += nt) {
        double x = 0.0;
        for (int f = 0; f < nx; ++f) x = fma(Qm[p + nx * f], tx[f], x);
        g.etah[(size_t)d * nx * nobs + p + (size_t)nx * t] = x;
      }
      for (int s = tid; s < ns; s += nt) { r0[s] = Ks[s]; r1[s] = Ki[s]; }
      __syncthreads();
      if (tid < 32) {                                        // the observations of period t, last to first
        for (int i = ny - 1; i >= 0; --i) {
          const double* st = sstore + 4 * ((size_t)ny * t + i);
          const int kind = (int)st[0];
          if (kind < 0) continue;
          const double v = st[1], Fi = st[2], Fs = st[3];
          const double* Kst = Kstore + 2 * (size_t)ns * ((size_t)ny * t + i);
          const int z = zi[i];
          double ds0 = 0.0, ds1 = 0.0, di0 = 0.0, di1 = 0.0;
          for (int s = tid; s < ns; s += 32) {
            const double ks = Kst[s];
            ds0 = fma(ks, r0[s], ds0); ds1 = fma(ks, r1[s], ds1);
            if (kind == 1) { const double ki = Kst[ns + s]; di0 = fma(ki, r0[s], di0); di1 = fma(ki, r1[s], di1); }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            ds0 += __shfl_xor_sync(0xffffffffu, ds0, o); ds1 += __shfl_xor_sync(0xffffffffu, ds1, o);
            di0 += __shfl_xor_sync(0xffffffffu, di0, o); di1 += __shfl_xor_sync(0xffffffffu, di1, o);
          }
          __syncwarp();
          if (tid == 0) {
            if (kind == 1) {
              r1[z] += v / Fi + ((di0 * (Fs / Fi) - ds0) - di1) / Fi;
              r0[z] -= di0 / Fi;
            } else {
              r0[z] += (v - ds0) / Fs;
              r1[z] -= ds1 / Fs;
            }
          }
          __syncwarp();
        }
      }
      __syncthreads();
      const double* Pt = Pstore + (size_t)ns * ns * t;
      const double* Pit = Pistore + (size_t)ns * ns * t;
      for (int s = tid; s < ns; s += nt) {                   // alpha_hat_t = a_t + Pstar_t r0 + Pinf_t r1
        double x = Astore[s + (size_t)ns * t];
        for (int p = 0; p < ns; ++p) x = fma(Pt[s + (size_t)ns * p], r0[p], x);
        if (dper) for (int p = 0; p < ns; ++p) x = fma(Pit[s + (size_t)ns * p], r1[p], x);
        Ks[s] = x;
        if (g.alphah) g.alphah[(size_t)d * ns * nobs + s + (size_t)ns * t] = x;
      }
      __syncthreads();
      if (g.epsilonh) for (int i = tid; i < ny; i += nt) {   // eps_hat = y - Z alpha_hat where H_ii > 0
        const double yi = yt[i];
        const bool has = !isnan(yi) && Hd && Hd[i] > 0.0;
        g.epsilonh[(size_t)d * ny * nobs + i + (size_t)ny * t] = has ? (yi - (yb ? yb[i] : 0.0)) - Ks[zi[i]] : 0.0;
      }
      __syncthreads();
    }
    if (g.a_pred) for (int q = tid; q < ns * (nobs + 1); q += nt) g.a_pred[(size_t)d * ns * (nobs + 1) + q] = Astore[q];
    if (tid == 0) {
      const double ll = -0.5 * (nseen * 1.8378770664093453 + acc);
      const bool fail = !isfinite(ll);
      if (g.loglik) g.loglik[d] = fail ? -INFINITY : ll;
      if (g.n_diffuse) g.n_diffuse[d] = ndiff;
      g.status[d] = fail ? ST_F_NOT_PD : ST_OK;
    }
  }
}

}  // namespace dyb
