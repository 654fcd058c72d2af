// Generic batched Kalman log-likelihood for dense state spaces of run-time size, SCALAR fp64 variant: one CTA per
// draw, the covariance recursion in shared memory when it fits (ns <= ~96), otherwise in a per-CTA global scratch
// that stays L2-resident.  Same recursion as kalman_kernel.cuh but with a dense transition matrix (no lagged-column
// structure assumed).  The tensor-core kernels (kalman_dmma.cuh, kalman_dmma_large.cuh) are the production path up
// to ns = 256; this one serves larger state spaces and is the A/B reference (dyb_set_kalman_variant(1)).  It also
// defines KalmanGenericArgs, shared by the three dense filter kernels.
//
// Replaces `kalman_likelihood(Y,Z,H,T,R,Q,a0,P,start,last,presample,ws)` (reference call site
// src/estimation/estimation.jl:687-700) for a selection matrix Z given by z_idx.
#pragma once
#include "model_traits.cuh"

namespace dyb {

struct KalmanGenericArgs {
  int ns, ny, nobs, presample;
  const int* z_idx;        // ny
  const double* Y;         // ny x nobs
  const double* T;         // ns x ns per draw, column-major
  const double* RQR;       // ns x ns per draw
  const double* P0;        // ns x ns per draw
  const double* a0;        // ns per draw or null
  const double* H;         // ny x ny per draw or null
  const double* ybar_obs;  // ny per draw or null: steady state of the observables, removed from Y (data.jl:355-363)
  const int* status_in;    // per draw or null: draws that already failed upstream are passed through
  long long n_draws;
  double* loglik;
  int* status;
  double* scratch;         // null => P and M live in shared memory; else 2*ns*ns doubles per CTA
};

// shared memory (doubles): [P ns*ns | M ns*ns]? (only if scratch == null) | U ny*ns | F ny*ny | a ns | au ns | v ny | w ny | red 32
__global__ void __launch_bounds__(256)
kalman_generic_kernel(KalmanGenericArgs g) {
  extern __shared__ double sm[];
  const int ns = g.ns, ny = g.ny;
  const int tid = threadIdx.x, nt = blockDim.x;
  double* P;
  double* Mx;
  double* base = sm;
  if (g.scratch) {
    P = g.scratch + (size_t)blockIdx.x * 2 * ns * ns;
    Mx = P + (size_t)ns * ns;
  } else {
    P = base; base += (size_t)ns * ns;
    Mx = base; base += (size_t)ns * ns;
  }
  double* U = base; base += (size_t)ny * ns;     // U[i + ny*s]
  double* F = base; base += (size_t)ny * ny;     // column-major lower Cholesky
  double* a = base; base += ns;
  double* au = base; base += ns;
  double* v = base; base += ny;
  double* w = base; base += ny;
  double* red = base; base += 32;
  __shared__ int bad;
  __shared__ int zi[64];

  for (long long d = blockIdx.x; d < g.n_draws; d += gridDim.x) {
    const double* Tm = g.T + (size_t)d * ns * ns;
    const double* QQ = g.RQR + (size_t)d * ns * ns;
    const double* P0 = g.P0 + (size_t)d * ns * ns;
    const double* Hm = g.H ? g.H + (size_t)d * ny * ny : nullptr;
    const double* yb = g.ybar_obs ? g.ybar_obs + (size_t)d * ny : nullptr;
    if (g.status_in && g.status_in[d] != ST_OK) {
      if (tid == 0) { g.loglik[d] = -INFINITY; g.status[d] = g.status_in[d]; }
      continue;
    }
    if (tid == 0) bad = 0;
    for (int i = tid; i < ny; i += nt) zi[i] = g.z_idx[i];
    for (int i = tid; i < ns * ns; i += nt) P[i] = P0[i];
    for (int i = tid; i < ns; i += nt) a[i] = g.a0 ? g.a0[(size_t)d * ns + i] : 0.0;
    __syncthreads();
    double acc = 0.0;          // only thread 0's copy is used
    int nseen = 0;

    for (int t = 0; t < g.nobs; ++t) {
      const double* yt = g.Y + (size_t)t * ny;
      // v, F
      for (int i = tid; i < ny; i += nt) {
        const double yi = yt[i];
        v[i] = isnan(yi) ? 0.0 : (yi - (yb ? yb[i] : 0.0)) - a[zi[i]];
      }
      for (int e = tid; e < ny * ny; e += nt) {
        const int i = e % ny, j = e / ny;
        const bool mi = isnan(yt[i]), mj = isnan(yt[j]);
        double f = P[zi[i] + (size_t)ns * zi[j]] + (Hm ? Hm[e] : 0.0);
        if (mi || mj) f = (i == j) ? 1.0 : 0.0;
        F[e] = f;
      }
      __syncthreads();
      // Cholesky by warp 0 (right-looking, column by column)
      if (tid < 32) {
        for (int j = 0; j < ny; ++j) {
          double djj = F[j + ny * j];
          if (tid == 0) {
            if (!(djj > 0.0)) bad = 1;
            F[j + ny * j] = sqrt(djj);
          }
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
        // w = L^-1 v (thread 0; ny is small)
        if (tid == 0) {
          double quad = 0.0, ld = 0.0;
          int cnt = 0;
          for (int i = 0; i < ny; ++i) {
            double x = v[i];
            for (int q = 0; q < i; ++q) x -= F[i + ny * q] * w[q];
            x /= F[i + ny * i];
            w[i] = x;
            quad += x * x;
            ld += log(F[i + ny * i]);
            cnt += isnan(yt[i]) ? 0 : 1;
          }
          if (t >= g.presample) { acc += 2.0 * ld + quad; nseen += cnt; }
        }
      }
      __syncthreads();
      if (t == g.nobs - 1) break;
      // U[:, s] = L^-1 (Z P)[:, s]
      for (int s = tid; s < ns; s += nt) {
        for (int i = 0; i < ny; ++i) {
          double x = isnan(yt[i]) ? 0.0 : P[zi[i] + (size_t)ns * s];
          for (int q = 0; q < i; ++q) x -= F[i + ny * q] * U[q + ny * s];
          U[i + ny * s] = x / F[i + ny * i];
        }
        double x = a[s];
        for (int i = 0; i < ny; ++i) x += U[i + ny * s] * w[i];
        au[s] = x;
      }
      __syncthreads();
      // P <- P - U'U
      for (int e = tid; e < ns * ns; e += nt) {
        const int i = e % ns, j = e / ns;
        double x = P[e];
        for (int q = 0; q < ny; ++q) x -= U[q + ny * i] * U[q + ny * j];
        P[e] = x;
      }
      __syncthreads();
      // a <- T au ; Mx <- T P
      for (int i = tid; i < ns; i += nt) {
        double x = 0.0;
        for (int q = 0; q < ns; ++q) x += Tm[i + (size_t)ns * q] * au[q];
        a[i] = x;
      }
      for (int e = tid; e < ns * ns; e += nt) {
        const int i = e % ns, j = e / ns;
        double x = 0.0;
        for (int q = 0; q < ns; ++q) x += Tm[i + (size_t)ns * q] * P[q + (size_t)ns * j];
        Mx[e] = x;
      }
      __syncthreads();
      // P <- Mx T' + QQ   (computed on the lower triangle, mirrored)
      for (int e = tid; e < ns * ns; e += nt) {
        const int i = e % ns, j = e / ns;
        if (i < j) continue;
        double x = QQ[e];
        for (int q = 0; q < ns; ++q) x += Mx[i + (size_t)ns * q] * Tm[j + (size_t)ns * q];
        P[e] = x;
        P[j + (size_t)ns * i] = x;
      }
      __syncthreads();
    }
    if (tid == 0) {
      const double ll = -0.5 * (nseen * 1.8378770664093453 + acc);
      if (bad || !isfinite(ll)) { g.loglik[d] = -INFINITY; g.status[d] = ST_F_NOT_PD; }
      else { g.loglik[d] = ll; g.status[d] = ST_OK; }
    }
    __syncthreads();
  }
  (void)red;
}

}  // namespace dyb
