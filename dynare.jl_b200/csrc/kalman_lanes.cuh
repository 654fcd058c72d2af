// Dense batched Kalman log-likelihood for SMALL state spaces of run-time size (ns <= 16, ny <= 8): G = four lanes per
// draw, eight draws per warp (eight lanes and four draws for 12 < ns <= 16, whose rows of T would not fit the registers
// of four lanes).  A CTA of two warps per draw with six block barriers per period (kalman_dmma.cuh) is the
// wrong mapping below ~16 states: the tiles are mostly padding and the warps mostly wait.  Here lane g of a draw owns
// the rows / columns i = g, g + 4, ... : it keeps those rows of T in registers, the covariance lives in a per-draw
// shared-memory tile, and the time update is two register-blocked products
//     W[:, j] = P T[j, :]'   (owned columns j)      P'[i, :] = T[i, :] W + (R Q R')[i, :]   (owned rows i)
// with one __syncwarp between them.  The innovation covariance (ny <= 8) is factorised redundantly by every lane in
// registers.  Control flow is uniform across the warp (sizes and the missing-data pattern are shared by all draws).
//
// Replaces `kalman_likelihood(Y,Z,H,T,R,Q,a0,P,start,last,presample,ws)` (reference call site
// src/estimation/estimation.jl:687-700) for a selection matrix Z; arguments shared with kalman_generic.cuh.
#pragma once
#include "kalman_generic.cuh"

namespace dyb {

constexpr int KLN_NY = 8;             // largest ny
constexpr int KLN_THREADS = 64;       // threads per CTA: 64 / G draws

template <int NSMAX, int G>
struct KalmanLanesCfg {
  static constexpr int R = NSMAX / G;                            // owned rows per lane
  static constexpr int DRAWS = KLN_THREADS / G;                  // draws per CTA
  // per-draw tile: P, W (NSMAX^2 each), U (KLN_NY x NSMAX), a, au; padded to an odd number of doubles so that the
  // eight draws of a warp sit in distinct banks
  static constexpr int RAW = 2 * NSMAX * NSMAX + KLN_NY * NSMAX + 2 * NSMAX;
  static constexpr int TILE = RAW | 1;
  static constexpr size_t SMEM_BYTES = sizeof(double) * (size_t)TILE * DRAWS;
};

template <int NSMAX, int G>
__global__ void __launch_bounds__(KLN_THREADS) kalman_lanes_kernel(KalmanGenericArgs g) {
  using Cfg = KalmanLanesCfg<NSMAX, G>;
  constexpr int KLN_G = G, KLN_DRAWS = Cfg::DRAWS;
  constexpr int R = Cfg::R;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ double sm[];
  const int ns = g.ns, ny = g.ny;
  const int lg = threadIdx.x & (KLN_G - 1);                       // lane inside the draw's group
  const int slot = threadIdx.x / KLN_G;                           // draw slot inside the CTA
  double* tile = sm + (size_t)slot * Cfg::TILE;
  double* P = tile;                                               // P[i + NSMAX * j]
  double* W = P + NSMAX * NSMAX;
  double* U = W + NSMAX * NSMAX;                                  // U[q + KLN_NY * s]
  double* a = U + KLN_NY * NSMAX;
  double* au = a + NSMAX;
  int zi[KLN_NY];
#pragma unroll
  for (int i = 0; i < KLN_NY; ++i) zi[i] = i < ny ? g.z_idx[i] : 0;

  const long long per_pass = (long long)gridDim.x * KLN_DRAWS;
  for (long long d0 = (long long)blockIdx.x * KLN_DRAWS; d0 < g.n_draws; d0 += per_pass) {
    const long long d = d0 + slot;
    const bool real = d < g.n_draws;
    const long long dd = real ? d : g.n_draws - 1;               // idle slots shadow the last draw (results discarded)
    const bool skip = g.status_in && g.status_in[dd] != ST_OK;
    const double* Tm = g.T + (size_t)dd * ns * ns;
    const double* QQ = g.RQR + (size_t)dd * ns * ns;
    const double* P0 = g.P0 + (size_t)dd * ns * ns;
    const double* Hm = g.H ? g.H + (size_t)dd * ny * ny : nullptr;
    const double* yb = g.ybar_obs ? g.ybar_obs + (size_t)dd * ny : nullptr;
    __syncwarp(FULL);
    // rows of T owned by this lane, in registers
    double Tr[R][NSMAX];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = lg + KLN_G * r;
#pragma unroll
      for (int k = 0; k < NSMAX; ++k) Tr[r][k] = (i < ns && k < ns) ? Tm[i + (size_t)ns * k] : 0.0;
    }
    for (int e = lg; e < NSMAX * NSMAX; e += KLN_G) {
      const int i = e % NSMAX, j = e / NSMAX;
      P[e] = (i < ns && j < ns) ? P0[i + (size_t)ns * j] : 0.0;
      W[e] = 0.0;
    }
    for (int e = lg; e < KLN_NY * NSMAX; e += KLN_G) U[e] = 0.0;
    for (int i = lg; i < NSMAX; i += KLN_G) { a[i] = (g.a0 && i < ns) ? g.a0[(size_t)dd * ns + i] : 0.0; au[i] = 0.0; }
    __syncwarp(FULL);
    double acc = 0.0;
    LogDetAcc detp;
    int nseen = 0;
    bool bad = false;

    for (int t = 0; t < g.nobs; ++t) {
      const double* yt = g.Y + (size_t)t * ny;
      // ---- innovation, its covariance and Cholesky factor: every lane, in registers --------------------------------
      double v[KLN_NY], F[KLN_NY][KLN_NY], rd[KLN_NY], w[KLN_NY];
      bool miss[KLN_NY];
      int nobs_t = 0;
#pragma unroll
      for (int i = 0; i < KLN_NY; ++i) {
        const double yi = i < ny ? yt[i] : 0.0;
        miss[i] = i < ny ? isnan(yi) : true;
        nobs_t += miss[i] ? 0 : 1;
        v[i] = miss[i] ? 0.0 : (yi - (yb ? yb[i] : 0.0)) - a[zi[i]];
      }
#pragma unroll
      for (int i = 0; i < KLN_NY; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          double f = (i < ny) ? P[zi[i] + NSMAX * zi[j]] + (Hm ? Hm[i + ny * j] : 0.0) : 0.0;
          if (miss[i] || miss[j]) f = (i == j) ? 1.0 : 0.0;
          F[i][j] = f;
        }
      double det = 1.0, quad = 0.0;
#pragma unroll
      for (int j = 0; j < KLN_NY; ++j) {
        double djj = F[j][j];
#pragma unroll
        for (int q = 0; q < j; ++q) djj = fma(-F[j][q], F[j][q], djj);
        if (j < ny && (!(djj > 0.0) || !(djj < 1e300))) bad = true;
        det *= djj;
        const double r = fast_rsqrt(djj);
        rd[j] = r;
        F[j][j] = djj * r;
#pragma unroll
        for (int i = j + 1; i < KLN_NY; ++i) {
          double x = F[i][j];
#pragma unroll
          for (int q = 0; q < j; ++q) x = fma(-F[i][q], F[j][q], x);
          F[i][j] = x * r;
        }
      }
#pragma unroll
      for (int i = 0; i < KLN_NY; ++i) {
        double x = v[i];
#pragma unroll
        for (int q = 0; q < i; ++q) x = fma(-F[i][q], w[q], x);
        w[i] = x * rd[i];
        quad = fma(w[i], w[i], quad);
      }
      if (t >= g.presample) {
        detp.mul(det);
        acc += quad;
        nseen += nobs_t;
      }
      if (t == g.nobs - 1) break;
      // ---- U[:, s] = L^-1 (Z P)[:, s], au[s] = a[s] + U[:, s]'w for the owned states s ------------------------------
      double Uown[R][KLN_NY];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int s = lg + KLN_G * r;
        double x = a[s];
#pragma unroll
        for (int i = 0; i < KLN_NY; ++i) {
          double u = miss[i] ? 0.0 : P[zi[i] + NSMAX * s];
#pragma unroll
          for (int q = 0; q < i; ++q) u = fma(-F[i][q], Uown[r][q], u);
          u *= rd[i];
          Uown[r][i] = u;
          U[i + KLN_NY * s] = u;
          x = fma(u, w[i], x);
        }
        au[s] = x;
      }
      __syncwarp(FULL);
      // ---- a' = T au (owned rows) ; P <- P - U'U (owned rows) -------------------------------------------------------
      double anew[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        double x = 0.0;
#pragma unroll
        for (int k = 0; k < NSMAX; ++k) x = fma(Tr[r][k], au[k], x);
        anew[r] = x;
      }
      for (int j = 0; j < ns; ++j) {
        double uj[KLN_NY];
#pragma unroll
        for (int q = 0; q < KLN_NY; ++q) uj[q] = U[q + KLN_NY * j];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = lg + KLN_G * r;
          double x = P[i + NSMAX * j];
#pragma unroll
          for (int q = 0; q < KLN_NY; ++q) x = fma(-Uown[r][q], uj[q], x);
          P[i + NSMAX * j] = x;
        }
      }
      __syncwarp(FULL);
#pragma unroll
      for (int r = 0; r < R; ++r) a[lg + KLN_G * r] = anew[r];
      // ---- W[:, j] = P T[j, :]' for the owned columns j -------------------------------------------------------------
      {
        double c[R][NSMAX];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int i = 0; i < NSMAX; ++i) c[r][i] = 0.0;
#pragma unroll
        for (int k = 0; k < NSMAX; ++k) {
          if (k < ns) {
            double pk[NSMAX];
#pragma unroll
            for (int i = 0; i < NSMAX; ++i) pk[i] = P[i + NSMAX * k];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
              for (int i = 0; i < NSMAX; ++i) c[r][i] = fma(pk[i], Tr[r][k], c[r][i]);
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int i = 0; i < NSMAX; ++i) W[i + NSMAX * (lg + KLN_G * r)] = c[r][i];
      }
      __syncwarp(FULL);
      // ---- P'[i, :] = T[i, :] W + (R Q R')[i, :] for the owned rows i -------------------------------------------------
      {
        double c[R][NSMAX];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = lg + KLN_G * r;
#pragma unroll
          for (int j = 0; j < NSMAX; ++j) c[r][j] = (i < ns && j < ns) ? QQ[i + (size_t)ns * j] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < NSMAX; ++k) {
          if (k < ns) {
            double wk[NSMAX];
#pragma unroll
            for (int j = 0; j < NSMAX; ++j) wk[j] = W[k + NSMAX * j];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
              for (int j = 0; j < NSMAX; ++j) c[r][j] = fma(Tr[r][k], wk[j], c[r][j]);
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int j = 0; j < NSMAX; ++j) P[(lg + KLN_G * r) + NSMAX * j] = c[r][j];
      }
      __syncwarp(FULL);
    }
    if (real && lg == 0) {
      if (skip) {
        g.loglik[d] = -INFINITY; g.status[d] = g.status_in[d];
      } else {
        acc += detp.value();
        const double ll = -0.5 * (nseen * 1.8378770664093453 + acc);
        if (bad || !isfinite(ll)) { g.loglik[d] = -INFINITY; g.status[d] = ST_F_NOT_PD; }
        else { g.loglik[d] = ll; g.status[d] = ST_OK; }
      }
    }
  }
}

}  // namespace dyb
