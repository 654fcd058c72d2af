// Model-specialised Kalman-filter log-likelihood for SMALL batches: G lanes per draw.
//
// kalman_kernel<M> (one thread per draw, everything in registers) is the throughput kernel: it needs >= ~16 384 draws to
// occupy every scheduler of a B200.  An SMC stage on a shard of 2 048 - 8 192 particles runs it at less than one warp per
// scheduler, i.e. at the latency of ONE thread walking the whole T-step recursion (ls2003: 0.23 ms whatever the batch).
// This kernel computes the same recursion (same structure exploited: Z a selection, T non-zero in the K lagged-state
// columns only, P symmetric, missing rows neutralised -- see kalman_kernel.cuh; reference call
// src/estimation/estimation.jl:665-701) with the draw's P, a, Tl, R Q R' in a shared-memory tile:
//   * the small dense algebra of a period -- v, F = Z P Z' + H, its Cholesky factor, w = L^-1 v, U = L^-1 (Z P)[:, lagged],
//     the updated mean al and covariance Pll on the K lagged states -- is done by EVERY lane of the draw in registers
//     (static indices, no exchange): it is O(NY^2 K + NY K^2) and its inputs are a few entries of the tile;
//   * the time update, the O(NS^2 K) bulk, is dealt by ROW: lane g updates rows s = g, g + G, ... of
//     a <- Tl al and P <- R Q R' + (Tl Pll) Tl' (lower triangle) and writes them back to the tile;
//   * two __syncwarp per period (tile read -> tile written -> next period); all draws of a warp run one instruction stream.
// Sums run in the same order as in kalman_kernel<M>: both kernels return the same bits.
#pragma once
#include "model_traits.cuh"
#include "solve_kernel.cuh"

namespace dyb {

template <class M, int G>
struct KalmanSubCfg {
  static constexpr int NS = M::NS, K = M::K, NY = M::NY;
  static constexpr int QQ_N = NS * (NS + 1) / 2, KK_N = K * (K + 1) / 2;
  static constexpr int OFF_P = 0, OFF_QQ = OFF_P + QQ_N, OFF_TL = OFF_QQ + QQ_N, OFF_A = OFF_TL + NS * K, TILE_RAW = OFF_A + NS;
  static constexpr int TILE = TILE_RAW | 1;                 // odd stride: the draws of a warp start in different banks
  static constexpr int BLOCK = 128, DRAWS = BLOCK / G;
  // doubles every lane keeps in registers: F, U, Pll, al, v, w, rdiag, one row of Tl and of Tl Pll
  static constexpr int REG_DOUBLES = NY * (NY + 1) / 2 + NY * K + KK_N + K + 3 * NY + 2 * K;
  static constexpr bool FITS = REG_DOUBLES <= 104 && NS <= 255;
  static constexpr size_t smem_bytes(int nobs) { return sizeof(double) * ((size_t)NY * nobs + (size_t)DRAWS * TILE); }
};

// the recursion of the CTA's draws; MISS = false: data without a single missing observation (uniform: the pattern is shared
// by every draw), the neutralising selects of the measurement update compiled out -- as kalman_draw<M, MISS>
template <class M, int G, bool MISS>
DYB_D void kalman_sublanes_draws(const double* __restrict__ ss, long long n_draws, double* sm_ks, int nobs,
                                 int presample, const int* __restrict__ status_in, double* __restrict__ loglik,
                                 int* __restrict__ status_out) {
  using C = KalmanSubCfg<M, G>;
  using L = SSLayout<M>;
  constexpr int NS = M::NS, K = M::K, NY = M::NY, QQ_N = C::QQ_N, KK_N = C::KK_N;
  constexpr unsigned FULL = 0xffffffffu;
  double* ysh = sm_ks;                                        // NY x nobs, filled by the kernel
  double* tiles = ysh + (size_t)NY * nobs;

  const int lg = threadIdx.x % G;
  long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const bool real = d < n_draws;
  if (!real) d = n_draws - 1;                                 // tail groups recompute the last draw and store nothing
  double* tl = tiles + (size_t)(threadIdx.x / G) * C::TILE;
  double* Ps = tl + C::OFF_P; double* Qs = tl + C::OFF_QQ; double* Ts = tl + C::OFF_TL; double* as = tl + C::OFF_A;
  const double* s_ = ss + d;
  const long long st = n_draws;
  const int st_in = status_in[d];
  // the draw's transition data; a failed draw's hand-off entries are undefined: it runs on zeros
  for (int e = lg; e < NS * K; e += G) Ts[e] = (st_in == ST_OK) ? s_[(L::OFF_T + e) * st] : 0.0;
  for (int e = lg; e < QQ_N; e += G) {
    Qs[e] = (st_in == ST_OK) ? s_[(L::OFF_QQ + e) * st] : 0.0;
    Ps[e] = (st_in == ST_OK) ? s_[(L::OFF_P0 + e) * st] : 0.0;
  }
  for (int e = lg; e < NS; e += G) as[e] = 0.0;
  double yb[NY], H[NY * (NY + 1) / 2];
#pragma unroll
  for (int i = 0; i < NY; ++i) yb[i] = (st_in == ST_OK) ? s_[(L::OFF_YB + i) * st] : 0.0;
#pragma unroll
  for (int i = 0; i < NY * (NY + 1) / 2; ++i) H[i] = (st_in == ST_OK) ? s_[(L::OFF_H + i) * st] : 0.0;
  __syncthreads();

  double acc = 0.0;
  LogDetAcc detp;
  int nseen = 0;
  bool bad = false;
  for (int t = 0; t < nobs; ++t) {
    // ---- measurement update, every lane in registers (as kalman_kernel<M>) ------------------------------------
    double v[NY], F[NY * (NY + 1) / 2], rdiag[NY], w[NY];
    bool miss[NY];
    int nobs_t = 0;
#pragma unroll
    for (int i = 0; i < NY; ++i) {
      const double yi = ysh[i + NY * t];
      miss[i] = MISS && isnan(yi);
      nobs_t += miss[i] ? 0 : 1;
      v[i] = miss[i] ? 0.0 : (yi - yb[i] - as[M::obs_idx_state(i)]);
    }
#pragma unroll
    for (int i = 0; i < NY; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        double f = Ps[sym(M::obs_idx_state(i), M::obs_idx_state(j))] + H[tri(i, j)];
        if (miss[i] || miss[j]) f = (i == j) ? 1.0 : 0.0;
        F[tri(i, j)] = f;
      }
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
      detp.mul(det);
      acc += quad;
      nseen += nobs_t;
    }
    if (t == nobs - 1) break;
    double U[NY * K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
#pragma unroll
      for (int i = 0; i < NY; ++i) {
        double x = miss[i] ? 0.0 : Ps[sym(M::obs_idx_state(i), M::lagged_state_ids(j))];
#pragma unroll
        for (int q = 0; q < i; ++q) x = fma(-F[tri(i, q)], U[q + NY * j], x);
        U[i + NY * j] = x * rdiag[i];
      }
    }
    double al[K], Pll[KK_N];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      double x = as[M::lagged_state_ids(j)];
#pragma unroll
      for (int i = 0; i < NY; ++i) x = fma(U[i + NY * j], w[i], x);
      al[j] = x;
#pragma unroll
      for (int q = 0; q <= j; ++q) {
        double pjq = Ps[sym(M::lagged_state_ids(j), M::lagged_state_ids(q))];
#pragma unroll
        for (int i = 0; i < NY; ++i) pjq = fma(-U[i + NY * j], U[i + NY * q], pjq);
        Pll[tri(j, q)] = pjq;
      }
    }
    __syncwarp(FULL);                       // every lane has read a and P of this period
    // ---- time update by rows: lane lg owns rows lg, lg + G, ... -------------------------------------------------
#pragma unroll
    for (int r0 = 0; r0 < NS; r0 += G) {
      const int s = r0 + lg;
      if (s < NS) {
        double trow[K], mrow[K];
#pragma unroll
        for (int j = 0; j < K; ++j) trow[j] = Ts[s * K + j];
        double x = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) x = fma(trow[j], al[j], x);
#pragma unroll
        for (int j = 0; j < K; ++j) {
          double m = 0.0;
#pragma unroll
          for (int q = 0; q < K; ++q) m = fma(trow[q], Pll[sym(q, j)], m);
          mrow[j] = m;
        }
        const int base = s * (s + 1) / 2;
#pragma unroll
        for (int t2 = 0; t2 < NS; ++t2) {
          if (t2 <= s && t2 < r0 + G) {
            double x2 = Qs[base + t2];
#pragma unroll
            for (int j = 0; j < K; ++j) x2 = fma(mrow[j], Ts[t2 * K + j], x2);
            Ps[base + t2] = x2;
          }
        }
        as[s] = x;
      }
    }
    __syncwarp(FULL);
  }
  acc += detp.value();
  const double ll = -0.5 * (nseen * 1.8378770664093453 + acc);
  if (real && lg == 0) {
    if (st_in != ST_OK) { loglik[d] = -INFINITY; status_out[d] = st_in; }
    else if (bad || !isfinite(ll)) { loglik[d] = -INFINITY; status_out[d] = ST_F_NOT_PD; }
    else { loglik[d] = ll; status_out[d] = ST_OK; }
  }
}

template <class M, int G>
__global__ void __launch_bounds__(KalmanSubCfg<M, G>::BLOCK)
kalman_sublanes_kernel(const double* __restrict__ ss, long long n_draws, const double* __restrict__ Y, int nobs,
                       int presample, const int* __restrict__ status_in, double* __restrict__ loglik,
                       int* __restrict__ status_out) {
  extern __shared__ double sm_ks[];
  int nan_seen = 0;
  for (int i = threadIdx.x; i < M::NY * nobs; i += blockDim.x) { const double y = Y[i]; sm_ks[i] = y; nan_seen |= isnan(y) ? 1 : 0; }
  const int any_missing = __syncthreads_or(nan_seen);
  if (any_missing) kalman_sublanes_draws<M, G, true>(ss, n_draws, sm_ks, nobs, presample, status_in, loglik, status_out);
  else kalman_sublanes_draws<M, G, false>(ss, n_draws, sm_ks, nobs, presample, status_in, loglik, status_out);
}

}  // namespace dyb
