// Model-specialised Kalman-filter log-likelihood for SMALL batches: G lanes per draw.
//
// kalman_kernel<M> (one thread per draw, everything in registers) is the throughput kernel: it needs >= ~32 768 draws to
// fill a B200.  An SMC stage on a shard of 2 048 - 16 384 particles runs it at one warp per scheduler, i.e. at the
// latency of ONE thread walking the whole T-step recursion (ls2003: 0.23 ms whatever the batch).  This kernel computes
// the same recursion (same structure exploited: Z a selection, T non-zero in the K lagged-state columns only, P
// symmetric, missing rows neutralised -- see kalman_kernel.cuh; reference call src/estimation/estimation.jl:665-701)
// with the draw's state in a shared-memory tile and the entries of every phase dealt round-robin to the G lanes of the
// draw, so a period costs ~1/G of the dependent work plus four __syncwarp.  All draws of a warp run one instruction
// stream (no data-dependent control flow), phases per period:
//   1. v, F = (Z P Z' + H), Cholesky, w = L^-1 v, log det, v'F^-1 v            every lane, in registers (NY is small)
//   2. U[:, j] = L^-1 (Z P)[:, lagged_j]                                      lane j mod G
//   3. Pll = P_ll - U'U  (K(K+1)/2 entries),  al = a_l + U'w                  entries mod G
//   4. Mrow = Tl Pll  (NS x K entries),  a <- Tl al                           entries mod G
//   5. P <- R Q R' + Mrow Tl'  (NS(NS+1)/2 entries)                           entries mod G
// Sums run in the same order as in kalman_kernel<M>, so both kernels agree to rounding of the Cholesky's rsqrt only.
#pragma once
#include "model_traits.cuh"
#include "solve_kernel.cuh"

namespace dyb {

template <class M, int G>
struct KalmanSubCfg {
  static constexpr int NS = M::NS, K = M::K, NY = M::NY;
  static constexpr int QQ_N = NS * (NS + 1) / 2, KK_N = K * (K + 1) / 2;
  static constexpr int OFF_P = 0, OFF_QQ = OFF_P + QQ_N, OFF_TL = OFF_QQ + QQ_N, OFF_MR = OFF_TL + NS * K,
                       OFF_U = OFF_MR + NS * K, OFF_PLL = OFF_U + NY * K, OFF_AL = OFF_PLL + KK_N, OFF_A = OFF_AL + K,
                       TILE_RAW = OFF_A + NS;
  static constexpr int TILE = TILE_RAW | 1;                 // odd stride: the draws of a warp start in different banks
  static constexpr int BLOCK = 128, DRAWS = BLOCK / G;
  static constexpr int NTAB = 2 * QQ_N + 2 * KK_N + K + NY; // decode tables (bytes, padded below)
  static constexpr size_t smem_bytes(int nobs) {
    return sizeof(double) * ((size_t)NY * nobs + (size_t)DRAWS * TILE) + ((NTAB + 15) / 16) * 16;
  }
};

template <class M, int G>
__global__ void __launch_bounds__(KalmanSubCfg<M, G>::BLOCK)
kalman_sublanes_kernel(const double* __restrict__ ss, long long n_draws, const double* __restrict__ Y, int nobs,
                       int presample, const int* __restrict__ status_in, double* __restrict__ loglik,
                       int* __restrict__ status_out) {
  using C = KalmanSubCfg<M, G>;
  using L = SSLayout<M>;
  constexpr int NS = M::NS, K = M::K, NY = M::NY, QQ_N = C::QQ_N, KK_N = C::KK_N;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ double sm_ks[];
  double* ysh = sm_ks;                                        // NY x nobs
  double* tiles = ysh + (size_t)NY * nobs;
  unsigned char* tab = reinterpret_cast<unsigned char*>(tiles + (size_t)C::DRAWS * C::TILE);
  unsigned char* ps_ = tab;             // row of packed NS entry
  unsigned char* pt_ = ps_ + QQ_N;      // column
  unsigned char* kj_ = pt_ + QQ_N;      // row of packed K entry
  unsigned char* kq_ = kj_ + KK_N;
  unsigned char* lag_ = kq_ + KK_N;     // lagged_state_ids
  unsigned char* obs_ = lag_ + K;       // obs_idx_state
  for (int i = threadIdx.x; i < NY * nobs; i += blockDim.x) ysh[i] = Y[i];
  for (int e = threadIdx.x; e < QQ_N; e += blockDim.x) {
    int r = 0, acc = 0;
    while (e >= acc + r + 1) { acc += r + 1; ++r; }
    ps_[e] = (unsigned char)r; pt_[e] = (unsigned char)(e - acc);
  }
  for (int e = threadIdx.x; e < KK_N; e += blockDim.x) {
    int r = 0, acc = 0;
    while (e >= acc + r + 1) { acc += r + 1; ++r; }
    kj_[e] = (unsigned char)r; kq_[e] = (unsigned char)(e - acc);
  }
  for (int j = threadIdx.x; j < K; j += blockDim.x) lag_[j] = (unsigned char)M::lagged_state_ids(j);
  for (int i = threadIdx.x; i < NY; i += blockDim.x) obs_[i] = (unsigned char)M::obs_idx_state(i);

  const int lg = threadIdx.x % G;
  long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const bool real = d < n_draws;
  if (!real) d = n_draws - 1;                                 // tail groups recompute the last draw and store nothing
  double* tl = tiles + (size_t)(threadIdx.x / G) * C::TILE;
  double* Ps = tl + C::OFF_P; double* Qs = tl + C::OFF_QQ; double* Ts = tl + C::OFF_TL; double* Ms = tl + C::OFF_MR;
  double* Us = tl + C::OFF_U; double* Pl = tl + C::OFF_PLL; double* als = tl + C::OFF_AL; double* as = tl + C::OFF_A;
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

  double acc = 0.0, detp = 1.0;
  int nseen = 0;
  bool bad = false;
  for (int t = 0; t < nobs; ++t) {
    // ---- 1. innovation, its covariance and Cholesky factor: every lane ---------------------------------
    double v[NY], F[NY * (NY + 1) / 2], rdiag[NY], w[NY];
    bool miss[NY];
    int nobs_t = 0;
#pragma unroll
    for (int i = 0; i < NY; ++i) {
      const double yi = ysh[i + NY * t];
      miss[i] = isnan(yi);
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
      detp *= det;
      if (!(detp > 1e-200 && detp < 1e200)) { acc += log(detp); detp = 1.0; }
      acc += quad;
      nseen += nobs_t;
    }
    if (t == nobs - 1) break;
    // ---- 2. U = L^-1 (Z P)[:, lagged]: column j on lane j mod G -------------------------------------------
    for (int j = lg; j < K; j += G) {
      const int lj = lag_[j];
      double u[NY];
#pragma unroll
      for (int i = 0; i < NY; ++i) {
        const int oi = M::obs_idx_state(i);
        double x = miss[i] ? 0.0 : Ps[oi >= lj ? tri(oi, lj) : tri(lj, oi)];
#pragma unroll
        for (int q = 0; q < i; ++q) x = fma(-F[tri(i, q)], u[q], x);
        u[i] = x * rdiag[i];
        Us[i + NY * j] = u[i];
      }
    }
    __syncwarp(FULL);
    // ---- 3. updated mean / covariance on the lagged states ------------------------------------------------------
    for (int e = lg; e < KK_N; e += G) {
      const int j = kj_[e], q = kq_[e];
      const int lj = lag_[j], lq = lag_[q];
      double pjq = Ps[lj >= lq ? tri(lj, lq) : tri(lq, lj)];
#pragma unroll
      for (int i = 0; i < NY; ++i) pjq = fma(-Us[i + NY * j], Us[i + NY * q], pjq);
      Pl[e] = pjq;
    }
    for (int j = lg; j < K; j += G) {
      double x = as[lag_[j]];
#pragma unroll
      for (int i = 0; i < NY; ++i) x = fma(Us[i + NY * j], w[i], x);
      als[j] = x;
    }
    __syncwarp(FULL);
    // ---- 4. Mrow = Tl Pll, a <- Tl al -----------------------------------------------------------------------------
    for (int e = lg; e < NS * K; e += G) {
      const int s = e / K, j = e - s * K;
      double m = 0.0;
#pragma unroll
      for (int q = 0; q < K; ++q) m = fma(Ts[s * K + q], Pl[q >= j ? tri(q, j) : tri(j, q)], m);
      Ms[e] = m;
    }
    for (int s = lg; s < NS; s += G) {
      double x = 0.0;
#pragma unroll
      for (int j = 0; j < K; ++j) x = fma(Ts[s * K + j], als[j], x);
      as[s] = x;
    }
    __syncwarp(FULL);
    // ---- 5. P <- R Q R' + Mrow Tl' -------------------------------------------------------------------------------
    for (int e = lg; e < QQ_N; e += G) {
      const int s = ps_[e], t2 = pt_[e];
      double x2 = Qs[e];
#pragma unroll
      for (int j = 0; j < K; ++j) x2 = fma(Ms[s * K + j], Ts[t2 * K + j], x2);
      Ps[e] = x2;
    }
    __syncwarp(FULL);
  }
  acc += log(detp);
  const double ll = -0.5 * (nseen * 1.8378770664093453 + acc);
  if (real && lg == 0) {
    if (st_in != ST_OK) { loglik[d] = -INFINITY; status_out[d] = st_in; }
    else if (bad || !isfinite(ll)) { loglik[d] = -INFINITY; status_out[d] = ST_F_NOT_PD; }
    else { loglik[d] = ll; status_out[d] = ST_OK; }
  }
}

}  // namespace dyb
