// Model-specialised Kalman-filter log-likelihood for the SMALLEST batches: G = 16 or 32 lanes per draw, every stage of a
// period dealt over the lanes.
//
// kalman_sublanes_kernel<M, 4 | 8> deals only the time update over its lanes and repeats the measurement update in every
// lane, so one period still costs every lane ~350 instructions of a serial chain (ls2003: 2.1 us per period whatever the
// batch).  An SMC stage on a shard of 2 048 particles (config 3 on 8 GPUs) sits on that latency.  Here a draw owns a whole
// warp (or half of one) and a period is five short cooperative phases on a shared-memory tile:
//   M  the augmented array [F | C | v] = [Z P Z' + H | (Z P)[:, lagged] | y - ybar - Z a] (NY rows, NY + K + 1 columns), ONE
//      COLUMN PER LANE in registers, is reduced by a right-looking Cholesky elimination: step j broadcasts row j of the F
//      block (shuffles), every lane scales its entry of row j by 1 / L_jj and updates its rows below.  Afterwards the C
//      lanes hold U = L^-1 (Z P)[:, lagged] and the v lane holds w = L^-1 v: the factorisation and both triangular solves
//      are one pass of NY steps;
//   B  U and w go to the tile; lane e computes entry e of P_ll - U'U (and the C lanes a_l + U'w);
//   C1 lane e computes entry e of [Tl P_ll | Tl a_l] (rows of Tl in registers for the whole filter);
//   C2 lane e computes entry e of the packed lower triangle of P' = (Tl P_ll) Tl' + R Q R'.
// Every scalar is produced by exactly the sequence of operations kalman_kernel<M> uses for it (same fma order, same
// operands), so the three kernels return the same bits and a sharded SMC run stays bit-identical to the single-GPU run.
// Reference call replaced: kalman_likelihood(...) at src/estimation/estimation.jl:665-701 (see kalman_kernel.cuh).
#pragma once
#include "model_traits.cuh"
#include "solve_kernel.cuh"

namespace dyb {

template <class M, int G>
struct KalmanCoopCfg {
  static constexpr int NS = M::NS, K = M::K, NY = M::NY;
  static constexpr int QQ_N = NS * (NS + 1) / 2, KK_N = K * (K + 1) / 2;
  static constexpr int NC = NY + K + 1;                         // columns of [F | C | v]
  static constexpr int NME = NS * (K + 1);                      // entries of [Tl P_ll | Tl a_l]
  static constexpr int MS = (NME + G - 1) / G, PS = (QQ_N + G - 1) / G, LS = (KK_N + G - 1) / G;
  // tile of one draw (doubles): packed P, a, [U | w] by columns, full K x K P_ll followed by a_l, Tl P_ll by rows
  static constexpr int OFF_P = 0, OFF_A = OFF_P + QQ_N, OFF_U = OFF_A + NS, OFF_PLL = OFF_U + NY * (K + 1),
                       OFF_M = OFF_PLL + K * K + K, TILE_RAW = OFF_M + NS * K;
  static constexpr int TILE = TILE_RAW | 1;
  static constexpr int BLOCK = 64, DRAWS = BLOCK / G;
  static constexpr int MIN_CTAS = G == 32 ? 7 : 4;             // 2 048 draws resident at once: 14 warps per SM with 32 lanes, 7 with 16
  static constexpr bool FITS = (G == 16 || G == 32) && NC <= G && NY <= 30 &&
                               (MS * K + PS * (K + 1) + 3 * NY + LS) <= 100;     // doubles a lane keeps in registers
  static constexpr size_t smem_bytes(int nobs) {
    return sizeof(double) * ((size_t)NY * nobs + (size_t)(nobs + 1) / 2 + (size_t)DRAWS * TILE);
  }
};

template <class M, int G>
__global__ void __launch_bounds__(KalmanCoopCfg<M, G>::BLOCK, KalmanCoopCfg<M, G>::MIN_CTAS)
kalman_coop_kernel(const double* __restrict__ ss, long long n_draws, const double* __restrict__ Y, int nobs,
                   int presample, const int* __restrict__ status_in, double* __restrict__ loglik,
                   int* __restrict__ status_out) {
  using C = KalmanCoopCfg<M, G>;
  using L = SSLayout<M>;
  constexpr int K = M::K, NY = M::NY, QQ_N = C::QQ_N, KK_N = C::KK_N;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ double sm_kc[];
  double* ysh = sm_kc;                                          // NY x nobs
  int* msk = reinterpret_cast<int*>(ysh + (size_t)NY * nobs);   // missing-observation pattern of every period
  double* tiles = ysh + (size_t)NY * nobs + (size_t)(nobs + 1) / 2;
  for (int i = threadIdx.x; i < NY * nobs; i += blockDim.x) ysh[i] = Y[i];
  __syncthreads();
  for (int t = threadIdx.x; t < nobs; t += blockDim.x) {
    int m = 0;
    for (int i = 0; i < NY; ++i) m |= isnan(ysh[i + NY * t]) ? (1 << i) : 0;
    msk[t] = m;
  }

  const int lg = threadIdx.x % G, grp = threadIdx.x / G;
  long long d = (long long)blockIdx.x * C::DRAWS + grp;
  const bool real = d < n_draws;
  if (!real) d = n_draws - 1;                                   // tail groups recompute the last draw and store nothing
  double* tl = tiles + (size_t)grp * C::TILE;
  const double* s_ = ss + d;
  const long long st = n_draws;
  const int st_in = status_in[d];
  const bool okin = st_in == ST_OK;                             // a failed draw's hand-off is undefined: it runs on zeros
  for (int e = lg; e < QQ_N; e += G) tl[C::OFF_P + e] = okin ? s_[(L::OFF_P0 + e) * st] : 0.0;
  for (int e = lg; e < C::TILE - QQ_N; e += G) tl[QQ_N + e] = 0.0;

  // ---- this lane's column of [F | C | v] ---------------------------------------------------------------------------
  const int c = lg;
  const bool isF = c < NY, isC = c >= NY && c < NY + K, isV = c == NY + K;
  const int xs = isF ? M::obs_idx_state(isF ? c : 0) : (isC ? M::lagged_state_ids(isC ? c - NY : 0) : 0);
  int goff[NY];
  double Hc[NY], yb[NY];
#pragma unroll
  for (int i = 0; i < NY; ++i) {
    goff[i] = isV ? C::OFF_A + M::obs_idx_state(i) : C::OFF_P + ((isF || isC) ? sym(M::obs_idx_state(i), xs) : 0);
    Hc[i] = (isF && okin) ? s_[(L::OFF_H + sym(i, isF ? c : 0)) * st] : 0.0;
    yb[i] = okin ? s_[(L::OFF_YB + i) * st] : 0.0;
  }
  const double sgn = isV ? -1.0 : 1.0;
  const int al_src = C::OFF_A + (isC ? M::lagged_state_ids(isC ? c - NY : 0) : 0);
  // ---- entries of P_ll this lane computes ------------------------------------------------------------------------
  int l_uj[C::LS], l_uq[C::LS], l_p[C::LS], l_d1[C::LS], l_d2[C::LS];
#pragma unroll
  for (int ei = 0; ei < C::LS; ++ei) {
    int e = lg + G * ei;
    if (e >= KK_N) e = KK_N - 1;
    int j = 0;
    while ((j + 1) * (j + 2) / 2 <= e) ++j;
    const int q = e - j * (j + 1) / 2;
    l_uj[ei] = C::OFF_U + NY * j; l_uq[ei] = C::OFF_U + NY * q;
    l_p[ei] = C::OFF_P + sym(M::lagged_state_ids(j), M::lagged_state_ids(q));
    l_d1[ei] = C::OFF_PLL + K * j + q; l_d2[ei] = C::OFF_PLL + K * q + j;
  }
  // ---- entries of [Tl P_ll | Tl a_l]: (s, j), j = K is the mean ---------------------------------------------------
  double trM[C::MS][K];
  int m_src[C::MS], m_dst[C::MS];
#pragma unroll
  for (int ei = 0; ei < C::MS; ++ei) {
    int e = lg + G * ei;
    if (e >= C::NME) e = C::NME - 1;
    const int s = e / (K + 1), j = e % (K + 1);
    m_src[ei] = C::OFF_PLL + K * j;
    m_dst[ei] = j < K ? C::OFF_M + s * K + j : C::OFF_A + s;
#pragma unroll
    for (int q = 0; q < K; ++q) trM[ei][q] = okin ? s_[(L::OFF_T + s * K + q) * st] : 0.0;
  }
  // ---- entries of the packed lower triangle of P' -----------------------------------------------------------------
  double trP[C::PS][K], qq[C::PS];
  int p_src[C::PS];
#pragma unroll
  for (int ei = 0; ei < C::PS; ++ei) {
    int e = lg + G * ei;
    if (e >= QQ_N) e = QQ_N - 1;
    int s = 0;
    while ((s + 1) * (s + 2) / 2 <= e) ++s;
    const int t2 = e - s * (s + 1) / 2;
    p_src[ei] = C::OFF_M + s * K;
    qq[ei] = okin ? s_[(L::OFF_QQ + e) * st] : 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) trP[ei][j] = okin ? s_[(L::OFF_T + t2 * K + j) * st] : 0.0;
  }
  __syncthreads();

  double acc = 0.0;
  LogDetAcc detp;
  int nseen = 0;
  bool bad = false;
  for (int t = 0; t < nobs; ++t) {
    // ---- phase M: this lane's column, then NY elimination steps -------------------------------------------------
    const int mm = msk[t];
    double g[NY];
#pragma unroll
    for (int i = 0; i < NY; ++i) {
      const double yv = ysh[i + NY * t] - yb[i];
      g[i] = fma(sgn, tl[goff[i]], isV ? yv : Hc[i]);
    }
    if (mm != 0) {                            // some observation is missing (uniform: the data are shared by every draw)
#pragma unroll
      for (int i = 0; i < NY; ++i) {
        const bool mi = (mm >> i) & 1, mc = isF && ((mm >> c) & 1);
        if (mi || mc) g[i] = (isF && i == c) ? 1.0 : 0.0;
      }
    }
    double det = 1.0;
#pragma unroll
    for (int j = 0; j < NY; ++j) {
      double xr[NY];                          // row j of the F block: entry (j, i) lives in lane i
#pragma unroll
      for (int i = j; i < NY; ++i) xr[i] = __shfl_sync(FULL, g[j], i, G);
      const double djj = xr[j];
      if (!(djj > 0.0) || !(djj < 1e300)) bad = true;
      det *= djj;
      const double r = fast_rsqrt(djj);
      g[j] *= r;
#pragma unroll
      for (int i = j + 1; i < NY; ++i) g[i] = fma(-(xr[i] * r), g[j], g[i]);
    }
    double quad = 0.0;
#pragma unroll
    for (int i = 0; i < NY; ++i) quad = fma(g[i], g[i], quad);       // meaningful in the v lane only
    if (t >= presample) {
      detp.mul(det);
      acc += quad;
      nseen += NY - __popc(mm);
    }
    if (t == nobs - 1) break;
    // ---- phase B: U, w to the tile; a_l and P_ll ------------------------------------------------------------------
    if (isC || isV) {
#pragma unroll
      for (int i = 0; i < NY; ++i) tl[C::OFF_U + NY * (c - NY) + i] = g[i];
    }
    __syncwarp(FULL);
    double xal = tl[al_src];                    // meaningful in the C lanes only
#pragma unroll
    for (int i = 0; i < NY; ++i) xal = fma(g[i], tl[C::OFF_U + NY * K + i], xal);
    // (loads, then the interleaved chains, then the stores: a store between two slots would order them, the tile being one
    // array to the compiler)
    double pl[C::LS];
#pragma unroll
    for (int ei = 0; ei < C::LS; ++ei) pl[ei] = tl[l_p[ei]];
#pragma unroll
    for (int i = 0; i < NY; ++i)
#pragma unroll
      for (int ei = 0; ei < C::LS; ++ei) pl[ei] = fma(-tl[l_uj[ei] + i], tl[l_uq[ei] + i], pl[ei]);
#pragma unroll
    for (int ei = 0; ei < C::LS; ++ei)
      if (lg + G * ei < KK_N) { tl[l_d1[ei]] = pl[ei]; tl[l_d2[ei]] = pl[ei]; }
    if (isC) tl[C::OFF_PLL + K * K + (c - NY)] = xal;
    __syncwarp(FULL);
    // ---- phase C1: [Tl P_ll | Tl a_l] ---------------------------------------------------------------------------------
    double mv[C::MS];
#pragma unroll
    for (int ei = 0; ei < C::MS; ++ei) mv[ei] = 0.0;
#pragma unroll
    for (int q = 0; q < K; ++q)
#pragma unroll
      for (int ei = 0; ei < C::MS; ++ei) mv[ei] = fma(trM[ei][q], tl[m_src[ei] + q], mv[ei]);
#pragma unroll
    for (int ei = 0; ei < C::MS; ++ei)
      if (lg + G * ei < C::NME) tl[m_dst[ei]] = mv[ei];
    __syncwarp(FULL);
    // ---- phase C2: P' = (Tl P_ll) Tl' + R Q R', packed lower triangle ---------------------------------------------------
    double pv[C::PS];
#pragma unroll
    for (int ei = 0; ei < C::PS; ++ei) pv[ei] = qq[ei];
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int ei = 0; ei < C::PS; ++ei) pv[ei] = fma(tl[p_src[ei] + j], trP[ei][j], pv[ei]);
#pragma unroll
    for (int ei = 0; ei < C::PS; ++ei)
      if (lg + G * ei < QQ_N) tl[C::OFF_P + lg + G * ei] = pv[ei];
    __syncwarp(FULL);
  }
  acc = __shfl_sync(FULL, acc, NY + K, G);
  acc += detp.value();
  const double ll = -0.5 * (nseen * 1.8378770664093453 + acc);
  if (real && lg == 0) {
    if (!okin) { loglik[d] = -INFINITY; status_out[d] = st_in; }
    else if (bad || !isfinite(ll)) { loglik[d] = -INFINITY; status_out[d] = ST_F_NOT_PD; }
    else { loglik[d] = ll; status_out[d] = ST_OK; }
  }
}

}  // namespace dyb
