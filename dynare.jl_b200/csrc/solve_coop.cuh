// Register-resident first-order solve in three kernels (the production path for models whose dynamic
// block fits the register file; solve_kernel.cuh keeps the size-agnostic one-thread-per-draw variant).
//
//   model_kernel<M>      one thread per draw: parameter scatter, steady state + residual check, Jacobian
//                        and the structure-aware static elimination generated at model set-up
//                        (M::static_elim).  Writes the non-zeros of the reduced system and of the pivot rows.
//   cr_kernel<M,G>       G lanes per draw, all draws of a warp in ONE instruction stream (full-mask shuffles, finished
//                        draws frozen by predicated commits).  Between iterations the state of a draw (A1, A0, A2,
//                        accumulated correction) lives in a shared tile; inside an iteration A1 and [X0 | X2] are
//                        distributed by COLUMN over the lanes in registers with compile-time indices only: the
//                        m x m solve is a Householder QR (no pivoting => no dynamic register indexing) followed by
//                        a backward elimination, pivot columns travel by warp shuffles, A0 / A2 columns are
//                        broadcast from the tile.  Then G = -A1hat^-1 A0 and g_u = -(A2 G + A1)^-1 F_u with the
//                        same solver.
//   assemble_kernel<M>   one thread per draw: static rows by sparse back-substitution, Lyapunov doubling,
//                        state-space assembly into the hand-off buffer of the filter kernel.
//
// Reference stages replaced: LRE.first_order_solver! (src/perturbations.jl:663-664), compute_variance!
// (src/estimation/estimation.jl:630), state-space assembly (estimation.jl:640-658).
#pragma once
#include <type_traits>
#include "model_traits.cuh"
#include "solve_kernel.cuh"

namespace dyb {

template <class M>
struct MidLayout {
  static constexpr int OFF_RED = 0;
  static constexpr int OFF_TOP = OFF_RED + M::RED_NNZ;
  static constexpr int OFF_YB = OFF_TOP + M::TOP_NNZ;                 // NY
  static constexpr int OFF_SE = OFF_YB + M::NY;                       // NX*NX
  static constexpr int OFF_H = OFF_SE + M::NX * M::NX;                // packed lower NY(NY+1)/2
  static constexpr int OFF_G = OFF_H + M::NY * (M::NY + 1) / 2;       // MM x K  (written by cr_kernel)
  static constexpr int OFF_GU = OFF_G + M::M * M::K;                  // MM x NX
  static constexpr int SIZE = OFF_GU + M::M * M::NX;
};

struct StridedOut {
  double* p; long long stride;
  DYB_D double& operator[](int e) const { return p[(long long)e * stride]; }
};

// ------------------------------------------------------------------------------------------------
// launch geometry of the two one-thread-per-draw kernels: at 65 536 draws on 148 SMs a kernel needs >= 443 resident threads per
// SM to finish in ONE wave (a second, nearly empty wave costs a full thread latency); DYB_*_MINB bounds the registers for that
#ifndef DYB_MODEL_BLOCK
#define DYB_MODEL_BLOCK 128
#endif
#ifndef DYB_MODEL_MINB
#define DYB_MODEL_MINB 4
#endif
#ifndef DYB_ASM_BLOCK
#define DYB_ASM_BLOCK 128
#endif
#ifndef DYB_ASM_MINB
#define DYB_ASM_MINB 1
#endif
template <class M>
__global__ void __launch_bounds__(DYB_MODEL_BLOCK, DYB_MODEL_MINB)
model_kernel(const SolveParams<M> prm, const double* __restrict__ theta, long long n_draws,
             double* __restrict__ mid, int* __restrict__ status, double* __restrict__ ybar_out) {
  constexpr int N = M::N, NX = M::NX, NY = M::NY;
  using L = MidLayout<M>;
  const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_draws) return;
  const double* th = theta + d * (long long)prm.np;
  double p[M::NPAR > 0 ? M::NPAR : 1];
  double Se[NX * NX];
  double Sm[NY * NY];
#pragma unroll
  for (int i = 0; i < M::NPAR; ++i) p[i] = prm.params0[i];
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) Se[i] = prm.sigma_e0[i];
#pragma unroll
  for (int i = 0; i < NY * NY; ++i) Sm[i] = prm.sigma_m0[i];
  for (int q = 0; q < prm.np; ++q) {
    const double v = th[q];
    const int ty = prm.est_type[q], i = prm.est_i[q], j = prm.est_j[q];
    // compile-time indexed stores (keeps p / Se / Sm in registers)
    if (ty == EST_PARAM) {
#pragma unroll
      for (int a = 0; a < M::NPAR; ++a) if (a == i) p[a] = v;
    } else if (ty == EST_SD_SHOCK || ty == EST_VAR_SHOCK) {
      const double x = (ty == EST_SD_SHOCK) ? v * v : v;
#pragma unroll
      for (int a = 0; a < NX * NX; ++a) if (a == i + NX * j) Se[a] = x;
    } else if (ty == EST_CORR_SHOCK) {
      double vi = 0.0, vj = 0.0;
#pragma unroll
      for (int a = 0; a < NX; ++a) { if (a == i) vi = Se[a + NX * a]; if (a == j) vj = Se[a + NX * a]; }
      const double c = v * sqrt(vi * vj);
#pragma unroll
      for (int a = 0; a < NX * NX; ++a) if (a == i + NX * j || a == j + NX * i) Se[a] = c;
    } else if (ty == EST_SD_MEAS || ty == EST_VAR_MEAS) {
      const double x = (ty == EST_SD_MEAS) ? v * v : v;
#pragma unroll
      for (int a = 0; a < NY * NY; ++a) if (a == i + NY * j) Sm[a] = x;
    } else if (ty == EST_CORR_MEAS) {
      double vi = 0.0, vj = 0.0;
#pragma unroll
      for (int a = 0; a < NY; ++a) { if (a == i) vi = Sm[a + NY * a]; if (a == j) vj = Sm[a + NY * a]; }
      const double c = v * sqrt(vi * vj);
#pragma unroll
      for (int a = 0; a < NY * NY; ++a) if (a == i + NY * j || a == j + NY * i) Sm[a] = c;
    }
  }
  double y[N];
  M::steady_state(p, y);
  bool ok = true;
#pragma unroll
  for (int i = 0; i < N; ++i) ok = ok && isfinite(y[i]);
  if (ybar_out) {
#pragma unroll
    for (int i = 0; i < N; ++i) ybar_out[d * N + i] = y[i];
  }
  if (ok) {
    const double r2 = M::static_resid_ss(p, y);
    ok = (r2 <= prm.steady_tol * prm.steady_tol);
  }
  int st = ok ? ST_OK : ST_STEADY;
  StridedOut out{mid + d, n_draws};
  if (ok) {
    double buf[M::RED_NNZ + M::TOP_NNZ];
    const bool nonsing = M::static_elim(p, y, buf);
    double chk = 0.0;
#pragma unroll
    for (int e = 0; e < M::RED_NNZ + M::TOP_NNZ; ++e) { chk = fma(buf[e], 0.0, chk); out[L::OFF_RED + e] = buf[e]; }
    if (!(chk == 0.0)) st = ST_STEADY;            // NaN / Inf in the Jacobian
    else if (!nonsing) st = ST_SOLVER;
  }
#pragma unroll
  for (int i = 0; i < NY; ++i) out[L::OFF_YB + i] = y[M::state_ids(M::obs_idx_state(i))];
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) out[L::OFF_SE + i] = Se[i];
#pragma unroll
  for (int i = 0; i < NY; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) out[L::OFF_H + tri(i, j)] = Sm[i + NY * j];
  status[d] = st;
}

// ------------------------------------------------------------------------------------------------
// model_kernel's counterpart for "sizes-only" models (M::JAC_FED): the host has evaluated the steady state and the compressed
// dynamic Jacobian of every draw (as the reference does per draw, src/perturbations.jl:526-561, 616-635); this kernel does
// the static elimination -- a dense Householder QR on the S static columns, applied to every column of the working matrix
// [static | lags | dynamic current | leads | shocks] -- and writes the same hand-off (`mid`) as model_kernel: the reduced rows
// (dense: RED_NNZ = M x (NCOL - S)), the S pivot rows (TOP_NNZ = S x NCOL), ybar of the observables, Sigma_e, H.
// One thread per draw, the working matrix in local memory (run-time loops: the sizes are compile-time constants but the
// code stays small).  jac: N x NCOL per draw, column-major, columns [lags of i_bkwrd_b | N current | leads of i_fwrd_b | NX].
template <class M>
__global__ void __launch_bounds__(128)
jac_model_kernel(const double* __restrict__ jac, const double* __restrict__ ybar, const double* __restrict__ sigma_e,
                 int se_per_draw, const double* __restrict__ sigma_m, int sm_per_draw, const int* __restrict__ status_in,
                 long long n_draws, double* __restrict__ mid, int* __restrict__ status) {
  constexpr int N = M::N, NX = M::NX, NY = M::NY, NCOL = M::NCOL, S = M::S, W = NCOL - S;
  using L = MidLayout<M>;
  static_assert(M::RED_NNZ == M::M * W && M::TOP_NNZ == S * NCOL, "dense hand-off pattern");
  const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_draws) return;
  int st = status_in ? status_in[d] : ST_OK;
  StridedOut out{mid + d, n_draws};
  double Wm[N * NCOL];
  const double* J = jac + d * (long long)(N * NCOL);
  double chk = 0.0;
#pragma unroll 1
  for (int c = 0; c < NCOL; ++c) {
    const int wc = M::jperm(c);
#pragma unroll 1
    for (int r = 0; r < N; ++r) { const double v = J[r + N * c]; Wm[r + N * wc] = v; chk = fma(v, 0.0, chk); }
  }
  if (st == ST_OK && !(chk == 0.0)) st = ST_STEADY;          // NaN / Inf in the Jacobian
  bool nonsing = true;
#pragma unroll 1
  for (int q = 0; q < S; ++q) {
    double s2 = 0.0;
#pragma unroll 1
    for (int r = q + 1; r < N; ++r) s2 = fma(Wm[r + N * q], Wm[r + N * q], s2);
    const double xq = Wm[q + N * q];
    const double nrm = sqrt(fma(xq, xq, s2));
    if (!(nrm > 0.0)) { nonsing = false; break; }
    const double alpha = xq >= 0.0 ? -nrm : nrm;
    const double vq = xq - alpha;
    const double beta = 1.0 / (-alpha * vq);                 // 2 / (v'v)
    Wm[q + N * q] = alpha;
#pragma unroll 1
    for (int c = q + 1; c < NCOL; ++c) {
      double dt = vq * Wm[q + N * c];
#pragma unroll 1
      for (int r = q + 1; r < N; ++r) dt = fma(Wm[r + N * q], Wm[r + N * c], dt);
      dt *= beta;
      Wm[q + N * c] = fma(-dt, vq, Wm[q + N * c]);
#pragma unroll 1
      for (int r = q + 1; r < N; ++r) Wm[r + N * c] = fma(-dt, Wm[r + N * q], Wm[r + N * c]);
    }
#pragma unroll 1
    for (int r = q + 1; r < N; ++r) Wm[r + N * q] = 0.0;
  }
  if (st == ST_OK && !nonsing) st = ST_SOLVER;
  if (st == ST_OK) {
#pragma unroll 1
    for (int r = 0; r < M::M; ++r)
#pragma unroll 1
      for (int c = 0; c < W; ++c) out[L::OFF_RED + r * W + c] = Wm[(S + r) + N * (S + c)];
#pragma unroll 1
    for (int q = 0; q < S; ++q)
#pragma unroll 1
      for (int c = 0; c < NCOL; ++c) out[L::OFF_TOP + q * NCOL + c] = Wm[q + N * c];
  }
  const double* yb = ybar ? ybar + d * N : nullptr;
#pragma unroll 1
  for (int i = 0; i < NY; ++i) out[L::OFF_YB + i] = yb ? yb[M::state_ids(M::obs_idx_state(i))] : 0.0;
  const double* Se = sigma_e + (se_per_draw ? d * (long long)(NX * NX) : 0);
#pragma unroll 1
  for (int i = 0; i < NX * NX; ++i) out[L::OFF_SE + i] = Se[i];
  const double* Sm = sigma_m ? sigma_m + (sm_per_draw ? d * (long long)(NY * NY) : 0) : nullptr;
#pragma unroll 1
  for (int i = 0; i < NY; ++i)
#pragma unroll 1
    for (int j = 0; j <= i; ++j) out[L::OFF_H + tri(i, j)] = Sm ? Sm[i + NY * j] : 0.0;
  status[d] = st;
}

// detection of sizes-only models
template <class T, class = void> struct is_jac_fed : std::false_type {};
template <class T> struct is_jac_fed<T, std::void_t<decltype(T::JAC_FED)>> : std::integral_constant<bool, T::JAC_FED> {};

// ------------------------------------------------------------------------------------------------
// cyclic reduction, G lanes per draw
// ------------------------------------------------------------------------------------------------
template <int G>
DYB_D double gshfl(unsigned mask, double v, int src) { return __shfl_sync(mask, v, src, G); }

template <int G>
DYB_D double gmax(unsigned mask, double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o, G));
  return v;
}

// Householder QR of the MM x MM matrix distributed in A (column c in lane c % G, slot c / G) applied on
// the fly to NR slots of right-hand-side columns B, followed by backward elimination:  B <- A^-1 B.
// Every register index is a compile-time constant.  Returns false on an exactly singular / non-finite column.
template <int MM, int G, int NR>
DYB_D bool qr_solve(double (&A)[(MM + G - 1) / G][MM], double (&B)[NR][MM], int lg, unsigned mask) {
  constexpr int SL = (MM + G - 1) / G;
  double rinv[MM];
  bool ok = true;
#pragma unroll
  for (int p = 0; p < MM; ++p) {
    double x[MM];
#pragma unroll
    for (int r = p; r < MM; ++r) x[r] = gshfl<G>(mask, A[p / G][r], p % G);
    double sg = 0.0;
#pragma unroll
    for (int r = p + 1; r < MM; ++r) sg = fma(x[r], x[r], sg);
    const double s = fma(x[p], x[p], sg);
    ok = ok && (s > 0.0) && (s < 1e300);
    const double rn = fast_rsqrt(s);
    const double nrm = s * rn;
    const double alpha = x[p] >= 0.0 ? -nrm : nrm;
    const double vp = x[p] - alpha;
    const double beta = fast_rcp(-alpha * vp);
    rinv[p] = x[p] >= 0.0 ? -rn : rn;
#pragma unroll
    for (int sl = p / G; sl < SL; ++sl) {
      const int col = sl * G + lg;
      if (col > p) {
        double dt = vp * A[sl][p];
#pragma unroll
        for (int r = p + 1; r < MM; ++r) dt = fma(x[r], A[sl][r], dt);
        dt *= beta;
        A[sl][p] = fma(-dt, vp, A[sl][p]);
#pragma unroll
        for (int r = p + 1; r < MM; ++r) A[sl][r] = fma(-dt, x[r], A[sl][r]);
      }
    }
#pragma unroll
    for (int sl = 0; sl < NR; ++sl) {
      double dt = vp * B[sl][p];
#pragma unroll
      for (int r = p + 1; r < MM; ++r) dt = fma(x[r], B[sl][r], dt);
      dt *= beta;
      B[sl][p] = fma(-dt, vp, B[sl][p]);
#pragma unroll
      for (int r = p + 1; r < MM; ++r) B[sl][r] = fma(-dt, x[r], B[sl][r]);
    }
  }
#pragma unroll
  for (int p = MM - 1; p >= 0; --p) {
    double rc[MM];
#pragma unroll
    for (int r = 0; r < p; ++r) rc[r] = gshfl<G>(mask, A[p / G][r], p % G);
#pragma unroll
    for (int sl = 0; sl < NR; ++sl) {
      const double yp = B[sl][p] * rinv[p];
      B[sl][p] = yp;
#pragma unroll
      for (int r = 0; r < p; ++r) B[sl][r] = fma(-rc[r], yp, B[sl][r]);
    }
  }
  return ok;
}

// Gauss-Jordan elimination with COLUMN pivoting on the same distributed layout:  B <- A^-1 B up to a row permutation.
// Step p works on ROW p: the pivot is the largest |A[p][c]| over the columns c not used yet (a 32-bit key -- upper word of
// |a| with the column number in its four low bits -- maximised over the G lanes by two shuffles), the pivot column is
// broadcast from its owner and every lane eliminates it from all other rows of its own columns.  Rows stay where they are
// (every register index is a compile-time constant) and only the owner of the pivot column depends on the data, so what
// comes out is row p of B = row pc[p] of the solution; the caller undoes the permutation through shared memory.
// (MM - 1) FMAs per column and step against 2 (MM - p) + back-substitution for the Householder variant: 0.7 x the
// arithmetic at MM = 9 with 3 + 3 slots, and no serial backward phase.  Returns false on a zero / non-finite pivot.
template <int MM, int G, int NR>
DYB_D bool gj_solve(double (&A)[(MM + G - 1) / G][MM], double (&B)[NR][MM], int lg, unsigned mask, int (&pc)[MM]) {
  constexpr int SL = (MM + G - 1) / G;
  static_assert(MM <= 16, "column number must fit the four key bits");
  unsigned used = 0u;                               // bit sl: own column sl * G + lg has served as a pivot
  bool ok = true;
#pragma unroll
  for (int p = 0; p < MM; ++p) {
    unsigned key = 0u;
    int msl = 0;
#pragma unroll
    for (int sl = 0; sl < SL; ++sl) {
      const int col = sl * G + lg;
      const unsigned k = (((unsigned)__double2hiint(A[sl][p]) & 0x7ffffff0u) + 16u) | (unsigned)col;
      const bool cand = (col < MM) && !((used >> sl) & 1u);
      if (cand && k > key) { key = k; msl = sl; }
    }
    double v[MM];
#pragma unroll
    for (int r = 0; r < MM; ++r) {
      double t = A[0][r];
#pragma unroll
      for (int sl = 1; sl < SL; ++sl) t = (msl == sl) ? A[sl][r] : t;
      v[r] = t;
    }
    unsigned kb = key;
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) kb = max(kb, __shfl_xor_sync(mask, kb, o, G));
    const int pcol = (int)(kb & 15u);
    const int pl = pcol % G;
    pc[p] = pcol;
    if (lg == pl) used |= 1u << msl;
    double x[MM];
#pragma unroll
    for (int r = 0; r < MM; ++r) x[r] = gshfl<G>(mask, v[r], pl);
    const double ap = fabs(x[p]);
    ok = ok && (ap > 0.0) && (ap < 1e150);
    const double ip = fast_rcp(x[p]);
#pragma unroll
    for (int sl = 0; sl < SL; ++sl) A[sl][p] *= ip;
#pragma unroll
    for (int sl = 0; sl < NR; ++sl) B[sl][p] *= ip;
#pragma unroll
    for (int r = 0; r < MM; ++r) {
      if (r == p) continue;
#pragma unroll
      for (int sl = 0; sl < SL; ++sl) A[sl][r] = fma(-x[r], A[sl][p], A[sl][r]);
#pragma unroll
      for (int sl = 0; sl < NR; ++sl) B[sl][r] = fma(-x[r], B[sl][p], B[sl][r]);
    }
  }
  return ok;
}

// Gaussian elimination WITHOUT pivoting on the same layout: half the multiply-adds of the Householder variant, the same
// shuffles, every index static (1 368 instructions per cyclic-reduction iteration at MM = 9 against 1 894).  The rows have
// been put in the pivot order of the calibration point by the caller; nothing here bounds the element growth, so the
// caller verifies the residual of what it finally returns (cr_kernel, CR_VERIFY_TOL).
template <int MM, int G, int NR>
DYB_D bool lu_solve_static(double (&A)[(MM + G - 1) / G][MM], double (&B)[NR][MM], int lg, unsigned mask) {
  constexpr int SL = (MM + G - 1) / G;
  double ipv[MM];
  bool ok = true;
#pragma unroll
  for (int p = 0; p < MM; ++p) {
    double x[MM];
#pragma unroll
    for (int r = p; r < MM; ++r) x[r] = gshfl<G>(mask, A[p / G][r], p % G);
    const double ap = fabs(x[p]);
    ok = ok && (ap > 0.0) && (ap < 1e150);
    const double ip = fast_rcp(x[p]);
    ipv[p] = ip;
    double l[MM];
#pragma unroll
    for (int r = p + 1; r < MM; ++r) l[r] = x[r] * ip;
#pragma unroll
    for (int sl = p / G; sl < SL; ++sl) {
#pragma unroll
      for (int r = p + 1; r < MM; ++r) A[sl][r] = fma(-l[r], A[sl][p], A[sl][r]);
    }
#pragma unroll
    for (int sl = 0; sl < NR; ++sl) {
#pragma unroll
      for (int r = p + 1; r < MM; ++r) B[sl][r] = fma(-l[r], B[sl][p], B[sl][r]);
    }
  }
#pragma unroll
  for (int p = MM - 1; p >= 0; --p) {
    double rc[MM];
#pragma unroll
    for (int r = 0; r < p; ++r) rc[r] = gshfl<G>(mask, A[p / G][r], p % G);
#pragma unroll
    for (int sl = 0; sl < NR; ++sl) {
      const double yp = B[sl][p] * ipv[p];
      B[sl][p] = yp;
#pragma unroll
      for (int r = 0; r < p; ++r) B[sl][r] = fma(-rc[r], yp, B[sl][r]);
    }
  }
  return ok;
}

// m x m solver of cr_kernel: 0 = Householder QR (default), 1 = Gauss-Jordan with column pivoting.  Measured on B200
// (profiles/r02_d_cr_solver_ab.txt, 65 536 draws): fs2000 0.410 ms (QR) against 0.485 ms (GJ), ls2003 0.324 / 0.360,
// hs_nk 0.159 / 0.155 -- the pivoted variant executes 0.71 x the fp64 instructions (881 against 1 243 per iteration at
// MM = 9) but 331 FSEL + LDS / STS for the dynamic slot and the un-permutation make its loop 2 165 instructions against
// 1 894, and the kernel's time follows the instruction count, not the fp64 count.
#ifndef DYB_CR_SOLVER
#define DYB_CR_SOLVER 0
#endif

#ifndef DYB_CR_BATCH_RMW
#define DYB_CR_BATCH_RMW 1
#endif
#ifndef DYB_CR_TREE_NORM
#define DYB_CR_TREE_NORM 1
#endif
struct CrOptions { double cr_tol; int cr_maxit; };

// sum_r |v[r]| by pairs (depth log2 MM instead of a chain of MM dependent additions; only the stopping norms use it)
template <int MM>
DYB_D double abs_sum_tree(const double (&v)[MM]) {
#if !DYB_CR_TREE_NORM
  double cs = 0.0;
#pragma unroll
  for (int i = 0; i < MM; ++i) cs += fabs(v[i]);
  return cs;
#endif
  double t[MM];
#pragma unroll
  for (int i = 0; i < MM; ++i) t[i] = fabs(v[i]);
#pragma unroll
  for (int n = MM; n > 1; n = (n + 1) / 2) {
#pragma unroll
    for (int i = 0; i < n / 2; ++i) t[i] = t[2 * i] + t[2 * i + 1];
    if (n & 1) t[n / 2] = t[n - 1];
  }
  return t[0];
}

#ifndef DYB_CR_LANES
#define DYB_CR_LANES 4
#endif
// (6 CTAs per SM = 168 registers with 128 bytes of spills in the fast pass; 5 CTAs = 200 registers, no spills, measured SLOWER:
// cr stage 0.313 -> 0.328 ms, headline 107.7 -> 105.7 M evals/s -- unlike the filter, this kernel gains more from the twelfth warp)
#ifndef DYB_CR_MIN_CTAS
#define DYB_CR_MIN_CTAS (DYB_CR_SOLVER == 1 ? 5 : 6)
#endif

// Geometry of cr_kernel: G lanes per draw, CTAs of BLOCK threads, one shared-memory tile per draw holding
// A1 (MM x MM) | acc_hat (MM x K) | A0 (MM x K) | A2 (MM x NF) | X (MM x (K + NF), the un-permuting buffer of the
// Gauss-Jordan solver), column-major.  The tile stride is padded to
// 4 (mod 16) doubles so that the "own column" accesses of a half-warp (4 draws x 4 lanes, column stride MM)
// fall into 16 distinct 8-byte banks when MM is odd.
template <class M>
struct CrCfg {
  static constexpr int G = DYB_CR_LANES, BLOCK = 64, GROUPS = BLOCK / G;
  static constexpr int MM = M::M, K = M::K, NF = M::NF;
  static constexpr int OFF_A1 = 0, OFF_ACC = MM * MM, OFF_A0 = OFF_ACC + MM * K, OFF_A2 = OFF_A0 + MM * K;
  static constexpr int OFF_X = OFF_A2 + MM * NF;
  static constexpr int TILE_RAW = OFF_X + (DYB_CR_SOLVER == 1 ? MM * (K + NF) : 0);
  static constexpr int PAD_TO = (G == 4) ? 4 : 8;       // tile stride mod 16 that spreads a half-warp over 16 banks
  static constexpr int TILE = TILE_RAW + ((PAD_TO - TILE_RAW % 16) + 16) % 16;
  static constexpr int MIN_CTAS = DYB_CR_MIN_CTAS;
  static constexpr int MAX_REGS = (65536 / (BLOCK * MIN_CTAS)) / 8 * 8;
  static constexpr bool FITS = (MM <= 12) && (K >= 1) && (NF >= 1) && (NF + K >= M::NX) &&
                               (sizeof(double) * GROUPS * TILE <= 48 * 1024);
  static constexpr bool FITS_LU = FITS && (M::NX <= 2 * K);          // F_u staged over the acc_hat | A0 regions of the tile
};

// B (own columns, rows in pivot order pc) -> natural row order through the draw's X buffer.  Slot s of the lane is
// column xcol[s] of the buffer (negative: a padding slot, nothing stored, whatever read); same thread writes and reads.
template <int MM, int NR>
DYB_D void unpermute_rows(double (&B)[NR][MM], const int (&pc)[MM], double* Xs, const int (&xcol)[NR]) {
#pragma unroll
  for (int s = 0; s < NR; ++s) {
    if (xcol[s] >= 0) {
#pragma unroll
      for (int p = 0; p < MM; ++p) Xs[pc[p] + MM * xcol[s]] = B[s][p];
    }
  }
#pragma unroll
  for (int s = 0; s < NR; ++s) {
    const double* xs = Xs + MM * (xcol[s] >= 0 ? xcol[s] : 0);
#pragma unroll
    for (int r = 0; r < MM; ++r) B[s][r] = xs[r];
  }
}

// All 32 lanes of a warp (8 draws) run the SAME instruction stream: shuffles use the full mask (a partial mask costs
// a WARPSYNC / BSSY / BSYNC sequence per shuffle), the iteration runs until every draw of the warp has stopped, and a
// draw that has stopped is frozen by predicating its commits (short divergent sections without shuffles), so its
// result does not depend on its neighbours.  Between iterations the whole state of a draw lives in its shared tile;
// registers only hold the transient copies of one iteration (A1 and [X0 | X2] by column, then the GEMM accumulators).
// SOLVER: the m x m solver (CR_QR Householder, CR_GJ Gauss-Jordan with column pivoting, CR_LU Gaussian elimination without
// pivoting on rows pre-ordered by `row_perm`).  The CR_LU pass is the fast one: it does not pivot, so it VERIFIES what it
// returns -- the residuals of A0 + (A1 + A2 G[fw,:] E_bk') G = 0 and of (A1 + A2 G[fw,:] E_bk') g_u = -F_u against their own
// scale -- and hands every draw it cannot vouch for (residual above CR_VERIFY_TOL, zero pivot, no convergence) to a second
// launch with REDO = true and SOLVER = CR_QR, which only touches draws marked ST_REDO (warps without one return at once).
constexpr int CR_QR = 0, CR_GJ = 1, CR_LU = 2;
constexpr double CR_VERIFY_TOL = 1e-12;

template <class M, int G, int SOLVER, bool REDO>
__global__ void __launch_bounds__(CrCfg<M>::BLOCK) __maxnreg__(CrCfg<M>::MAX_REGS)
cr_kernel(double* __restrict__ mid, long long n_draws, CrOptions opt,
          int* __restrict__ status, int* __restrict__ cr_iters, const int* __restrict__ row_perm,
          unsigned long long* __restrict__ redo_count, long long* __restrict__ redo_list) {
  using C = CrCfg<M>;
  static_assert(G == C::G, "lane count");
  static_assert(SOLVER != CR_GJ || DYB_CR_SOLVER == 1, "the Gauss-Jordan variant needs the X buffer of a -DDYB_CR_SOLVER=1 build");
  constexpr int MM = M::M, K = M::K, NF = M::NF, NX = M::NX;
  constexpr int SL1 = (MM + G - 1) / G, SL0 = (K + G - 1) / G, SL2 = (NF + G - 1) / G, SLU = (NX + G - 1) / G;
  constexpr int NR = SL0 + SL2;
  constexpr unsigned FULL = 0xffffffffu;
  using L = MidLayout<M>;
  __shared__ double tile[C::GROUPS * C::TILE];

  const int lg = threadIdx.x % G;
  long long d = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  bool real = d < n_draws;
  if constexpr (REDO) {
    // second pass: group g takes entry g of the list the fast pass packed (redo_count[1] entries); warps beyond it leave at
    // once (the whole warp: there is no block-level barrier below)
    const long long nred = (long long)redo_count[1];
    if ((d / (32 / G)) * (32 / G) >= nred) return;
    real = d < nred;
    d = redo_list[real ? d : nred - 1];
  }
  if (!real) d = n_draws - 1;                       // tail groups recompute the last draw and store nothing
  double* A1s = tile + (threadIdx.x / G) * C::TILE + C::OFF_A1;
  double* accs = A1s + (C::OFF_ACC - C::OFF_A1);
  double* A0s = A1s + (C::OFF_A0 - C::OFF_A1);
  double* A2s = A1s + (C::OFF_A2 - C::OFF_A1);
  [[maybe_unused]] double* Xs = A1s + (C::OFF_X - C::OFF_A1);
  [[maybe_unused]] int pc[MM];
  const double* md = mid + d;
  const long long st = n_draws;
  int st_in = status[d];
  constexpr bool skip = false;
  if constexpr (SOLVER == CR_LU) {
    // no usable pivot order (the calibration point itself failed): nothing to gain from eliminating in the natural order,
    // every draw goes straight to the Householder pass
    if (row_perm[MM] == 0) {
      if (real && lg == 0 && st_in == ST_OK) { status[d] = ST_REDO; redo_list[atomicAdd(redo_count + 1, 1ULL)] = d; }
      return;
    }
  }
  if constexpr (REDO) {
    st_in = ST_OK;                                    // listed draws entered the fast pass with ST_OK
    if (real && lg == 0) atomicAdd(redo_count, 1ULL);
  }
  // tile row of equation r: the pivot order found once at the calibration (cr_perm_kernel), identity otherwise
  int prow[MM];
#pragma unroll
  for (int r = 0; r < MM; ++r) prow[r] = (SOLVER == CR_LU && row_perm) ? row_perm[r] : r;

  // (re)load A1(0), A0(0), A2(0) of the draw into the tile; acc_hat is left alone unless zero_acc
  auto load_tile = [&](bool zero_acc) {
    for (int e = lg; e < C::OFF_ACC; e += G) A1s[e] = 0.0;
    for (int e = lg; e < MM * K + MM * NF; e += G) A0s[e] = 0.0;
    if (zero_acc) for (int e = lg; e < MM * K; e += G) accs[e] = 0.0;
    __syncwarp(FULL);
#pragma unroll
    for (int e = 0; e < M::RED_NNZ; ++e) {
      if (e % G != lg) continue;
      const int rr = prow[M::red_row(e)], c = M::red_col(e);
      if (c < K) A0s[rr + MM * c] = md[(L::OFF_RED + e) * st];
      else if (c < K + MM) A1s[rr + MM * (c - K)] = md[(L::OFF_RED + e) * st];
      else if (c < K + MM + NF) A2s[rr + MM * (c - K - MM)] = md[(L::OFF_RED + e) * st];
    }
    __syncwarp(FULL);
  };
  load_tile(true);

  int it = 0;
  int result = st_in;                 // draws that already failed are frozen from the start
  bool active = (st_in == ST_OK) && !skip, conv = false, stalled = false;
  double n0 = 0.0, n2 = 0.0;
  for (int iter = 1; iter <= opt.cr_maxit; ++iter) {
    if (!__any_sync(FULL, active)) break;
    // own columns of A1 and of [A0 | A2]
    double A1w[SL1][MM], XB[NR][MM];
#pragma unroll
    for (int s = 0; s < SL1; ++s) {
      const int col = s * G + lg;
#pragma unroll
      for (int r = 0; r < MM; ++r) A1w[s][r] = (col < MM) ? A1s[r + MM * col] : 0.0;
    }
#pragma unroll
    for (int s = 0; s < SL0; ++s) {
      const int col = s * G + lg;
#pragma unroll
      for (int r = 0; r < MM; ++r) XB[s][r] = (col < K) ? A0s[r + MM * col] : 0.0;
    }
#pragma unroll
    for (int s = 0; s < SL2; ++s) {
      const int col = s * G + lg;
#pragma unroll
      for (int r = 0; r < MM; ++r) XB[SL0 + s][r] = (col < NF) ? A2s[r + MM * col] : 0.0;
    }
    // [X0 | X2] = A1^-1 [A0 | A2]
    bool okq;
    if constexpr (SOLVER == CR_LU) {
      okq = lu_solve_static<MM, G, NR>(A1w, XB, lg, FULL);
    } else if constexpr (SOLVER == CR_GJ) {
#if DYB_CR_SOLVER == 1
      okq = gj_solve<MM, G, NR>(A1w, XB, lg, FULL, pc);
      int xcol[NR];
#pragma unroll
      for (int s = 0; s < SL0; ++s) xcol[s] = (s * G + lg < K) ? s * G + lg : -1;
#pragma unroll
      for (int s = 0; s < SL2; ++s) xcol[SL0 + s] = (s * G + lg < NF) ? K + s * G + lg : -1;
      unpermute_rows<MM, NR>(XB, pc, Xs, xcol);
#else
      okq = false;
#endif
    } else {
      okq = qr_solve<MM, G, NR>(A1w, XB, lg, FULL);
    }
    const bool commit = active && okq;

    // pass 1: own columns j of [W00 | W01] = A0 * X[bk, j], columns of A0 broadcast from the tile
    double acc[NR][MM];
#pragma unroll
    for (int s = 0; s < NR; ++s)
#pragma unroll
      for (int r = 0; r < MM; ++r) acc[s][r] = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      double a[MM];
#pragma unroll
      for (int r = 0; r < MM; ++r) a[r] = A0s[r + MM * i];
#pragma unroll
      for (int s = 0; s < NR; ++s) {
        const double xv = XB[s][M::bkwrd_in_dyn(i)];
#pragma unroll
        for (int r = 0; r < MM; ++r) acc[s][r] = fma(a[r], xv, acc[s][r]);
      }
    }
    double m0 = 0.0;
#pragma unroll
    for (int s = 0; s < SL0; ++s) m0 = fmax(m0, abs_sum_tree<MM>(acc[s]));
    __syncwarp(FULL);                               // every lane has read A0 before anybody overwrites it
    if (commit) {
      // A0 <- -W00 (own columns) ; A1[:, fw(j)] -= W01[:, j]
#pragma unroll
      for (int s = 0; s < SL0; ++s) {
        const int col = s * G + lg;
        if (col < K) {
#pragma unroll
          for (int r = 0; r < MM; ++r) A0s[r + MM * col] = -acc[s][r];
        }
      }
      // read-modify-write of the tile in two sweeps (all loads, then all stores): the regions alias as far as the compiler
      // knows, and load -> add -> store chains per element stalled on the shared-memory latency
#if DYB_CR_BATCH_RMW
      double old1[SL2][MM];
#pragma unroll
      for (int s = 0; s < SL2; ++s) {
        const int col = s * G + lg;
        const int tc = M::fwrd_in_dyn(col < NF ? col : 0);
#pragma unroll
        for (int r = 0; r < MM; ++r) old1[s][r] = A1s[r + MM * tc];
      }
#pragma unroll
      for (int s = 0; s < SL2; ++s) {
        const int col = s * G + lg;
        if (col < NF) {
          const int tc = M::fwrd_in_dyn(col);
#pragma unroll
          for (int r = 0; r < MM; ++r) A1s[r + MM * tc] = old1[s][r] - acc[SL0 + s][r];
        }
      }
#else
#pragma unroll
      for (int s = 0; s < SL2; ++s) {
        const int col = s * G + lg;
        if (col < NF) {
          const int tc = M::fwrd_in_dyn(col);
#pragma unroll
          for (int r = 0; r < MM; ++r) A1s[r + MM * tc] -= acc[SL0 + s][r];
        }
      }
#endif
    }
    __syncwarp(FULL);
    // pass 2: own columns j of [W10 | W11] = A2 * X[fw, j]
#pragma unroll
    for (int s = 0; s < NR; ++s)
#pragma unroll
      for (int r = 0; r < MM; ++r) acc[s][r] = 0.0;
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      double a[MM];
#pragma unroll
      for (int r = 0; r < MM; ++r) a[r] = A2s[r + MM * i];
#pragma unroll
      for (int s = 0; s < NR; ++s) {
        const double xv = XB[s][M::fwrd_in_dyn(i)];
#pragma unroll
        for (int r = 0; r < MM; ++r) acc[s][r] = fma(a[r], xv, acc[s][r]);
      }
    }
    double m2 = 0.0;
#pragma unroll
    for (int s = 0; s < SL2; ++s) m2 = fmax(m2, abs_sum_tree<MM>(acc[SL0 + s]));
    __syncwarp(FULL);
    if (commit) {
#pragma unroll
      for (int s = 0; s < SL2; ++s) {
        const int col = s * G + lg;
        if (col < NF) {
#pragma unroll
          for (int r = 0; r < MM; ++r) A2s[r + MM * col] = -acc[SL0 + s][r];
        }
      }
#if DYB_CR_BATCH_RMW
      double old1[SL0][MM], olda[SL0][MM];
#pragma unroll
      for (int s = 0; s < SL0; ++s) {
        const int col = (s * G + lg < K) ? s * G + lg : 0;
        const int tc = M::bkwrd_in_dyn(col);
#pragma unroll
        for (int r = 0; r < MM; ++r) { old1[s][r] = A1s[r + MM * tc]; olda[s][r] = accs[r + MM * col]; }
      }
#pragma unroll
      for (int s = 0; s < SL0; ++s) {
        const int col = s * G + lg;
        if (col < K) {
          const int tc = M::bkwrd_in_dyn(col);
#pragma unroll
          for (int r = 0; r < MM; ++r) { A1s[r + MM * tc] = old1[s][r] - acc[s][r]; accs[r + MM * col] = olda[s][r] + acc[s][r]; }
        }
      }
#else
#pragma unroll
      for (int s = 0; s < SL0; ++s) {
        const int col = s * G + lg;
        if (col < K) {
          const int tc = M::bkwrd_in_dyn(col);
#pragma unroll
          for (int r = 0; r < MM; ++r) { A1s[r + MM * tc] -= acc[s][r]; accs[r + MM * col] += acc[s][r]; }
        }
      }
#endif
    }
    __syncwarp(FULL);
    m0 = gmax<G>(FULL, m0);
    m2 = gmax<G>(FULL, m2);
    if (active) {
      it = iter;
      if (!okq || !(m0 <= 1e150) || !(m2 <= 1e150)) {
        if (iter == 1) result = ST_SOLVER; else stalled = true;
        active = false;
      } else {
        const double n0_prev = n0;
        n0 = m0; n2 = m2;
        // converged: both sequences vanished, or A2(j) vanished while A0(j) has settled on a non-zero limit -- a unit
        // root among the stable roots (the reference's generalised Schur solver counts |lambda| < 1 + 1e-6 as stable);
        // the correction of A1hat, and with it the error of G, is then O(n2): hence the squared tolerance on n2
        if ((n0 < opt.cr_tol && n2 < opt.cr_tol) || (n2 < opt.cr_tol * opt.cr_tol && fabs(n0 - n0_prev) <= CR_UNIT_ROOT_EPS * n0)) { conv = true; active = false; }
      }
    }
  }
  if (active) it = opt.cr_maxit + 1;               // ran out of iterations
  if (result == ST_OK && !conv) {
    // Blanchard-Kahn reading of a reduction that does not converge (see solve_kernel.cuh)
    if (stalled) result = (n0 < n2) ? ST_INDETERMINATE : ST_UNSTABLE;
    else result = (n0 >= opt.cr_tol) ? ST_UNSTABLE : ST_INDETERMINATE;
  }

  // ---- G (dynamic rows, backward columns) = -(A1(0) - acc_hat)^-1 A0(0); executed by every lane, stored if ok --------
  load_tile(false);
#pragma unroll
  for (int s = 0; s < SL0; ++s) {
    const int col = s * G + lg;
    if (col < K) {
      const int tc = M::bkwrd_in_dyn(col);
#pragma unroll
      for (int r = 0; r < MM; ++r) A1s[r + MM * tc] -= accs[r + MM * col];
    }
  }
  __syncwarp(FULL);
  double A1h[SL1][MM], Gc[SL0][MM];
#pragma unroll
  for (int s = 0; s < SL1; ++s) {
    const int col = s * G + lg;
#pragma unroll
    for (int r = 0; r < MM; ++r) A1h[s][r] = (col < MM) ? A1s[r + MM * col] : 0.0;
  }
#pragma unroll
  for (int s = 0; s < SL0; ++s) {
    const int col = s * G + lg;
#pragma unroll
    for (int r = 0; r < MM; ++r) Gc[s][r] = (col < K) ? -A0s[r + MM * col] : 0.0;
  }
  bool okq;
  if constexpr (SOLVER == CR_LU) {
    okq = lu_solve_static<MM, G, SL0>(A1h, Gc, lg, FULL);
  } else if constexpr (SOLVER == CR_GJ) {
#if DYB_CR_SOLVER == 1
    okq = gj_solve<MM, G, SL0>(A1h, Gc, lg, FULL, pc);
    int xcol[SL0];
#pragma unroll
    for (int s = 0; s < SL0; ++s) xcol[s] = (s * G + lg < K) ? s * G + lg : -1;
    unpermute_rows<MM, SL0>(Gc, pc, Xs, xcol);
#else
    okq = false;
#endif
  } else {
    okq = qr_solve<MM, G, SL0>(A1h, Gc, lg, FULL);
  }
  // Gc now holds G's own columns
  // ---- g_u = -(A1(0) + A2(0) G[fw, :])^-1 F_u ---------------------------------------------------------------------------
  double w[SL0][MM];
#pragma unroll
  for (int s = 0; s < SL0; ++s)
#pragma unroll
    for (int r = 0; r < MM; ++r) w[s][r] = 0.0;
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    double a[MM];
#pragma unroll
    for (int r = 0; r < MM; ++r) a[r] = A2s[r + MM * i];
#pragma unroll
    for (int s = 0; s < SL0; ++s) {
      const double xv = Gc[s][M::fwrd_in_dyn(i)];
#pragma unroll
      for (int r = 0; r < MM; ++r) w[s][r] = fma(a[r], xv, w[s][r]);
    }
  }
  __syncwarp(FULL);
  // A1 tile <- A1(0) + acc_hat on the backward columns (undo the subtraction above), then + w
#pragma unroll
  for (int s = 0; s < SL0; ++s) {
    const int col = s * G + lg;
    if (col < K) {
      const int tc = M::bkwrd_in_dyn(col);
#pragma unroll
      for (int r = 0; r < MM; ++r) A1s[r + MM * tc] = (A1s[r + MM * tc] + accs[r + MM * col]) + w[s][r];
    }
  }
  __syncwarp(FULL);
  // own columns c of Atilde = A1(0) + A2(0) G[fw,:] E_bk' and, per own right-hand-side column x with constant term f0, the
  // residual f0 + Atilde x against its scale |f0| + |Atilde| |x| (largest row of each): true when within CR_VERIFY_TOL
  [[maybe_unused]] auto verify = [&](const double* f0, const double* x) {
    double res[MM], sc[MM];
#pragma unroll
    for (int r = 0; r < MM; ++r) { res[r] = f0[r]; sc[r] = fabs(f0[r]); }
#pragma unroll
    for (int c = 0; c < MM; ++c) {
#pragma unroll
      for (int r = 0; r < MM; ++r) {
        const double a = A1s[r + MM * c];
        res[r] = fma(a, x[c], res[r]);
        sc[r] = fma(fabs(a), fabs(x[c]), sc[r]);
      }
    }
    // norm-wise: rows that are structurally zero carry rounding residue of their own size
    double rmax = 0.0, smax = 0.0;
    bool finite = true;
#pragma unroll
    for (int r = 0; r < MM; ++r) {
      rmax = fmax(rmax, fabs(res[r])); smax = fmax(smax, sc[r]);
      finite = finite && (sc[r] < 1e300);               // false for NaN / Inf (fmax would drop a NaN)
    }
    return finite && rmax <= CR_VERIFY_TOL * smax;
  };
  bool verified = true;
  if constexpr (SOLVER == CR_LU) {
#pragma unroll
    for (int s = 0; s < SL0; ++s) {
      const int col = s * G + lg;
      if (col < K) verified = verified && verify(A0s + MM * col, Gc[s]);
    }
  }
#pragma unroll
  for (int s = 0; s < SL1; ++s) {
    const int col = s * G + lg;
#pragma unroll
    for (int r = 0; r < MM; ++r) A1h[s][r] = (col < MM) ? A1s[r + MM * col] : 0.0;
  }
  double U[SLU][MM];
#pragma unroll
  for (int s = 0; s < SLU; ++s)
#pragma unroll
    for (int r = 0; r < MM; ++r) U[s][r] = 0.0;
  [[maybe_unused]] double* Fs = accs;                  // CR_LU: F_u in tile-row order, over acc_hat | A0 (both dead by now)
  if constexpr (SOLVER == CR_LU) {
    static_assert(NX <= 2 * K, "F_u is staged over the acc_hat and A0 regions of the tile");
    __syncwarp(FULL);                                  // the verification above has read A0
    for (int e = lg; e < MM * NX; e += G) Fs[e] = 0.0;
    __syncwarp(FULL);
#pragma unroll
    for (int e = 0; e < M::RED_NNZ; ++e) {
      if (e % G != lg) continue;
      const int rr = prow[M::red_row(e)], c = M::red_col(e);
      if (c >= K + MM + NF) Fs[rr + MM * (c - K - MM - NF)] = md[(L::OFF_RED + e) * st];
    }
    __syncwarp(FULL);
#pragma unroll
    for (int s = 0; s < SLU; ++s) {
      const int col = s * G + lg;
#pragma unroll
      for (int r = 0; r < MM; ++r) U[s][r] = (col < NX) ? -Fs[r + MM * col] : 0.0;
    }
    okq = lu_solve_static<MM, G, SLU>(A1h, U, lg, FULL) && okq;
#pragma unroll
    for (int s = 0; s < SLU; ++s) {
      const int col = s * G + lg;
      if (col < NX) verified = verified && verify(Fs + MM * col, U[s]);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) verified = verified && (__shfl_xor_sync(FULL, (int)verified, o, G) != 0);
  } else {
#pragma unroll
    for (int e = 0; e < M::RED_NNZ; ++e) {
      const int rr = M::red_row(e), c = M::red_col(e);
      if (c >= K + MM + NF) {
        const int cu = c - K - MM - NF;
        if (lg == cu % G) U[cu / G][rr] = -md[(L::OFF_RED + e) * st];
      }
    }
    if constexpr (SOLVER == CR_GJ) {
#if DYB_CR_SOLVER == 1
      okq = gj_solve<MM, G, SLU>(A1h, U, lg, FULL, pc) && okq;
      static_assert(NX <= K + NF, "X buffer holds the shock columns");
      int xcol[SLU];
#pragma unroll
      for (int s = 0; s < SLU; ++s) xcol[s] = (s * G + lg < NX) ? s * G + lg : -1;
      unpermute_rows<MM, SLU>(U, pc, Xs, xcol);
#endif
    } else {
      okq = qr_solve<MM, G, SLU>(A1h, U, lg, FULL) && okq;
    }
  }
#ifdef DYB_CR_DEBUG
  if (SOLVER == CR_LU && real && d < 2 && lg == 0)
    printf("draw %lld: st_in %d result %d conv %d stalled %d it %d okq %d verified %d n0 %.3e n2 %.3e\n", d, st_in, result, (int)conv,
           (int)stalled, it, (int)okq, (int)verified, n0, n2);
#endif
  if (result == ST_OK && !okq) result = ST_SOLVER;
  if constexpr (SOLVER == CR_LU) {
    // the fast pass only vouches for verified successes; every other draw that entered the solver goes to the second pass
    if (st_in == ST_OK && !(result == ST_OK && verified)) {
      result = ST_REDO;
      if (real && lg == 0) redo_list[atomicAdd(redo_count + 1, 1ULL)] = d;
    }
  }
  if (real && !skip && result == ST_OK) {
#pragma unroll
    for (int s = 0; s < SL0; ++s) {
      const int col = s * G + lg;
      if (col < K) {
#pragma unroll
        for (int r = 0; r < MM; ++r) mid[(L::OFF_G + r + MM * col) * st + d] = Gc[s][r];
      }
    }
#pragma unroll
    for (int s = 0; s < SLU; ++s) {
      const int col = s * G + lg;
      if (col < NX) {
#pragma unroll
        for (int r = 0; r < MM; ++r) mid[(L::OFF_GU + r + MM * col) * st + d] = U[s][r];
      }
    }
  }
  if (real && !skip && lg == 0) {
    status[d] = result;
    if (cr_iters) cr_iters[d] = it;
  }
}

// Pivot order of the equations of the dynamic block, found ONCE per configuration at the calibration point: partial pivoting
// on A1(0) of the draw in `mid1` (one draw, written by model_kernel with no estimated parameter), row_perm[r] = position of
// equation r, row_perm[M] = 1.  When that draw fails row_perm[M] = 0 and the fast pass hands every draw to the Householder pass.
template <class M>
__global__ void cr_perm_kernel(const double* __restrict__ mid1, const int* __restrict__ status1, int* __restrict__ row_perm) {
  constexpr int MM = M::M, K = M::K;
  using L = MidLayout<M>;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double A[MM * MM];
  int orig[MM];
  for (int i = 0; i < MM * MM; ++i) A[i] = 0.0;
  for (int r = 0; r < MM; ++r) { orig[r] = r; row_perm[r] = r; }
  row_perm[MM] = 0;                                  // [MM]: 1 once a pivot order has been found
  if (status1[0] != ST_OK) return;
  for (int e = 0; e < M::RED_NNZ; ++e) {
    const int rr = M::red_row(e), c = M::red_col(e);
    if (c >= K && c < K + MM) A[rr + MM * (c - K)] = mid1[L::OFF_RED + e];
  }
  for (int p = 0; p < MM; ++p) {
    int best = p;
    for (int r = p + 1; r < MM; ++r) if (fabs(A[r + MM * p]) > fabs(A[best + MM * p])) best = r;
    if (best != p) {
      for (int c = 0; c < MM; ++c) { const double t = A[p + MM * c]; A[p + MM * c] = A[best + MM * c]; A[best + MM * c] = t; }
      const int t = orig[p]; orig[p] = orig[best]; orig[best] = t;
    }
    const double piv = A[p + MM * p];
    if (!(fabs(piv) > 0.0)) { for (int r = 0; r < MM; ++r) row_perm[r] = r; return; }
    for (int r = p + 1; r < MM; ++r) {
      const double l = A[r + MM * p] / piv;
      for (int c = p; c < MM; ++c) A[r + MM * c] -= l * A[p + MM * c];
    }
  }
  for (int p = 0; p < MM; ++p) row_perm[orig[p]] = p;
  row_perm[MM] = 1;
}

// ------------------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(DYB_ASM_BLOCK, DYB_ASM_MINB)
assemble_kernel(const double* __restrict__ mid, long long n_draws, double lyap_tol, int lyap_maxit,
                double* __restrict__ ss, int* __restrict__ status, double* __restrict__ g1_out,
                int* __restrict__ lyap_iters) {
  constexpr int N = M::N, NX = M::NX, K = M::K, NF = M::NF, S = M::S, MM = M::M;
  constexpr int NS = M::NS, NY = M::NY;
  constexpr int cC = S, cB = S + K, cA = S + K + MM, cU = S + K + MM + NF;
  using L = SSLayout<M>;
  using ML = MidLayout<M>;
  const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_draws) return;
  if (status[d] != ST_OK) return;
  const double* md = mid + d;
  const long long st = n_draws;

  double Gd[MM * K], gud[MM * NX], Se[NX * NX];
#pragma unroll
  for (int i = 0; i < MM * K; ++i) Gd[i] = md[(ML::OFF_G + i) * st];
#pragma unroll
  for (int i = 0; i < MM * NX; ++i) gud[i] = md[(ML::OFF_GU + i) * st];
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) Se[i] = md[(ML::OFF_SE + i) * st];

  // ---- static rows: R_s y_s = -(C_top + B_top G + A_top G[fw,:] G[bk,:]) etc., sparse pivot rows -------
  double gst[(S > 0 ? S : 1) * (K + NX)];
  bool ok = true;
  if (S > 0) {
    double GG[NF * (K + NX)];
#pragma unroll
    for (int i = 0; i < NF; ++i) {
#pragma unroll
      for (int j = 0; j < K; ++j) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < K; ++q) s = fma(Gd[M::fwrd_in_dyn(i) + MM * q], Gd[M::bkwrd_in_dyn(q) + MM * j], s);
        GG[i + NF * j] = s;
      }
#pragma unroll
      for (int e = 0; e < NX; ++e) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < K; ++q) s = fma(Gd[M::fwrd_in_dyn(i) + MM * q], gud[M::bkwrd_in_dyn(q) + MM * e], s);
        GG[i + NF * (K + e)] = s;
      }
    }
#pragma unroll
    for (int i = 0; i < S * (K + NX); ++i) gst[i] = 0.0;
    double rdiag[S > 0 ? S : 1];
#pragma unroll
    for (int q = 0; q < S; ++q) rdiag[q] = 0.0;
    // accumulate the right-hand sides from the non-static entries of the pivot rows
#pragma unroll
    for (int e = 0; e < M::TOP_NNZ; ++e) {
      const int q = M::top_row(e), c = M::top_col(e);
      const double v = md[(ML::OFF_TOP + e) * st];
      if (c < S) {
        if (c == q) rdiag[q] = v;
      } else if (c < cB) {
        gst[q + S * (c - cC)] -= v;
      } else if (c < cA) {
#pragma unroll
        for (int j = 0; j < K; ++j) gst[q + S * j] = fma(-v, Gd[(c - cB) + MM * j], gst[q + S * j]);
#pragma unroll
        for (int x = 0; x < NX; ++x) gst[q + S * (K + x)] = fma(-v, gud[(c - cB) + MM * x], gst[q + S * (K + x)]);
      } else if (c < cU) {
#pragma unroll
        for (int j = 0; j < K + NX; ++j) gst[q + S * j] = fma(-v, GG[(c - cA) + NF * j], gst[q + S * j]);
      } else {
        gst[q + S * (K + (c - cU))] -= v;
      }
    }
    // back-substitution with the upper-triangular R_s (entries c > q of the pivot rows), last row first
#pragma unroll
    for (int q = S - 1; q >= 0; --q) {
#pragma unroll
      for (int e = 0; e < M::TOP_NNZ; ++e) {
        if (M::top_row(e) == q && M::top_col(e) > q && M::top_col(e) < S) {
          const double v = md[(ML::OFF_TOP + e) * st];
#pragma unroll
          for (int j = 0; j < K + NX; ++j) gst[q + S * j] = fma(-v, gst[M::top_col(e) + S * j], gst[q + S * j]);
        }
      }
      ok = ok && (rdiag[q] != 0.0);
      const double inv = 1.0 / rdiag[q];
#pragma unroll
      for (int j = 0; j < K + NX; ++j) gst[q + S * j] *= inv;
    }
  }
  if (!ok) { status[d] = ST_SOLVER; return; }
  if (g1_out) {
    double* g1 = g1_out + d * (long long)(N * (K + NX));
#pragma unroll
    for (int c = 0; c < K + NX; ++c)
#pragma unroll
      for (int v = 0; v < N; ++v) {
        const int pos = M::var_pos(v);
        double val;
        if (M::var_static(v)) val = gst[pos + S * c];
        else val = (c < K) ? Gd[pos + MM * c] : gud[pos + MM * (c - K)];
        g1[v + N * c] = val;
      }
  }

  // ---- Lyapunov doubling on the K x K predetermined block ---------------------------------------------
  double Ak[K * K], Sg[K * K], Tm[K * K];
  {
    double Bs[K * NX], BQ[K * NX];
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int i = 0; i < K; ++i) Ak[i + K * j] = Gd[M::bkwrd_in_dyn(i) + MM * j];
#pragma unroll
    for (int e = 0; e < NX; ++e)
#pragma unroll
      for (int i = 0; i < K; ++i) Bs[i + K * e] = gud[M::bkwrd_in_dyn(i) + MM * e];
#pragma unroll
    for (int e = 0; e < NX; ++e)
#pragma unroll
      for (int i = 0; i < K; ++i) {
        double s = 0.0;
#pragma unroll
        for (int f = 0; f < NX; ++f) s = fma(Bs[i + K * f], Se[f + NX * e], s);
        BQ[i + K * e] = s;
      }
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int i = 0; i < K; ++i) {
        double s = 0.0;
#pragma unroll
        for (int e = 0; e < NX; ++e) s = fma(BQ[i + K * e], Bs[j + K * e], s);
        Sg[i + K * j] = s;
      }
  }
  bool lconv = false;
  int li_done = 0;
  for (int li = 0; li < lyap_maxit; ++li) {
    li_done = li + 1;
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int i = 0; i < K; ++i) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < K; ++q) s = fma(Ak[i + K * q], Sg[q + K * j], s);
        Tm[i + K * j] = s;
      }
    double dmax = 0.0, smax = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int i = 0; i < K; ++i) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < K; ++q) s = fma(Tm[i + K * q], Ak[j + K * q], s);
        dmax = fmax(dmax, fabs(s));
        const double v = Sg[i + K * j] + s;
        Sg[i + K * j] = v;
        smax = fmax(smax, fabs(v));
      }
    if (!isfinite(smax)) break;
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int i = 0; i < K; ++i) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < K; ++q) s = fma(Ak[i + K * q], Ak[q + K * j], s);
        Tm[i + K * j] = s;
      }
    double amax = 0.0;
#pragma unroll
    for (int i = 0; i < K * K; ++i) { Ak[i] = Tm[i]; amax = fmax(amax, fabs(Tm[i])); }
    if (dmax <= lyap_tol * smax && amax < 1.0) { lconv = true; break; }
    if (li == LYAP_UNIT_ROOT_IT && amax >= LYAP_UNIT_ROOT_NORM) break;      // a root within 1e-6 of the unit circle
    if (!isfinite(amax)) break;
  }
  if (lyap_iters) lyap_iters[d] = li_done;
  if (!lconv) { status[d] = ST_NONSTATIONARY; return; }
#pragma unroll
  for (int j = 0; j < K; ++j)
#pragma unroll
    for (int i = j + 1; i < K; ++i) {
      const double v = 0.5 * (Sg[i + K * j] + Sg[j + K * i]);
      Sg[i + K * j] = v; Sg[j + K * i] = v;
    }

  // ---- state-space assembly (estimation.jl:640-658) ---------------------------------------------------
  double* so = ss + d;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int v = M::state_ids(s);
    const int pos = M::var_pos(v);
    double g1r[K], g2r[NX];
#pragma unroll
    for (int j = 0; j < K; ++j) g1r[j] = M::var_static(v) ? gst[pos + S * j] : Gd[pos + MM * j];
#pragma unroll
    for (int e = 0; e < NX; ++e) g2r[e] = M::var_static(v) ? gst[pos + S * (K + e)] : gud[pos + MM * e];
    double rq[NX], gs[K];
#pragma unroll
    for (int e = 0; e < NX; ++e) {
      double a = 0.0;
#pragma unroll
      for (int f = 0; f < NX; ++f) a = fma(g2r[f], Se[f + NX * e], a);
      rq[e] = a;
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
      double a = 0.0;
#pragma unroll
      for (int q = 0; q < K; ++q) a = fma(g1r[q], Sg[q + K * j], a);
      gs[j] = a;
    }
#pragma unroll
    for (int j = 0; j < K; ++j) so[(L::OFF_T + s * K + j) * st] = g1r[j];
#pragma unroll
    for (int t = 0; t <= s; ++t) {
      const int v2 = M::state_ids(t);
      const int pos2 = M::var_pos(v2);
      double qq = 0.0;
#pragma unroll
      for (int e = 0; e < NX; ++e) qq = fma(rq[e], M::var_static(v2) ? gst[pos2 + S * (K + e)] : gud[pos2 + MM * e], qq);
      double pp = qq;
#pragma unroll
      for (int j = 0; j < K; ++j) pp = fma(gs[j], M::var_static(v2) ? gst[pos2 + S * j] : Gd[pos2 + MM * j], pp);
      so[(L::OFF_QQ + tri(s, t)) * st] = qq;
      so[(L::OFF_P0 + tri(s, t)) * st] = pp;
    }
  }
#pragma unroll
  for (int i = 0; i < NY; ++i) so[(L::OFF_YB + i) * st] = md[(ML::OFF_YB + i) * st];
#pragma unroll
  for (int i = 0; i < NY * (NY + 1) / 2; ++i) so[(L::OFF_H + i) * st] = md[(ML::OFF_H + i) * st];
}

}  // namespace dyb
