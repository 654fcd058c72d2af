// Batched first-order solve: theta -> state space of one draw, one thread per draw.
//
// Stages (SURVEY.md section 8a rows a-1 .. a-8; reference call sites in brackets):
//   K1  parameter scatter, steady state + residual check, Jacobian   [estimation.jl:517-601, SteadyState.jl:183-200,476-482, model.jl:120-139]
//   K2  Householder elimination of the static variables               [LRE.first_order_solver! -> remove_static!, perturbations.jl:663]
//   K3  cyclic reduction on the dynamic block + Blanchard-Kahn status  [PolynomialMatrixEquations CR solver]
//   K4  shock response g_u by pivoted LU, static rows by back-substitution [add_static!]
//   K5  doubling Lyapunov for the predetermined block, variance of the Kalman states [compute_variance!, estimation.jl:630]
//       and assembly of T (lagged columns only), R Q R', P0, ybar[varobs], H  [estimation.jl:640-658]
//
// All sizes are compile-time constants of the generated model traits M; every matrix of a draw is
// private to its thread.  Column structure is exploited: A0 = df/dy(-1) only has the K backward
// columns and A2 = df/dy(+1) only the NF forward columns, and cyclic reduction preserves both.
#pragma once
#include "model_traits.cuh"
#include <cmath>

namespace dyb {

template <class M>
struct SolveParams {
  double params0[M::NPAR > 0 ? M::NPAR : 1];
  double sigma_e0[M::NX * M::NX];
  double sigma_m0[M::NY * M::NY];
  int np;
  int est_type[MAX_EST];
  int est_i[MAX_EST];
  int est_j[MAX_EST];
  double cr_tol;
  int cr_maxit;
  double steady_tol;      // on the 2-norm of the static residuals
  double lyap_tol;
  int lyap_maxit;
};

// number of doubles one draw hands from the solve stage to the filter stage
template <class M>
struct SSLayout {
  static constexpr int NS = M::NS, K = M::K, NY = M::NY;
  static constexpr int OFF_T = 0;                               // NS x K, row-major
  static constexpr int OFF_QQ = OFF_T + NS * K;                 // packed lower, NS(NS+1)/2
  static constexpr int OFF_P0 = OFF_QQ + NS * (NS + 1) / 2;     // packed lower
  static constexpr int OFF_YB = OFF_P0 + NS * (NS + 1) / 2;     // NY steady-state values of the observables
  static constexpr int OFF_H = OFF_YB + NY;                     // packed lower NY(NY+1)/2
  static constexpr int SIZE = OFF_H + NY * (NY + 1) / 2;
};

// accessor that lets the generated Jacobian (compressed column order) write straight into the
// working matrix whose columns are ordered [static | lags | dynamic current | leads | shocks]
template <class M>
struct JacPerm {
  double* w;
  DYB_D double& operator[](int idx) const {
    return w[(idx % M::N) + M::N * M::jperm(idx / M::N)];
  }
};

// LU factorisation with partial pivoting of an n x n column-major matrix, in place.
// Returns false on a zero / non-finite pivot.
template <int n>
DYB_D bool lu_factor(double* a, int* piv) {
  for (int p = 0; p < n; ++p) {
    int r = p;
    double best = fabs(a[p + n * p]);
    for (int i = p + 1; i < n; ++i) {
      double v = fabs(a[i + n * p]);
      if (v > best) { best = v; r = i; }
    }
    piv[p] = r;
    if (!(best > 0.0) || !isfinite(best)) return false;
    if (r != p) {
      for (int c = 0; c < n; ++c) { double t = a[p + n * c]; a[p + n * c] = a[r + n * c]; a[r + n * c] = t; }
    }
    const double inv = 1.0 / a[p + n * p];
    for (int i = p + 1; i < n; ++i) a[i + n * p] *= inv;
    for (int c = p + 1; c < n; ++c) {
      const double u = a[p + n * c];
      for (int i = p + 1; i < n; ++i) a[i + n * c] = fma(-a[i + n * p], u, a[i + n * c]);
    }
  }
  return true;
}

// solve L U x = P b for nrhs right-hand sides stored column-major (n x nrhs), in place
template <int n>
DYB_D void lu_solve(const double* lu, const int* piv, double* b, int nrhs) {
  for (int c = 0; c < nrhs; ++c) {
    double* x = b + n * c;
    for (int p = 0; p < n; ++p) {
      const int r = piv[p];
      if (r != p) { double t = x[p]; x[p] = x[r]; x[r] = t; }
    }
    for (int p = 0; p < n; ++p) {
      const double xp = x[p];
      for (int i = p + 1; i < n; ++i) x[i] = fma(-lu[i + n * p], xp, x[i]);
    }
    for (int p = n - 1; p >= 0; --p) {
      const double xp = x[p] / lu[p + n * p];
      x[p] = xp;
      for (int i = 0; i < p; ++i) x[i] = fma(-lu[i + n * p], xp, x[i]);
    }
  }
}

// Optional per-draw outputs of the stage-level entry points (dyb_first_order_batch)
struct SolveExtra {
  double* g1;        // n x (k+nx) column-major per draw (AoS), or nullptr
  double* ybar;      // n per draw, or nullptr
  int* cr_iters;     // per draw, or nullptr
  int* lyap_iters;   // per draw, or nullptr
};

template <class M>
__device__ int solve_draw(const SolveParams<M>& prm, const double* __restrict__ theta,
                          double* __restrict__ ss, long long stride,   // ss[e * stride]
                          double* g1_out, double* ybar_out, int* cr_it_out, int* lyap_it_out) {
  constexpr int N = M::N, NX = M::NX, K = M::K, NF = M::NF, S = M::S, MM = M::M, NCOL = M::NCOL;
  constexpr int NS = M::NS, NY = M::NY;
  constexpr int cC = S, cB = S + K, cA = S + K + MM, cU = S + K + MM + NF;
  using L = SSLayout<M>;

  // ---- K1: parameter scatter (estimation.jl:517-601) ----------------------------------------
  double p[M::NPAR > 0 ? M::NPAR : 1];
  double Se[NX * NX];
  double Sm[NY * NY];
  for (int i = 0; i < M::NPAR; ++i) p[i] = prm.params0[i];
  for (int i = 0; i < NX * NX; ++i) Se[i] = prm.sigma_e0[i];
  for (int i = 0; i < NY * NY; ++i) Sm[i] = prm.sigma_m0[i];
  for (int q = 0; q < prm.np; ++q) {
    const double v = theta[q];
    const int ty = prm.est_type[q], i = prm.est_i[q], j = prm.est_j[q];
    switch (ty) {
      case EST_PARAM: p[i] = v; break;
      case EST_SD_SHOCK: Se[i + NX * j] = v * v; break;
      case EST_VAR_SHOCK: Se[i + NX * j] = v; break;
      case EST_CORR_SHOCK: { double c = v * sqrt(Se[i + NX * i] * Se[j + NX * j]); Se[i + NX * j] = c; Se[j + NX * i] = c; } break;
      case EST_SD_MEAS: Sm[i + NY * j] = v * v; break;
      case EST_VAR_MEAS: Sm[i + NY * j] = v; break;
      case EST_CORR_MEAS: { double c = v * sqrt(Sm[i + NY * i] * Sm[j + NY * j]); Sm[i + NY * j] = c; Sm[j + NY * i] = c; } break;
      default: break;
    }
  }

  // ---- steady state + residual check ----------------------------------------------------------
  double y[N];
  M::steady_state(p, y);
  bool ok = true;
  for (int i = 0; i < N; ++i) ok = ok && isfinite(y[i]);
  if (ybar_out) for (int i = 0; i < N; ++i) ybar_out[i] = y[i];
  if (!ok) return ST_STEADY;
  {
    const double r2 = M::static_resid_ss(p, y);
    if (!(r2 <= prm.steady_tol * prm.steady_tol)) return ST_STEADY;
  }

  // ---- Jacobian into the working matrix -----------------------------------------------------
  double W[N * NCOL];
  for (int i = 0; i < N * NCOL; ++i) W[i] = 0.0;
  {
    JacPerm<M> acc{W};
    M::jacobian(p, y, acc);
  }
  {
    double chk = 0.0;
    for (int i = 0; i < N * NCOL; ++i) chk += W[i] * 0.0;     // NaN/Inf anywhere -> NaN
    if (!(chk == 0.0)) return ST_STEADY;
  }

  // ---- K2: Householder QR on the S static columns, applied to every other column ----------------
  for (int j = 0; j < S; ++j) {
    double* col = W + N * j;
    double nrm2 = 0.0;
    for (int i = j; i < N; ++i) nrm2 = fma(col[i], col[i], nrm2);
    const double nrm = sqrt(nrm2);
    if (!(nrm > 0.0)) return ST_SOLVER;
    const double alpha = col[j] >= 0.0 ? -nrm : nrm;
    // v = x - alpha e1 ; H = I - 2 v v' / (v'v) ; v'v = 2 (nrm2 - alpha x_j)
    const double vj = col[j] - alpha;
    const double vtv = 2.0 * (nrm2 - alpha * col[j]);
    const double beta = 2.0 / vtv;
    for (int c = j + 1; c < NCOL; ++c) {
      double* x = W + N * c;
      double dot = vj * x[j];
      for (int i = j + 1; i < N; ++i) dot = fma(col[i], x[i], dot);
      dot *= beta;
      x[j] = fma(-dot, vj, x[j]);
      for (int i = j + 1; i < N; ++i) x[i] = fma(-dot, col[i], x[i]);
    }
    col[j] = alpha;
    for (int i = j + 1; i < N; ++i) col[i] = 0.0;
  }

  // ---- K3: cyclic reduction on the MM x MM dynamic block -----------------------------------------
  double A0[MM * K], A2[MM * NF], A1[MM * MM], LU[MM * MM];
  double X[MM * (K + NF)];                 // [X0 | X2] = A1^-1 [A0 | A2]
  double acc_hat[MM * K];                  // A1hat = A1(0) - acc_hat on the backward columns
  int piv[MM];
  for (int j = 0; j < K; ++j) for (int r = 0; r < MM; ++r) { A0[r + MM * j] = W[S + r + N * (cC + j)]; acc_hat[r + MM * j] = 0.0; }
  for (int j = 0; j < NF; ++j) for (int r = 0; r < MM; ++r) A2[r + MM * j] = W[S + r + N * (cA + j)];
  for (int c = 0; c < MM; ++c) for (int r = 0; r < MM; ++r) A1[r + MM * c] = W[S + r + N * (cB + c)];

  int it = 0;
  int status = ST_OK;
  bool conv = false, stalled = false;
  double n0 = 0.0, n2 = 0.0;          // ||A0||_1, ||A2||_1 of the last iteration that stayed finite
  for (it = 1; it <= prm.cr_maxit; ++it) {
    for (int i = 0; i < MM * MM; ++i) LU[i] = A1[i];
    if (!lu_factor<MM>(LU, piv)) { if (it == 1) status = ST_SOLVER; else stalled = true; break; }
    for (int i = 0; i < MM * K; ++i) X[i] = A0[i];
    for (int i = 0; i < MM * NF; ++i) X[MM * K + i] = A2[i];
    lu_solve<MM>(LU, piv, X, K + NF);
    const double* X0 = X;
    const double* X2 = X + MM * K;
    // W01 = A0 X2[bk,:] -> A1 forward columns ;  W10 = A2 X0[fw,:] -> A1 backward columns, acc_hat
    for (int j = 0; j < NF; ++j)
      for (int r = 0; r < MM; ++r) {
        double s = 0.0;
        for (int i = 0; i < K; ++i) s = fma(A0[r + MM * i], X2[M::bkwrd_in_dyn(i) + MM * j], s);
        A1[r + MM * M::fwrd_in_dyn(j)] -= s;
      }
    for (int j = 0; j < K; ++j)
      for (int r = 0; r < MM; ++r) {
        double s = 0.0;
        for (int i = 0; i < NF; ++i) s = fma(A2[r + MM * i], X0[M::fwrd_in_dyn(i) + MM * j], s);
        A1[r + MM * M::bkwrd_in_dyn(j)] -= s;
        acc_hat[r + MM * j] += s;
      }
    // A0 <- -A0 X0[bk,:] ; A2 <- -A2 X2[fw,:]   (row by row so the update can be done in place)
    double m0 = 0.0, m2 = 0.0;
    {
      double coln[K > NF ? K : NF];
      for (int j = 0; j < K; ++j) coln[j] = 0.0;
      for (int r = 0; r < MM; ++r) {
        double row[K];
        for (int j = 0; j < K; ++j) {
          double s = 0.0;
          for (int i = 0; i < K; ++i) s = fma(A0[r + MM * i], X0[M::bkwrd_in_dyn(i) + MM * j], s);
          row[j] = -s;
        }
        for (int j = 0; j < K; ++j) { A0[r + MM * j] = row[j]; coln[j] += fabs(row[j]); }
      }
      for (int j = 0; j < K; ++j) m0 = fmax(m0, coln[j]);
      for (int j = 0; j < NF; ++j) coln[j] = 0.0;
      for (int r = 0; r < MM; ++r) {
        double row[NF];
        for (int j = 0; j < NF; ++j) {
          double s = 0.0;
          for (int i = 0; i < NF; ++i) s = fma(A2[r + MM * i], X2[M::fwrd_in_dyn(i) + MM * j], s);
          row[j] = -s;
        }
        for (int j = 0; j < NF; ++j) { A2[r + MM * j] = row[j]; coln[j] += fabs(row[j]); }
      }
      for (int j = 0; j < NF; ++j) m2 = fmax(m2, coln[j]);
    }
    if (!(m0 <= 1e150) || !(m2 <= 1e150)) {          // NaN / overflow: one of the two sequences diverges
      if (it == 1) status = ST_SOLVER; else stalled = true;
      break;
    }
    const double n0_prev = n0;
    n0 = m0; n2 = m2;
    // unit roots among the stable roots: A0(j) settles on a non-zero limit while A2(j) vanishes (see solve_coop.cuh)
    if ((n0 < prm.cr_tol && n2 < prm.cr_tol) || (n2 < prm.cr_tol * prm.cr_tol && fabs(n0 - n0_prev) <= CR_UNIT_ROOT_EPS * n0)) { conv = true; break; }
  }
  if (cr_it_out) *cr_it_out = it;
  if (status == ST_OK && !conv) {
    // Blanchard-Kahn reading of a reduction that does not converge: A0(j) ~ lambda_m^(2^j) fails to vanish
    // when fewer than MM roots lie inside the unit circle (no stable solution: UnstableSystemException);
    // A2(j) ~ lambda_(m+1)^-(2^j) fails to vanish when more than MM do (UndeterminateSystemException).
    // When the iteration blows up the sequence that was still large decides; at maxit the tolerance does.
    if (stalled) status = (n0 < n2) ? ST_INDETERMINATE : ST_UNSTABLE;
    else status = (n0 >= prm.cr_tol) ? ST_UNSTABLE : ST_INDETERMINATE;
  }
  if (status != ST_OK) return status;

  // G (dynamic rows, backward columns) = -A1hat^-1 A0(0)
  double Gd[MM * K];
  for (int c = 0; c < MM; ++c) for (int r = 0; r < MM; ++r) LU[r + MM * c] = W[S + r + N * (cB + c)];
  for (int j = 0; j < K; ++j) for (int r = 0; r < MM; ++r) LU[r + MM * M::bkwrd_in_dyn(j)] -= acc_hat[r + MM * j];
  if (!lu_factor<MM>(LU, piv)) return ST_SOLVER;
  for (int j = 0; j < K; ++j) for (int r = 0; r < MM; ++r) Gd[r + MM * j] = -W[S + r + N * (cC + j)];
  lu_solve<MM>(LU, piv, Gd, K);

  // ---- K4: g_u = -(A G + B)^-1 F_u on the dynamic block -----------------------------------------
  double gud[MM * NX];
  for (int c = 0; c < MM; ++c) for (int r = 0; r < MM; ++r) LU[r + MM * c] = W[S + r + N * (cB + c)];
  for (int j = 0; j < K; ++j)
    for (int r = 0; r < MM; ++r) {
      double s = 0.0;
      for (int i = 0; i < NF; ++i) s = fma(W[S + r + N * (cA + i)], Gd[M::fwrd_in_dyn(i) + MM * j], s);
      LU[r + MM * M::bkwrd_in_dyn(j)] += s;
    }
  if (!lu_factor<MM>(LU, piv)) return ST_SOLVER;
  for (int e = 0; e < NX; ++e) for (int r = 0; r < MM; ++r) gud[r + MM * e] = -W[S + r + N * (cU + e)];
  lu_solve<MM>(LU, piv, gud, NX);

  // static rows:  R_s y_s + B_top y_d + A_top y_f(+1) + C_top y_b(-1) + Fu_top u = 0
  double gst[(S > 0 ? S : 1) * (K + NX)];
  if (S > 0) {
    // GG = G[fw,:] G[bk,:] (NF x K) and Gg = G[fw,:] gu[bk,:] (NF x NX)
    double GG[NF * (K + NX)];
    for (int i = 0; i < NF; ++i) {
      for (int j = 0; j < K; ++j) {
        double s = 0.0;
        for (int q = 0; q < K; ++q) s = fma(Gd[M::fwrd_in_dyn(i) + MM * q], Gd[M::bkwrd_in_dyn(q) + MM * j], s);
        GG[i + NF * j] = s;
      }
      for (int e = 0; e < NX; ++e) {
        double s = 0.0;
        for (int q = 0; q < K; ++q) s = fma(Gd[M::fwrd_in_dyn(i) + MM * q], gud[M::bkwrd_in_dyn(q) + MM * e], s);
        GG[i + NF * (K + e)] = s;
      }
    }
    for (int c = 0; c < K + NX; ++c) {
      for (int q = 0; q < S; ++q) {
        double s = (c < K) ? W[q + N * (cC + c)] : W[q + N * (cU + (c - K))];
        const double* gcol = (c < K) ? (Gd + MM * c) : (gud + MM * (c - K));
        for (int d = 0; d < MM; ++d) s = fma(W[q + N * (cB + d)], gcol[d], s);
        for (int i = 0; i < NF; ++i) s = fma(W[q + N * (cA + i)], GG[i + NF * c], s);
        gst[q + S * c] = -s;
      }
      for (int q = S - 1; q >= 0; --q) {
        double s = gst[q + S * c];
        for (int r = q + 1; r < S; ++r) s = fma(-W[q + N * r], gst[r + S * c], s);
        const double d = W[q + N * q];
        if (d == 0.0) return ST_SOLVER;
        gst[q + S * c] = s / d;
      }
    }
  }
  if (g1_out) {
    for (int c = 0; c < K + NX; ++c)
      for (int v = 0; v < N; ++v) {
        const int pos = M::var_pos(v);
        double val;
        if (M::var_static(v)) val = gst[pos + S * c];
        else val = (c < K) ? Gd[pos + MM * c] : gud[pos + MM * (c - K)];
        g1_out[v + N * c] = val;
      }
  }

  // ---- K5: Lyapunov doubling on the K x K predetermined block ---------------------------------------
  double Ak[K * K], Sg[K * K], Tm[K * K], Bs[K * NX];
  for (int j = 0; j < K; ++j) for (int i = 0; i < K; ++i) Ak[i + K * j] = Gd[M::bkwrd_in_dyn(i) + MM * j];
  for (int e = 0; e < NX; ++e) for (int i = 0; i < K; ++i) Bs[i + K * e] = gud[M::bkwrd_in_dyn(i) + MM * e];
  // Sg = Bs Se Bs'
  {
    double BQ[K * NX];
    for (int e = 0; e < NX; ++e) for (int i = 0; i < K; ++i) {
      double s = 0.0;
      for (int f = 0; f < NX; ++f) s = fma(Bs[i + K * f], Se[f + NX * e], s);
      BQ[i + K * e] = s;
    }
    for (int j = 0; j < K; ++j) for (int i = 0; i < K; ++i) {
      double s = 0.0;
      for (int e = 0; e < NX; ++e) s = fma(BQ[i + K * e], Bs[j + K * e], s);
      Sg[i + K * j] = s;
    }
  }
  bool lconv = false;
  int li_done = 0;
  for (int li = 0; li < prm.lyap_maxit; ++li) {
    li_done = li + 1;
    // Tm = Ak Sg ; Sg += Tm Ak' ; Ak = Ak Ak
    for (int j = 0; j < K; ++j) for (int i = 0; i < K; ++i) {
      double s = 0.0;
      for (int q = 0; q < K; ++q) s = fma(Ak[i + K * q], Sg[q + K * j], s);
      Tm[i + K * j] = s;
    }
    double dmax = 0.0, smax = 0.0;
    for (int j = 0; j < K; ++j) for (int i = 0; i < K; ++i) {
      double s = 0.0;
      for (int q = 0; q < K; ++q) s = fma(Tm[i + K * q], Ak[j + K * q], s);
      dmax = fmax(dmax, fabs(s));
      const double v = Sg[i + K * j] + s;
      Sg[i + K * j] = v;
      smax = fmax(smax, fabs(v));
    }
    if (!isfinite(smax)) break;
    for (int j = 0; j < K; ++j) for (int i = 0; i < K; ++i) {
      double s = 0.0;
      for (int q = 0; q < K; ++q) s = fma(Ak[i + K * q], Ak[q + K * j], s);
      Tm[i + K * j] = s;
    }
    double amax = 0.0;
    for (int i = 0; i < K * K; ++i) { Ak[i] = Tm[i]; amax = fmax(amax, fabs(Tm[i])); }
    if (dmax <= prm.lyap_tol * smax && amax < 1.0) { lconv = true; break; }
    if (li == LYAP_UNIT_ROOT_IT && amax >= LYAP_UNIT_ROOT_NORM) break;      // a root within 1e-6 of the unit circle
    if (!isfinite(amax)) break;
  }
  if (lyap_it_out) *lyap_it_out = li_done;
  if (!lconv) return ST_NONSTATIONARY;
  // symmetrise
  for (int j = 0; j < K; ++j) for (int i = j + 1; i < K; ++i) {
    const double v = 0.5 * (Sg[i + K * j] + Sg[j + K * i]);
    Sg[i + K * j] = v; Sg[j + K * i] = v;
  }

  // ---- state-space assembly (estimation.jl:640-658) -------------------------------------------------
  // rows of g1 for the Kalman states
  double g1s[NS * K], g2s[NS * NX];
  for (int s = 0; s < NS; ++s) {
    const int v = M::state_ids(s);
    const int pos = M::var_pos(v);
    if (M::var_static(v)) {
      for (int j = 0; j < K; ++j) g1s[s + NS * j] = gst[pos + S * j];
      for (int e = 0; e < NX; ++e) g2s[s + NS * e] = gst[pos + S * (K + e)];
    } else {
      for (int j = 0; j < K; ++j) g1s[s + NS * j] = Gd[pos + MM * j];
      for (int e = 0; e < NX; ++e) g2s[s + NS * e] = gud[pos + MM * e];
    }
  }
  for (int s = 0; s < NS; ++s) for (int j = 0; j < K; ++j) ss[(L::OFF_T + s * K + j) * stride] = g1s[s + NS * j];
  {
    double RQ[NS * NX], GS[NS * K];
    for (int e = 0; e < NX; ++e) for (int s = 0; s < NS; ++s) {
      double a = 0.0;
      for (int f = 0; f < NX; ++f) a = fma(g2s[s + NS * f], Se[f + NX * e], a);
      RQ[s + NS * e] = a;
    }
    for (int j = 0; j < K; ++j) for (int s = 0; s < NS; ++s) {
      double a = 0.0;
      for (int q = 0; q < K; ++q) a = fma(g1s[s + NS * q], Sg[q + K * j], a);
      GS[s + NS * j] = a;
    }
    for (int s = 0; s < NS; ++s) for (int t = 0; t <= s; ++t) {
      double qq = 0.0;
      for (int e = 0; e < NX; ++e) qq = fma(RQ[s + NS * e], g2s[t + NS * e], qq);
      double pp = qq;
      for (int j = 0; j < K; ++j) pp = fma(GS[s + NS * j], g1s[t + NS * j], pp);
      ss[(L::OFF_QQ + tri(s, t)) * stride] = qq;
      ss[(L::OFF_P0 + tri(s, t)) * stride] = pp;
    }
  }
  for (int i = 0; i < NY; ++i) ss[(L::OFF_YB + i) * stride] = y[M::state_ids(M::obs_idx_state(i))];
  for (int i = 0; i < NY; ++i) for (int j = 0; j <= i; ++j) ss[(L::OFF_H + tri(i, j)) * stride] = Sm[i + NY * j];
  return ST_OK;
}

template <class M>
__global__ void __launch_bounds__(128)
solve_kernel(const SolveParams<M> prm, const double* __restrict__ theta, long long n_draws,
             double* __restrict__ ss, int* __restrict__ status, SolveExtra ex) {
  const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_draws) return;
  double* g1 = ex.g1 ? ex.g1 + d * (long long)(M::N * (M::K + M::NX)) : nullptr;
  double* yb = ex.ybar ? ex.ybar + d * (long long)M::N : nullptr;
  int* ci = ex.cr_iters ? ex.cr_iters + d : nullptr;
  int* li = ex.lyap_iters ? ex.lyap_iters + d : nullptr;
  const int st = solve_draw<M>(prm, theta + d * (long long)prm.np, ss + d, n_draws, g1, yb, ci, li);
  status[d] = st;
}

}  // namespace dyb
