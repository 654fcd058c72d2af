/*
 * TEST INFRASTRUCTURE (CPU oracle port) -- never linked into or called by the product library.
 *
 * Plain-C restatement of the reference's per-draw likelihood loop
 *   for j in 1:N  loglikelihood(theta_j, context, observations, ssws)
 * (reference src/estimation/estimation.jl:603-702; serial particle loop src/estimation/smc.jl:172-233),
 * used (a) as a second CPU checker of the numpy oracle and (b) as the CPU baseline timed by
 * bench.py on the host cores (one draw per OpenMP thread).  Dense algebra throughout, as the
 * reference does through BLAS/LAPACK: no structural shortcuts.  The first-order solve uses
 * cyclic reduction (the reference's dr_algo "CR" option, src/perturbations.jl:27-28) because
 * LAPACK's dgges is not available to plain C here; oracle/pipeline.py carries the QZ ("GS") variant
 * and tests/test_oracle_c.py checks that the two agree.
 *
 * PARITY UNPINNED by the reference (no stored likelihoods); pinned against oracle/pipeline.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle_model.h"

extern const oracle_model oracle_model_fs2000;
extern const oracle_model oracle_model_ls2003;
extern const oracle_model oracle_model_hs_nk;

enum { ST_OK = 0, ST_STEADY, ST_INDETERMINATE, ST_UNSTABLE, ST_SOLVER, ST_NONSTATIONARY, ST_F_NOT_PD };

static const oracle_model* find_model(const char* name) {
  if (!strcmp(name, "fs2000")) return &oracle_model_fs2000;
  if (!strcmp(name, "ls2003")) return &oracle_model_ls2003;
  if (!strcmp(name, "hs_nk")) return &oracle_model_hs_nk;
  return 0;
}

#define A_(M, ld, i, j) (M)[(i) + (size_t)(ld) * (j)]

/* LU with partial pivoting, column-major n x n; returns 0 on success */
static int lu_factor(int n, double* a, int* piv) {
  for (int p = 0; p < n; ++p) {
    int r = p; double best = fabs(A_(a, n, p, p));
    for (int i = p + 1; i < n; ++i) { double v = fabs(A_(a, n, i, p)); if (v > best) { best = v; r = i; } }
    piv[p] = r;
    if (!(best > 0.0) || !isfinite(best)) return 1;
    if (r != p) for (int c = 0; c < n; ++c) { double t = A_(a, n, p, c); A_(a, n, p, c) = A_(a, n, r, c); A_(a, n, r, c) = t; }
    double inv = 1.0 / A_(a, n, p, p);
    for (int i = p + 1; i < n; ++i) A_(a, n, i, p) *= inv;
    for (int c = p + 1; c < n; ++c) {
      double u = A_(a, n, p, c);
      for (int i = p + 1; i < n; ++i) A_(a, n, i, c) -= A_(a, n, i, p) * u;
    }
  }
  return 0;
}
static void lu_solve(int n, const double* lu, const int* piv, double* b, int nrhs) {
  for (int c = 0; c < nrhs; ++c) {
    double* x = b + (size_t)n * c;
    for (int p = 0; p < n; ++p) { int r = piv[p]; if (r != p) { double t = x[p]; x[p] = x[r]; x[r] = t; } }
    for (int p = 0; p < n; ++p) { double xp = x[p]; for (int i = p + 1; i < n; ++i) x[i] -= A_(lu, n, i, p) * xp; }
    for (int p = n - 1; p >= 0; --p) { double xp = x[p] / A_(lu, n, p, p); x[p] = xp; for (int i = 0; i < p; ++i) x[i] -= A_(lu, n, i, p) * xp; }
  }
}
/* C (m x n) = alpha * A (m x k) * B (k x n) + beta * C, column-major; tb != 0 uses B' (B is n x k) */
static void gemm(int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, int tb,
                 double beta, double* Cm, int ldc) {
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < m; ++i) A_(Cm, ldc, i, j) = beta == 0.0 ? 0.0 : beta * A_(Cm, ldc, i, j);
    for (int q = 0; q < k; ++q) {
      double b = alpha * (tb ? A_(B, ldb, j, q) : A_(B, ldb, q, j));
      if (b == 0.0) continue;
      for (int i = 0; i < m; ++i) A_(Cm, ldc, i, j) += A_(A, lda, i, q) * b;
    }
  }
}

typedef struct {
  double *p, *Se, *Sm, *y, *J, *W, *A0, *A1, *A2, *A1h, *LU, *X, *Wk, *G, *gu, *M1, *g1, *Ak, *Sg, *Tm, *T2, *Sy;
  double *T, *R, *QQ, *P, *a, *Yd, *F, *PZ, *K, *v, *tmp, *tmp2;
  int* piv;
} ws_t;

static double* dalloc(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }

static void ws_init(ws_t* w, const oracle_model* M, int nobs) {
  int n = M->n, nx = M->nx, ns = M->ns, ny = M->ny, k = M->k, m = n - M->s;
  w->p = dalloc(M->npar); w->Se = dalloc(nx * nx); w->Sm = dalloc(ny * ny); w->y = dalloc(n);
  w->J = dalloc((size_t)n * M->ncol); w->W = dalloc((size_t)n * M->ncol);
  w->A0 = dalloc(m * m); w->A1 = dalloc(m * m); w->A2 = dalloc(m * m); w->A1h = dalloc(m * m); w->LU = dalloc(n * n);
  w->X = dalloc((size_t)m * 2 * m); w->Wk = dalloc((size_t)4 * m * m); w->G = dalloc(n * n); w->gu = dalloc(n * nx);
  w->M1 = dalloc(n * n); w->g1 = dalloc((size_t)n * (k + nx));
  w->Ak = dalloc(k * k); w->Sg = dalloc(k * k); w->Tm = dalloc(k * k); w->T2 = dalloc(k * k); w->Sy = dalloc(n * n);
  w->T = dalloc(ns * ns); w->R = dalloc(ns * nx); w->QQ = dalloc(ns * ns); w->P = dalloc(ns * ns); w->a = dalloc(ns);
  w->Yd = dalloc((size_t)ny * nobs); w->F = dalloc(ny * ny); w->PZ = dalloc(ns * ny); w->K = dalloc(ns * ny); w->v = dalloc(ny);
  w->tmp = dalloc((size_t)(n > ns ? n : ns) * (n > ns ? n : ns) + n * nx); w->tmp2 = dalloc((size_t)(n > ns ? n : ns) * (n > ns ? n : ns) + n * nx);
  w->piv = (int*)calloc(n, sizeof(int));
}
static void ws_free(ws_t* w) {
  double** all[] = {&w->p, &w->Se, &w->Sm, &w->y, &w->J, &w->W, &w->A0, &w->A1, &w->A2, &w->A1h, &w->LU, &w->X, &w->Wk, &w->G,
                    &w->gu, &w->M1, &w->g1, &w->Ak, &w->Sg, &w->Tm, &w->T2, &w->Sy, &w->T, &w->R, &w->QQ, &w->P, &w->a, &w->Yd,
                    &w->F, &w->PZ, &w->K, &w->v, &w->tmp, &w->tmp2};
  for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); ++i) free(*all[i]);
  free(w->piv);
}

#define CR_UNIT_ROOT_EPS 1e-6
typedef struct { double cr_tol; int cr_maxit; double steady_tol; double lyap_tol; int lyap_maxit; } opts_t;

/* theta -> g1 (n x (k+nx)), ybar, Se, Sm; returns status */
static int solve_draw(const oracle_model* M, const opts_t* o, int np, const int* et, const int* ei, const int* ej,
                      const double* theta, ws_t* w) {
  const int n = M->n, nx = M->nx, ny = M->ny, k = M->k, nf = M->nf, s = M->s, m = n - s, ncol = M->ncol;
  memcpy(w->p, M->params0, sizeof(double) * M->npar);
  memcpy(w->Se, M->sigma_e0, sizeof(double) * nx * nx);
  memcpy(w->Sm, M->sigma_m0, sizeof(double) * ny * ny);
  for (int q = 0; q < np; ++q) {
    double v = theta[q]; int i = ei[q], j = ej[q];
    switch (et[q]) {
      case 0: w->p[i] = v; break;
      case 1: A_(w->Se, nx, i, j) = v * v; break;
      case 2: A_(w->Se, nx, i, j) = v; break;
      case 3: { double c = v * sqrt(A_(w->Se, nx, i, i) * A_(w->Se, nx, j, j)); A_(w->Se, nx, i, j) = c; A_(w->Se, nx, j, i) = c; } break;
      case 4: A_(w->Sm, ny, i, j) = v * v; break;
      case 5: A_(w->Sm, ny, i, j) = v; break;
      case 6: { double c = v * sqrt(A_(w->Sm, ny, i, i) * A_(w->Sm, ny, j, j)); A_(w->Sm, ny, i, j) = c; A_(w->Sm, ny, j, i) = c; } break;
    }
  }
  M->steady_state(w->p, w->y);
  for (int i = 0; i < n; ++i) if (!isfinite(w->y[i])) return ST_STEADY;
  { double r2 = M->static_resid_ss(w->p, w->y); if (!(sqrt(r2) <= o->steady_tol)) return ST_STEADY; }
  memset(w->J, 0, sizeof(double) * n * ncol);
  M->jacobian(w->p, w->y, w->J);
  for (int i = 0; i < n * ncol; ++i) if (!isfinite(w->J[i])) return ST_STEADY;

  /* full-width blocks in W: [static cols | C (k) | B dyn (m) | A (nf) | Fu (nx)] */
  int cC = s, cB = s + k, cA = s + k + m, cU = s + k + m + nf;
  int* dynpos = (int*)malloc(sizeof(int) * n);
  int* statpos = (int*)malloc(sizeof(int) * n);
  for (int i = 0; i < n; ++i) { dynpos[i] = -1; statpos[i] = -1; }
  for (int j = 0; j < m; ++j) dynpos[M->i_dyn[j]] = j;
  for (int j = 0; j < s; ++j) statpos[M->i_static[j]] = j;
  double* W = w->W;
  for (int j = 0; j < k; ++j) memcpy(W + (size_t)n * (cC + j), w->J + (size_t)n * j, sizeof(double) * n);
  for (int v = 0; v < n; ++v) {
    int dst = statpos[v] >= 0 ? statpos[v] : cB + dynpos[v];
    memcpy(W + (size_t)n * dst, w->J + (size_t)n * (k + v), sizeof(double) * n);
  }
  for (int j = 0; j < nf + nx; ++j) memcpy(W + (size_t)n * (cA + j), w->J + (size_t)n * (k + n + j), sizeof(double) * n);
  /* Householder QR on the static columns */
  int status = ST_OK;
  for (int j = 0; j < s && status == ST_OK; ++j) {
    double* col = W + (size_t)n * j;
    double nrm2 = 0; for (int i = j; i < n; ++i) nrm2 += col[i] * col[i];
    double nrm = sqrt(nrm2);
    if (!(nrm > 0)) { status = ST_SOLVER; break; }
    double alpha = col[j] >= 0 ? -nrm : nrm;
    double vj = col[j] - alpha, beta = 2.0 / (2.0 * (nrm2 - alpha * col[j]));
    for (int c = j + 1; c < ncol; ++c) {
      double* x = W + (size_t)n * c;
      double dot = vj * x[j]; for (int i = j + 1; i < n; ++i) dot += col[i] * x[i];
      dot *= beta; x[j] -= dot * vj; for (int i = j + 1; i < n; ++i) x[i] -= dot * col[i];
    }
    col[j] = alpha; for (int i = j + 1; i < n; ++i) col[i] = 0;
  }
  if (status != ST_OK) { free(dynpos); free(statpos); return status; }
  /* dense m x m blocks of the reduced system */
  double *A0 = w->A0, *A1 = w->A1, *A2 = w->A2, *A1h = w->A1h;
  memset(A0, 0, sizeof(double) * m * m); memset(A2, 0, sizeof(double) * m * m);
  for (int j = 0; j < k; ++j) { int c = dynpos[M->i_bkwrd_b[j]]; for (int r = 0; r < m; ++r) A_(A0, m, r, c) = A_(W, n, s + r, cC + j); }
  for (int j = 0; j < nf; ++j) { int c = dynpos[M->i_fwrd_b[j]]; for (int r = 0; r < m; ++r) A_(A2, m, r, c) = A_(W, n, s + r, cA + j); }
  for (int c = 0; c < m; ++c) for (int r = 0; r < m; ++r) A_(A1, m, r, c) = A_(W, n, s + r, cB + c);
  memcpy(A1h, A1, sizeof(double) * m * m);
  double* A00 = w->tmp;   /* copies of the initial A0 (and of B, A for g_u) */
  memcpy(A00, A0, sizeof(double) * m * m);
  double* A20 = w->tmp2; memcpy(A20, A2, sizeof(double) * m * m);
  double* B0 = w->M1; memcpy(B0, A1, sizeof(double) * m * m);
  int conv = 0, stalled = 0; double n0 = 0, n2 = 0;
  for (int it = 1; it <= o->cr_maxit; ++it) {
    memcpy(w->LU, A1, sizeof(double) * m * m);
    if (lu_factor(m, w->LU, w->piv)) { if (it == 1) status = ST_SOLVER; else stalled = 1; break; }
    memcpy(w->X, A0, sizeof(double) * m * m); memcpy(w->X + (size_t)m * m, A2, sizeof(double) * m * m);
    lu_solve(m, w->LU, w->piv, w->X, 2 * m);
    double *W00 = w->Wk, *W01 = w->Wk + m * m, *W10 = w->Wk + 2 * m * m, *W11 = w->Wk + 3 * m * m;
    gemm(m, m, m, 1.0, A0, m, w->X, m, 0, 0.0, W00, m);
    gemm(m, m, m, 1.0, A0, m, w->X + (size_t)m * m, m, 0, 0.0, W01, m);
    gemm(m, m, m, 1.0, A2, m, w->X, m, 0, 0.0, W10, m);
    gemm(m, m, m, 1.0, A2, m, w->X + (size_t)m * m, m, 0, 0.0, W11, m);
    double m0 = 0, m2 = 0;
    for (int c = 0; c < m; ++c) {
      double c0 = 0, c2 = 0;
      for (int r = 0; r < m; ++r) {
        A_(A1, m, r, c) -= A_(W01, m, r, c) + A_(W10, m, r, c);
        A_(A1h, m, r, c) -= A_(W10, m, r, c);
        A_(A0, m, r, c) = -A_(W00, m, r, c); A_(A2, m, r, c) = -A_(W11, m, r, c);
        c0 += fabs(A_(A0, m, r, c)); c2 += fabs(A_(A2, m, r, c));
      }
      if (!(c0 <= m0)) m0 = c0;
      if (!(c2 <= m2)) m2 = c2;
    }
    if (!(m0 <= 1e150) || !(m2 <= 1e150)) { if (it == 1) status = ST_SOLVER; else stalled = 1; break; }
    {
      const double n0_prev = n0;
      n0 = m0; n2 = m2;
      /* unit roots among the stable roots: A0(j) settles on a non-zero limit while A2(j) vanishes */
      if ((n0 < o->cr_tol && n2 < o->cr_tol) || (n2 < o->cr_tol * o->cr_tol && fabs(n0 - n0_prev) <= CR_UNIT_ROOT_EPS * n0)) { conv = 1; break; }
    }
  }
  /* Blanchard-Kahn reading of a non-converging reduction: A0(j) not vanishing = too few stable roots
   * (UnstableSystemException); A2(j) not vanishing = too many (UndeterminateSystemException) */
  if (status == ST_OK && !conv) {
    if (stalled) status = (n0 < n2) ? ST_INDETERMINATE : ST_UNSTABLE;
    else status = (n0 >= o->cr_tol) ? ST_UNSTABLE : ST_INDETERMINATE;
  }
  if (status != ST_OK) { free(dynpos); free(statpos); return status; }
  /* Gd = -A1h^-1 A0(0)  (m x m, non-zero columns at the backward positions) */
  double* Gd = w->G;
  memcpy(w->LU, A1h, sizeof(double) * m * m);
  if (lu_factor(m, w->LU, w->piv)) { free(dynpos); free(statpos); return ST_SOLVER; }
  for (int i = 0; i < m * m; ++i) Gd[i] = -A00[i];
  lu_solve(m, w->LU, w->piv, Gd, m);
  /* g_u (dynamic rows) = -(A2(0) Gd + B0)^-1 Fu_d */
  double* gud = w->gu;
  memcpy(w->LU, B0, sizeof(double) * m * m);
  gemm(m, m, m, 1.0, A20, m, Gd, m, 0, 1.0, w->LU, m);
  if (lu_factor(m, w->LU, w->piv)) { free(dynpos); free(statpos); return ST_SOLVER; }
  for (int e = 0; e < nx; ++e) for (int r = 0; r < m; ++r) A_(gud, m, r, e) = -A_(W, n, s + r, cU + e);
  lu_solve(m, w->LU, w->piv, gud, nx);
  /* assemble g1: dynamic rows, then static rows by back-substitution */
  double* g1 = w->g1;
  for (int j = 0; j < k; ++j) { int c = dynpos[M->i_bkwrd_b[j]]; for (int d = 0; d < m; ++d) A_(g1, n, M->i_dyn[d], j) = A_(Gd, m, d, c); }
  for (int e = 0; e < nx; ++e) for (int d = 0; d < m; ++d) A_(g1, n, M->i_dyn[d], k + e) = A_(gud, m, d, e);
  if (s > 0) {
    /* y_f(+1) response: GG = G[fw,:] * [G[bk,:] | gu[bk,:]] */
    for (int c = 0; c < k + nx; ++c) {
      double* rhs = w->X;   /* s entries */
      for (int q = 0; q < s; ++q) {
        double acc = (c < k) ? A_(W, n, q, cC + c) : A_(W, n, q, cU + (c - k));
        for (int d = 0; d < m; ++d) acc += A_(W, n, q, cB + d) * A_(g1, n, M->i_dyn[d], c);
        for (int i = 0; i < nf; ++i) {
          double gg = 0;
          for (int jj = 0; jj < k; ++jj) gg += A_(g1, n, M->i_fwrd_b[i], jj) * A_(g1, n, M->i_bkwrd_b[jj], c);
          acc += A_(W, n, q, cA + i) * gg;
        }
        rhs[q] = -acc;
      }
      for (int q = s - 1; q >= 0; --q) {
        double acc = rhs[q];
        for (int r = q + 1; r < s; ++r) acc -= A_(W, n, q, r) * rhs[r];
        if (A_(W, n, q, q) == 0.0) { free(dynpos); free(statpos); return ST_SOLVER; }
        rhs[q] = acc / A_(W, n, q, q);
      }
      for (int q = 0; q < s; ++q) A_(g1, n, M->i_static[q], c) = rhs[q];
    }
  }
  free(dynpos); free(statpos);
  return ST_OK;
}

/* unconditional variance of all endogenous variables (compute_variance!): doubling on the k x k block */
static int variance(const oracle_model* M, const opts_t* o, ws_t* w) {
  const int n = M->n, nx = M->nx, k = M->k;
  double *Ak = w->Ak, *Sg = w->Sg, *Tm = w->Tm, *T2 = w->T2;
  double* Bs = w->tmp;      /* k x nx */
  for (int j = 0; j < k; ++j) for (int i = 0; i < k; ++i) A_(Ak, k, i, j) = A_(w->g1, n, M->i_bkwrd_b[i], j);
  for (int e = 0; e < nx; ++e) for (int i = 0; i < k; ++i) A_(Bs, k, i, e) = A_(w->g1, n, M->i_bkwrd_b[i], k + e);
  double* BQ = w->tmp2;
  gemm(k, nx, nx, 1.0, Bs, k, w->Se, nx, 0, 0.0, BQ, k);
  gemm(k, k, nx, 1.0, BQ, k, Bs, k, 1, 0.0, Sg, k);
  int conv = 0;
  for (int it = 0; it < o->lyap_maxit; ++it) {
    gemm(k, k, k, 1.0, Ak, k, Sg, k, 0, 0.0, Tm, k);
    gemm(k, k, k, 1.0, Tm, k, Ak, k, 1, 0.0, T2, k);
    double dmax = 0, smax = 0;
    for (int i = 0; i < k * k; ++i) { Sg[i] += T2[i]; if (fabs(T2[i]) > dmax) dmax = fabs(T2[i]); if (fabs(Sg[i]) > smax) smax = fabs(Sg[i]); }
    if (!isfinite(smax)) break;
    gemm(k, k, k, 1.0, Ak, k, Ak, k, 0, 0.0, Tm, k);
    double amax = 0;
    for (int i = 0; i < k * k; ++i) { Ak[i] = Tm[i]; if (fabs(Tm[i]) > amax) amax = fabs(Tm[i]); }
    if (dmax <= o->lyap_tol * smax && amax < 1.0) { conv = 1; break; }
    if (it == 19 && amax >= 0.35) break;   /* a root within 1e-6 of the unit circle: (1 - 1e-6)^(2^20) = 0.35 */
    if (!isfinite(amax)) break;
  }
  if (!conv) return ST_NONSTATIONARY;
  for (int j = 0; j < k; ++j) for (int i = j + 1; i < k; ++i) { double v = 0.5 * (A_(Sg, k, i, j) + A_(Sg, k, j, i)); A_(Sg, k, i, j) = v; A_(Sg, k, j, i) = v; }
  /* Sy = g1_1 Sg g1_1' + g1_2 Se g1_2' */
  double* t1 = w->tmp;
  gemm(n, k, k, 1.0, w->g1, n, Sg, k, 0, 0.0, t1, n);
  gemm(n, n, k, 1.0, t1, n, w->g1, n, 1, 0.0, w->Sy, n);
  gemm(n, nx, nx, 1.0, w->g1 + (size_t)n * k, n, w->Se, nx, 0, 0.0, t1, n);
  gemm(n, n, nx, 1.0, t1, n, w->g1 + (size_t)n * k, n, 1, 1.0, w->Sy, n);
  return ST_OK;
}

/* dense Kalman filter (kalman_likelihood): returns status, *ll */
static int kalman(const oracle_model* M, ws_t* w, const double* Y, int nobs, int presample, double* ll) {
  const int n = M->n, nx = M->nx, ns = M->ns, ny = M->ny, k = M->k;
  double *T = w->T, *R = w->R, *QQ = w->QQ, *P = w->P, *a = w->a;
  memset(T, 0, sizeof(double) * ns * ns);
  for (int j = 0; j < k; ++j) for (int i = 0; i < ns; ++i) A_(T, ns, i, M->lagged_state_ids[j]) = A_(w->g1, n, M->state_ids[i], j);
  for (int e = 0; e < nx; ++e) for (int i = 0; i < ns; ++i) A_(R, ns, i, e) = A_(w->g1, n, M->state_ids[i], k + e);
  gemm(ns, nx, nx, 1.0, R, ns, w->Se, nx, 0, 0.0, w->tmp, ns);
  gemm(ns, ns, nx, 1.0, w->tmp, ns, R, ns, 1, 0.0, QQ, ns);
  for (int j = 0; j < ns; ++j) for (int i = 0; i < ns; ++i) A_(P, ns, i, j) = A_(w->Sy, n, M->state_ids[i], M->state_ids[j]);
  memset(a, 0, sizeof(double) * ns);
  double acc = 0; int nseen = 0;
  int* obs = (int*)malloc(sizeof(int) * ny);
  for (int t = 0; t < nobs; ++t) {
    int no = 0;
    for (int i = 0; i < ny; ++i) if (!isnan(Y[i + (size_t)ny * t])) obs[no++] = i;
    if (no > 0) {
      double *F = w->F, *PZ = w->PZ, *K = w->K, *v = w->v;
      for (int q = 0; q < no; ++q) {
        int i = obs[q]; int zi = M->obs_idx_state[i];
        v[q] = Y[i + (size_t)ny * t] - w->y[M->state_ids[zi]] - a[zi];
        for (int r = 0; r < ns; ++r) A_(PZ, ns, r, q) = A_(P, ns, r, zi);
      }
      for (int q = 0; q < no; ++q) for (int r = 0; r < no; ++r)
        A_(F, no, r, q) = A_(PZ, ns, M->obs_idx_state[obs[r]], q) + A_(w->Sm, ny, obs[r], obs[q]);
      /* Cholesky F = L L' */
      for (int j = 0; j < no; ++j) {
        double d = A_(F, no, j, j);
        for (int q = 0; q < j; ++q) d -= A_(F, no, j, q) * A_(F, no, j, q);
        if (!(d > 0.0)) { free(obs); return ST_F_NOT_PD; }
        d = sqrt(d); A_(F, no, j, j) = d;
        for (int i = j + 1; i < no; ++i) {
          double x = A_(F, no, i, j);
          for (int q = 0; q < j; ++q) x -= A_(F, no, i, q) * A_(F, no, j, q);
          A_(F, no, i, j) = x / d;
        }
      }
      double* wv = w->tmp2;
      double quad = 0, ld = 0;
      for (int i = 0; i < no; ++i) {
        double x = v[i]; for (int q = 0; q < i; ++q) x -= A_(F, no, i, q) * wv[q];
        wv[i] = x / A_(F, no, i, i); quad += wv[i] * wv[i]; ld += log(A_(F, no, i, i));
      }
      if (t >= presample) { acc += 2 * ld + quad; nseen += no; }
      /* U = L^-1 (Z P) stored as K (ns x no): K[r,:] = solve */
      for (int r = 0; r < ns; ++r) {
        for (int i = 0; i < no; ++i) {
          double x = A_(PZ, ns, r, i); for (int q = 0; q < i; ++q) x -= A_(F, no, i, q) * A_(K, ns, r, q);
          A_(K, ns, r, i) = x / A_(F, no, i, i);
        }
      }
      for (int r = 0; r < ns; ++r) { double x = a[r]; for (int i = 0; i < no; ++i) x += A_(K, ns, r, i) * wv[i]; a[r] = x; }
      gemm(ns, ns, no, -1.0, K, ns, K, ns, 1, 1.0, P, ns);
    }
    /* a = T a ; P = T P T' + QQ */
    double* t1 = w->tmp;
    for (int i = 0; i < ns; ++i) { double x = 0; for (int q = 0; q < ns; ++q) x += A_(T, ns, i, q) * a[q]; t1[i] = x; }
    memcpy(a, t1, sizeof(double) * ns);
    gemm(ns, ns, ns, 1.0, T, ns, P, ns, 0, 0.0, t1, ns);
    memcpy(P, QQ, sizeof(double) * ns * ns);
    gemm(ns, ns, ns, 1.0, t1, ns, T, ns, 1, 1.0, P, ns);
  }
  free(obs);
  *ll = -0.5 * (nseen * 1.8378770664093453 + acc);
  return isfinite(*ll) ? ST_OK : ST_F_NOT_PD;
}

/* ---- exported ------------------------------------------------------------------------------------ */
int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* theta: np x N column-major; np <= 0 uses the model file's estimated-parameter map.
 * g1_out (n x (k+nx) per draw) may be NULL.  Returns the number of threads used, < 0 on error. */
int oracle_loglik_batch(const char* model, const double* theta, long n_draws, int np,
                        const int* est_type, const int* est_i, const int* est_j,
                        const double* Y, int ny, int nobs, int presample, double cr_tol,
                        double* loglik, int* status, double* g1_out, int nthreads) {
  const oracle_model* M = find_model(model);
  if (!M || ny != M->ny) return -1;
  if (np <= 0) { np = M->np_default; est_type = M->est_type; est_i = M->est_i; est_j = M->est_j; }
  opts_t o = {cr_tol > 0 ? cr_tol : 1e-8, 100, cbrt(2.220446049250313e-16), 1e-16, 60};
  int used = 1;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  used = nthreads;
#pragma omp parallel num_threads(nthreads)
#endif
  {
    ws_t w; ws_init(&w, M, nobs);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
    for (long d = 0; d < n_draws; ++d) {
      double ll = -INFINITY;
      memset(w.g1, 0, sizeof(double) * M->n * (M->k + M->nx));
      int st = solve_draw(M, &o, np, est_type, est_i, est_j, theta + (size_t)np * d, &w);
      if (st == ST_OK && g1_out) memcpy(g1_out + (size_t)M->n * (M->k + M->nx) * d, w.g1, sizeof(double) * M->n * (M->k + M->nx));
      if (st == ST_OK) st = variance(M, &o, &w);
      if (st == ST_OK) st = kalman(M, &w, Y, nobs, presample, &ll);
      if (st != ST_OK) ll = -INFINITY;
      loglik[d] = ll; status[d] = st;
    }
    ws_free(&w);
  }
  return used;
}
