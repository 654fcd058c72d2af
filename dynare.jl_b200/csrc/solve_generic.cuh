// Run-time-size first-order solve for models WITHOUT a generated device function: the host (Julia) evaluates
// steady state and Jacobian per draw (src/steady_state/SteadyState.jl:183-252, src/model.jl:120-139,
// src/perturbations.jl:616-635) and uploads the compressed Jacobians; this kernel does everything after that
// with every matrix of the draw in shared memory, one warp per draw (small models) or one CTA per draw:
//   static elimination (Householder) -> cyclic reduction + Blanchard-Kahn status -> g_u -> static rows
//   -> doubling Lyapunov -> dense state space (T, R Q R', P0, ybar[varobs], H) for the dense filter kernels.
// Same algorithm and failure semantics as solve_kernel.cuh (one thread per draw, compile-time sizes); the loops
// are distributed over the threads of the group.  Reference stages replaced: LRE.first_order_solver!
// (src/perturbations.jl:663-664), compute_variance! (src/estimation/estimation.jl:630), state-space assembly
// (estimation.jl:640-658).  Also hosts the stand-alone batched doubling Lyapunov solver (dyb_lyapunov_batch).
#pragma once
#include "model_traits.cuh"

namespace dyb {

struct GenericModel {
  int n, nx, ny, s, k, nf, m, ncol, ns;
  const int* jperm;          // ncol: compressed Jacobian column -> working column [static | lags | dyn | leads | shocks]
  const int* bkwrd_in_dyn;   // k: position of backward variable j among the dynamic variables
  const int* fwrd_in_dyn;    // nf
  const int* var_static;     // n: 1 if static
  const int* var_pos;        // n: position among the static / the dynamic variables
  const int* state_ids;      // ns: variable of each Kalman state
  const int* obs_var;        // ny: observed variable of each row of Y
  const int* lagged_state;   // k: Kalman state holding backward variable j
};

struct GenericSolveArgs {
  const double* jac;         // n x ncol per draw, column-major, compressed column order
  const double* ybar;        // n per draw (steady state), may be null (=> ybar_obs = 0)
  const double* sigma_e;     // nx x nx per draw (stride_se = nx*nx) or shared by all draws (stride_se = 0)
  const double* sigma_m;     // ny x ny per draw / shared / null (= 0)
  long long stride_se, stride_sm;
  const int* status_in;      // optional: draws already failed on the host (status != 0) are skipped
  long long n_draws;
  double cr_tol; int cr_maxit; double lyap_tol; int lyap_maxit;
  // outputs (any of the dense ones may be null)
  double* T; double* RQR; double* P0; double* ybar_obs; double* H;   // per draw ns*ns, ns*ns, ns*ns, ny, ny*ny
  double* g1;                // n x (k+nx) per draw
  int* status; int* cr_iters; int* lyap_iters;
  double* wscratch;          // solve_block_kernel only: SbLayout::wdoubles per CTA (set by the launcher)
};

// A draw is worked on by a GROUP of GSZ threads: one warp (GSZ = 32, several draws per CTA, __syncwarp between phases)
// for small models, the whole CTA (GSZ = blockDim, __syncthreads) for large ones.
template <int GSZ>
struct Grp {
  int tid, nt;
  DYB_D Grp() : tid(GSZ == 32 ? (threadIdx.x & 31) : threadIdx.x), nt(GSZ == 32 ? 32 : blockDim.x) {}
  DYB_D void sync() const { if (GSZ == 32) __syncwarp(); else __syncthreads(); }
};

// ---- group-cooperative dense helpers (column-major, all threads of the group must call) ---------------------
template <int GSZ>
__device__ __noinline__ void blk_lu_factor(double* a, int n, int* piv, int* okflag) {
  const Grp<GSZ> gp;
  const int tid = gp.tid, nt = gp.nt;
  #pragma unroll 1
  for (int p = 0; p < n; ++p) {
    if (tid == 0) {
      int r = p;
      double best = fabs(a[p + n * p]);
      #pragma unroll 1
      for (int i = p + 1; i < n; ++i) { const double v = fabs(a[i + n * p]); if (v > best) { best = v; r = i; } }
      piv[p] = r;
      if (!(best > 0.0) || !isfinite(best)) *okflag = 0;
    }
    gp.sync();
    if (!*okflag) return;
    const int r = piv[p];
    if (r != p) for (int c = tid; c < n; c += nt) { const double t = a[p + n * c]; a[p + n * c] = a[r + n * c]; a[r + n * c] = t; }
    gp.sync();
    const double inv = 1.0 / a[p + n * p];
    #pragma unroll 1
    for (int i = p + 1 + tid; i < n; i += nt) a[i + n * p] *= inv;
    gp.sync();
    const int w = n - p - 1;
    #pragma unroll 1
    for (int e = tid; e < w * w; e += nt) {
      const int i = p + 1 + e % w, c = p + 1 + e / w;
      a[i + n * c] = fma(-a[i + n * p], a[p + n * c], a[i + n * c]);
    }
    gp.sync();
  }
}

// solve L U x = P b for nrhs column-major right-hand sides in place: one thread per column
template <int GSZ>
__device__ __noinline__ void blk_lu_solve(const double* lu, const int* piv, double* b, int n, int nrhs) {
  const Grp<GSZ> gp;
  #pragma unroll 1
  for (int c = gp.tid; c < nrhs; c += gp.nt) {
    double* x = b + n * c;
    #pragma unroll 1
    for (int p = 0; p < n; ++p) { const int r = piv[p]; if (r != p) { const double t = x[p]; x[p] = x[r]; x[r] = t; } }
    #pragma unroll 1
    for (int p = 0; p < n; ++p) { const double xp = x[p]; for (int i = p + 1; i < n; ++i) x[i] = fma(-lu[i + n * p], xp, x[i]); }
    #pragma unroll 1
    for (int p = n - 1; p >= 0; --p) {
      const double xp = x[p] / lu[p + n * p];
      x[p] = xp;
      #pragma unroll 1
      for (int i = 0; i < p; ++i) x[i] = fma(-lu[i + n * p], xp, x[i]);
    }
  }
  gp.sync();
}

// doubling solve of X = A X A' + B (k x k): on entry Ak = A, Sg = B; on exit Sg = X (symmetrised).
// red: >= 4 doubles of shared scratch.  Returns 1 on convergence (same stopping rule as solve_kernel.cuh).
template <int GSZ>
__device__ __noinline__ int blk_lyapunov(double* Ak, double* Sg, double* Tm, int k, double tol, int maxit, double* red, int* iters) {
  const Grp<GSZ> gp;
  const int tid = gp.tid, nt = gp.nt;
  int conv = 0, done = 0;
  #pragma unroll 1
  for (int li = 0; li < maxit; ++li) {
    done = li + 1;
    #pragma unroll 1
    for (int e = tid; e < k * k; e += nt) {
      const int i = e % k, j = e / k;
      double s = 0.0;
      #pragma unroll 1
      for (int q = 0; q < k; ++q) s = fma(Ak[i + k * q], Sg[q + k * j], s);
      Tm[e] = s;
    }
    if (tid == 0) { red[0] = 0.0; red[1] = 0.0; red[2] = 0.0; }
    gp.sync();
    double dmax = 0.0, smax = 0.0;
    bool nan_seen = false;
    #pragma unroll 1
    for (int e = tid; e < k * k; e += nt) {
      const int i = e % k, j = e / k;
      double s = 0.0;
      #pragma unroll 1
      for (int q = 0; q < k; ++q) s = fma(Tm[i + k * q], Ak[j + k * q], s);
      const double v = Sg[e] + s;
      Sg[e] = v;
      dmax = fmax(dmax, fabs(s)); smax = fmax(smax, fabs(v));
      nan_seen = nan_seen || !isfinite(v);
    }
    // non-negative doubles order like their bit patterns: atomicMax on the integer view
    atomicMax(reinterpret_cast<unsigned long long*>(&red[0]), (unsigned long long)__double_as_longlong(dmax));
    atomicMax(reinterpret_cast<unsigned long long*>(&red[1]), (unsigned long long)__double_as_longlong(nan_seen ? INFINITY : smax));
    gp.sync();
    #pragma unroll 1
    for (int e = tid; e < k * k; e += nt) {
      const int i = e % k, j = e / k;
      double s = 0.0;
      #pragma unroll 1
      for (int q = 0; q < k; ++q) s = fma(Ak[i + k * q], Ak[q + k * j], s);
      Tm[e] = s;
    }
    gp.sync();
    double amax = 0.0;
    #pragma unroll 1
    for (int e = tid; e < k * k; e += nt) { const double v = Tm[e]; Ak[e] = v; amax = fmax(amax, isfinite(v) ? fabs(v) : INFINITY); }
    atomicMax(reinterpret_cast<unsigned long long*>(&red[2]), (unsigned long long)__double_as_longlong(amax));
    gp.sync();
    const double D = red[0], S = red[1], A = red[2];
    gp.sync();
    if (!isfinite(S)) break;
    if (D <= tol * S && A < 1.0) { conv = 1; break; }
    if (li == LYAP_UNIT_ROOT_IT && A >= LYAP_UNIT_ROOT_NORM) break;         // a root within 1e-6 of the unit circle
    if (!isfinite(A)) break;
  }
  if (iters) *iters = done;
  if (conv) {
    #pragma unroll 1
    for (int e = tid; e < k * k; e += nt) {
      const int i = e % k, j = e / k;
      if (i > j) Tm[e] = 0.5 * (Sg[i + k * j] + Sg[j + k * i]);
    }
    gp.sync();
    #pragma unroll 1
    for (int e = tid; e < k * k; e += nt) {
      const int i = e % k, j = e / k;
      if (i > j) { Sg[i + k * j] = Tm[e]; Sg[j + k * i] = Tm[e]; }
    }
    gp.sync();
  }
  return conv;
}

// shared-memory footprint (doubles) of solve_generic_kernel for a model
inline size_t generic_solve_smem_doubles(int n, int nx, int ny, int s, int k, int nf, int ns) {
  const size_t m = (size_t)(n - s), ncol = (size_t)(s + k) + m + nf + nx;
  size_t d = 0;
  d += (size_t)n * ncol;                 // W
  d += m * k + m * nf + 2 * m * m;       // A0, A2, A1, LU
  d += m * (k + nf) + m * k;             // X, acc_hat
  d += m * k + m * nx;                   // Gd, gud
  d += (size_t)(s > 0 ? s : 1) * (k + nx) + (size_t)nf * (k + nx);   // gst, GG
  d += 3 * (size_t)k * k + 2 * (size_t)k * nx;                       // Ak, Sg, Tm, Bs, BQ
  d += 2 * (size_t)ns * k + 2 * (size_t)ns * nx;                     // g1s, GS, g2s, RQ
  d += (size_t)nx * nx + (size_t)(ncol > 8 ? ncol : 8);              // Se, red
  d += (m + 1) / 2 + 1;                                              // piv (ints)
  d += 8;                                                            // per-group scalars (flags, norms)
  return d;
}

template <int GSZ>
__global__ void __launch_bounds__(128)
solve_generic_kernel(GenericModel gm, GenericSolveArgs a, int per_draw_doubles) {
  extern __shared__ double sm_all[];
  const Grp<GSZ> gp;
  const int ngroups = GSZ == 32 ? (int)(blockDim.x >> 5) : 1, grp = GSZ == 32 ? (int)(threadIdx.x >> 5) : 0;
  double* sm = sm_all + (size_t)grp * per_draw_doubles;
  const int n = gm.n, nx = gm.nx, ny = gm.ny, S = gm.s, K = gm.k, NF = gm.nf, MM = gm.m, NCOL = gm.ncol, NS = gm.ns;
  const int cC = S, cB = S + K, cA = S + K + MM, cU = S + K + MM + NF;
  const int tid = gp.tid, nt = gp.nt;
  double* W = sm;
  double* A0 = W + (size_t)n * NCOL;
  double* A2 = A0 + MM * K;
  double* A1 = A2 + MM * NF;
  double* LU = A1 + MM * MM;
  double* X = LU + MM * MM;
  double* acc_hat = X + MM * (K + NF);
  double* Gd = acc_hat + MM * K;
  double* gud = Gd + MM * K;
  double* gst = gud + MM * nx;
  double* GG = gst + (S > 0 ? S : 1) * (K + nx);
  double* Ak = GG + NF * (K + nx);
  double* Sg = Ak + K * K;
  double* Tm = Sg + K * K;
  double* Bs = Tm + K * K;
  double* BQ = Bs + K * nx;
  double* g1s = BQ + K * nx;
  double* GS = g1s + NS * K;
  double* g2s = GS + NS * K;
  double* RQ = g2s + NS * nx;
  double* Se = RQ + NS * nx;
  double* red = Se + nx * nx;
  int* piv = reinterpret_cast<int*>(red + (NCOL > 8 ? NCOL : 8));
  double* scal = reinterpret_cast<double*>(piv) + (MM + 1) / 2 + 1;       // per-group scalars
  int& okflag = *reinterpret_cast<int*>(scal);
  double& n0_sh = scal[1];
  double& n2_sh = scal[2];

  #pragma unroll 1
  for (long long d = (long long)blockIdx.x * ngroups + grp; d < a.n_draws; d += (long long)gridDim.x * ngroups) {
    gp.sync();
    if (a.status_in && a.status_in[d] != ST_OK) {
      if (tid == 0) { a.status[d] = a.status_in[d]; if (a.cr_iters) a.cr_iters[d] = 0; if (a.lyap_iters) a.lyap_iters[d] = 0; }
      continue;
    }
    int status = ST_OK;
    const double* J = a.jac + (size_t)d * n * NCOL;
    // ---- Jacobian into the working column order; NaN / Inf anywhere -> steady-state class failure -------------
    if (tid == 0) { okflag = 1; red[0] = 0.0; }
    gp.sync();
    double chk = 0.0;
    #pragma unroll 1
    for (int e = tid; e < n * NCOL; e += nt) {
      const double v = J[e];
      chk = fma(v, 0.0, chk);
      W[(e % n) + n * gm.jperm[e / n]] = v;
    }
    #pragma unroll 1
    for (int e = tid; e < nx * nx; e += nt) { const double v = a.sigma_e[(size_t)d * a.stride_se + e]; Se[e] = v; chk = fma(v, 0.0, chk); }
    if (!(chk == 0.0)) okflag = 0;
    gp.sync();
    if (!okflag) status = ST_STEADY;

    // ---- Householder QR on the static columns, applied to every other column ------------------------------
    #pragma unroll 1
    for (int j = 0; j < S && status == ST_OK; ++j) {
      double* col = W + n * j;
      if (tid == 0) {
        double nrm2 = 0.0;
        #pragma unroll 1
        for (int i = j; i < n; ++i) nrm2 = fma(col[i], col[i], nrm2);
        red[0] = nrm2;
      }
      gp.sync();
      const double nrm2 = red[0];
      const double nrm = sqrt(nrm2);
      if (!(nrm > 0.0) || !isfinite(nrm)) { status = ST_SOLVER; break; }      // uniform: every thread reads red[0]
      const double alpha = col[j] >= 0.0 ? -nrm : nrm;
      const double vj = col[j] - alpha;
      const double beta = 2.0 / (2.0 * (nrm2 - alpha * col[j]));
      #pragma unroll 1
      for (int c = j + 1 + tid; c < NCOL; c += nt) {
        double* x = W + n * c;
        double dot = vj * x[j];
        #pragma unroll 1
        for (int i = j + 1; i < n; ++i) dot = fma(col[i], x[i], dot);
        dot *= beta;
        x[j] = fma(-dot, vj, x[j]);
        #pragma unroll 1
        for (int i = j + 1; i < n; ++i) x[i] = fma(-dot, col[i], x[i]);
      }
      gp.sync();
      if (tid == 0) { col[j] = alpha; for (int i = j + 1; i < n; ++i) col[i] = 0.0; }
      gp.sync();
    }

    // ---- cyclic reduction on the MM x MM dynamic block --------------------------------------------------------
    int it = 0;
    if (status == ST_OK) {
      #pragma unroll 1
      for (int e = tid; e < MM * K; e += nt) { const int r = e % MM, j = e / MM; A0[e] = W[S + r + n * (cC + j)]; acc_hat[e] = 0.0; }
      #pragma unroll 1
      for (int e = tid; e < MM * NF; e += nt) { const int r = e % MM, j = e / MM; A2[e] = W[S + r + n * (cA + j)]; }
      #pragma unroll 1
      for (int e = tid; e < MM * MM; e += nt) { const int r = e % MM, c = e / MM; A1[e] = W[S + r + n * (cB + c)]; }
      if (tid == 0) { n0_sh = 0.0; n2_sh = 0.0; }
      gp.sync();
      bool conv = false, stalled = false;
      double n0 = 0.0, n2 = 0.0;
      #pragma unroll 1
      for (it = 1; it <= a.cr_maxit; ++it) {
        #pragma unroll 1
        for (int e = tid; e < MM * MM; e += nt) LU[e] = A1[e];
        #pragma unroll 1
        for (int e = tid; e < MM * K; e += nt) X[e] = A0[e];
        #pragma unroll 1
        for (int e = tid; e < MM * NF; e += nt) X[MM * K + e] = A2[e];
        if (tid == 0) okflag = 1;
        gp.sync();
        blk_lu_factor<GSZ>(LU, MM, piv, &okflag);
        gp.sync();
        if (!okflag) { if (it == 1) status = ST_SOLVER; else stalled = true; break; }
        blk_lu_solve<GSZ>(LU, piv, X, MM, K + NF);
        const double* X0 = X;
        const double* X2 = X + MM * K;
        // A1 forward columns -= A0 X2[bk,:] ; A1 backward columns -= A2 X0[fw,:] (also accumulated in acc_hat)
        #pragma unroll 1
        for (int e = tid; e < MM * NF; e += nt) {
          const int r = e % MM, j = e / MM;
          double s = 0.0;
          #pragma unroll 1
          for (int i = 0; i < K; ++i) s = fma(A0[r + MM * i], X2[gm.bkwrd_in_dyn[i] + MM * j], s);
          LU[e] = s;                                     // LU is free after the solve: staging
        }
        gp.sync();
        #pragma unroll 1
        for (int e = tid; e < MM * NF; e += nt) { const int r = e % MM, j = e / MM; A1[r + MM * gm.fwrd_in_dyn[j]] -= LU[e]; }
        gp.sync();
        #pragma unroll 1
        for (int e = tid; e < MM * K; e += nt) {
          const int r = e % MM, j = e / MM;
          double s = 0.0;
          #pragma unroll 1
          for (int i = 0; i < NF; ++i) s = fma(A2[r + MM * i], X0[gm.fwrd_in_dyn[i] + MM * j], s);
          LU[e] = s;
        }
        gp.sync();
        #pragma unroll 1
        for (int e = tid; e < MM * K; e += nt) { const int r = e % MM, j = e / MM; A1[r + MM * gm.bkwrd_in_dyn[j]] -= LU[e]; acc_hat[e] += LU[e]; }
        gp.sync();
        // A0 <- -A0 X0[bk,:] ; A2 <- -A2 X2[fw,:]
        #pragma unroll 1
        for (int e = tid; e < MM * K; e += nt) {
          const int r = e % MM, j = e / MM;
          double s = 0.0;
          #pragma unroll 1
          for (int i = 0; i < K; ++i) s = fma(A0[r + MM * i], X0[gm.bkwrd_in_dyn[i] + MM * j], s);
          LU[e] = -s;
        }
        gp.sync();
        #pragma unroll 1
        for (int e = tid; e < MM * K; e += nt) A0[e] = LU[e];
        gp.sync();
        #pragma unroll 1
        for (int e = tid; e < MM * NF; e += nt) {
          const int r = e % MM, j = e / MM;
          double s = 0.0;
          #pragma unroll 1
          for (int i = 0; i < NF; ++i) s = fma(A2[r + MM * i], X2[gm.fwrd_in_dyn[i] + MM * j], s);
          LU[e] = -s;
        }
        gp.sync();
        #pragma unroll 1
        for (int e = tid; e < MM * NF; e += nt) A2[e] = LU[e];
        gp.sync();
        // 1-norms (largest column sum), sequential per column so the value does not depend on the launch geometry
        #pragma unroll 1
        for (int j = tid; j < K + NF; j += nt) {
          const double* c = j < K ? A0 + MM * j : A2 + MM * (j - K);
          double cs = 0.0;
          #pragma unroll 1
          for (int r = 0; r < MM; ++r) cs += fabs(c[r]);
          red[j] = cs;
        }
        gp.sync();
        if (tid == 0) {
          double m0 = 0.0, m2 = 0.0;
          #pragma unroll 1
          for (int j = 0; j < K; ++j) m0 = fmax(m0, red[j]);
          #pragma unroll 1
          for (int j = 0; j < NF; ++j) m2 = fmax(m2, red[K + j]);
          bool bad = false;
          #pragma unroll 1
          for (int j = 0; j < K + NF; ++j) bad = bad || !(red[j] <= 1e150);
          n0_sh = bad ? NAN : m0; n2_sh = m2;
        }
        gp.sync();
        const double m0 = n0_sh, m2 = n2_sh;
        gp.sync();
        if (!(m0 <= 1e150) || !(m2 <= 1e150)) { if (it == 1) status = ST_SOLVER; else stalled = true; break; }
        const double n0_prev = n0;
        n0 = m0; n2 = m2;
        if ((n0 < a.cr_tol && n2 < a.cr_tol) || (n2 < a.cr_tol * a.cr_tol && fabs(n0 - n0_prev) <= CR_UNIT_ROOT_EPS * n0)) { conv = true; break; }
      }
      if (status == ST_OK && !conv) {
        if (stalled) status = (n0 < n2) ? ST_INDETERMINATE : ST_UNSTABLE;
        else status = (n0 >= a.cr_tol) ? ST_UNSTABLE : ST_INDETERMINATE;
      }
    }
    gp.sync();

    if (status == ST_OK) {
      // ---- G (dynamic rows, backward columns) = -A1hat^-1 A0(0) ------------------------------------------------
      #pragma unroll 1
      for (int e = tid; e < MM * MM; e += nt) { const int r = e % MM, c = e / MM; LU[e] = W[S + r + n * (cB + c)]; }
      gp.sync();
      #pragma unroll 1
      for (int e = tid; e < MM * K; e += nt) { const int r = e % MM, j = e / MM; LU[r + MM * gm.bkwrd_in_dyn[j]] -= acc_hat[e]; Gd[e] = -W[S + r + n * (cC + j)]; }
      if (tid == 0) okflag = 1;
      gp.sync();
      blk_lu_factor<GSZ>(LU, MM, piv, &okflag);
      gp.sync();
      if (!okflag) status = ST_SOLVER;
      else blk_lu_solve<GSZ>(LU, piv, Gd, MM, K);
    }
    gp.sync();
    if (status == ST_OK) {
      // ---- g_u = -(A G + B)^-1 F_u on the dynamic block ---------------------------------------------------------
      #pragma unroll 1
      for (int e = tid; e < MM * MM; e += nt) { const int r = e % MM, c = e / MM; LU[e] = W[S + r + n * (cB + c)]; }
      gp.sync();
      #pragma unroll 1
      for (int e = tid; e < MM * K; e += nt) {
        const int r = e % MM, j = e / MM;
        double s = 0.0;
        #pragma unroll 1
        for (int i = 0; i < NF; ++i) s = fma(W[S + r + n * (cA + i)], Gd[gm.fwrd_in_dyn[i] + MM * j], s);
        LU[r + MM * gm.bkwrd_in_dyn[j]] += s;
      }
      #pragma unroll 1
      for (int e = tid; e < MM * nx; e += nt) { const int r = e % MM, q = e / MM; gud[e] = -W[S + r + n * (cU + q)]; }
      if (tid == 0) okflag = 1;
      gp.sync();
      blk_lu_factor<GSZ>(LU, MM, piv, &okflag);
      gp.sync();
      if (!okflag) status = ST_SOLVER;
      else blk_lu_solve<GSZ>(LU, piv, gud, MM, nx);
    }
    gp.sync();
    if (status == ST_OK && S > 0) {
      // ---- static rows by back-substitution: R_s y_s + B_top y_d + A_top y_f(+1) + C_top y_b(-1) + Fu_top u = 0 ---
      #pragma unroll 1
      for (int e = tid; e < NF * (K + nx); e += nt) {
        const int i = e % NF, c = e / NF;
        double s = 0.0;
        #pragma unroll 1
        for (int q = 0; q < K; ++q) {
          const double g = (c < K) ? Gd[gm.bkwrd_in_dyn[q] + MM * c] : gud[gm.bkwrd_in_dyn[q] + MM * (c - K)];
          s = fma(Gd[gm.fwrd_in_dyn[i] + MM * q], g, s);
        }
        GG[e] = s;
      }
      if (tid == 0) okflag = 1;
      gp.sync();
      #pragma unroll 1
      for (int c = tid; c < K + nx; c += nt) {
        const double* gcol = (c < K) ? (Gd + MM * c) : (gud + MM * (c - K));
        #pragma unroll 1
        for (int q = 0; q < S; ++q) {
          double s = (c < K) ? W[q + n * (cC + c)] : W[q + n * (cU + (c - K))];
          #pragma unroll 1
          for (int dd = 0; dd < MM; ++dd) s = fma(W[q + n * (cB + dd)], gcol[dd], s);
          #pragma unroll 1
          for (int i = 0; i < NF; ++i) s = fma(W[q + n * (cA + i)], GG[i + NF * c], s);
          gst[q + S * c] = -s;
        }
        #pragma unroll 1
        for (int q = S - 1; q >= 0; --q) {
          double s = gst[q + S * c];
          #pragma unroll 1
          for (int r = q + 1; r < S; ++r) s = fma(-W[q + n * r], gst[r + S * c], s);
          const double dg = W[q + n * q];
          if (dg == 0.0) okflag = 0;
          gst[q + S * c] = s / dg;
        }
      }
      gp.sync();
      if (!okflag) status = ST_SOLVER;
    }
    gp.sync();
    if (status == ST_OK && a.g1) {
      double* g1 = a.g1 + (size_t)d * n * (K + nx);
      #pragma unroll 1
      for (int e = tid; e < n * (K + nx); e += nt) {
        const int v = e % n, c = e / n;
        const int pos = gm.var_pos[v];
        g1[e] = gm.var_static[v] ? gst[pos + S * c] : ((c < K) ? Gd[pos + MM * c] : gud[pos + MM * (c - K)]);
      }
    }

    // ---- Lyapunov doubling on the K x K predetermined block -----------------------------------------------------
    int lyap_it = 0;
    if (status == ST_OK) {
      #pragma unroll 1
      for (int e = tid; e < K * K; e += nt) { const int i = e % K, j = e / K; Ak[e] = Gd[gm.bkwrd_in_dyn[i] + MM * j]; }
      #pragma unroll 1
      for (int e = tid; e < K * nx; e += nt) { const int i = e % K, q = e / K; Bs[e] = gud[gm.bkwrd_in_dyn[i] + MM * q]; }
      gp.sync();
      #pragma unroll 1
      for (int e = tid; e < K * nx; e += nt) {
        const int i = e % K, q = e / K;
        double s = 0.0;
        #pragma unroll 1
        for (int f = 0; f < nx; ++f) s = fma(Bs[i + K * f], Se[f + nx * q], s);
        BQ[e] = s;
      }
      gp.sync();
      #pragma unroll 1
      for (int e = tid; e < K * K; e += nt) {
        const int i = e % K, j = e / K;
        double s = 0.0;
        #pragma unroll 1
        for (int q = 0; q < nx; ++q) s = fma(BQ[i + K * q], Bs[j + K * q], s);
        Sg[e] = s;
      }
      gp.sync();
      if (!blk_lyapunov<GSZ>(Ak, Sg, Tm, K, a.lyap_tol, a.lyap_maxit, red, &lyap_it)) status = ST_NONSTATIONARY;
    }
    gp.sync();

    // ---- dense state space (estimation.jl:640-658) -----------------------------------------------------------------
    if (status == ST_OK) {
      #pragma unroll 1
      for (int e = tid; e < NS * (K + nx); e += nt) {
        const int sI = e % NS, c = e / NS;
        const int v = gm.state_ids[sI];
        const int pos = gm.var_pos[v];
        const double val = gm.var_static[v] ? gst[pos + S * c] : ((c < K) ? Gd[pos + MM * c] : gud[pos + MM * (c - K)]);
        if (c < K) g1s[sI + NS * c] = val; else g2s[sI + NS * (c - K)] = val;
      }
      gp.sync();
      #pragma unroll 1
      for (int e = tid; e < NS * nx; e += nt) {
        const int sI = e % NS, q = e / NS;
        double s = 0.0;
        #pragma unroll 1
        for (int f = 0; f < nx; ++f) s = fma(g2s[sI + NS * f], Se[f + nx * q], s);
        RQ[e] = s;
      }
      #pragma unroll 1
      for (int e = tid; e < NS * K; e += nt) {
        const int sI = e % NS, j = e / NS;
        double s = 0.0;
        #pragma unroll 1
        for (int q = 0; q < K; ++q) s = fma(g1s[sI + NS * q], Sg[q + K * j], s);
        GS[e] = s;
      }
      gp.sync();
      double* To = a.T ? a.T + (size_t)d * NS * NS : nullptr;
      double* Qo = a.RQR ? a.RQR + (size_t)d * NS * NS : nullptr;
      double* Po = a.P0 ? a.P0 + (size_t)d * NS * NS : nullptr;
      if (To) for (int e = tid; e < NS * NS; e += nt) To[e] = 0.0;
      gp.sync();
      if (To) for (int e = tid; e < NS * K; e += nt) { const int sI = e % NS, j = e / NS; To[sI + NS * gm.lagged_state[j]] = g1s[e]; }
      #pragma unroll 1
      for (int e = tid; e < NS * NS; e += nt) {
        const int sI = e % NS, t = e / NS;
        if (sI < t) continue;
        double qq = 0.0;
        #pragma unroll 1
        for (int q = 0; q < nx; ++q) qq = fma(RQ[sI + NS * q], g2s[t + NS * q], qq);
        double pp = qq;
        #pragma unroll 1
        for (int j = 0; j < K; ++j) pp = fma(GS[sI + NS * j], g1s[t + NS * j], pp);
        if (Qo) { Qo[sI + NS * t] = qq; Qo[t + NS * sI] = qq; }
        if (Po) { Po[sI + NS * t] = pp; Po[t + NS * sI] = pp; }
      }
      if (a.ybar_obs) for (int i = tid; i < ny; i += nt) a.ybar_obs[(size_t)d * ny + i] = a.ybar ? a.ybar[(size_t)d * n + gm.obs_var[i]] : 0.0;
      if (a.H) for (int e = tid; e < ny * ny; e += nt) a.H[(size_t)d * ny * ny + e] = a.sigma_m ? a.sigma_m[(size_t)d * a.stride_sm + e] : 0.0;
    }
    if (tid == 0) {
      a.status[d] = status;
      if (a.cr_iters) a.cr_iters[d] = it;
      if (a.lyap_iters) a.lyap_iters[d] = lyap_it;
    }
  }
}

// ---- stand-alone batched doubling Lyapunov: X = A X A' + B, one CTA per draw (compute_variance!'s solver) ---------
struct LyapunovArgs {
  int k; const double* A; const double* B; double* X; int* iters; int* status; long long n_draws; double tol; int maxit;
};
__global__ void __launch_bounds__(128) lyapunov_generic_kernel(LyapunovArgs a) {
  extern __shared__ double sm[];
  const int k = a.k, tid = threadIdx.x, nt = blockDim.x;
  double* Ak = sm; double* Sg = Ak + k * k; double* Tm = Sg + k * k; double* red = Tm + k * k;
  #pragma unroll 1
  for (long long d = blockIdx.x; d < a.n_draws; d += gridDim.x) {
    __syncthreads();
    #pragma unroll 1
    for (int e = tid; e < k * k; e += nt) { Ak[e] = a.A[(size_t)d * k * k + e]; Sg[e] = a.B[(size_t)d * k * k + e]; }
    __syncthreads();
    int it = 0;
    const int conv = blk_lyapunov<128>(Ak, Sg, Tm, k, a.tol, a.maxit, red, &it);
    #pragma unroll 1
    for (int e = tid; e < k * k; e += nt) a.X[(size_t)d * k * k + e] = conv ? Sg[e] : NAN;
    if (tid == 0) { a.status[d] = conv ? ST_OK : ST_NONSTATIONARY; if (a.iters) a.iters[d] = it; }
  }
}

}  // namespace dyb
