// Run-time-size first-order solve for MEDIUM models (dynamic block 12 < m <= 8 RPT, Smets-Wouters scale: BASELINE configs[3]),
// one CTA of 128 threads per draw.  Same stages, stopping rules and failure semantics as solve_generic_kernel
// (solve_generic.cuh), which stays the kernel for small models (one warp per draw) and for anything larger; what changes is
// the mapping of the dense algebra:
//   * the threads form an 8 x 16 grid; thread (ti, tj) owns rows ti + 8a and columns tj + 16b of whatever matrix is being
//     worked on, in REGISTERS (static indices only).  The linear solves of the cyclic reduction, A1^-1 [A0 | A2], are a
//     Gauss-Jordan elimination of the m x (m + k + nf) array with implicit row pivoting: per step one column and one row travel
//     through shared memory (two barriers), the rank-one update of the whole array is register-local (RPT x CPT fused
//     multiply-adds per thread), the pivot search is one REDUX.MAX over packed (magnitude, row) keys per warp.  No row is ever
//     swapped: the solution rows come out permuted and are put back in order when they are stored.
//   * the products of an iteration, A0 X[bk,:] and A2 X[fw,:], and the three products of a Lyapunov doubling step are
//     register-blocked (RPT x 3 / RPT x 2 accumulators per thread, operands broadcast from shared memory with odd leading
//     dimensions => conflict-free).
// Reference stages replaced: LRE.first_order_solver! (src/perturbations.jl:663-664), compute_variance!
// (src/estimation/estimation.jl:630), state-space assembly (estimation.jl:640-658).
#pragma once
#include "solve_generic.cuh"

namespace dyb {

constexpr int SB_THREADS = 128;      // 8 x 16 thread grid
constexpr int SB_GRID = 148 * 6;     // CTAs of a launch (each walks over its draws); sizes the global scratch

struct SbLayout {
  int ldw, ldm, lds, ldk, ldn;                                   // odd leading dimensions
  int oW, oU, oGd, oGud, oGst, oSe, oCol, oRow, oRed, oInt;      // offsets in doubles (oW == oU: one region, three lives)
  int uA1, uA0, uA2, uACC, uX;                                   // inside U while the cyclic reduction runs
  int uWt, uGG;                                                  // inside U for the static rows: top S rows of W, G[fw,:] [G | g_u]
  int uAk, uSg, uTm, uBs, uBQ, uG1s, uGS, uG2s, uRQ;             // inside U afterwards
  int total;                                                     // doubles
  int wdoubles;                                                  // global scratch per CTA: the working Jacobian after the QR
};

DYB_HD int sb_odd(int x) { return x | 1; }

// maxm = 8 RPT, maxc = 16 CPT of the instantiation that will run.
// The working Jacobian W (n x ncol) is in shared memory only while it is read in and reduced by the Householder QR on the
// static columns; it is then written to a per-CTA global scratch (L2-resident: 39 KB per CTA for mc4) and the same region
// holds, in turn, the cyclic-reduction state, the top rows of W for the static back-substitution and the Lyapunov / state-space
// work arrays.  112 -> 66 KB per draw for mc4, i.e. three CTAs per SM instead of two for a kernel that is latency-bound.
inline __host__ __device__ SbLayout sb_layout(int n, int nx, int s, int k, int nf, int ns, int maxm, int maxc) {
  SbLayout L;
  const int m = n - s, ncol = s + k + m + nf + nx;
  L.ldw = sb_odd(n); L.ldm = sb_odd(m); L.lds = sb_odd(s > 0 ? s : 1); L.ldk = sb_odd(k); L.ldn = sb_odd(ns);
  L.wdoubles = L.ldw * ncol;
  int o = 0;
  L.oW = o;
  L.oU = o;
  {
    int u = 0;
    L.uA1 = u; u += L.ldm * m;
    L.uA0 = u; u += L.ldm * k;
    L.uA2 = u; u += L.ldm * nf;
    L.uACC = u; u += L.ldm * k;
    L.uX = u; u += L.ldm * (k + nf);
    const int cr = u;
    u = 0;
    L.uWt = u; u += L.lds * ncol;
    L.uGG = u; u += (nf > 0 ? nf : 1) * (k + nx);
    const int st = u;
    u = 0;
    L.uAk = u; u += L.ldk * k;
    L.uSg = u; u += L.ldk * k;
    L.uTm = u; u += L.ldk * k;
    L.uBs = u; u += L.ldk * nx;
    L.uBQ = u; u += L.ldk * nx;
    L.uG1s = u; u += L.ldn * k;
    L.uGS = u; u += L.ldn * k;
    L.uG2s = u; u += L.ldn * nx;
    L.uRQ = u; u += L.ldn * nx;
    int x = L.wdoubles;
    x = cr > x ? cr : x; x = st > x ? st : x; x = u > x ? u : x;
    o += x;
  }
  L.oGd = o; o += L.ldm * k;
  L.oGud = o; o += L.ldm * nx;
  L.oGst = o; o += L.lds * (k + nx);
  L.oSe = o; o += nx * nx;
  L.oCol = o; o += 2 * maxm;
  L.oRow = o; o += maxc;
  L.oRed = o; o += ncol + 32;
  L.oInt = o; o += (k + nf + m + 1) / 2 + 4;                      // bk, fw, bk_of_dyn (ints)
  L.total = o;
  return L;
}

// non-negative doubles order like their bit patterns; NaN maps to +Inf
DYB_D double sb_warp_max(double v) {
  unsigned long long x = (unsigned long long)__double_as_longlong(v == v ? fabs(v) : INFINITY);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, o);
    x = y > x ? y : x;
  }
  return __longlong_as_double((long long)x);
}

// CTA-wide maximum of up to three non-negative quantities per thread (NaN -> Inf); red: 12 doubles; two barriers
DYB_D void sb_block_max3(double& a, double& b, double& c, double* red, int warp, int lane) {
  a = sb_warp_max(a); b = sb_warp_max(b); c = sb_warp_max(c);
  if (lane == 0) { red[warp] = a; red[4 + warp] = b; red[8 + warp] = c; }
  __syncthreads();
  a = fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
  b = fmax(fmax(red[4], red[5]), fmax(red[6], red[7]));
  c = fmax(fmax(red[8], red[9]), fmax(red[10], red[11]));
  __syncthreads();
}

// Gauss-Jordan elimination with implicit row pivoting of the m x nc array R = [A | B] (A = first m columns), element
// (ti + 8a, tj + 16b) in R[a][b]; rows >= m and columns >= nc must hold zeros.  A pivot row is NOT normalised when it is
// used (its multiplier is simply zero in its own step): it stays a multiple of the solution row through the later
// eliminations and is scaled by rinv[a] = 1 / pivot when it is read out, so the rank-one update of a step is RPT x CPT
// plain multiply-adds per thread without a select.  Columns already eliminated (c <= p) are updated too: they are never
// read again.  On exit the B part of row r, times rinv[a], holds row dest[a] of A^-1 B.  colbuf: 2 x 8 RPT doubles,
// prow: 16 CPT doubles.  Returns false (uniformly) on a zero / non-finite pivot.  Ends with a barrier.
template <int RPT, int CPT>
DYB_D bool sb_gauss_jordan(double (&R)[RPT][CPT], int (&dest)[RPT], double (&rinv)[RPT], int m, int nc, double* colbuf,
                           double* prow, int ti, int tj, int lane) {
  constexpr int MAXM = 8 * RPT;
  unsigned long long used = 0;
  bool ok = true;
#pragma unroll
  for (int a = 0; a < RPT; ++a) { dest[a] = 0; rinv[a] = 0.0; }
  if (tj == 0) {
#pragma unroll
    for (int a = 0; a < RPT; ++a) colbuf[ti + 8 * a] = R[a][0];
  }
  __syncthreads();
#pragma unroll
  for (int b0 = 0; b0 < CPT; ++b0) {
#pragma unroll 1
    for (int pp = 0; pp < 16; ++pp) {
      const int p = 16 * b0 + pp;
      if (p >= m || !ok) break;
      const double* cb = colbuf + (p & 1) * MAXM;
      double* cbn = colbuf + ((p + 1) & 1) * MAXM;
      // pivot row: largest magnitude among the rows not used yet (top 25 bits of |x|, row index in the low 6 bits of the key).
      // Every lane also takes the reciprocal of ITS candidate while the keys are being reduced, so that the winner's
      // reciprocal is one shuffle away instead of a dependent load + MUFU + two Newton steps after the reduction
      unsigned key = 0;
      double cand[(MAXM + 31) / 32], cinv[(MAXM + 31) / 32];
#pragma unroll
      for (int h = 0; h < (MAXM + 31) / 32; ++h) {
        const int r = lane + 32 * h;
        cand[h] = r < m ? cb[r] : 1.0;
        cinv[h] = fast_rcp(cand[h]);
        if (r < m && !((used >> r) & 1ull)) {
          const unsigned hi = (unsigned)__double2hiint(cand[h]) & 0x7fffffffu;
          const unsigned kk = (hi & ~63u) | (unsigned)(63 - r);
          key = kk > key ? kk : key;
        }
      }
      double lraw[RPT];
#pragma unroll
      for (int a = 0; a < RPT; ++a) lraw[a] = -cb[ti + 8 * a];
      key = __reduce_max_sync(0xffffffffu, key);
      const unsigned expo = key >> 20;
      if (expo == 0u || expo >= 0x7ffu) { ok = false; break; }          // zero / subnormal / Inf / NaN pivot column
      const int rs = 63 - (int)(key & 63u);
      double inv = __shfl_sync(0xffffffffu, cinv[0], rs & 31);
#pragma unroll
      for (int h = 1; h < (MAXM + 31) / 32; ++h) { const double o = __shfl_sync(0xffffffffu, cinv[h], rs & 31); if ((rs >> 5) == h) inv = o; }
      used |= 1ull << rs;
      const int as = rs >> 3;                                             // uniform over the CTA
      const bool mine = (ti == (rs & 7));
      if (mine) {
#pragma unroll
        for (int a = 0; a < RPT; ++a) {
          if (as == a) {
#pragma unroll
            for (int b = b0; b < CPT; ++b) prow[tj + 16 * b] = R[a][b];
          }
        }
      }
      double l[RPT];
#pragma unroll
      for (int a = 0; a < RPT; ++a) {
        const bool isp = mine && (as == a);
        if (isp) { dest[a] = p; rinv[a] = inv; }
        l[a] = isp ? 0.0 : lraw[a] * inv;
      }
      __syncthreads();
#pragma unroll
      for (int b = b0; b < CPT; ++b) {
        const double pr = prow[tj + 16 * b];
#pragma unroll
        for (int a = 0; a < RPT; ++a) R[a][b] = fma(l[a], pr, R[a][b]);
      }
      // next pivot column for the following step
      if (p + 1 < m) {
        if (pp < 15) {
          if (tj == pp + 1) {
#pragma unroll
            for (int a = 0; a < RPT; ++a) cbn[ti + 8 * a] = R[a][b0];
          }
        } else if (b0 + 1 < CPT) {
          if (tj == 0) {
#pragma unroll
            for (int a = 0; a < RPT; ++a) cbn[ti + 8 * a] = R[a][b0 + 1 < CPT ? b0 + 1 : b0];
          }
        }
      }
      __syncthreads();
    }
  }
  return ok;
}

// acc[a][b] = sum_{i < kk} A[(ti + 8a) + lda i] * B[(idx ? idx[i] : i) + ldb (tj + 16b)]; rows / columns beyond mrows / ncols
// are computed from clamped addresses and must be ignored by the caller
template <int RA, int CB>
DYB_D void sb_gemm(double (&acc)[RA][CB], const double* A, int lda, const double* B, int ldb, const int* idx, int kk,
                   int mrows, int ncols, int ti, int tj) {
  int ro[RA], co[CB];
#pragma unroll
  for (int a = 0; a < RA; ++a) { const int r = ti + 8 * a; ro[a] = r < mrows ? r : mrows - 1; }
#pragma unroll
  for (int b = 0; b < CB; ++b) { const int c = tj + 16 * b; co[b] = ldb * (c < ncols ? c : ncols - 1); }
#pragma unroll
  for (int a = 0; a < RA; ++a)
#pragma unroll
    for (int b = 0; b < CB; ++b) acc[a][b] = 0.0;
#pragma unroll 2
  for (int i = 0; i < kk; ++i) {
    const int bi = idx ? idx[i] : i;
    double av[RA], bv[CB];
#pragma unroll
    for (int a = 0; a < RA; ++a) av[a] = A[ro[a] + lda * i];
#pragma unroll
    for (int b = 0; b < CB; ++b) bv[b] = B[bi + co[b]];
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
      for (int b = 0; b < CB; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
  }
}

// acc[a][b] = sum_{i < kk} A[(ti + 8a) + lda i] * B[(tj + 16b) + ldb i]      (A B')
template <int RA, int CB>
DYB_D void sb_gemm_nt(double (&acc)[RA][CB], const double* A, int lda, const double* B, int ldb, int kk, int mrows, int ncols,
                      int ti, int tj) {
  int ro[RA], co[CB];
#pragma unroll
  for (int a = 0; a < RA; ++a) { const int r = ti + 8 * a; ro[a] = r < mrows ? r : mrows - 1; }
#pragma unroll
  for (int b = 0; b < CB; ++b) { const int c = tj + 16 * b; co[b] = c < ncols ? c : ncols - 1; }
#pragma unroll
  for (int a = 0; a < RA; ++a)
#pragma unroll
    for (int b = 0; b < CB; ++b) acc[a][b] = 0.0;
#pragma unroll 2
  for (int i = 0; i < kk; ++i) {
    double av[RA], bv[CB];
#pragma unroll
    for (int a = 0; a < RA; ++a) av[a] = A[ro[a] + lda * i];
#pragma unroll
    for (int b = 0; b < CB; ++b) bv[b] = B[co[b] + ldb * i];
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
      for (int b = 0; b < CB; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
  }
}

// doubling solve of X = A X A' + B (k x k, leading dimension ldk): on entry Ak = A, Sg = B; on exit Sg = X (symmetrised).
// Same recursion and stopping rule as blk_lyapunov; k <= 8 RPT and k <= 16 LCB.
template <int RPT, int LCB>
DYB_D int sb_lyapunov(double* Ak, double* Sg, double* Tm, int k, int ldk, double tol, int maxit, double* red, int* iters,
                      int ti, int tj, int warp, int lane) {
  int conv = 0, done = 0;
#pragma unroll 1
  for (int li = 0; li < maxit; ++li) {
    done = li + 1;
    double t1[RPT][LCB], a2[RPT][LCB];
    {
      // t1 = Ak Sg and a2 = Ak Ak share the left operand
      int ro[RPT], co[LCB];
#pragma unroll
      for (int a = 0; a < RPT; ++a) { const int r = ti + 8 * a; ro[a] = r < k ? r : k - 1; }
#pragma unroll
      for (int b = 0; b < LCB; ++b) { const int c = tj + 16 * b; co[b] = ldk * (c < k ? c : k - 1); }
#pragma unroll
      for (int a = 0; a < RPT; ++a)
#pragma unroll
        for (int b = 0; b < LCB; ++b) { t1[a][b] = 0.0; a2[a][b] = 0.0; }
#pragma unroll 2
      for (int q = 0; q < k; ++q) {
        double av[RPT], sv[LCB], cv[LCB];
#pragma unroll
        for (int a = 0; a < RPT; ++a) av[a] = Ak[ro[a] + ldk * q];
#pragma unroll
        for (int b = 0; b < LCB; ++b) { sv[b] = Sg[q + co[b]]; cv[b] = Ak[q + co[b]]; }
#pragma unroll
        for (int a = 0; a < RPT; ++a)
#pragma unroll
          for (int b = 0; b < LCB; ++b) { t1[a][b] = fma(av[a], sv[b], t1[a][b]); a2[a][b] = fma(av[a], cv[b], a2[a][b]); }
      }
    }
#pragma unroll
    for (int a = 0; a < RPT; ++a)
#pragma unroll
      for (int b = 0; b < LCB; ++b) {
        const int r = ti + 8 * a, c = tj + 16 * b;
        if (r < k && c < k) Tm[r + ldk * c] = t1[a][b];
      }
    __syncthreads();
    sb_gemm_nt<RPT, LCB>(t1, Tm, ldk, Ak, ldk, k, k, k, ti, tj);           // t1 = (Ak Sg) Ak'
    double dmax = 0.0, smax = 0.0, amax = 0.0;
#pragma unroll
    for (int a = 0; a < RPT; ++a)
#pragma unroll
      for (int b = 0; b < LCB; ++b) {
        const int r = ti + 8 * a, c = tj + 16 * b;
        if (r < k && c < k) {
          const double v = Sg[r + ldk * c] + t1[a][b];
          Sg[r + ldk * c] = v;
          dmax = fmax(dmax, fabs(t1[a][b]));
          smax = fmax(smax, v == v ? fabs(v) : INFINITY);
          const double w = a2[a][b];
          amax = fmax(amax, (w - w == 0.0) ? fabs(w) : INFINITY);
        }
      }
    __syncthreads();                                                      // every read of Ak is done
#pragma unroll
    for (int a = 0; a < RPT; ++a)
#pragma unroll
      for (int b = 0; b < LCB; ++b) {
        const int r = ti + 8 * a, c = tj + 16 * b;
        if (r < k && c < k) Ak[r + ldk * c] = a2[a][b];
      }
    sb_block_max3(dmax, smax, amax, red, warp, lane);
    if (!isfinite(smax)) break;
    if (dmax <= tol * smax && amax < 1.0) { conv = 1; break; }
    if (li == LYAP_UNIT_ROOT_IT && amax >= LYAP_UNIT_ROOT_NORM) break;       // a root within 1e-6 of the unit circle
    if (!isfinite(amax)) break;
  }
  if (iters) *iters = done;
  __syncthreads();
  if (conv) {
    double sy[RPT][LCB];
#pragma unroll
    for (int a = 0; a < RPT; ++a)
#pragma unroll
      for (int b = 0; b < LCB; ++b) {
        const int r = ti + 8 * a, c = tj + 16 * b;
        sy[a][b] = (r < k && c < k && r > c) ? 0.5 * (Sg[r + ldk * c] + Sg[c + ldk * r]) : 0.0;
      }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < RPT; ++a)
#pragma unroll
      for (int b = 0; b < LCB; ++b) {
        const int r = ti + 8 * a, c = tj + 16 * b;
        if (r < k && c < k && r > c) { Sg[r + ldk * c] = sy[a][b]; Sg[c + ldk * r] = sy[a][b]; }
      }
    __syncthreads();
  }
  return conv;
}

// RPT / CPT: row / column slots of the Gauss-Jordan array (m <= 8 RPT, m + k + nf <= 16 CPT, m + nx <= 16 CPT); XCB: column
// slots of the k + nf wide products (k + nf <= 16 XCB); all three are checked on the host, which picks the smallest instantiation
template <int RPT, int CPT, int XCB>
__global__ void __launch_bounds__(SB_THREADS, 3)
solve_block_kernel(GenericModel gm, GenericSolveArgs a, SbLayout L) {
  extern __shared__ double sm[];
  constexpr int LCB = (RPT + 1) / 2;            // column slots of k x k blocks (k <= 8 RPT)
  const int tid = threadIdx.x, nt = SB_THREADS, lane = tid & 31, warp = tid >> 5;
  const int ti = tid >> 4, tj = tid & 15;
  const int n = gm.n, nx = gm.nx, ny = gm.ny, S = gm.s, K = gm.k, NF = gm.nf, MM = gm.m, NCOL = gm.ncol, NS = gm.ns;
  const int cC = S, cB = S + K, cA = S + K + MM, cU = S + K + MM + NF;
  const int ldw = L.ldw, ldm = L.ldm, lds = L.lds, ldk = L.ldk, ldn = L.ldn;
  double* W = sm + L.oW;
  double* U = sm + L.oU;
  double* A1 = U + L.uA1; double* A0 = U + L.uA0; double* A2 = U + L.uA2; double* ACC = U + L.uACC; double* X = U + L.uX;
  double* Ak = U + L.uAk; double* Sg = U + L.uSg; double* Tm = U + L.uTm; double* Bs = U + L.uBs; double* BQ = U + L.uBQ;
  double* g1s = U + L.uG1s; double* GS = U + L.uGS; double* g2s = U + L.uG2s; double* RQ = U + L.uRQ;
  double* Gd = sm + L.oGd; double* gud = sm + L.oGud; double* gst = sm + L.oGst; double* GG = U + L.uGG;
  double* Wt = U + L.uWt;                       // top S rows of W (leading dimension lds) while the static rows are solved
  double* Wg = a.wscratch + (size_t)blockIdx.x * L.wdoubles;    // W after the QR (leading dimension ldw)
  double* Se = sm + L.oSe; double* colbuf = sm + L.oCol; double* prow = sm + L.oRow; double* red = sm + L.oRed;
  int* bk = reinterpret_cast<int*>(sm + L.oInt);
  int* fw = bk + K;
  int* bkof = fw + NF;                          // dynamic variable -> its index among the backward variables, or -1
  __shared__ int okflag;

  for (int i = tid; i < K; i += nt) bk[i] = gm.bkwrd_in_dyn[i];
  for (int i = tid; i < NF; i += nt) fw[i] = gm.fwrd_in_dyn[i];
  for (int i = tid; i < MM; i += nt) bkof[i] = -1;
  __syncthreads();
  for (int i = tid; i < K; i += nt) bkof[bk[i]] = i;

#pragma unroll 1
  for (long long d = blockIdx.x; d < a.n_draws; d += gridDim.x) {
    __syncthreads();
    if (a.status_in && a.status_in[d] != ST_OK) {
      if (tid == 0) { a.status[d] = a.status_in[d]; if (a.cr_iters) a.cr_iters[d] = 0; if (a.lyap_iters) a.lyap_iters[d] = 0; }
      continue;
    }
    int status = ST_OK;
    const double* J = a.jac + (size_t)d * n * NCOL;
    // ---- Jacobian into the working column order; NaN / Inf anywhere -> steady-state class failure -------------
    if (tid == 0) okflag = 1;
    __syncthreads();
    double chk = 0.0;
    {
      // eight coalesced loads in flight per thread (the Jacobian is the one large global read of a draw: n x ncol doubles)
      const int tot = n * NCOL;
#pragma unroll 1
      for (int e0 = tid; e0 < tot; e0 += 8 * nt) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int e = e0 + u * nt; v[u] = e < tot ? J[e] : 0.0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = e0 + u * nt;
          if (e < tot) {
            const int c = e / n, r = e - c * n;
            chk = fma(v[u], 0.0, chk);
            W[r + ldw * gm.jperm[c]] = v[u];
          }
        }
      }
    }
#pragma unroll 1
    for (int e = tid; e < nx * nx; e += nt) { const double v = a.sigma_e[(size_t)d * a.stride_se + e]; Se[e] = v; chk = fma(v, 0.0, chk); }
    if (!(chk == 0.0)) okflag = 0;
    __syncthreads();
    if (!okflag) status = ST_STEADY;

    // ---- Householder QR on the static columns, applied to every other column ------------------------------
#pragma unroll 1
    for (int j = 0; j < S && status == ST_OK; ++j) {
      double* col = W + ldw * j;
      // squared norm of col[j:], the same association order in every warp
      double part = 0.0;
#pragma unroll 1
      for (int i = j + lane; i < n; i += 32) part = fma(col[i], col[i], part);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      const double nrm2 = part;
      const double nrm = sqrt(nrm2);
      if (!(nrm > 0.0) || !isfinite(nrm)) { status = ST_SOLVER; break; }
      const double cj = col[j];
      const double alpha = cj >= 0.0 ? -nrm : nrm;
      const double vj = cj - alpha;
      const double beta = 2.0 / (2.0 * (nrm2 - alpha * cj));
#pragma unroll 1
      for (int c = j + 1 + tid; c < NCOL; c += nt) {
        double* x = W + ldw * c;
        double dot = vj * x[j];
#pragma unroll 4
        for (int i = j + 1; i < n; ++i) dot = fma(col[i], x[i], dot);
        dot *= beta;
        x[j] = fma(-dot, vj, x[j]);
        // eight elements at a time, loads before stores: x and col are one array to the compiler, so a store to x[i]
        // inside the loop orders the loads of x[i + 1], col[i + 1] behind it (one shared-memory round trip per element)
#pragma unroll 1
        for (int i0 = j + 1; i0 < n; i0 += 8) {
          double xv[8], cv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { const int ii = i0 + u < n ? i0 + u : n - 1; xv[u] = x[ii]; cv[u] = col[ii]; }
#pragma unroll
          for (int u = 0; u < 8; ++u) if (i0 + u < n) x[i0 + u] = fma(-dot, cv[u], xv[u]);
        }
      }
      __syncthreads();
      if (tid == 0) col[j] = alpha;
    }
    __syncthreads();
    // W leaves shared memory: its region is the cyclic reduction's from here on
    if (status == ST_OK) {
#pragma unroll 4
      for (int e = tid; e < L.wdoubles; e += nt) Wg[e] = W[e];
    }
    __syncthreads();

    // ---- cyclic reduction on the MM x MM dynamic block --------------------------------------------------------
    int it = 0;
    if (status == ST_OK) {
#pragma unroll 1
      for (int e = tid; e < MM * K; e += nt) { const int r = e % MM, j = e / MM; A0[r + ldm * j] = Wg[S + r + ldw * (cC + j)]; ACC[r + ldm * j] = 0.0; }
#pragma unroll 1
      for (int e = tid; e < MM * NF; e += nt) { const int r = e % MM, j = e / MM; A2[r + ldm * j] = Wg[S + r + ldw * (cA + j)]; }
#pragma unroll 1
      for (int e = tid; e < MM * MM; e += nt) { const int r = e % MM, c = e / MM; A1[r + ldm * c] = Wg[S + r + ldw * (cB + c)]; }
      __syncthreads();
      bool conv = false, stalled = false;
      double n0 = 0.0, n2 = 0.0;
      const int nc = MM + K + NF;
#pragma unroll 1
      for (it = 1; it <= a.cr_maxit; ++it) {
        double R[RPT][CPT];
        int dest[RPT];
        double rinv[RPT];
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
          for (int b = 0; b < CPT; ++b) {
            const int r = ti + 8 * q, c = tj + 16 * b;
            double v = 0.0;
            if (r < MM) {
              if (c < MM) v = A1[r + ldm * c];
              else if (c < MM + K) v = A0[r + ldm * (c - MM)];
              else if (c < nc) v = A2[r + ldm * (c - MM - K)];
            }
            R[q][b] = v;
          }
        const bool okq = sb_gauss_jordan<RPT, CPT>(R, dest, rinv, MM, nc, colbuf, prow, ti, tj, lane);
        if (!okq) { if (it == 1) status = ST_SOLVER; else stalled = true; break; }
        // X = A1^-1 [A0 | A2], rows back in order
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
          for (int b = 0; b < CPT; ++b) {
            const int r = ti + 8 * q, c = tj + 16 * b;
            if (r < MM && c >= MM && c < nc) X[dest[q] + ldm * (c - MM)] = R[q][b] * rinv[q];
          }
        __syncthreads();
        // [W00 | W01] = A0 X[bk, :],  [W10 | W11] = A2 X[fw, :]
        double w0[RPT][XCB], w2[RPT][XCB];
        sb_gemm<RPT, XCB>(w0, A0, ldm, X, ldm, bk, K, MM, K + NF, ti, tj);
        sb_gemm<RPT, XCB>(w2, A2, ldm, X, ldm, fw, NF, MM, K + NF, ti, tj);
        __syncthreads();
        // A0 <- -W00, A2 <- -W11, acc_hat += W10, A1[:, bk] -= W10
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
          for (int b = 0; b < XCB; ++b) {
            const int r = ti + 8 * q, c = tj + 16 * b;
            if (r < MM) {
              if (c < K) { A0[r + ldm * c] = -w0[q][b]; ACC[r + ldm * c] += w2[q][b]; A1[r + ldm * bk[c]] -= w2[q][b]; }
              else if (c < K + NF) A2[r + ldm * (c - K)] = -w2[q][b];
            }
          }
        __syncthreads();
        // A1[:, fw] -= W01
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
          for (int b = 0; b < XCB; ++b) {
            const int r = ti + 8 * q, c = tj + 16 * b;
            if (r < MM && c >= K && c < K + NF) A1[r + ldm * fw[c - K]] -= w0[q][b];
          }
        // 1-norms (largest column sum), sequential per column
#pragma unroll 1
        for (int j = tid; j < K + NF; j += nt) {
          const double* c = j < K ? A0 + ldm * j : A2 + ldm * (j - K);
          double cs = 0.0;
#pragma unroll 4
          for (int r = 0; r < MM; ++r) cs += fabs(c[r]);
          red[16 + j] = cs;
        }
        __syncthreads();
        double m0 = 0.0, m2 = 0.0, bad = 0.0;
#pragma unroll 1
        for (int j = lane; j < K + NF; j += 32) {
          const double v = red[16 + j];
          if (!(v <= 1e150)) bad = 1.0;
          else if (j < K) m0 = fmax(m0, v);
          else m2 = fmax(m2, v);
        }
        m0 = sb_warp_max(m0); m2 = sb_warp_max(m2); bad = sb_warp_max(bad);
        __syncthreads();
        if (bad != 0.0) { if (it == 1) status = ST_SOLVER; else stalled = true; break; }
        const double n0_prev = n0;
        n0 = m0; n2 = m2;
        if ((n0 < a.cr_tol && n2 < a.cr_tol) || (n2 < a.cr_tol * a.cr_tol && fabs(n0 - n0_prev) <= CR_UNIT_ROOT_EPS * n0)) { conv = true; break; }
      }
      if (status == ST_OK && !conv) {
        if (stalled) status = (n0 < n2) ? ST_INDETERMINATE : ST_UNSTABLE;
        else status = (n0 >= a.cr_tol) ? ST_UNSTABLE : ST_INDETERMINATE;
      }
    }
    __syncthreads();

    if (status == ST_OK) {
      // ---- G (dynamic rows, backward columns) = -A1hat^-1 A0(0),  A1hat = A1(0) - acc_hat on the backward columns ------
      double R[RPT][CPT];
      int dest[RPT];
      double rinv[RPT];
      const int nc = MM + K;
#pragma unroll
      for (int q = 0; q < RPT; ++q)
#pragma unroll
        for (int b = 0; b < CPT; ++b) {
          const int r = ti + 8 * q, c = tj + 16 * b;
          double v = 0.0;
          if (r < MM) {
            if (c < MM) { v = Wg[S + r + ldw * (cB + c)]; const int jb = bkof[c]; if (jb >= 0) v -= ACC[r + ldm * jb]; }
            else if (c < nc) v = -Wg[S + r + ldw * (cC + (c - MM))];
          }
          R[q][b] = v;
        }
      if (!sb_gauss_jordan<RPT, CPT>(R, dest, rinv, MM, nc, colbuf, prow, ti, tj, lane)) status = ST_SOLVER;
      else {
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
          for (int b = 0; b < CPT; ++b) {
            const int r = ti + 8 * q, c = tj + 16 * b;
            if (r < MM && c >= MM && c < nc) Gd[dest[q] + ldm * (c - MM)] = R[q][b] * rinv[q];
          }
      }
    }
    __syncthreads();
    if (status == ST_OK) {
      // ---- g_u = -(A G + B)^-1 F_u on the dynamic block ---------------------------------------------------------
      double t[RPT][LCB];
      sb_gemm<RPT, LCB>(t, Wg + S + ldw * cA, ldw, Gd, ldm, fw, NF, MM, K, ti, tj);       // A2(0) G[fw, :] -> X as staging
#pragma unroll
      for (int q = 0; q < RPT; ++q)
#pragma unroll
        for (int b = 0; b < LCB; ++b) {
          const int r = ti + 8 * q, c = tj + 16 * b;
          if (r < MM && c < K) X[r + ldm * c] = t[q][b];
        }
      __syncthreads();
      double R[RPT][CPT];
      int dest[RPT];
      double rinv[RPT];
      const int nc = MM + nx;
#pragma unroll
      for (int q = 0; q < RPT; ++q)
#pragma unroll
        for (int b = 0; b < CPT; ++b) {
          const int r = ti + 8 * q, c = tj + 16 * b;
          double v = 0.0;
          if (r < MM) {
            if (c < MM) { v = Wg[S + r + ldw * (cB + c)]; const int jb = bkof[c]; if (jb >= 0) v += X[r + ldm * jb]; }
            else if (c < nc) v = -Wg[S + r + ldw * (cU + (c - MM))];
          }
          R[q][b] = v;
        }
      if (!sb_gauss_jordan<RPT, CPT>(R, dest, rinv, MM, nc, colbuf, prow, ti, tj, lane)) status = ST_SOLVER;
      else {
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
          for (int b = 0; b < CPT; ++b) {
            const int r = ti + 8 * q, c = tj + 16 * b;
            if (r < MM && c >= MM && c < nc) gud[dest[q] + ldm * (c - MM)] = R[q][b] * rinv[q];
          }
      }
    }
    __syncthreads();
    if (status == ST_OK && S > 0) {
      // ---- static rows by back-substitution: R_s y_s + B_top y_d + A_top y_f(+1) + C_top y_b(-1) + Fu_top u = 0 ---
      // the top S rows of W come back from the scratch into the region the cyclic reduction has left
#pragma unroll 4
      for (int e = tid; e < S * NCOL; e += nt) { const int q = e % S, c = e / S; Wt[q + lds * c] = Wg[q + ldw * c]; }
#pragma unroll 1
      for (int e = tid; e < NF * (K + nx); e += nt) {
        const int i = e % NF, c = e / NF;
        const double* gc = (c < K) ? Gd + ldm * c : gud + ldm * (c - K);
        const int fi = fw[i];
        double s = 0.0;
#pragma unroll 4
        for (int q = 0; q < K; ++q) s = fma(Gd[fi + ldm * q], gc[bk[q]], s);
        GG[e] = s;
      }
      if (tid == 0) okflag = 1;
      __syncthreads();
#pragma unroll 1
      for (int e = tid; e < S * (K + nx); e += nt) {
        const int q = e % S, c = e / S;
        const double* gcol = (c < K) ? (Gd + ldm * c) : (gud + ldm * (c - K));
        double s = (c < K) ? Wt[q + lds * (cC + c)] : Wt[q + lds * (cU + (c - K))];
#pragma unroll 4
        for (int dd = 0; dd < MM; ++dd) s = fma(Wt[q + lds * (cB + dd)], gcol[dd], s);
#pragma unroll 4
        for (int i = 0; i < NF; ++i) s = fma(Wt[q + lds * (cA + i)], GG[i + NF * c], s);
        gst[q + lds * c] = -s;
      }
      __syncthreads();
#pragma unroll 1
      for (int c = tid; c < K + nx; c += nt) {
#pragma unroll 1
        for (int q = S - 1; q >= 0; --q) {
          double s = gst[q + lds * c];
#pragma unroll 1
          for (int r = q + 1; r < S; ++r) s = fma(-Wt[q + lds * r], gst[r + lds * c], s);
          const double dg = Wt[q + lds * q];
          if (dg == 0.0) okflag = 0;
          gst[q + lds * c] = s / dg;
        }
      }
      __syncthreads();
      if (!okflag) status = ST_SOLVER;
    }
    __syncthreads();
    if (status == ST_OK && a.g1) {
      double* g1 = a.g1 + (size_t)d * n * (K + nx);
#pragma unroll 1
      for (int e = tid; e < n * (K + nx); e += nt) {
        const int v = e % n, c = e / n;
        const int pos = gm.var_pos[v];
        g1[e] = gm.var_static[v] ? gst[pos + lds * c] : ((c < K) ? Gd[pos + ldm * c] : gud[pos + ldm * (c - K)]);
      }
    }

    // ---- Lyapunov doubling on the K x K predetermined block -----------------------------------------------------
    int lyap_it = 0;
    if (status == ST_OK) {
#pragma unroll 1
      for (int e = tid; e < K * K; e += nt) { const int i = e % K, j = e / K; Ak[i + ldk * j] = Gd[bk[i] + ldm * j]; }
#pragma unroll 1
      for (int e = tid; e < K * nx; e += nt) { const int i = e % K, q = e / K; Bs[i + ldk * q] = gud[bk[i] + ldm * q]; }
      __syncthreads();
#pragma unroll 1
      for (int e = tid; e < K * nx; e += nt) {
        const int i = e % K, q = e / K;
        double s = 0.0;
#pragma unroll 4
        for (int f = 0; f < nx; ++f) s = fma(Bs[i + ldk * f], Se[f + nx * q], s);
        BQ[i + ldk * q] = s;
      }
      __syncthreads();
#pragma unroll 1
      for (int e = tid; e < K * K; e += nt) {
        const int i = e % K, j = e / K;
        double s = 0.0;
#pragma unroll 4
        for (int q = 0; q < nx; ++q) s = fma(BQ[i + ldk * q], Bs[j + ldk * q], s);
        Sg[i + ldk * j] = s;
      }
      __syncthreads();
      if (!sb_lyapunov<RPT, LCB>(Ak, Sg, Tm, K, ldk, a.lyap_tol, a.lyap_maxit, red, &lyap_it, ti, tj, warp, lane)) status = ST_NONSTATIONARY;
    }
    __syncthreads();

    // ---- dense state space (estimation.jl:640-658) -----------------------------------------------------------------
    if (status == ST_OK) {
#pragma unroll 1
      for (int e = tid; e < NS * (K + nx); e += nt) {
        const int sI = e % NS, c = e / NS;
        const int v = gm.state_ids[sI];
        const int pos = gm.var_pos[v];
        const double val = gm.var_static[v] ? gst[pos + lds * c] : ((c < K) ? Gd[pos + ldm * c] : gud[pos + ldm * (c - K)]);
        if (c < K) g1s[sI + ldn * c] = val; else g2s[sI + ldn * (c - K)] = val;
      }
      __syncthreads();
#pragma unroll 1
      for (int e = tid; e < NS * nx; e += nt) {
        const int sI = e % NS, q = e / NS;
        double s = 0.0;
#pragma unroll 4
        for (int f = 0; f < nx; ++f) s = fma(g2s[sI + ldn * f], Se[f + nx * q], s);
        RQ[sI + ldn * q] = s;
      }
#pragma unroll 1
      for (int e = tid; e < NS * K; e += nt) {
        const int sI = e % NS, j = e / NS;
        double s = 0.0;
#pragma unroll 4
        for (int q = 0; q < K; ++q) s = fma(g1s[sI + ldn * q], Sg[q + ldk * j], s);
        GS[sI + ldn * j] = s;
      }
      __syncthreads();
      double* To = a.T ? a.T + (size_t)d * NS * NS : nullptr;
      double* Qo = a.RQR ? a.RQR + (size_t)d * NS * NS : nullptr;
      double* Po = a.P0 ? a.P0 + (size_t)d * NS * NS : nullptr;
      if (To) for (int e = tid; e < NS * NS; e += nt) To[e] = 0.0;
      __syncthreads();
      if (To) for (int e = tid; e < NS * K; e += nt) { const int sI = e % NS, j = e / NS; To[sI + NS * gm.lagged_state[j]] = g1s[sI + ldn * j]; }
#pragma unroll 1
      for (int e = tid; e < NS * NS; e += nt) {
        const int sI = e % NS, t = e / NS;
        if (sI < t) continue;
        double qq = 0.0;
#pragma unroll 4
        for (int q = 0; q < nx; ++q) qq = fma(RQ[sI + ldn * q], g2s[t + ldn * q], qq);
        double pp = qq;
#pragma unroll 4
        for (int j = 0; j < K; ++j) pp = fma(GS[sI + ldn * j], g1s[t + ldn * j], pp);
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

}  // namespace dyb
