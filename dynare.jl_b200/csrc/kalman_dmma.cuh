// Dense batched Kalman log-likelihood for medium state spaces (8 < ns <= 88: Smets-Wouters scale, BASELINE config 4)
// with the covariance recursion on the fp64 tensor cores.
//
// One CTA per draw; T, P and T P stay in shared memory for all nobs periods (no HBM traffic inside the recursion
// except the R Q R' tile added once per period, which is L2-resident).  The three dense updates of a period,
//     P <- P - U'U        (U = L^-1 Z P, rank-ny downdate)
//     M  = T P
//     P  = M T' + R Q R'  (lower block triangle, mirrored)
// are issued as mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4, 256 FMA per warp instruction): warp w owns the 8-row strip w of
// the result and reuses its A fragment across the strip, so a k-step costs 1 + ntile shared loads for ntile DMMAs.
// Leading dimensions are padded to ld = 4 or 12 (mod 16) doubles, which makes every fragment load conflict-free.
// Everything else (innovation, Cholesky of F, U, state update) is the scalar code of kalman_generic.cuh.
//
// Replaces `kalman_likelihood(Y,Z,H,T,R,Q,a0,P,start,last,presample,ws)` (reference call site
// src/estimation/estimation.jl:687-700); KalmanGenericArgs is shared with kalman_generic.cuh.
#pragma once
#include "kalman_generic.cuh"

namespace dyb {

constexpr int DMMA_MAX_TILES = 11;     // ns <= 88

DYB_D void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

DYB_HD int dmma_pad_ld(int n) {        // smallest ld >= n with ld mod 16 in {4, 12}
  int ld = n;
  while ((ld & 15) != 4 && (ld & 15) != 12) ++ld;
  return ld;
}

// Cholesky of the ny x ny innovation covariance (ny <= 8) and w = L^-1 v by ONE warp with the matrix in registers:
// lane i owns row i (ny <= NY <= 32), pivots and multipliers travel by shuffles (no shared-memory round trips on the critical path).
// In: F (column-major, lower part used), v.  Out: F <- L (lower incl. diagonal), w, and on every lane the sums
// quad = w'w and ldet = sum_i log L_ii; returns false if a pivot is not positive.
template <int NY>
DYB_D bool chol_warp(double* F, int ny, const double* v, double* w, double* rdv, double& quad, double& ldet) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double f[NY];
#pragma unroll
  for (int j = 0; j < NY; ++j) f[j] = (lane < ny && j <= lane) ? F[lane + ny * j] : 0.0;
  double rd = 1.0, ldiag = 1.0;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < NY; ++j) {
    if (j < ny) {
      const double d = __shfl_sync(FULL, f[j], j);
      ok = ok && (d > 0.0);
      const double r = rsqrt(d);
      const double lij = (lane == j) ? d * r : f[j] * r;
      f[j] = lij;
      if (lane == j) { rd = r; ldiag = lij; }
#pragma unroll
      for (int c = j + 1; c < NY; ++c) {
        if (c < ny) {
          const double lcj = __shfl_sync(FULL, lij, c);
          f[c] = fma(-lij, lcj, f[c]);
        }
      }
    }
  }
  double x = lane < ny ? v[lane] : 0.0;
#pragma unroll
  for (int j = 0; j < NY; ++j) {
    if (j < ny) {
      const double wj = __shfl_sync(FULL, x * rd, j);
      if (lane > j) x = fma(-f[j], wj, x);
    }
  }
  const double wi = x * rd;
  double q2 = lane < ny ? wi * wi : 0.0, lg = lane < ny ? log(ldiag) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { q2 += __shfl_xor_sync(FULL, q2, o); lg += __shfl_xor_sync(FULL, lg, o); }
  if (lane < ny) {
    w[lane] = wi;
    rdv[lane] = rd;
#pragma unroll
    for (int j = 0; j < NY; ++j) if (j <= lane) F[lane + ny * j] = f[j];
  }
  quad = q2; ldet = lg;
  return ok;
}

DYB_D bool chol8_warp(double* F, int ny, const double* v, double* w, double* rdv, double& quad, double& ldet) {
  return chol_warp<8>(F, ny, v, w, rdv, quad, ldet);
}
// The same factorisation for ny <= NY <= 32 as a LOOP over the pivot columns whose body has static register indices:
// after column j is eliminated every lane shifts its row one place to the left, so the pivot column is always f[0].
// (A fully unrolled version is ~100 KB of straight-line code that is executed once per period and misses the
// instruction cache every time: measured 47k cycles at ny = 20 against ~4k for this loop.)
template <int NY>
DYB_D bool chol_warp_rot(double* F, int ny, const double* v, double* w, double* rdv, double& quad, double& ldet) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double f[NY];
#pragma unroll
  for (int c = 0; c < NY; ++c) f[c] = (lane < ny && c <= lane) ? F[lane + ny * c] : 0.0;
  double x = lane < ny ? v[lane] : 0.0;
  double rd = 1.0, ldiag = 1.0, wi = 0.0;
  bool ok = true;
#pragma unroll 1
  for (int j = 0; j < ny; ++j) {
    const double d = __shfl_sync(FULL, f[0], j);
    ok = ok && (d > 0.0);
    const double r = rsqrt(d);
    const double lij = (lane == j) ? d * r : f[0] * r;
    if (lane >= j && lane < ny) F[lane + ny * j] = lij;
    const double wj = __shfl_sync(FULL, x * r, j);
    if (lane == j) { rd = r; ldiag = lij; wi = wj; }
    if (lane > j) x = fma(-lij, wj, x);
#pragma unroll
    for (int c = 1; c < NY; ++c) {
      const double lcj = __shfl_sync(FULL, lij, j + c);       // lanes >= ny hold zeros; upper-part entries are never used
      f[c - 1] = fma(-lij, lcj, f[c]);
    }
    f[NY - 1] = 0.0;
  }
  double q2 = lane < ny ? wi * wi : 0.0, lg = lane < ny ? log(ldiag) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { q2 += __shfl_xor_sync(FULL, q2, o); lg += __shfl_xor_sync(FULL, lg, o); }
  if (lane < ny) { w[lane] = wi; rdv[lane] = rd; }
  quad = q2; ldet = lg;
  return ok;
}
// out of line, so that the NY doubles of a matrix row do not compete with the caller's accumulators for registers
template <int NY>
__device__ __noinline__ bool chol_warp_call(double* F, int ny, const double* v, double* w, double* rdv, double* quad_ldet) {
  double quad, ldet;
  const bool ok = chol_warp_rot<NY>(F, ny, v, w, rdv, quad, ldet);
  quad_ldet[0] = quad; quad_ldet[1] = ldet;
  return ok;
}

struct KalmanDmmaGeom { int nsp, ld, nyp, ldu; };

inline KalmanDmmaGeom kalman_dmma_geom(int ns, int ny) {
  KalmanDmmaGeom q;
  q.nsp = (ns + 7) / 8 * 8; q.ld = dmma_pad_ld(q.nsp);
  q.nyp = (ny + 3) / 4 * 4; q.ldu = dmma_pad_ld(q.nyp);
  return q;
}
inline size_t kalman_dmma_smem_doubles(int ns, int ny) {
  const KalmanDmmaGeom q = kalman_dmma_geom(ns, ny);
  return 3 * (size_t)q.ld * q.nsp + (size_t)q.ldu * q.nsp + (size_t)ny * ny + 2 * (size_t)q.nsp + 3 * (size_t)ny + 8;
}

// blockDim.x = 32 * nsp / 8 <= 32 * MAXT; MAXT bounds the accumulator tiles a warp holds (register budget)
template <int MAXT>
__global__ void __launch_bounds__(32 * MAXT, (MAXT <= 3 ? 8 : (MAXT <= 5 ? 4 : (MAXT <= 8 ? 2 : 1))))
kalman_dmma_kernel(KalmanGenericArgs g, KalmanDmmaGeom q) {
  extern __shared__ double sm[];
  const int ns = g.ns, ny = g.ny, nsp = q.nsp, ld = q.ld, nyp = q.nyp, ldu = q.ldu;
  const int ntile = nsp / 8;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  double* Ts = sm;                              // Ts[i + ld * j]
  double* Ps = Ts + (size_t)ld * nsp;
  double* Ms = Ps + (size_t)ld * nsp;
  double* U = Ms + (size_t)ld * nsp;            // U[q + ldu * s], q < nyp
  double* F = U + (size_t)ldu * nsp;            // ny x ny lower Cholesky, column-major
  double* a = F + (size_t)ny * ny;
  double* au = a + nsp;
  double* v = au + nsp;
  double* w = v + ny;
  double* rdv = w + ny;                         // 1 / L_ii of the current period
  __shared__ int bad;
  __shared__ int zi[64];
  const int i0 = 8 * warp;
  const int wu = __reduce_max_sync(0xffffffffu, warp);   // = warp, as a value that ptxas knows to be warp-uniform

  for (long long d = blockIdx.x; d < g.n_draws; d += gridDim.x) {
    __syncthreads();
    if (g.status_in && g.status_in[d] != ST_OK) {
      if (tid == 0) { g.loglik[d] = -INFINITY; g.status[d] = g.status_in[d]; }
      continue;
    }
    const double* Tm = g.T + (size_t)d * ns * ns;
    const double* QQ = g.RQR + (size_t)d * ns * ns;
    const double* P0 = g.P0 + (size_t)d * ns * ns;
    const double* Hm = g.H ? g.H + (size_t)d * ny * ny : nullptr;
    const double* yb = g.ybar_obs ? g.ybar_obs + (size_t)d * ny : nullptr;
    if (tid == 0) bad = 0;
    for (int i = tid; i < ny; i += nt) zi[i] = g.z_idx[i];
    for (int e = tid; e < ld * nsp; e += nt) {
      const int i = e % ld, j = e / ld;
      const bool in = i < ns && j < ns;
      Ts[e] = in ? Tm[i + (size_t)ns * j] : 0.0;
      Ps[e] = in ? P0[i + (size_t)ns * j] : 0.0;
      Ms[e] = 0.0;
    }
    for (int e = tid; e < ldu * nsp; e += nt) U[e] = 0.0;
    for (int i = tid; i < nsp; i += nt) { a[i] = (g.a0 && i < ns) ? g.a0[(size_t)d * ns + i] : 0.0; au[i] = 0.0; }
    __syncthreads();
    double acc = 0.0;          // only thread 0's copy is used
    int nseen = 0;

    for (int t = 0; t < g.nobs; ++t) {
      const double* yt = g.Y + (size_t)t * ny;
      for (int i = tid; i < ny; i += nt) {
        const double yi = yt[i];
        v[i] = isnan(yi) ? 0.0 : (yi - (yb ? yb[i] : 0.0)) - a[zi[i]];
      }
      for (int e = tid; e < ny * ny; e += nt) {
        const int i = e % ny, j = e / ny;
        const bool mi = isnan(yt[i]), mj = isnan(yt[j]);
        double f = Ps[zi[i] + ld * zi[j]] + (Hm ? Hm[e] : 0.0);
        if (mi || mj) f = (i == j) ? 1.0 : 0.0;
        F[e] = f;
      }
      __syncthreads();
      if (tid < 32 && ny <= 8) {              // small ny: register / shuffle Cholesky
        double quad, ldet;
        const bool okc = chol8_warp(F, ny, v, w, rdv, quad, ldet);
        const int cnt = __popc(__ballot_sync(0xffffffffu, tid < ny && !isnan(yt[tid < ny ? tid : 0])));
        if (tid == 0) {
          if (!okc) bad = 1;
          if (t >= g.presample) { acc += 2.0 * ldet + quad; nseen += cnt; }
        }
      } else if (tid < 32 && ny <= 32) {      // loop-form register Cholesky (chol_warp_rot)
        double ql[2];
        bool okc;
        if (ny <= 16) okc = chol_warp_call<16>(F, ny, v, w, rdv, ql);
        else if (ny <= 24) okc = chol_warp_call<24>(F, ny, v, w, rdv, ql);
        else okc = chol_warp_call<32>(F, ny, v, w, rdv, ql);
        const int cnt = __popc(__ballot_sync(0xffffffffu, tid < ny && !isnan(yt[tid < ny ? tid : 0])));
        if (tid == 0) {
          if (!okc) bad = 1;
          if (t >= g.presample) { acc += 2.0 * ql[1] + ql[0]; nseen += cnt; }
        }
      } else if (tid < 32) {                  // ny > 32: Cholesky of F by warp 0 in shared memory, then w = L^-1 v
        for (int j = 0; j < ny; ++j) {
          const double djj = F[j + ny * j];
          __syncwarp();
          if (tid == 0) { if (!(djj > 0.0)) bad = 1; F[j + ny * j] = sqrt(djj); }
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
        if (tid == 0) {
          double quad = 0.0, ldet = 0.0;
          int cnt = 0;
          for (int i = 0; i < ny; ++i) {
            double x = v[i];
            for (int p = 0; p < i; ++p) x -= F[i + ny * p] * w[p];
            const double ri = 1.0 / F[i + ny * i];
            rdv[i] = ri;
            x *= ri;
            w[i] = x;
            quad += x * x;
            ldet += log(F[i + ny * i]);
            cnt += isnan(yt[i]) ? 0 : 1;
          }
          if (t >= g.presample) { acc += 2.0 * ldet + quad; nseen += cnt; }
        }
      }
      __syncthreads();
      if (t == g.nobs - 1) break;
      // U[:, s] = L^-1 (Z P)[:, s] ; au = a + U'w
      for (int s = tid; s < ns; s += nt) {
        double x = a[s];
        for (int i = 0; i < ny; ++i) {
          double u = isnan(yt[i]) ? 0.0 : Ps[zi[i] + ld * s];
          for (int p = 0; p < i; ++p) u -= F[i + ny * p] * U[p + ldu * s];
          u *= rdv[i];
          U[i + ldu * s] = u;
          x += u * w[i];
        }
        au[s] = x;
      }
      __syncthreads();
      // ---- P <- P - U'U  (strip of warp `warp`) ; a <- T au -------------------------------------------------------
      for (int base = 0; base < 4 * ns; base += nt) {       // 4 lanes per row, combined by shuffles (uniform loop)
        const int idx = base + tid, i = idx >> 2, part = idx & 3;
        double x = 0.0;
        if (idx < 4 * ns) for (int p = part; p < ns; p += 4) x = fma(Ts[i + ld * p], au[p], x);
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        if (idx < 4 * ns && part == 0) a[i] = x;            // nobody reads a[] before the next barrier
      }
      {
        double c[MAXT][2];
#pragma unroll
        for (int jt = 0; jt < MAXT; ++jt)
          if (jt < ntile) { c[jt][0] = Ps[(i0 + gq) + ld * (8 * jt + 2 * tq)]; c[jt][1] = Ps[(i0 + gq) + ld * (8 * jt + 2 * tq + 1)]; }
        for (int k0 = 0; k0 < nyp; k0 += 4) {
          const double av = -U[(k0 + tq) + ldu * (i0 + gq)];
#pragma unroll
          for (int jt = 0; jt < MAXT; ++jt)
            if (jt < ntile) dmma884(c[jt][0], c[jt][1], av, U[(k0 + tq) + ldu * (8 * jt + gq)]);
        }
#pragma unroll
        for (int jt = 0; jt < MAXT; ++jt)
          if (jt < ntile) { Ps[(i0 + gq) + ld * (8 * jt + 2 * tq)] = c[jt][0]; Ps[(i0 + gq) + ld * (8 * jt + 2 * tq + 1)] = c[jt][1]; }
      }
      __syncthreads();
      // ---- M = T P ----------------------------------------------------------------------------------------------------
      {
        double c[MAXT][2];
#pragma unroll
        for (int jt = 0; jt < MAXT; ++jt) { c[jt][0] = 0.0; c[jt][1] = 0.0; }
        for (int k0 = 0; k0 < nsp; k0 += 4) {
          const double av = Ts[(i0 + gq) + ld * (k0 + tq)];
#pragma unroll
          for (int jt = 0; jt < MAXT; ++jt)
            if (jt < ntile) dmma884(c[jt][0], c[jt][1], av, Ps[(k0 + tq) + ld * (8 * jt + gq)]);
        }
#pragma unroll
        for (int jt = 0; jt < MAXT; ++jt)
          if (jt < ntile) { Ms[(i0 + gq) + ld * (8 * jt + 2 * tq)] = c[jt][0]; Ms[(i0 + gq) + ld * (8 * jt + 2 * tq + 1)] = c[jt][1]; }
      }
      __syncthreads();
      // ---- P = M T' + R Q R'  (tiles j <= warp, mirrored) ------------------------------------------------------------------
      {
        double c[MAXT][2];
#pragma unroll
        for (int jt = 0; jt < MAXT; ++jt) {
          c[jt][0] = 0.0; c[jt][1] = 0.0;
          if (jt <= wu) {
            const int i = i0 + gq, j = 8 * jt + 2 * tq;
            if (i < ns && j < ns) c[jt][0] = QQ[i + (size_t)ns * j];
            if (i < ns && j + 1 < ns) c[jt][1] = QQ[i + (size_t)ns * (j + 1)];
          }
        }
        for (int k0 = 0; k0 < nsp; k0 += 4) {
          const double av = Ms[(i0 + gq) + ld * (k0 + tq)];
#pragma unroll
          for (int jt = 0; jt < MAXT; ++jt)
            if (jt <= wu) dmma884(c[jt][0], c[jt][1], av, Ts[(8 * jt + gq) + ld * (k0 + tq)]);
        }
#pragma unroll
        for (int jt = 0; jt < MAXT; ++jt)
          if (jt <= wu) {
            const int i = i0 + gq, j = 8 * jt + 2 * tq;
            if (jt < wu) {
              Ps[i + ld * j] = c[jt][0]; Ps[i + ld * (j + 1)] = c[jt][1];
              Ps[j + ld * i] = c[jt][0]; Ps[(j + 1) + ld * i] = c[jt][1];
            } else {                            // diagonal tile: keep the lower triangle, mirror it
              if (i >= j) { Ps[i + ld * j] = c[jt][0]; Ps[j + ld * i] = c[jt][0]; }
              if (i >= j + 1) { Ps[i + ld * (j + 1)] = c[jt][1]; Ps[(j + 1) + ld * i] = c[jt][1]; }
            }
          }
      }
      __syncthreads();
    }
    if (tid == 0) {
      const double ll = -0.5 * (nseen * 1.8378770664093453 + acc);
      if (bad || !isfinite(ll)) { g.loglik[d] = -INFINITY; g.status[d] = ST_F_NOT_PD; }
      else { g.loglik[d] = ll; g.status[d] = ST_OK; }
    }
  }
}

}  // namespace dyb
