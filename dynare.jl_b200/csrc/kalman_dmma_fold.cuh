// Dense batched Kalman log-likelihood for medium state spaces (8 < ns <= 88, ny <= 32: BASELINE config 4) on the fp64
// tensor cores, with the innovation algebra taken OFF the critical path of the covariance recursion.
//
// kalman_dmma_kernel (kalman_dmma.cuh) runs a period as  F, chol(F) -> U = L^-1 Z P -> P -= U'U -> M = T P -> P = M T' + RQR':
// five dependent phases, the first two on one warp while the others wait (measured 55 % of a period at ns = 40, ny = 7).
// Here the rank-ny correction is moved behind the products (Z is a selection, so T P Z' is a set of columns of M = T P):
//     M  = T P                                 (tensor-core warps)      ||   F = P[z,z] + H, L = chol(F), w = L^-1 v   (scalar warp)
//     P' = M T' + R Q R'  (lower triangle)     (tensor-core warps)      ||   V = M[:, z] L^-T,  a' = T a + V w         (scalar warp)
//     P' -= V V'          (two k-steps), stored mirrored
// i.e. P' = T P T' - (T P Z') F^-1 (T P Z')' + R Q R' and a' = T a + T P Z' F^-1 v: the same recursion, three barriers per
// period, and the Cholesky / triangular solves overlap the two big products.  One CTA per draw: ns/8 tensor-core warps
// (warp w owns the 8-row strip w) plus one scalar warp; T, P, M in shared memory for all periods.
//
// Replaces `kalman_likelihood(Y,Z,H,T,R,Q,a0,P,start,last,presample,ws)` (reference call site
// src/estimation/estimation.jl:687-700); KalmanGenericArgs is shared with kalman_generic.cuh.
#pragma once
#include "kalman_dmma.cuh"

namespace dyb {

inline size_t kalman_fold_smem_doubles(int ns, int ny) {
  const KalmanDmmaGeom q = kalman_dmma_geom(ns, ny);
  return 3 * (size_t)q.ld * q.nsp + (size_t)q.ldu * q.nsp + 2 * (size_t)ny * ny + 2 * (size_t)q.nsp + 3 * (size_t)ny + 8;
}

// V[:, s] = L^-1 M[s, z]' for state s; returns V[:, s]' w.  NYR > 0: ny <= NYR, the column of V stays in registers.
template <int NYR>
DYB_D double fold_gain_row(int s, int ny, int ld, int ldu, unsigned miss, const int* zi, const double* Ms, const double* F,
                           const double* rdv, const double* w, double* V) {
  double xa = 0.0;
  if (NYR > 0) {
    double u[NYR > 0 ? NYR : 1];
#pragma unroll
    for (int i = 0; i < NYR; ++i) {
      if (i < ny) {
        double x = ((miss >> i) & 1u) ? 0.0 : Ms[s + ld * zi[i]];
#pragma unroll
        for (int p = 0; p < i; ++p) x = fma(-F[i + ny * p], u[p], x);
        x *= rdv[i];
        u[i] = x;
        V[i + ldu * s] = x;
        xa = fma(x, w[i], xa);
      }
    }
  } else {
#pragma unroll 1
    for (int i = 0; i < ny; ++i) {
      double x = ((miss >> i) & 1u) ? 0.0 : Ms[s + ld * zi[i]];
#pragma unroll 1
      for (int p = 0; p < i; ++p) x = fma(-F[i + ny * p], V[p + ldu * s], x);
      x *= rdv[i];
      V[i + ldu * s] = x;
      xa = fma(x, w[i], xa);
    }
  }
  return xa;
}

DYB_D void fold_bar() { asm volatile("bar.sync 0;" ::: "memory"); }

struct FoldScalarArgs {
  const double* Ps; const double* Hs; const double* abuf; double* F; double* w; double* rdv; int* miss_sh; const int* zi;
  const double* Y; const double* ybar_obs_d;      // observations (ny x nobs), steady state of the observables of this draw or null
  int nsp, ld, nobs, presample;
  double* loglik_d; int* status_d;
};

// Scalar role of kalman_dmma_fold_kernel for NY <= 8 observables, all periods of one draw, executed by ONE warp:
// lane i owns row i of F = P[z,z] + H in registers; Cholesky and w = L^-1 v by shuffles; the lane's share of the
// quadratic form and of det(L) (a running product with its exponent split off) is accumulated per lane, so that no
// cross-lane reduction, logarithm or division is on the per-period path.  Straight-line code for the exact NY.
// Meets the tensor-core warps at the CTA barrier three times per period.
template <int NY>
__device__ __noinline__ void fold_scalar_role(const FoldScalarArgs A) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const bool act = lane < NY;
  int poff[NY];                  // offsets of this lane's row of P[z, z]
  {
    const int zme = act ? A.zi[lane] : 0;
#pragma unroll
    for (int j = 0; j < NY; ++j) poff[j] = zme + A.ld * A.zi[j];
  }
  const double* hrow = A.Hs + (act ? lane : 0);
  const int aoff = act ? A.zi[lane] : 0;
  const double ybi = (act && A.ybar_obs_d) ? A.ybar_obs_d[lane] : 0.0;
  double ynext = (act && A.nobs > 0) ? A.Y[lane] : 0.0;
  double qacc = 0.0, pr = 1.0;
  int nseen = 0, eacc = 0;
  bool okall = true;
#pragma unroll 1
  for (int t = 0; t < A.nobs; ++t) {
    const double* a = A.abuf + (t & 1) * A.nsp;
    const bool last = (t == A.nobs - 1);
    const double yi = ynext;
    if (act && !last) ynext = A.Y[(size_t)(t + 1) * NY + lane];
    const unsigned miss = __ballot_sync(FULL, act && isnan(yi));
    if (lane == 0) *A.miss_sh = (int)miss;
    const bool mi = (miss >> lane) & 1u;
    double x = (act && !mi) ? (yi - ybi) - a[aoff] : 0.0;
    double f[NY];
#pragma unroll
    for (int j = 0; j < NY; ++j) {
      double val = (j == lane) ? 1.0 : 0.0;                 // rows beyond NY and missing observations: identity
      if (act && j <= lane && !(mi || ((miss >> j) & 1u))) val = A.Ps[poff[j]] + hrow[NY * j];
      f[j] = val;
    }
    double rd = 1.0, ldiag = 1.0;
#pragma unroll
    for (int j = 0; j < NY; ++j) {
      const double dj = __shfl_sync(FULL, f[j], j);
      okall = okall && (dj > 0.0);
      const double r = fast_rsqrt(dj);
      const double lij = (lane == j) ? dj * r : f[j] * r;
      f[j] = lij;
      if (lane == j) { rd = r; ldiag = lij; }
      // forward substitution w = L^-1 v in step with the factorisation (lane j's x is final at step j): off the critical path
      const double wj = __shfl_sync(FULL, x * r, j);
      if (lane > j) x = fma(-lij, wj, x);
#pragma unroll
      for (int cc = j + 1; cc < NY; ++cc) {
        const double lcj = __shfl_sync(FULL, lij, cc);
        f[cc] = fma(-lij, lcj, f[cc]);
      }
    }
    const double wi = x * rd;
    if (act) {
      A.w[lane] = wi;
      A.rdv[lane] = rd;
#pragma unroll
      for (int j = 0; j < NY; ++j) if (j <= lane) A.F[lane + NY * j] = f[j];
      if (t >= A.presample) {
        qacc = fma(wi, wi, qacc);
        pr *= ldiag;
        const int hi = __double2hiint(pr);
        const int ex = ((hi >> 20) & 0x7ff) - 1023;
        eacc += ex;
        pr = __hiloint2double(hi - (ex << 20), __double2loint(pr));
      }
    }
    if (t >= A.presample) nseen += NY - __popc(miss);
    if (last) break;
    fold_bar();                                      // M, L and the missing-data pattern are ready
    fold_bar();                                      // V and V w are ready
    fold_bar();                                      // P' and a' are ready
  }
  double tot = act ? fma(2.0, fma((double)eacc, 0.6931471805599453, log(pr)), qacc) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
  if (lane == 0) {
    const double ll = -0.5 * (nseen * 1.8378770664093453 + tot);
    if (!okall || !isfinite(ll)) { *A.loglik_d = -INFINITY; *A.status_d = ST_F_NOT_PD; }
    else { *A.loglik_d = ll; *A.status_d = ST_OK; }
  }
}

// blockDim.x = 32 * (NT + 1): warps 0 .. NT - 1 issue the tensor-core products, the last warp does the scalar work.
// NT = nsp / 8 is a template parameter: every fragment address is then base + immediate (the run-time-size version spent
// 11 instructions per DMMA on address arithmetic and predicates; ncu: 6 % of the issued instructions were DMMAs).
// The two roles run their own period loops and meet at the CTA barrier three times per period (bar.sync 0 from two
// program locations); their register sets do not add up.

template <int NT>
__global__ void __launch_bounds__(32 * (NT + 1), (NT <= 3 ? 6 : (NT <= 4 ? 5 : (NT <= 5 ? 4 : (NT <= 8 ? 2 : 1)))))
kalman_dmma_fold_kernel(KalmanGenericArgs g, KalmanDmmaGeom q) {
  extern __shared__ double sm[];
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NSP = 8 * NT;
  constexpr int LD = NSP + 4;      // = dmma_pad_ld(NSP): NSP mod 16 is 0 or 8, so NSP + 4 is 4 or 12 (mod 16)
  const int ns = g.ns, ny = g.ny, nyp = q.nyp, ldu = q.ldu;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  double* Ts = sm;                              // Ts[i + LD * j]
  double* Ps = Ts + LD * NSP;
  double* Ms = Ps + LD * NSP;
  double* V = Ms + LD * NSP;                    // V[q + ldu * s], q < nyp:  (T P Z' L^-T)'
  double* F = V + (size_t)ldu * NSP;            // ny x ny lower Cholesky, column-major
  double* Hs = F + (size_t)ny * ny;             // measurement-error covariance of the draw
  double* abuf = Hs + (size_t)ny * ny;          // 2 x NSP: predicted state, double-buffered
  double* v = abuf + 2 * NSP;
  double* w = v + ny;
  double* rdv = w + ny;                         // 1 / L_ii of the current period
  __shared__ int zi[64];
  __shared__ int miss_sh;
  const int wu = __reduce_max_sync(FULL, warp);   // = warp, as a value that ptxas knows to be warp-uniform
  const bool scalar_warp = (wu == NT);
  const bool small = ny <= 8;

  for (long long d = blockIdx.x; d < g.n_draws; d += gridDim.x) {
    __syncthreads();
    if (g.status_in && g.status_in[d] != ST_OK) {
      if (tid == 0) { g.loglik[d] = -INFINITY; g.status[d] = g.status_in[d]; }
      continue;
    }
    {
      const double* Tm = g.T + (size_t)d * ns * ns;
      const double* P0 = g.P0 + (size_t)d * ns * ns;
      const double* Hm = g.H ? g.H + (size_t)d * ny * ny : nullptr;
      for (int i = tid; i < ny; i += nt) zi[i] = g.z_idx[i];
      for (int j = warp; j < NSP; j += NT + 1)
        for (int i = lane; i < LD; i += 32) {
          const bool in = i < ns && j < ns;
          Ts[i + LD * j] = in ? Tm[i + (size_t)ns * j] : 0.0;
          Ps[i + LD * j] = in ? P0[i + (size_t)ns * j] : 0.0;
          Ms[i + LD * j] = 0.0;
        }
      for (int e = tid; e < ldu * NSP; e += nt) V[e] = 0.0;
      for (int e = tid; e < ny * ny; e += nt) Hs[e] = Hm ? Hm[e] : 0.0;
      for (int i = tid; i < 2 * NSP; i += nt) abuf[i] = (g.a0 && i < ns) ? g.a0[(size_t)d * ns + i] : 0.0;
    }
    __syncthreads();

    if (scalar_warp) {
      // ================= scalar role: innovation, F = P[z,z] + H, L = chol(F), w = L^-1 v, likelihood terms =================
      if (small) {
        FoldScalarArgs A{Ps, Hs, abuf, F, w, rdv, &miss_sh, zi, g.Y, g.ybar_obs ? g.ybar_obs + (size_t)d * ny : nullptr,
                         NSP, LD, g.nobs, g.presample, g.loglik + d, g.status + d};
        switch (ny) {
          case 1: fold_scalar_role<1>(A); break;
          case 2: fold_scalar_role<2>(A); break;
          case 3: fold_scalar_role<3>(A); break;
          case 4: fold_scalar_role<4>(A); break;
          case 5: fold_scalar_role<5>(A); break;
          case 6: fold_scalar_role<6>(A); break;
          case 7: fold_scalar_role<7>(A); break;
          default: fold_scalar_role<8>(A); break;
        }
      } else {
        double acc = 0.0;
        int nseen = 0, bad = 0;
        const double ybi = (lane < ny && g.ybar_obs) ? g.ybar_obs[(size_t)d * ny + lane] : 0.0;
        double ynext = (lane < ny && g.nobs > 0) ? g.Y[lane] : 0.0;
#pragma unroll 1
        for (int t = 0; t < g.nobs; ++t) {
          const double* a = abuf + (t & 1) * NSP;
          const bool last = (t == g.nobs - 1);
          const double yi = ynext;
          if (lane < ny && !last) ynext = g.Y[(size_t)(t + 1) * ny + lane];
          const unsigned miss = __ballot_sync(FULL, lane < ny && isnan(yi));
          if (lane == 0) miss_sh = (int)miss;
          if (lane < ny) v[lane] = isnan(yi) ? 0.0 : (yi - ybi) - a[zi[lane]];
          for (int e0 = 0; e0 < ny * ny; e0 += 32) {
            const int e = e0 + lane;
            if (e < ny * ny) {
              const int jj = e / ny, ii = e - jj * ny;
              double f = Ps[zi[ii] + LD * zi[jj]] + Hs[e];
              if (((miss >> ii) | (miss >> jj)) & 1u) f = (ii == jj) ? 1.0 : 0.0;
              F[e] = f;
            }
          }
          __syncwarp();
          double ql[2];
          bool okc;
          if (ny <= 16) okc = chol_warp_call<16>(F, ny, v, w, rdv, ql);
          else if (ny <= 24) okc = chol_warp_call<24>(F, ny, v, w, rdv, ql);
          else okc = chol_warp_call<32>(F, ny, v, w, rdv, ql);
          if (!okc) bad = 1;
          if (t >= g.presample) { acc += 2.0 * ql[1] + ql[0]; nseen += ny - __popc(miss); }
          if (last) break;
          fold_bar();                                      // M, L and the missing-data pattern are ready
          fold_bar();                                      // V and V w are ready
          fold_bar();                                      // P' and a' are ready
        }
        if (lane == 0) {
          const double ll = -0.5 * (nseen * 1.8378770664093453 + acc);
          if (bad || !isfinite(ll)) { g.loglik[d] = -INFINITY; g.status[d] = ST_F_NOT_PD; }
          else { g.loglik[d] = ll; g.status[d] = ST_OK; }
        }
      }
    } else {
      // ================= tensor-core role ======================================================================================
      // strip of this warp: for NT <= 4 rotated with the CTA index, so that the warps with the most tiles of the triangular
      // second product (strip NT - 1: NT tiles, strip 0: one) do not all sit on the same SM sub-partition (warp w of every
      // CTA does); measured: ns = 24 / 30 6 % / 3 % faster, ns = 40 (NT = 5: five tensor warps already straddle the four
      // sub-partitions) 1.4 % slower, hence the limit
      const int ws = __reduce_max_sync(FULL, NT <= 4 ? (warp + (int)(blockIdx.x % NT)) % NT : warp);
      const int gq = lane >> 2, tq = lane & 3, i0 = 8 * ws;
      const int vtid = 32 * ws + lane;                  // rows of V: the strips with the fewest tiles first
      const double* aT = Ts + (i0 + gq) + LD * tq;      // A fragment of T:  aT[LD k0]
      const double* bP = Ps + tq + LD * gq;             // B fragment of P:  bP[k0 + 8 LD jt]
      const double* aM = Ms + (i0 + gq) + LD * tq;      // A fragment of M:  aM[LD k0]
      const double* bT = Ts + gq + LD * tq;             // B fragment of T': bT[8 jt + LD k0]
      double* sM = Ms + (i0 + gq) + LD * 2 * tq;        // accumulator pair of tile jt: sM[8 LD jt], sM[8 LD jt + LD]
      double* sP = Ps + (i0 + gq) + LD * 2 * tq;
      double* sPm = Ps + 2 * tq + LD * (i0 + gq);       // mirrored: sPm[8 jt], sPm[8 jt + 1]
      const double* vA = V + tq + ldu * (i0 + gq);
      const double* vB = V + tq + ldu * gq;
      const int qi = i0 + gq, qj = 2 * tq;              // R Q R' elements of this thread's accumulators
      const double* qp = g.RQR + (size_t)d * ns * ns + qi + (size_t)ns * qj;
      const double* tr = Ts + (tid >> 2) + LD * (tid & 3);
#pragma unroll 1
      for (int t = 0; t + 1 < g.nobs; ++t) {
        const double* a = abuf + (t & 1) * NSP;
        double* anew = abuf + ((t + 1) & 1) * NSP;
        double c[NT][2];
        // ---- M = T P ----------------------------------------------------------------------------------------------------
#pragma unroll
        for (int jt = 0; jt < NT; ++jt) { c[jt][0] = 0.0; c[jt][1] = 0.0; }
#pragma unroll
        for (int k0 = 0; k0 < NSP; k0 += 4) {
          const double av = aT[LD * k0];
#pragma unroll
          for (int jt = 0; jt < NT; ++jt) dmma884(c[jt][0], c[jt][1], av, bP[k0 + 8 * LD * jt]);
        }
#pragma unroll
        for (int jt = 0; jt < NT; ++jt) { sM[8 * LD * jt] = c[jt][0]; sM[8 * LD * jt + LD] = c[jt][1]; }
        // R Q R' tiles of the second product: the loads fly while the barrier is waited for
#pragma unroll
        for (int jt = 0; jt < NT; ++jt) {
          c[jt][0] = 0.0; c[jt][1] = 0.0;
          if (jt <= ws) {
            const int j = 8 * jt + qj;
            if (qi < ns && j < ns) c[jt][0] = qp[(size_t)ns * (8 * jt)];
            if (qi < ns && j + 1 < ns) c[jt][1] = qp[(size_t)ns * (8 * jt + 1)];
          }
        }
        fold_bar();                                      // M, L and the missing-data pattern are ready
        // ---- V = M[:, z] L^-T: thread s < ns, i.e. the warps with the fewest tiles of the second product --------------------
        if (vtid < ns) {
          const unsigned ms = (unsigned)miss_sh;
          const double xa = small ? fold_gain_row<8>(vtid, ny, LD, ldu, ms, zi, Ms, F, rdv, w, V)
                                  : fold_gain_row<0>(vtid, ny, LD, ldu, ms, zi, Ms, F, rdv, w, V);
          anew[vtid] = xa;
        }
        // ---- T a: four lanes per row ----------------------------------------------------------------------------------------
        double ta = 0.0;
#pragma unroll
        for (int p = 0; p < NSP; p += 4) ta = fma(tr[LD * p], a[p + (tid & 3)], ta);
        ta += __shfl_xor_sync(FULL, ta, 1);
        ta += __shfl_xor_sync(FULL, ta, 2);
        // ---- P' = M T' + R Q R'  (tiles j <= warp) -----------------------------------------------------------------------
#pragma unroll
        for (int k0 = 0; k0 < NSP; k0 += 4) {
          const double av = aM[LD * k0];
#pragma unroll
          for (int jt = 0; jt < NT; ++jt)
            if (jt <= ws) dmma884(c[jt][0], c[jt][1], av, bT[8 * jt + LD * k0]);
        }
        fold_bar();                                      // V and V w are ready
        if ((tid & 3) == 0 && (tid >> 2) < ns) anew[tid >> 2] += ta;
        // ---- P' -= V V', stored mirrored ------------------------------------------------------------------------------------
        for (int k0 = 0; k0 < nyp; k0 += 4) {
          const double av = -vA[k0];
#pragma unroll
          for (int jt = 0; jt < NT; ++jt)
            if (jt <= ws) dmma884(c[jt][0], c[jt][1], av, vB[k0 + 8 * ldu * jt]);
        }
#pragma unroll
        for (int jt = 0; jt < NT; ++jt)
          if (jt <= ws) {
            if (jt < ws) {
              sP[8 * LD * jt] = c[jt][0]; sP[8 * LD * jt + LD] = c[jt][1];
              sPm[8 * jt] = c[jt][0]; sPm[8 * jt + 1] = c[jt][1];
            } else {                            // diagonal tile: keep the lower triangle, mirror it
              if (gq >= 2 * tq) { sP[8 * LD * jt] = c[jt][0]; sPm[8 * jt] = c[jt][0]; }
              if (gq >= 2 * tq + 1) { sP[8 * LD * jt + LD] = c[jt][1]; sPm[8 * jt + 1] = c[jt][1]; }
            }
          }
        fold_bar();                                      // P' and a' are ready
      }
    }
  }
}

}  // namespace dyb
