// Dense batched Kalman log-likelihood for LARGE state spaces (88 < ns <= 256: BASELINE config 5, ns = 200, ny = 20)
// on the fp64 tensor cores.  P and T P of a draw do not fit in shared memory (2 x 320 KB at ns = 200), so they live
// in a per-CTA global scratch that stays L2-resident; one persistent CTA per SM walks over the draws.  The two big
// products of a period,  M = T P  and  P = M T' + R Q R'  (lower block triangle, mirrored), are blocked GEMMs:
// k-panels of 8 columns of A (all rows) and 8 rows of B (a pass of up to NTP tile columns) are staged through shared
// memory with register prefetch (global loads of panel p+1 are in flight while panel p is multiplied), every warp owns
// up to two 8-row strips of the result and keeps their accumulators in registers, fragments are read conflict-free
// (leading dimensions = 4 or 12 mod 16) and multiplied with mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).
// The rank-ny downdate P -= U'U uses the same instruction with U in shared memory.
//
// Replaces `kalman_likelihood(Y,Z,H,T,R,Q,a0,P,start,last,presample,ws)` (reference call site
// src/estimation/estimation.jl:687-700) for a selection matrix Z; arguments shared with kalman_generic.cuh.
#pragma once
#include "kalman_dmma.cuh"

namespace dyb {

constexpr int DL_BK = 8;          // k-panel depth
constexpr int DL_NTP = 9;         // tile columns per pass (accumulators: 2 strips x 9 tiles x 2 doubles per lane)
constexpr int DL_LDB = 12;        // leading dimension of the staged B panel (8 rows, = 12 mod 16)

struct KalmanLargeGeom { int nsp, lda, nyp, ldu, nwarps, nxp; };

inline KalmanLargeGeom kalman_large_geom(int ns, int ny) {
  KalmanLargeGeom q;
  q.nsp = (ns + 7) / 8 * 8; q.lda = dmma_pad_ld(q.nsp);
  q.nyp = (ny + 3) / 4 * 4; q.ldu = dmma_pad_ld(q.nyp);
  q.nwarps = (q.nsp / 8 + 1) / 2;
  q.nxp = (ny + 1 + 7) / 8 * 8;          // extension columns of the bulk path: U' (ny), a + U'w (1), zero padding
  return q;
}
inline size_t kalman_large_smem_doubles(int ns, int ny, bool bulk = false);
inline size_t kalman_large_scratch_doubles(int ns, int ny) {
  const KalmanLargeGeom q = kalman_large_geom(ns, ny);
  return 2 * (size_t)q.nsp * (q.nsp + q.nxp) + (size_t)q.nsp * q.nxp;      // [P | U' au] , [T P | K a'] , -K
}

// C = A B (+ Cinit) on the nsp x nsp tile grid, block-cooperative.
//   A(i,k) = Ag[i + lda_g k]                         (valid for i, k < nv; zero beyond)
//   B(k,j) = BT ? Bg[j + ldb_g k] : Bg[k + ldb_g j]  (same validity)
//   SYMM: only tiles on or below the block diagonal are computed, the result is mirrored (C symmetric)
//   Cinit (may be null): nv x nv column-major with leading dimension nv, added to the valid part
// (no __restrict__ on purpose: A / B may have been written earlier in the same kernel, the read-only data path must not
// be used for them)
template <bool BT, bool SYMM>
DYB_D void block_gemm_dmma(const double* Ag, int lda_g, const double* Bg, int ldb_g, int nv,
                           int nsp, int lda, const double* Cinit, double* Cg, int ldc,
                           double* As0, double* Bs0, int panel_doubles) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int nw = nt >> 5, ntile = nsp / 8, npanel = nsp / DL_BK;
  constexpr int MAXR = 8;                 // staged doubles per thread and panel (>= (nsp*8 + 8*72) / threads)
  const int bs_off = (int)(Bs0 - As0);    // the B panel follows the A panel inside a stage buffer
  for (int jt0 = 0; jt0 < ntile; jt0 += DL_NTP) {
    const int ntw = (ntile - jt0 < DL_NTP) ? ntile - jt0 : DL_NTP;
    const int ncolw = 8 * ntw, jbase = 8 * jt0;
    const int nA = nsp * DL_BK, nB = DL_BK * ncolw;
    // tiles of strip s that this pass computes form a prefix 0 .. nact[s]-1 (SYMM: up to the block diagonal)
    int nact[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int si = warp + s * nw;
      int na = si < ntile ? ntw : 0;
      if (SYMM && si < ntile) { const int lim = si - jt0 + 1; na = lim < na ? (lim > 0 ? lim : 0) : na; }
      nact[s] = na;
    }
    double c[2][DL_NTP][2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int si = warp + s * nw;
#pragma unroll
      for (int jt = 0; jt < DL_NTP; ++jt) {
        c[s][jt][0] = 0.0; c[s][jt][1] = 0.0;
        if (Cinit && jt < nact[s]) {
          const int i = 8 * si + gq, j = jbase + 8 * jt + 2 * tq;
          if (i < nv && j < nv) c[s][jt][0] = Cinit[i + (size_t)nv * j];
          if (i < nv && j + 1 < nv) c[s][jt][1] = Cinit[i + (size_t)nv * (j + 1)];
        }
      }
    }
    // staging descriptors of this thread, fixed for the pass: global offset at panel 0 (-1: always zero), k index inside
    // the panel, shared-memory offset, A or B
    int gofs[MAXR], sofs[MAXR];
    unsigned kks = 0, isb = 0;
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      const int e = tid + r * nt;
      int go = -1, so = -1, kk = 0;
      if (e < nA) {
        const int i = e % nsp; kk = e / nsp;
        so = i + lda * kk;
        if (i < nv) go = i + lda_g * kk;
      } else if (e < nA + nB) {
        const int f = e - nA;
        int n;
        if (BT) { n = f % ncolw; kk = f / ncolw; } else { kk = f % DL_BK; n = f / DL_BK; }
        so = bs_off + kk + DL_LDB * n;
        isb |= 1u << r;
        if (jbase + n < nv) go = BT ? (jbase + n) + ldb_g * kk : kk + ldb_g * (jbase + n);
      }
      gofs[r] = go; sofs[r] = so; kks |= (unsigned)kk << (4 * r);
    }
    const int stepA = lda_g * DL_BK, stepB = BT ? ldb_g * DL_BK : DL_BK;
    double stage[MAXR];
    auto fetch = [&](int p) {
      const int k0 = p * DL_BK;
#pragma unroll
      for (int r = 0; r < MAXR; ++r) {
        double v = 0.0;
        const int kk = (kks >> (4 * r)) & 15;
        if (gofs[r] >= 0 && k0 + kk < nv)
          v = ((isb >> r) & 1) ? Bg[(size_t)gofs[r] + (size_t)p * stepB] : Ag[(size_t)gofs[r] + (size_t)p * stepA];
        stage[r] = v;
      }
    };
    auto commit = [&](int buf) {
      double* dst = As0 + buf * panel_doubles;
#pragma unroll
      for (int r = 0; r < MAXR; ++r) if (sofs[r] >= 0) dst[sofs[r]] = stage[r];
    };
    const bool full = (nact[0] == DL_NTP) && (nact[1] == DL_NTP || nact[1] == 0);
    const bool two = nact[1] > 0;
    fetch(0);
    commit(0);
    __syncthreads();
    for (int p = 0; p < npanel; ++p) {
      if (p + 1 < npanel) fetch(p + 1);
      const double* As = As0 + (p & 1) * panel_doubles;
      const double* Bs = Bs0 + (p & 1) * panel_doubles;
      const double* Ap0 = As + (8 * warp + gq) + lda * tq;
      const double* Ap1 = Ap0 + 8 * nw;
      const double* Bp = Bs + tq + DL_LDB * gq;
      if (full) {                                   // warp-uniform: no per-tile predicates in the common case
#pragma unroll
        for (int ks = 0; ks < DL_BK / 4; ++ks) {
          const double a0 = Ap0[lda * 4 * ks];
          const double a1 = two ? Ap1[lda * 4 * ks] : 0.0;
#pragma unroll
          for (int jt = 0; jt < DL_NTP; ++jt) {
            const double b = Bp[4 * ks + DL_LDB * 8 * jt];
            dmma884(c[0][jt][0], c[0][jt][1], a0, b);
            if (two) dmma884(c[1][jt][0], c[1][jt][1], a1, b);
          }
        }
      } else {
#pragma unroll
        for (int ks = 0; ks < DL_BK / 4; ++ks) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            if (nact[s] > 0) {
              const double av = (s == 0 ? Ap0 : Ap1)[lda * 4 * ks];
#pragma unroll
              for (int jt = 0; jt < DL_NTP; ++jt)
                if (jt < nact[s]) dmma884(c[s][jt][0], c[s][jt][1], av, Bp[4 * ks + DL_LDB * 8 * jt]);
            }
          }
        }
      }
      if (p + 1 < npanel) commit((p + 1) & 1);
      __syncthreads();
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int si = warp + s * nw;
#pragma unroll
      for (int jt = 0; jt < DL_NTP; ++jt) {
        if (jt >= nact[s]) continue;
        const int i = 8 * si + gq, j = jbase + 8 * jt + 2 * tq;
        if (!SYMM) {
          Cg[i + (size_t)ldc * j] = c[s][jt][0]; Cg[i + (size_t)ldc * (j + 1)] = c[s][jt][1];
        } else if (jt0 + jt < si) {
          Cg[i + (size_t)ldc * j] = c[s][jt][0]; Cg[i + (size_t)ldc * (j + 1)] = c[s][jt][1];
          Cg[j + (size_t)ldc * i] = c[s][jt][0]; Cg[(j + 1) + (size_t)ldc * i] = c[s][jt][1];
        } else {
          if (i >= j) { Cg[i + (size_t)ldc * j] = c[s][jt][0]; Cg[j + (size_t)ldc * i] = c[s][jt][0]; }
          if (i >= j + 1) { Cg[i + (size_t)ldc * (j + 1)] = c[s][jt][1]; Cg[(j + 1) + (size_t)ldc * i] = c[s][jt][1]; }
        }
      }
    }
    __syncthreads();
  }
}

// ---- the same blocked GEMM with the panels moved by the bulk-copy engine (cp.async.bulk = SASS UBLKCP, completion on an
// mbarrier): no thread spends instructions on staging.  Needs ns = nsp (a multiple of 8, so that every column of an
// operand is a 16-byte aligned contiguous run); k-panels of 16.  Layout of a stage: A panel As[i + lda kk] (16 columns
// of nsp rows), B panel either Bs[kk + 20 n] (B read by columns: runs of 16 k-values) or BsT[n + ldn kk] (B = T': runs
// of the pass's n values), both conflict-free for the fragment loads.
constexpr int DB_BK = 32;
constexpr int DB_LDB = 36;                 // >= DB_BK, = 4 mod 16
DYB_HD int bulk_stage_doubles(int nsp) {
  return nsp * DB_BK + dmma_pad_ld(8 * DL_NTP) * DB_BK;        // A panel (leading dimension nsp) + B panel by rows
}

DYB_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
DYB_D void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
DYB_D void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DYB_D void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
DYB_D void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// k-steps of one staged panel for a warp whose strips own N (<= NTP) tiles of the pass: strip 1 (if TWO) all N tiles,
// strip 0 its first n0 <= N tiles (n0 warp-uniform).  N is a template parameter so that the B fragments of a k-step are
// loaded together ahead of an unpredicated DMMA sequence.
template <int N, bool TWO>
DYB_D void dmma_panel_steps(double (&c)[2][DL_NTP][2], int n0, const double* Ap0, const double* Ap1, const double* Bp,
                            int nsp, int ldn, int kc) {
#pragma unroll 2
  for (int ks = 0; ks < kc / 4; ++ks) {
    double b[N];
#pragma unroll
    for (int jt = 0; jt < N; ++jt) b[jt] = Bp[4 * ldn * ks + 8 * jt];
    if (TWO) {
      const double a1 = Ap1[nsp * 4 * ks];
#pragma unroll
      for (int jt = 0; jt < N; ++jt) dmma884(c[1][jt][0], c[1][jt][1], a1, b[jt]);
    }
    if (n0 == N) {
      const double a0 = Ap0[nsp * 4 * ks];
#pragma unroll
      for (int jt = 0; jt < N; ++jt) dmma884(c[0][jt][0], c[0][jt][1], a0, b[jt]);
    } else if (n0 > 0) {
      const double a0 = Ap0[nsp * 4 * ks];
#pragma unroll
      for (int jt = 0; jt < N; ++jt)
        if (jt < n0) dmma884(c[0][jt][0], c[0][jt][1], a0, b[jt]);
    }
  }
}

// C = A B (+ Cinit) with C nsp x ncols and inner dimension kdim (all multiples of 8):
//   A(i,k) = Ag[i + nsp k]  -- the leading dimension IS nsp, so a k-panel is ONE contiguous run and ONE bulk copy
//            (the copy engine's cost is per copy, not per byte: 104 copies per panel cost ~7.5k cycles);
//            the staged panel keeps leading dimension nsp, which costs a 2- or 4-way bank conflict on the two A
//            fragment loads of a k-step (the nine B loads stay conflict-free)
//   B(k,j) = k < ksplit ? Bg[j + ldb_g k] : Bg2[j + ldb_g2 (k - ksplit)]   (B given by rows: one copy per k)
//   SYMM: only tiles on or below the block diagonal are computed, the result is mirrored
//   Cneg (may be null): -C is also written there (leading dimension nsp), its column zero_col as zero
// The filter's rank-ny correction rides inside the products:  K = T U' and a' = T (a + U'w) are one thin product
// T [U' | a + U'w], and  P' = [T P | K] [T | -K]' + R Q R'  is one product with inner dimension nsp + nxp.
// SYMM is a run-time flag and the kernel calls this function from ONE site inside a loop over its three products: three
// inlined copies of the tile-path switch were ~400 KB of code and the kernel stalled on instruction fetch.
DYB_D void block_gemm_dmma_bulk(const bool SYMM, const double* Ag, const double* Bg, int ldb_g, const double* Bg2, int ldb_g2,
                                int ksplit, int nsp, int ncols, int kdim,
                                const double* Cinit, double* Cg, int ldc, double* Cneg, int zero_col,
                                double* stage0, int stage_doubles, unsigned long long* bars, unsigned& phases) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int nw = nt >> 5, ntm = nsp / 8, ntile = ncols / 8, npanel = (kdim + DB_BK - 1) / DB_BK;
  const int ldn = dmma_pad_ld(8 * DL_NTP);
  // operands may have been written with ordinary stores by this CTA: make them visible to the bulk-copy (async) proxy
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  for (int jt0 = 0; jt0 < ntile; jt0 += DL_NTP) {
    const int ntw = (ntile - jt0 < DL_NTP) ? ntile - jt0 : DL_NTP;
    const int ncolw = 8 * ntw, jbase = 8 * jt0;
    int nact[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int si = warp + s * nw;
      int na = si < ntm ? ntw : 0;
      if (SYMM && si < ntm) { const int lim = si - jt0 + 1; na = lim < na ? (lim > 0 ? lim : 0) : na; }
      nact[s] = na;
    }
    double c[2][DL_NTP][2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int si = warp + s * nw;
#pragma unroll
      for (int jt = 0; jt < DL_NTP; ++jt) {
        c[s][jt][0] = 0.0; c[s][jt][1] = 0.0;
        if (Cinit && jt < nact[s]) {
          const int i = 8 * si + gq, j = jbase + 8 * jt + 2 * tq;
          c[s][jt][0] = Cinit[i + (size_t)nsp * j];
          c[s][jt][1] = Cinit[i + (size_t)nsp * (j + 1)];
        }
      }
    }
    // producer: warp 0 issues the copies of panel p into stage (p & 1)
    auto issue = [&](int p) {
      if (warp != 0) return;
      const int k0 = p * DB_BK, kc = (kdim - k0 < DB_BK) ? kdim - k0 : DB_BK;
      double* As = stage0 + (size_t)(p & 1) * stage_doubles;
      double* Bs = As + nsp * DB_BK;
      unsigned long long* bar = bars + (p & 1);
      if (lane == 0) {
        mbar_expect_tx(bar, (unsigned)(8 * (kc * nsp + kc * ncolw)));
        bulk_g2s(As, Ag + (size_t)nsp * k0, 8u * nsp * kc, bar);
      }
      __syncwarp();
      for (int kk = lane; kk < kc; kk += 32) {
        const int k = k0 + kk;
        const double* src = k < ksplit ? Bg + jbase + (size_t)ldb_g * k : Bg2 + jbase + (size_t)ldb_g2 * (k - ksplit);
        bulk_g2s(Bs + ldn * kk, src, 8u * ncolw, bar);
      }
    };
    // warp-uniform by construction; taken from a vote so that ptxas knows it too and the mma.sync sequences below are
    // not interleaved with WARPSYNC / NOP pairs
    const bool full = __all_sync(0xffffffffu, (nact[0] == DL_NTP) && (nact[1] == DL_NTP || nact[1] == 0));
    const bool two = __all_sync(0xffffffffu, nact[1] > 0);
    const int n0u = __reduce_max_sync(0xffffffffu, nact[0]), n1u = __reduce_max_sync(0xffffffffu, nact[1]);
    issue(0);
    for (int p = 0; p < npanel; ++p) {
      const int st = p & 1;
      if (p + 1 < npanel) issue(p + 1);           // its stage was released by the barrier that ended iteration p - 1
      mbar_wait(bars + st, (phases >> st) & 1u);
      phases ^= 1u << st;
      const int k0 = p * DB_BK, kc = (kdim - k0 < DB_BK) ? kdim - k0 : DB_BK;
      const double* As = stage0 + (size_t)st * stage_doubles;
      const double* Bs = As + nsp * DB_BK;
      {
      const double* Ap0 = As + (8 * warp + gq) + nsp * tq;
      const double* Ap1 = Ap0 + 8 * nw;
      const double* Bp = Bs + gq + ldn * tq;
      if (full && two) {
#pragma unroll
        for (int ks = 0; ks < DB_BK / 4; ++ks) {
          if (4 * ks < kc) {
            const double a0 = Ap0[nsp * 4 * ks];
            const double a1 = Ap1[nsp * 4 * ks];
#pragma unroll
            for (int jt = 0; jt < DL_NTP; ++jt) {
              const double b = Bp[4 * ldn * ks + 8 * jt];
              dmma884(c[0][jt][0], c[0][jt][1], a0, b);
              dmma884(c[1][jt][0], c[1][jt][1], a1, b);
            }
          }
        }
      } else if (full) {
#pragma unroll
        for (int ks = 0; ks < DB_BK / 4; ++ks) {
          if (4 * ks < kc) {
            const double a0 = Ap0[nsp * 4 * ks];
#pragma unroll
            for (int jt = 0; jt < DL_NTP; ++jt) dmma884(c[0][jt][0], c[0][jt][1], a0, Bp[4 * ldn * ks + 8 * jt]);
          }
        }
      } else if (n1u >= n0u && n1u > 0) {           // two strips, strip 1 owns at least as many tiles (always so here)
        switch (n1u) {
          case 1: dmma_panel_steps<1, true>(c, n0u, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 2: dmma_panel_steps<2, true>(c, n0u, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 3: dmma_panel_steps<3, true>(c, n0u, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 4: dmma_panel_steps<4, true>(c, n0u, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 5: dmma_panel_steps<5, true>(c, n0u, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 6: dmma_panel_steps<6, true>(c, n0u, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 7: dmma_panel_steps<7, true>(c, n0u, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 8: dmma_panel_steps<8, true>(c, n0u, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          default: dmma_panel_steps<9, true>(c, n0u, Ap0, Ap1, Bp, nsp, ldn, kc); break;
        }
      } else if (n1u == 0 && n0u > 0) {             // one strip
        switch (n0u) {
          case 1: dmma_panel_steps<1, false>(c, 1, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 2: dmma_panel_steps<2, false>(c, 2, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 3: dmma_panel_steps<3, false>(c, 3, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 4: dmma_panel_steps<4, false>(c, 4, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 5: dmma_panel_steps<5, false>(c, 5, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 6: dmma_panel_steps<6, false>(c, 6, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 7: dmma_panel_steps<7, false>(c, 7, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          case 8: dmma_panel_steps<8, false>(c, 8, Ap0, Ap1, Bp, nsp, ldn, kc); break;
          default: dmma_panel_steps<9, false>(c, 9, Ap0, Ap1, Bp, nsp, ldn, kc); break;
        }
      } else if (n0u > 0) {                         // strip 0 owns more tiles than strip 1: predicated fallback
#pragma unroll
        for (int ks = 0; ks < DB_BK / 4; ++ks) {
          if (4 * ks < kc) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              if (nact[s] > 0) {
                const double av = (s == 0 ? Ap0 : Ap1)[nsp * 4 * ks];
#pragma unroll
                for (int jt = 0; jt < DL_NTP; ++jt)
                  if (jt < nact[s]) dmma884(c[s][jt][0], c[s][jt][1], av, Bp[4 * ldn * ks + 8 * jt]);
              }
            }
          }
        }
      }
      }
      __syncthreads();                              // every warp is done with stage st
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int si = warp + s * nw;
#pragma unroll
      for (int jt = 0; jt < DL_NTP; ++jt) {
        if (jt >= nact[s]) continue;
        const int i = 8 * si + gq, j = jbase + 8 * jt + 2 * tq;
        if (!SYMM) {
          Cg[i + (size_t)ldc * j] = c[s][jt][0]; Cg[i + (size_t)ldc * (j + 1)] = c[s][jt][1];
          if (Cneg) {
            Cneg[i + (size_t)nsp * j] = j == zero_col ? 0.0 : -c[s][jt][0];
            Cneg[i + (size_t)nsp * (j + 1)] = j + 1 == zero_col ? 0.0 : -c[s][jt][1];
          }
        } else if (jt0 + jt < si) {
          Cg[i + (size_t)ldc * j] = c[s][jt][0]; Cg[i + (size_t)ldc * (j + 1)] = c[s][jt][1];
          Cg[j + (size_t)ldc * i] = c[s][jt][0]; Cg[(j + 1) + (size_t)ldc * i] = c[s][jt][1];
        } else {
          if (i >= j) { Cg[i + (size_t)ldc * j] = c[s][jt][0]; Cg[j + (size_t)ldc * i] = c[s][jt][0]; }
          if (i >= j + 1) { Cg[i + (size_t)ldc * (j + 1)] = c[s][jt][1]; Cg[(j + 1) + (size_t)ldc * i] = c[s][jt][1]; }
        }
      }
    }
  }
}

inline size_t kalman_large_smem_doubles(int ns, int ny, bool bulk) {
  const KalmanLargeGeom q = kalman_large_geom(ns, ny);
  const size_t stage = bulk ? (size_t)bulk_stage_doubles(q.nsp) : (size_t)q.lda * DL_BK + (size_t)DL_LDB * 8 * DL_NTP;
  return 2 * stage + (size_t)q.ldu * q.nsp + (size_t)ny * ny + 2 * (size_t)q.nsp + 3 * (size_t)ny + 8;
}

// U[:, s] = L^-1 (Z P)[:, s] and au = a + U'w, one state s per thread, with the column in registers: right-looking
// substitution whose body has static register indices (the pending right-hand sides shift one place per step, as in
// chol_warp_rot).  Writes U' and a + U'w by rows into PxT for the thin product.
template <int NY>
DYB_D void usolve_regs(int tid, int nt, int ns, int ny, const double* yt, const int* zi, const double* P, int ldP,
                       const double* F, const double* rdv, const double* w, const double* a, double* au,
                       double* PxT, int ldx) {
  for (int s = tid; s < ns; s += nt) {
    double x[NY];
#pragma unroll
    for (int i = 0; i < NY; ++i) x[i] = (i < ny && !isnan(yt[i])) ? P[zi[i] + (size_t)ldP * s] : 0.0;
    double acc = a[s];
    double* row = PxT + (size_t)ldx * s;
#pragma unroll 1
    for (int i = 0; i < ny; ++i) {
      const double u = x[0] * rdv[i];
      row[i] = u;
      acc = fma(u, w[i], acc);
      const double* Fi = F + i + (size_t)ny * i;                 // F[i + c, i] = Fi[c]
#pragma unroll
      for (int c = 1; c < NY; ++c) x[c - 1] = (i + c < ny) ? fma(-Fi[c], u, x[c]) : 0.0;
      x[NY - 1] = 0.0;
    }
    au[s] = acc;
    row[ny] = acc;
  }
}

// blockDim.x = 32 * nwarps (nwarps = ceil(ntile / 2) <= 16); g.scratch: kalman_large_scratch_doubles per CTA
template <bool BULK>
__global__ void __launch_bounds__(512, 1) kalman_dmma_large_kernel(KalmanGenericArgs g, KalmanLargeGeom q) {
  extern __shared__ __align__(16) double sm[];
  const int ns = g.ns, ny = g.ny, nsp = q.nsp, lda = q.lda, nyp = q.nyp, ldu = q.ldu;
  const int ntile = nsp / 8;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3, nw = nt >> 5;
  const int panel_doubles = BULK ? bulk_stage_doubles(nsp) : lda * DL_BK + DL_LDB * 8 * DL_NTP;
  double* As0 = sm;                                  // two stages: [As | Bs] [As | Bs]
  double* Bs0 = sm + lda * DL_BK;
  double* U = sm + 2 * panel_doubles;                // U[q + ldu * s], q < nyp
  __shared__ __align__(8) unsigned long long bars[2];
  unsigned phases = 0;                               // parity of the next completion of each stage barrier
  if (BULK) {
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
  }
  double* F = U + (size_t)ldu * nsp;
  double* a = F + (size_t)ny * ny;
  double* au = a + nsp;
  double* v = au + nsp;
  double* w = v + ny;
  double* rdv = w + ny;
  __shared__ int bad;
  __shared__ int zi[64];
  const int nxp = q.nxp;
  double* P = g.scratch + (size_t)blockIdx.x * (2 * (size_t)nsp * (nsp + nxp) + (size_t)nsp * nxp);   // P[i + nsp * j]
  double* Mx = P + (size_t)nsp * (nsp + nxp);        // T P, then (bulk path) K = T U' and a' in the extension columns
  double* Kneg = Mx + (size_t)nsp * (nsp + nxp);     // -K, the extra rows of the second product's B operand
  double* PxT = P + (size_t)nsp * nsp;               // [U' | a + U'w | 0] by rows: PxT[q + nxp s]

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
    #pragma unroll 1
    for (int i = tid; i < ny; i += nt) zi[i] = g.z_idx[i];
    #pragma unroll 1
    for (int e = tid; e < nsp * nsp; e += nt) {
      const int i = e % nsp, j = e / nsp;
      // lower triangle mirrored: the bulk path reads P by rows (P = P')
      P[e] = (i < ns && j < ns) ? (i >= j ? P0[i + (size_t)ns * j] : P0[j + (size_t)ns * i]) : 0.0;
    }
    #pragma unroll 1
    for (int e = tid; e < ldu * nsp; e += nt) U[e] = 0.0;
    if (BULK) for (int e = tid; e < nsp * nxp; e += nt) PxT[e] = 0.0;
    #pragma unroll 1
    for (int i = tid; i < nsp; i += nt) { a[i] = (g.a0 && i < ns) ? g.a0[(size_t)d * ns + i] : 0.0; au[i] = 0.0; }
    __syncthreads();
    double acc = 0.0;          // only thread 0's copy is used
    int nseen = 0;

#ifdef DYB_KL_PROFILE
    long long pc[6] = {0, 0, 0, 0, 0, 0}, c0 = 0;
#define KL_MARK(k) { __syncthreads(); const long long c1 = clock64(); pc[k] += c1 - c0; c0 = c1; }
#else
#define KL_MARK(k)
#endif
    for (int t = 0; t < g.nobs; ++t) {
#ifdef DYB_KL_PROFILE
      __syncthreads(); c0 = clock64();
#endif
      const double* yt = g.Y + (size_t)t * ny;
      #pragma unroll 1
      for (int i = tid; i < ny; i += nt) {
        const double yi = yt[i];
        v[i] = isnan(yi) ? 0.0 : (yi - (yb ? yb[i] : 0.0)) - a[zi[i]];
      }
      #pragma unroll 1
      for (int e = tid; e < ny * ny; e += nt) {
        const int i = e % ny, j = e / ny;
        const bool mi = isnan(yt[i]), mj = isnan(yt[j]);
        double f = P[zi[i] + (size_t)nsp * zi[j]] + (Hm ? Hm[e] : 0.0);
        if (mi || mj) f = (i == j) ? 1.0 : 0.0;
        F[e] = f;
      }
      __syncthreads();
      KL_MARK(4)
      if (tid < 32 && ny <= 32) {             // Cholesky of F and w = L^-1 v by warp 0, F in registers (one row per lane)
        double ql[2];
        bool okc;
#ifdef DYB_KL_PROFILE
        const long long cc0 = clock64();
#endif
        if (ny <= 8) okc = chol_warp_call<8>(F, ny, v, w, rdv, ql);
        else if (ny <= 16) okc = chol_warp_call<16>(F, ny, v, w, rdv, ql);
        else if (ny <= 24) okc = chol_warp_call<24>(F, ny, v, w, rdv, ql);
        else okc = chol_warp_call<32>(F, ny, v, w, rdv, ql);
        const double quad = ql[0], ldet = ql[1];
#ifdef DYB_KL_PROFILE
        pc[5] += clock64() - cc0;
#endif
        const int cnt = __popc(__ballot_sync(0xffffffffu, tid < ny && !isnan(yt[tid < ny ? tid : 0])));
        if (tid == 0) {
          if (!okc) bad = 1;
          if (t >= g.presample) { acc += 2.0 * ldet + quad; nseen += cnt; }
        }
      } else if (tid < 32) {                  // ny > 32: right-looking Cholesky in shared memory
        for (int j = 0; j < ny; ++j) {
          const double djj = F[j + ny * j];
          __syncwarp();
          if (tid == 0) { if (!(djj > 0.0)) bad = 1; F[j + ny * j] = sqrt(djj); }
          __syncwarp();
          const double ljj = F[j + ny * j];
          #pragma unroll 1
          for (int i = j + 1 + tid; i < ny; i += 32) F[i + ny * j] /= ljj;
          __syncwarp();
          for (int c = j + 1; c < ny; ++c) {
            const double lcj = F[c + ny * j];
            #pragma unroll 1
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
      KL_MARK(0)
      if (t == g.nobs - 1) break;
      // U[:, s] = L^-1 (Z P)[:, s] ; au = a + U'w
      if (BULK && ny <= 32) {
        if (ny <= 8) usolve_regs<8>(tid, nt, ns, ny, yt, zi, P, nsp, F, rdv, w, a, au, PxT, nxp);
        else if (ny <= 16) usolve_regs<16>(tid, nt, ns, ny, yt, zi, P, nsp, F, rdv, w, a, au, PxT, nxp);
        else if (ny <= 24) usolve_regs<24>(tid, nt, ns, ny, yt, zi, P, nsp, F, rdv, w, a, au, PxT, nxp);
        else usolve_regs<32>(tid, nt, ns, ny, yt, zi, P, nsp, F, rdv, w, a, au, PxT, nxp);
      } else {
        #pragma unroll 1
        for (int s = tid; s < ns; s += nt) {
          double x = a[s];
          for (int i = 0; i < ny; ++i) {
            double u = isnan(yt[i]) ? 0.0 : P[zi[i] + (size_t)nsp * s];
            for (int p = 0; p < i; ++p) u -= F[i + ny * p] * U[p + ldu * s];
            u *= rdv[i];
            U[i + ldu * s] = u;
            if (BULK) PxT[i + (size_t)nxp * s] = u;
            x += u * w[i];
          }
          au[s] = x;
          if (BULK) PxT[ny + (size_t)nxp * s] = x;
        }
      }
      __syncthreads();
      KL_MARK(1)
      if (BULK) {
        // M = T P (P symmetric: B by rows = P itself) ; [K | a'] = T [U' | a + U'w] ; P' = [M | K] [T | -K]' + R Q R'
#pragma unroll 1
        for (int prod = 0; prod < 3; ++prod) {
          const bool symm = prod == 2;
          const double* Ag = prod == 2 ? Mx : Tm;
          const double* Bg = prod == 0 ? P : (prod == 1 ? PxT : Tm);
          const int ldb = prod == 0 ? nsp : (prod == 1 ? nxp : ns);
          const int ncols = prod == 1 ? nxp : nsp;
          const int kdim = prod == 2 ? nsp + nxp : nsp;
          double* Cg = prod == 0 ? Mx : (prod == 1 ? Mx + (size_t)nsp * nsp : P);
          block_gemm_dmma_bulk(symm, Ag, Bg, ldb, Kneg, nsp, nsp, nsp, ncols, kdim, prod == 2 ? QQ : nullptr, Cg, nsp,
                               prod == 1 ? Kneg : nullptr, prod == 1 ? ny : -1, As0, panel_doubles, bars, phases);
          if (prod == 1) {
            __syncthreads();
            KL_MARK(2)
#pragma unroll 1
            for (int i = tid; i < ns; i += nt) a[i] = Mx[i + (size_t)nsp * (nsp + ny)];
          }
        }
        __syncthreads();
        KL_MARK(3)
        continue;
      }
      // a <- T au
      #pragma unroll 1
      for (int i = tid; i < ns; i += nt) {
        double x = 0.0;
        for (int p = 0; p < ns; ++p) x = fma(Tm[i + (size_t)ns * p], au[p], x);
        a[i] = x;
      }
      // P <- P - U'U, tile by tile (full matrix; both triangles get identical sums)
      #pragma unroll 1
      for (int tl = warp; tl < ntile * ntile; tl += nw) {
        const int i0 = 8 * (tl % ntile), j0 = 8 * (tl / ntile);
        double c0 = P[(i0 + gq) + (size_t)nsp * (j0 + 2 * tq)], c1 = P[(i0 + gq) + (size_t)nsp * (j0 + 2 * tq + 1)];
        for (int k0 = 0; k0 < nyp; k0 += 4)
          dmma884(c0, c1, -U[(k0 + tq) + ldu * (i0 + gq)], U[(k0 + tq) + ldu * (j0 + gq)]);
        P[(i0 + gq) + (size_t)nsp * (j0 + 2 * tq)] = c0; P[(i0 + gq) + (size_t)nsp * (j0 + 2 * tq + 1)] = c1;
      }
      __syncthreads();
      // M = T P ; P = M T' + R Q R'
      block_gemm_dmma<false, false>(Tm, ns, P, nsp, ns, nsp, lda, nullptr, Mx, nsp, As0, Bs0, panel_doubles);
      block_gemm_dmma<true, true>(Mx, nsp, Tm, ns, ns, nsp, lda, QQ, P, nsp, As0, Bs0, panel_doubles);
    }
#ifdef DYB_KL_PROFILE
    if (tid == 0 && d == 0) printf("cycles per period: (chol call %lld) F build %lld  chol %lld  U-solve %lld  product1 %lld  product2 %lld\n",
                                   pc[5] / g.nobs, pc[4] / g.nobs, pc[0] / g.nobs, pc[1] / g.nobs, pc[2] / g.nobs, pc[3] / g.nobs);
#endif
    if (tid == 0) {
      const double ll = -0.5 * (nseen * 1.8378770664093453 + acc);
      if (bad || !isfinite(ll)) { g.loglik[d] = -INFINITY; g.status[d] = ST_F_NOT_PD; }
      else { g.loglik[d] = ll; g.status[d] = ST_OK; }
    }
  }
}

}  // namespace dyb
