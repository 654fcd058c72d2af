// SMC particle-weight primitives on device (reference src/estimation/smc.jl:118-167, 340-361).
//
// Every sum over particles uses ONE canonical association order, independent of launch geometry and
// of how particles are sharded over GPUs (three levels):
//   sequential fp64 sum of each run of 32 consecutive particles; sequential sum of the 32 run totals of each
//   block of SMC_BLOCK = 1024 consecutive particles; sequential sum over block totals.
// The cumulative weight used for resampling is  cw[b*1024 + i] = offset_b + local_b[i]  with local_b the
// sequential inclusive scan of block b and offset_b the sequential sum of the previous blocks' scan totals.
// The CPU oracle (oracle/smc.py) uses the same orders, which is what makes resampling indices bit-exact.
#pragma once
#include "model_traits.cuh"

namespace dyb {

constexpr int SMC_BLOCK = 1024;

// phase 1 of reweight: wt[i] = w[i] * exp(dphi * loglh[i]); block totals of wt (canonical order)
__global__ void smc_incweight_kernel(const double* __restrict__ w, const double* __restrict__ loglh,
                                     double dphi, long long n, double* __restrict__ wt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) wt[i] = w[i] * exp(dphi * loglh[i]);
}

constexpr int SMC_RUN = 32;     // SMC_BLOCK = SMC_RUN * 32

// sequential combination of the 32 lane values of a warp, lane 0 first (every lane gets the result)
DYB_D double warp_sum_in_order(double s) {
  double t = 0.0;
#pragma unroll
  for (int j = 0; j < 32; ++j) t = __dadd_rn(t, __shfl_sync(0xffffffffu, s, j));
  return t;
}

// canonical total of f(x) over each 1024-block: one warp per block, lane j owns run j. mode 0: x ; 1: x*x
// launch <<<number of blocks, 32>>>
__global__ void smc_block_sum_kernel(const double* __restrict__ x, long long n, int mode,
                                     double* __restrict__ block_tot) {
  const long long lo = (long long)blockIdx.x * SMC_BLOCK + (long long)threadIdx.x * SMC_RUN;
  double s = 0.0;
  for (int i = 0; i < SMC_RUN; ++i) {
    const long long e = lo + i;
    if (e < n) {
      const double v = x[e];
      s = mode == 0 ? __dadd_rn(s, v) : __dadd_rn(s, __dmul_rn(v, v));   // no FMA contraction: same bits as the CPU order
    }
  }
  const double t = warp_sum_in_order(s);
  if (threadIdx.x == 0) block_tot[blockIdx.x] = t;
}

// sequential sum over block totals (single thread); out[0] = total
__global__ void smc_total_kernel(const double* __restrict__ block_tot, long long nb, double* __restrict__ out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (long long b = 0; b < nb; ++b) s += block_tot[b];
    out[0] = s;
  }
}

// w_out = wt / zhat
__global__ void smc_normalise_kernel(const double* __restrict__ wt, const double* __restrict__ zhat,
                                     long long n, double* __restrict__ w_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w_out[i] = wt[i] / zhat[0];
}

// local inclusive scan inside each 1024-block + block totals.  The scan is sequential by definition (canonical order);
// one warp per block stages the weights through shared memory with coalesced loads and stores, so that the 1024
// dependent additions of lane 0 wait on shared memory instead of on one global load each (70 us -> 8 us at 16 blocks).
// launch <<<number of blocks, 32>>>
__global__ void smc_block_scan_kernel(const double* __restrict__ w, long long n,
                                      double* __restrict__ local, double* __restrict__ block_tot) {
  __shared__ double ws[SMC_BLOCK];
  const long long lo = (long long)blockIdx.x * SMC_BLOCK;
  const int cnt = (int)((lo + SMC_BLOCK < n) ? SMC_BLOCK : n - lo);
  const int lane = threadIdx.x;
  for (int k = lane; k < cnt; k += 32) ws[k] = w[lo + k];
  __syncwarp();
  if (lane == 0) {
    double s = 0.0;
#pragma unroll 8
    for (int i = 0; i < cnt; ++i) { s += ws[i]; ws[i] = s; }
    block_tot[blockIdx.x] = s;
  }
  __syncwarp();
  for (int k = lane; k < cnt; k += 32) local[lo + k] = ws[k];
}

// exclusive sequential scan of block totals (single thread): offs[b] = sum_{b' < b} tot[b']
__global__ void smc_offsets_kernel(const double* __restrict__ block_tot, long long nb, double* __restrict__ offs) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (long long b = 0; b < nb; ++b) { offs[b] = s; s += block_tot[b]; }
  }
}

__global__ void smc_add_offsets_kernel(double* __restrict__ cw, const double* __restrict__ offs, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cw[i] = offs[i / SMC_BLOCK] + cw[i];
}

// idx[i] = #{ j : cw[j] <= u_i }  (= first j with u_i < cw[j], 0-based), clamped to n-1.
// systematic != 0: u_i = (i + u0) / n, else u_i = u[i].  Output rows [lo, hi).
__global__ void smc_search_kernel(const double* __restrict__ cw, long long n,
                                  const double* __restrict__ u, double u0, int systematic,
                                  long long lo, long long hi,
                                  long long* __restrict__ idx, unsigned long long* __restrict__ n_clamped) {
  const long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const double ui = systematic ? ((double)i + u0) / (double)n : u[i];
  long long a = 0, b = n;                 // first position with cw[pos] > ui
  while (a < b) {
    const long long mid = (a + b) >> 1;
    if (cw[mid] <= ui) a = mid + 1; else b = mid;
  }
  if (a >= n) { a = n - 1; atomicAdd(n_clamped, 1ULL); }
  idx[i - lo] = a;
}

// weighted moments of the particles in the canonical order.  mu == nullptr: first moment,
//   partial[b*np + p] = sum_{i in block b} w[i] * para[p + np*i];
// else second central moment, partial[b*np*np + p + np*q] = sum_i (w_i (x_ip - mu_p)) (x_iq - mu_q).
// launch <<<dim3(blocks, ceil(ncomp / SMC_MOM_TILE)), dim3(32, SMC_MOM_TILE)>>>: one warp per (block, component),
// lane j sums run j sequentially, then the 32 run totals are combined in order.
constexpr int SMC_MOM_TILE = 8;
__global__ void smc_moment_kernel(const double* __restrict__ para, const double* __restrict__ w,
                                  const double* __restrict__ mu, long long n, int np,
                                  double* __restrict__ partial) {
  const int ncomp = mu ? np * np : np;
  const int c = blockIdx.y * SMC_MOM_TILE + threadIdx.y;
  if (c >= ncomp) return;                                  // whole warp leaves together
  const int p = c % np, q = c / np;
  const double mp = mu ? mu[p] : 0.0, mq = mu ? mu[q] : 0.0;
  const long long lo = (long long)blockIdx.x * SMC_BLOCK + (long long)threadIdx.x * SMC_RUN;
  double s = 0.0;
  for (int i = 0; i < SMC_RUN; ++i) {
    const long long e = lo + i;
    if (e < n) {
      const double* x = para + (long long)np * e;
      if (mu) s = __dadd_rn(s, __dmul_rn(__dmul_rn(w[e], x[p] - mp), x[q] - mq));
      else s = __dadd_rn(s, __dmul_rn(w[e], x[p]));
    }
  }
  const double t = warp_sum_in_order(s);
  if (threadIdx.x == 0) partial[(long long)blockIdx.x * ncomp + c] = t;
}

// The second moment again, same arithmetic in the same order, for np small enough that a 1024-particle block fits in
// shared memory ((np + 1) * 1056 doubles): the block is read ONCE per CTA with coalesced loads instead of once per
// component from 32 scattered rows per warp-load (the scattered version moves ~0.5 GB through L2 at 16 384 x 17).
// Particle e of the block sits at column e + e / 32, so that the 32 lanes of a warp (runs 32 apart) hit distinct banks.
// launch <<<dim3(blocks, groups), dim3(32, SMC_MOM_TILE), smc_moment2_smem_bytes(np)>>>; group g owns components
// c = g, g + groups, ... spread over the CTA's warps.
constexpr int SMC_MOM_LD = SMC_BLOCK + SMC_BLOCK / 32 + 1;     // odd: the staging stores of one particle's components (stride LD) hit distinct banks (an even LD made them np-way conflicts)
inline size_t smc_moment2_smem_bytes(int np) { return sizeof(double) * (size_t)(np + 1) * SMC_MOM_LD; }
__global__ void smc_moment2_staged_kernel(const double* __restrict__ para, const double* __restrict__ w,
                                          const double* __restrict__ mu, long long n, int np,
                                          double* __restrict__ partial) {
  extern __shared__ double sm_mom[];
  double* xs = sm_mom;                         // xs[p * LD + col]
  double* wsm = sm_mom + (size_t)np * SMC_MOM_LD;
  const int tid = threadIdx.y * 32 + threadIdx.x, nt = 32 * SMC_MOM_TILE;
  const long long lo = (long long)blockIdx.x * SMC_BLOCK;
  const int cnt = (int)((lo + SMC_BLOCK < n) ? SMC_BLOCK : n - lo);
  for (int k = tid; k < cnt * np; k += nt) {
    const int e = k / np, p = k - e * np;
    xs[(size_t)p * SMC_MOM_LD + e + e / 32] = para[(long long)np * lo + k];
  }
  for (int e = tid; e < cnt; e += nt) wsm[e + e / 32] = w[lo + e];
  __syncthreads();
  const int ncomp = np * np, lane = threadIdx.x;
  for (int c = blockIdx.y + gridDim.y * threadIdx.y; c < ncomp; c += gridDim.y * SMC_MOM_TILE) {
    const int p = c % np, q = c / np;
    const double mp = mu[p], mq = mu[q];
    const double* xp = xs + (size_t)p * SMC_MOM_LD + 33 * lane;
    const double* xq = xs + (size_t)q * SMC_MOM_LD + 33 * lane;
    const double* wl = wsm + 33 * lane;
    double s = 0.0;
    for (int i = 0; i < SMC_RUN; ++i)
      if (lane * SMC_RUN + i < cnt) s = __dadd_rn(s, __dmul_rn(__dmul_rn(wl[i], xp[i] - mp), xq[i] - mq));
    const double t = warp_sum_in_order(s);
    if (lane == 0) partial[(long long)blockIdx.x * ncomp + c] = t;
  }
}

// out[c] = sequential sum over blocks of partial[b*ncomp + c]
__global__ void smc_reduce_partials_kernel(const double* __restrict__ partial, long long nb, int ncomp,
                                           double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncomp) return;
  double s = 0.0;
  for (long long b = 0; b < nb; ++b) s += partial[b * ncomp + c];
  out[c] = s;
}

// gather particle payload rows: dst[:, i] = src[:, idx[i]]  (ncomp doubles per particle)
__global__ void smc_gather_kernel(const double* __restrict__ src, const long long* __restrict__ idx,
                                  long long n_out, int ncomp, double* __restrict__ dst) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_out * ncomp) return;
  const long long i = e / ncomp;
  const int c = (int)(e % ncomp);
  dst[e] = src[idx[i] * ncomp + c];
}

}  // namespace dyb
