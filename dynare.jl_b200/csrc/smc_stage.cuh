// One SMC tempering stage (reference src/estimation/smc.jl:110-243) as seven fused kernels around the likelihood
// batch, all driven by a device-resident control block so that the SAME launch sequence (a CUDA graph) serves every
// stage, and with the ONE exchange a stage needs between GPUs -- the mutated particles -- written straight into the
// peers' memory by the kernel that produces them.
//
//   particle payload      row j = (para[np], loglh, logpost) of global particle j, ALL npart rows on every rank, double
//                         buffered: stage i reads pay[(i-1)&1] and writes pay[i&1].  With several ranks the buffers live
//                         in one cudaMalloc slab per rank whose CUDA-IPC mapping every peer holds; smc_accept_publish_kernel
//                         stores the rows of its CTA into the slab of EVERY rank over NVLink (coalesced, staged through
//                         shared memory), and its last CTA then raises this rank's epoch flag on every peer
//                         (__threadfence_system + st.release.sys); the first kernel of the next stage spins on the
//                         flags (ld.acquire.sys).  Without peer mappings (DYB_SMC_TRANSPORT=nccl, or IPC unavailable)
//                         the kernel writes its own slab only and the host enqueues one in-place ncclAllGather.
//   replicated algebra    everything that is O(npart) or O(npart np^2) -- incremental weights, normalisation, ESS,
//                         cumulative weights, resampling indices of ALL particles, particle mean and covariance,
//                         Cholesky factor -- is computed redundantly by every rank from the gathered payload, in the
//                         canonical association order of smc_kernels.cuh, so no further collective is needed and the
//                         result is bit-identical for any number of GPUs.  Only proposal, likelihood and accept are
//                         sharded (rank r owns particles [r n, (r+1) n)).
//   serial tails          a kernel whose per-block results feed a short serial computation (sum over block totals, ESS
//                         and scale update, offsets, moment reduction, Cholesky) finishes it in its LAST block
//                         (__threadfence + ticket), instead of a <<<1,32>>> launch of its own.
//
// Arithmetic and order of every sum are those of the kernels in smc_kernels.cuh / sampler_kernels.cuh (which remain as
// the stand-alone primitives behind dyb_smc_reweight / _resample_* / _moments), so particles and records are unchanged.
#pragma once
#include "sampler_kernels.cuh"

namespace dyb {

constexpr int SMC_MAX_RANKS = 16;
constexpr int SMC_NTICKET = 6;

struct SmcCtl {
  int stage;                         // tempering stage running / about to run (index into phi, >= 1)
  int flag;                          // the running stage resamples
  int err;                           // the particle covariance was not positive definite at some stage
  unsigned int ticket[SMC_NTICKET];  // last-block-done counters
  double zhat, sumw2;
  SmcTune tune;
  unsigned long long nclamp;         // resampling indices clamped so far
  unsigned long long accepted;       // accepted proposals of this rank in the running stage
  unsigned long long epoch_base;     // flags carry epoch_base + (stage whose payload is complete) + 1
};

struct SmcStageArgs {
  SmcCtl* ctl;
  const double* phi;                 // tempering schedule on the device
  char* slab[SMC_MAX_RANKS];         // payload slab of every rank (own slab at [rank]); only [rank] valid when npub == 1
  size_t off_pay1, off_pub, off_cnt; // byte offsets inside a slab: second payload buffer, flags, acceptance counts
  int nranks, rank;
  int npub;                          // ranks written to directly: nranks (peer stores) or 1 (own slab only)
  int nwait;                         // flags to wait for: nranks (peer stores) or 0
  long long N, n, first;             // particles in total, local count, first local particle
  int np;
  double *w, *wt, *cwl, *bt, *bt2, *btw, *offs;
  long long* idx;                    // source particle of every slot (only written when the stage resamples)
  double* xsel;                      // the selected rows, contiguous (npart x (np + 2)), when the stage resamples
  double *part1, *part2, *mu, *R, *L, *rdiag, *aux;
  double* stats;                     // nphi x SMC_STATS
  double *prop, *qx, *q0;            // local proposals and their densities
  const double* lx; const int* status;
  unsigned long long seed; int systematic;
  double trgt, alp;
};

DYB_D void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
DYB_D unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

DYB_D const double* smc_pay_in(const SmcStageArgs& a, int stage) {
  return reinterpret_cast<const double*>(a.slab[a.rank] + (((stage - 1) & 1) ? a.off_pay1 : 0));
}
DYB_D double* smc_pay_out(const SmcStageArgs& a, int stage, int r) {
  return reinterpret_cast<double*>(a.slab[r] + ((stage & 1) ? a.off_pay1 : 0));
}
// rows the mutation step starts from: the gathered payload itself, or its resampled copy (smc.jl:130-145)
DYB_D const double* smc_rows(const SmcStageArgs& a, int stage) { return a.ctl->flag ? a.xsel : smc_pay_in(a, stage); }
DYB_D unsigned long long* smc_pub(const SmcStageArgs& a, int r) { return reinterpret_cast<unsigned long long*>(a.slab[r] + a.off_pub); }
DYB_D unsigned long long* smc_cnt(const SmcStageArgs& a, int r, int stage) {
  return reinterpret_cast<unsigned long long*>(a.slab[r] + a.off_cnt) + (stage & 1) * SMC_MAX_RANKS;
}

// true in every thread of the LAST block of the grid to get here; all blocks' earlier global writes are then visible
DYB_D bool smc_last_block(unsigned int* ticket, unsigned int nblocks) {
  __shared__ int last_flag;
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    last_flag = (t == nblocks - 1);
    if (last_flag) { *ticket = 0; __threadfence(); }
  }
  __syncthreads();
  return last_flag != 0;
}

DYB_D void smc_wait_peers(const SmcStageArgs& a, unsigned long long want) {
  if (a.nwait > 1) {
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    if (t < a.nwait) {
      const unsigned long long* f = smc_pub(a, a.rank) + t;
      while (ld_acquire_sys_u64(f) < want) {}
    }
    __syncthreads();
  }
}

// acpt = mean(accepted) over ALL particles of `stage` (smc.jl:236) from the per-rank counts
DYB_D void smc_acpt_from_counts(const SmcStageArgs& a, int stage) {
  const unsigned long long* c = smc_cnt(a, a.rank, stage);
  unsigned long long s = 0;
  for (int r = 0; r < a.nranks; ++r) s += __ldcg(c + r);
  a.ctl->tune.acpt = (double)s / (double)a.N;
  a.stats[(size_t)SMC_STATS * stage + 4] = a.ctl->tune.acpt;
}

// ---- K1: incremental weights (smc.jl:118-120), canonical block totals, zhat in the tail ------------------------------
// launch <<<blocks of 1024 particles, 256>>>
__global__ void __launch_bounds__(256) smc_w1_kernel(SmcStageArgs a) {
  __shared__ double sm[SMC_BLOCK + SMC_BLOCK / 32];
  const int stage = a.ctl->stage;
  smc_wait_peers(a, a.ctl->epoch_base + (unsigned long long)stage);          // payload of stage - 1 complete everywhere
  const double dphi = a.phi[stage] - a.phi[stage - 1];
  const double* pin = smc_pay_in(a, stage);
  const int PS = a.np + 2;
  const long long lo = (long long)blockIdx.x * SMC_BLOCK;
  const int cnt = (int)((lo + SMC_BLOCK < a.N) ? SMC_BLOCK : a.N - lo);
  for (int e = threadIdx.x; e < cnt; e += 256) {
    const double v = a.w[lo + e] * exp(dphi * pin[(lo + e) * PS + a.np]);
    a.wt[lo + e] = v;
    sm[e + e / 32] = v;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double s = 0.0;
    for (int i = 0; i < SMC_RUN; ++i) {
      const int e = threadIdx.x * SMC_RUN + i;
      if (e < cnt) s = __dadd_rn(s, sm[e + e / 32]);
    }
    const double t = warp_sum_in_order(s);
    if (threadIdx.x == 0) a.bt[blockIdx.x] = t;
  }
  if (smc_last_block(&a.ctl->ticket[0], gridDim.x) && threadIdx.x == 0) {
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(a.bt + b);
    a.ctl->zhat = s;
    if (stage > 1) smc_acpt_from_counts(a, stage - 1);
  }
}

// ---- K2: normalised weights (smc.jl:121-123), sum of squares, local cumulative weights; ESS (:129), selection flag
// (:130), scale update (:150-151) and block offsets in the tail.  launch <<<blocks, 256>>>
__global__ void __launch_bounds__(256) smc_w2_kernel(SmcStageArgs a) {
  __shared__ double sm[SMC_BLOCK + SMC_BLOCK / 32];
  __shared__ double sc[SMC_BLOCK];
  const int stage = a.ctl->stage;
  const double zhat = a.ctl->zhat;
  const long long lo = (long long)blockIdx.x * SMC_BLOCK;
  const int cnt = (int)((lo + SMC_BLOCK < a.N) ? SMC_BLOCK : a.N - lo);
  for (int e = threadIdx.x; e < cnt; e += 256) {
    const double v = a.wt[lo + e] / zhat;
    a.w[lo + e] = v;
    sm[e + e / 32] = v;
    sc[e] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    double s = 0.0;
    for (int i = 0; i < SMC_RUN; ++i) {
      const int e = lane * SMC_RUN + i;
      if (e < cnt) { const double v = sm[e + e / 32]; s = __dadd_rn(s, __dmul_rn(v, v)); }
    }
    const double t = warp_sum_in_order(s);
    if (lane == 0) a.bt2[blockIdx.x] = t;
  } else if (warp == 1 && lane == 0) {
    double s = 0.0;                           // sequential by definition of the canonical cumulative sum
#pragma unroll 8
    for (int i = 0; i < cnt; ++i) { s += sc[i]; sc[i] = s; }
    a.btw[blockIdx.x] = s;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < cnt; e += 256) a.cwl[lo + e] = sc[e];
  if (smc_last_block(&a.ctl->ticket[1], gridDim.x) && threadIdx.x == 0) {
    double s2 = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s2 += __ldcg(a.bt2 + b);
    a.ctl->sumw2 = s2;
    const double ess = 1.0 / s2;
    const int f = ess < (double)a.N / 2.0;
    a.ctl->flag = f;
    SmcTune* tune = &a.ctl->tune;
    const double e = exp(16.0 * (tune->acpt - a.trgt));
    tune->c = tune->c * (0.95 + 0.10 * e / (1.0 + e));
    double* st = a.stats + (size_t)SMC_STATS * stage;
    st[0] = zhat; st[1] = ess; st[2] = (double)f; st[3] = tune->c;
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) { a.offs[b] = s; s += __ldcg(a.btw + b); }
  }
}

// ---- K3: source particle of EVERY slot (multinomial_resampling smc.jl:340-361 / systematic), the selected rows copied
// into a contiguous array (smc.jl:137-141) and equal weights (:143).  A stage that does not resample leaves at once: its
// rows are the payload itself.  launch <<<ceil(npart / 256), 256>>>
__global__ void __launch_bounds__(256) smc_select_take_kernel(SmcStageArgs a) {
  __shared__ long long sidx[256];
  if (!a.ctl->flag) return;
  const unsigned stage = (unsigned)a.ctl->stage;
  const long long j0 = (long long)blockIdx.x * 256, j = j0 + threadIdx.x;
  if (j < a.N) {
    double ui;
    if (a.systematic) ui = ((double)j + rng_uniform(a.seed, 0ULL, stage, RNG_RESAMPLE)) / (double)a.N;
    else ui = rng_uniform(a.seed, (unsigned long long)j, stage, RNG_RESAMPLE);
    long long lo = 0, hi = a.N;             // first position with cw[pos] > ui, cw[pos] = offs[pos / 1024] + cwl[pos]
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (a.offs[mid / SMC_BLOCK] + a.cwl[mid] <= ui) lo = mid + 1; else hi = mid;
    }
    if (lo >= a.N) { lo = a.N - 1; atomicAdd(&a.ctl->nclamp, 1ULL); }
    a.idx[j] = lo;
    sidx[threadIdx.x] = lo;
    a.w[j] = 1.0 / (double)a.N;
  }
  __syncthreads();
  const int PS = a.np + 2;
  const int cnt = (int)((j0 + 256 <= a.N) ? 256 : a.N - j0);
  const double* pin = smc_pay_in(a, (int)stage);
  double* dst = a.xsel + j0 * PS;
  for (int e = threadIdx.x; e < cnt * PS; e += 256) {
    const int r = e / PS, c = e - r * PS;
    dst[e] = pin[sidx[r] * PS + c];
  }
}

// prox!(RPSD, IndPSD(), R) (smc.jl:163; ProximalOperators' IndPSD: symmetric eigen-decomposition, negative eigenvalues set
// to zero): cyclic Jacobi by one warp on the full symmetric matrix W (np x np), eigenvectors accumulated in V, eigenvalues
// left in lam; on return W = V max(Lambda, 0) V'.  Only reached when the Cholesky factorisation of R has failed -- on a
// positive-definite matrix the projection is the identity and is skipped.  W, V, lam: global or shared memory.
DYB_D void smc_psd_project_warp(double* W, double* V, double* lam, int np, int lane) {
  for (int e = lane; e < np * np; e += 32) V[e] = (e % np == e / np) ? 1.0 : 0.0;
  __syncwarp();
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int e = lane; e < np * np; e += 32) {
      const double v = W[e];
      if (e % np == e / np) dg = fma(v, v, dg); else off = fma(v, v, off);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); dg += __shfl_xor_sync(0xffffffffu, dg, o); }
    if (!(off > 1e-32 * dg)) break;                      // also leaves on NaN
    for (int p = 0; p < np - 1; ++p) {
      for (int q = p + 1; q < np; ++q) {
        const double apq = W[p + np * q];
        if (apq == 0.0) continue;                         // the same value in every lane
        const double theta = (W[q + np * q] - W[p + np * p]) / (2.0 * apq);
        const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
        const double c = rsqrt(fma(t, t, 1.0)), sn = t * c;
        __syncwarp();
        for (int k = lane; k < np; k += 32) {             // W <- W J, V <- V J
          const double wp = W[k + np * p], wq = W[k + np * q];
          W[k + np * p] = c * wp - sn * wq; W[k + np * q] = sn * wp + c * wq;
          const double vp = V[k + np * p], vq = V[k + np * q];
          V[k + np * p] = c * vp - sn * vq; V[k + np * q] = sn * vp + c * vq;
        }
        __syncwarp();
        for (int k = lane; k < np; k += 32) {             // W <- J' W
          const double wp = W[p + np * k], wq = W[q + np * k];
          W[p + np * k] = c * wp - sn * wq; W[q + np * k] = sn * wp + c * wq;
        }
        __syncwarp();
      }
    }
  }
  for (int k = lane; k < np; k += 32) lam[k] = W[k + np * k];
  __syncwarp();
  for (int e = lane; e < np * np; e += 32) {
    const int i = e % np, j = e / np;
    double sacc = 0.0;
    for (int k = 0; k < np; ++k) sacc = fma(V[i + np * k] * fmax(lam[k], 0.0), V[j + np * k], sacc);
    W[e] = sacc;
  }
  __syncwarp();
  for (int e = lane; e < np * np; e += 32) {              // exactly symmetric
    const int i = e % np, j = e / np;
    if (i > j) { const double m = 0.5 * (W[i + np * j] + W[j + np * i]); W[i + np * j] = m; W[j + np * i] = m; }
  }
  __syncwarp();
}

// lower Cholesky factor of R (lower triangle read) in A (np*np doubles of shared memory); true when a pivot is not positive
DYB_D bool smc_chol_try(const double* R, int np, double* A, int lane) {
  for (int e = lane; e < np * np; e += 32) {
    const int i = e % np, j = e / np;
    A[e] = (i >= j) ? R[i + np * j] : 0.0;
  }
  __syncwarp();
  bool bad = false;
  for (int j = 0; j < np; ++j) {
    for (int i = j + lane; i < np; i += 32) {
      double s = A[i + np * j];
      for (int k = 0; k < j; ++k) s -= A[i + np * k] * A[j + np * k];
      A[i + np * j] = s;
    }
    __syncwarp();
    const double d = A[j + np * j];
    if (!(d > 0.0) || !isfinite(d)) bad = true;
    const double ljj = sqrt(d);
    __syncwarp();
    for (int i = j + lane; i < np; i += 32) A[i + np * j] = (i == j) ? ljj : A[i + np * j] / ljj;
    __syncwarp();
  }
  return bad;
}

// Covariance of the particles -> RPSD, its diagonal and its Cholesky factor by one warp (smc.jl:161-165):
//     prox!(tune.RPSD, IndPSD(), tune.R); tune.Rdiag = diag(tune.RPSD); tune.Rchol .= cholesky(tune.RPSD).L
// R is factorised directly; only if that fails is it replaced by its projection on the positive semi-definite cone and
// factorised again (a matrix that is STILL not positive definite makes the reference's `cholesky` throw: *err = 1).
// A: np*np doubles of shared memory; work: 2 np*np doubles (global); aux[0..1] log-determinant terms, aux[2] counts the
// projections.  R may be overwritten by RPSD.
DYB_D void smc_chol_warp(double* R, int np, double* A, double* __restrict__ L, double* __restrict__ rdiag,
                         double* __restrict__ aux, double* work, int* __restrict__ err, int lane) {
  bool bad = smc_chol_try(R, np, A, lane);
  if (bad && work) {
    double* W = work; double* V = work + np * np;
    for (int e = lane; e < np * np; e += 32) {
      const int i = e % np, j = e / np;
      W[e] = (i >= j) ? R[i + np * j] : R[j + np * i];
    }
    __syncwarp();
    smc_psd_project_warp(W, V, rdiag, np, lane);
    for (int e = lane; e < np * np; e += 32) R[e] = W[e];
    __syncwarp();
    bad = smc_chol_try(R, np, A, lane);
    if (lane == 0) aux[2] += 1.0;
  }
  for (int e = lane; e < np * np; e += 32) L[e] = A[e];
  for (int k = lane; k < np; k += 32) rdiag[k] = R[k + np * k];
  if (lane == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (int k = 0; k < np; ++k) { s0 += log(A[k + np * k]); s1 += log(R[k + np * k]); }
    aux[0] = s0; aux[1] = s1;
    if (bad) *err = 1;
  }
}

// stand-alone entry (dyb_smc_covariance_factor): <<<1, 32, np*np doubles>>>; R np x np in/out, work 2 np*np, out3 = aux
__global__ void smc_covariance_factor_kernel(double* R, int np, double* L, double* rdiag, double* aux, double* work, int* err) {
  extern __shared__ double sm_cf[];
  for (int e = threadIdx.x; e < np * np; e += 32) {     // only the lower triangle of the input is meaningful
    const int i = e % np, j = e / np;
    if (i < j) R[e] = R[j + np * i];
  }
  __syncwarp();
  smc_chol_warp(R, np, sm_cf, L, rdiag, aux, work, err, threadIdx.x);
}

// ---- K4 / K5: particle mean (smc.jl:154) and covariance (:155-160) of the selected particles in the canonical order.
// SECOND = false: partial[b*np + p] = sum_{i in block b} w_i x_ip, mu in the tail; SECOND = true: partial[b*np*np + p + np*q]
// = sum_i (w_i (x_ip - mu_p)) (x_iq - mu_q), then R and its Cholesky factor in the tail.
// STAGED: the 1024 rows of the block are read ONCE per CTA, coalesced, into shared memory ((np + 1) * 1056 doubles;
// particle e at column e + e / 32 so that the 32 lanes of a warp, whose runs start 32 apart, hit distinct banks); one warp
// per component, lane j sums run j sequentially, the 32 run totals are combined in order.  The grid is one wave:
// <<<dim3(blocks, groups), dim3(32, SMC_MOM_WARPS), smem>>>, group g owns components g, g + groups, ...
constexpr int SMC_MOM_WARPS = 16;
template <bool SECOND, bool STAGED>
__global__ void __launch_bounds__(32 * SMC_MOM_WARPS) smc_moments_kernel(SmcStageArgs a) {
  extern __shared__ double sm_mom[];
  const int np = a.np, PS = np + 2, ncomp = SECOND ? np * np : np;
  const double* x = smc_rows(a, a.ctl->stage);
  const int tid = threadIdx.y * 32 + threadIdx.x, nt = 32 * SMC_MOM_WARPS, lane = threadIdx.x;
  const long long lo = (long long)blockIdx.x * SMC_BLOCK;
  const int cnt = (int)((lo + SMC_BLOCK < a.N) ? SMC_BLOCK : a.N - lo);
  double* part = SECOND ? a.part2 : a.part1;
  if (STAGED) {
    double* xs = sm_mom;
    double* wsm = sm_mom + (size_t)np * SMC_MOM_LD;
    const double* src = x + lo * PS;
    for (int k = tid; k < cnt * PS; k += nt) {
      const int e = k / PS, p = k - e * PS;
      if (p < np) xs[(size_t)p * SMC_MOM_LD + e + e / 32] = src[k];
    }
    for (int e = tid; e < cnt; e += nt) wsm[e + e / 32] = a.w[lo + e];
    __syncthreads();
    for (int c = blockIdx.y + gridDim.y * threadIdx.y; c < ncomp; c += gridDim.y * SMC_MOM_WARPS) {
      const int p = c % np, q = c / np;
      const double* xp = xs + (size_t)p * SMC_MOM_LD + 33 * lane;
      const double* wl = wsm + 33 * lane;
      double s = 0.0;
      if (SECOND) {
        const double mp = a.mu[p], mq = a.mu[q];
        const double* xq = xs + (size_t)q * SMC_MOM_LD + 33 * lane;
        for (int i = 0; i < SMC_RUN; ++i)
          if (lane * SMC_RUN + i < cnt) s = __dadd_rn(s, __dmul_rn(__dmul_rn(wl[i], xp[i] - mp), xq[i] - mq));
      } else {
        for (int i = 0; i < SMC_RUN; ++i)
          if (lane * SMC_RUN + i < cnt) s = __dadd_rn(s, __dmul_rn(wl[i], xp[i]));
      }
      const double t = warp_sum_in_order(s);
      if (lane == 0) part[(long long)blockIdx.x * ncomp + c] = t;
    }
  } else {
    for (int c = blockIdx.y + gridDim.y * threadIdx.y; c < ncomp; c += gridDim.y * SMC_MOM_WARPS) {
      const int p = c % np, q = c / np;
      const double mp = SECOND ? a.mu[p] : 0.0, mq = SECOND ? a.mu[q] : 0.0;
      double s = 0.0;
      for (int i = 0; i < SMC_RUN; ++i) {
        const int e = lane * SMC_RUN + i;
        if (e < cnt) {
          const double* xr = x + (lo + e) * PS;
          if (SECOND) s = __dadd_rn(s, __dmul_rn(__dmul_rn(a.w[lo + e], xr[p] - mp), xr[q] - mq));
          else s = __dadd_rn(s, __dmul_rn(a.w[lo + e], xr[p]));
        }
      }
      const double t = warp_sum_in_order(s);
      if (lane == 0) part[(long long)blockIdx.x * ncomp + c] = t;
    }
  }
  if (smc_last_block(&a.ctl->ticket[SECOND ? 3 : 2], gridDim.x * gridDim.y)) {
    double* out = SECOND ? a.R : a.mu;
    for (int k = tid; k < ncomp; k += nt) {
      double s = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(part + (size_t)b * ncomp + k);
      out[k] = s;
    }
    if (SECOND) {
      __syncthreads();
      if (threadIdx.y == 0) smc_chol_warp(a.R, np, sm_mom, a.L, a.rdiag, a.aux, a.aux + 4, &a.ctl->err, lane);
    }
  }
}

// ---- K6: one mutation proposal per local particle (smc.jl:176-200) with the proposal log-densities (:326-338) ----------
// `MvNormal(mu, c^2 * Rdiag)` with a VECTOR second argument is Distributions' (deprecated) standard-deviation constructor,
// so the diagonal component's density uses sigma_k = c^2 Rdiag_k; reproduced as coded.  c L is formed once per CTA; the two
// whitening passes (density of the proposal seen from the particle and back) are independent and run in one loop.
__global__ void __launch_bounds__(128) smc_propose2_kernel(SmcStageArgs a) {
  extern __shared__ double sh[];            // c L (np*np), mu (np), rdiag (np)
  const int np = a.np, PS = np + 2;
  double* Ls = sh; double* mus = Ls + np * np; double* rds = mus + np;
  const double c = a.ctl->tune.c;
  for (int e = threadIdx.x; e < np * np; e += blockDim.x) Ls[e] = c * a.L[e];
  for (int e = threadIdx.x; e < np; e += blockDim.x) { mus[e] = a.mu[e]; rds[e] = a.rdiag[e]; }
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const unsigned stage = (unsigned)a.ctl->stage;
  const unsigned long long g = (unsigned long long)(a.first + i);
  const double LOG2PI = 1.8378770664093453;
  double z[MAX_EST], px[MAX_EST], d1[MAX_EST], d0[MAX_EST];
  rng_normal_vector(a.seed, g, stage, np, z);
  const double mix = rng_uniform(a.seed, g, stage, RNG_MIX);
  const int mixsel = mix < a.alp ? 1 : (mix < a.alp + (1.0 - a.alp) / 2.0 ? 2 : 3);
  const double* p0 = smc_rows(a, (int)stage) + g * PS;
  for (int k = 0; k < np; ++k) {
    double s;
    if (mixsel == 2) s = (c * sqrt(rds[k])) * z[k];
    else { s = 0.0; for (int j = 0; j <= k; ++j) s += Ls[k + np * j] * z[j]; }
    px[k] = (mixsel == 3 ? mus[k] : p0[k]) + s;
    a.prop[i * np + k] = px[k];
  }
  double qx, q0;
  if (mixsel == 2) {
    double ld = 0.0, m1 = 0.0, m0 = 0.0;
    for (int k = 0; k < np; ++k) {
      const double sg = c * c * rds[k];
      ld += log(sg);
      const double e1 = (px[k] - p0[k]) / sg, e0 = (p0[k] - px[k]) / sg;
      m1 += e1 * e1; m0 += e0 * e0;
    }
    const double c0 = -0.5 * (np * LOG2PI + 2.0 * ld);
    qx = c0 - 0.5 * m1; q0 = c0 - 0.5 * m0;
  } else {
    const double c0 = -0.5 * (np * LOG2PI + 2.0 * np * log(c) + 2.0 * a.aux[0]);
    // whitened residuals by forward substitution with c L: pass "1" (proposal around the particle / the mean) and
    // pass "0" (the particle around the proposal / the mean)
    double m1 = 0.0, m0 = 0.0;
    for (int k = 0; k < np; ++k) {
      double s1 = px[k] - ((mixsel == 3) ? mus[k] : p0[k]);
      double s0 = p0[k] - ((mixsel == 3) ? mus[k] : px[k]);
      for (int j = 0; j < k; ++j) { const double l = Ls[k + np * j]; s1 -= l * d1[j]; s0 -= l * d0[j]; }
      const double dg = Ls[k + np * k];
      d1[k] = s1 / dg; d0[k] = s0 / dg;
      m1 += d1[k] * d1[k]; m0 += d0[k] * d0[k];
    }
    qx = c0 - 0.5 * m1; q0 = c0 - 0.5 * m0;
  }
  a.qx[i] = qx; a.q0[i] = q0;
}

// ---- K7: prior of the proposals, accept / reject (smc.jl:202-232), and the exchange: the new row of every local
// particle goes into the next payload buffer of EVERY rank; the last block publishes the acceptance count and the epoch
// flag, records the clamp count and advances the stage.  launch <<<ceil(n / 128), 128, 128 * (np + 2) doubles>>>
constexpr int SMC_ACC_BLOCK = 128;
__global__ void __launch_bounds__(SMC_ACC_BLOCK) smc_accept_publish_kernel(SmcStageArgs a, PriorSetDev ps) {
  extern __shared__ double rows[];
  const int np = a.np, PS = np + 2;
  const int stage = a.ctl->stage;
  const double phi = a.phi[stage], dphi = a.phi[stage] - a.phi[stage - 1];
  const double* pin = smc_rows(a, stage);
  const long long i0 = (long long)blockIdx.x * SMC_ACC_BLOCK;
  const long long i = i0 + threadIdx.x;
  bool acc = false;
  if (i < a.n) {
    const long long g = a.first + i;
    const double* src = pin + g * PS;
    const double l0 = src[np];
    const double post0 = src[np + 1] + dphi * l0;
    const double lx = a.status[i] != 0 ? -INFINITY : a.lx[i];
    const double* px = a.prop + i * np;
    double pr = 0.0;
    for (int k = 0; k < np; ++k) pr += prior_logpdf(ps.p[k], px[k]);
    const double postx = phi * lx + pr;
    const double alp = exp((postx - a.qx[i]) - (post0 - a.q0[i]));
    const double u = rng_uniform(a.seed, (unsigned long long)g, (unsigned)stage, RNG_ACCEPT);
    acc = u < alp;
    double* row = rows + (size_t)threadIdx.x * PS;
    for (int k = 0; k < np; ++k) row[k] = acc ? px[k] : src[k];
    row[np] = acc ? lx : l0;
    row[np + 1] = acc ? postx : post0;
  }
  const unsigned m = __ballot_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&a.ctl->accepted, (unsigned long long)__popc(m));
  __syncthreads();
  const long long cnt = (i0 + SMC_ACC_BLOCK <= a.n) ? SMC_ACC_BLOCK : (a.n - i0);
  const long long nel = cnt * PS, off = (a.first + i0) * PS;
  for (int q = 0; q < a.npub; ++q) {
    const int r = (a.npub == 1) ? a.rank : (a.rank + q) % a.nranks;      // start with the own slab, then round the ring
    double* dst = smc_pay_out(a, stage, r) + off;
    for (long long k = threadIdx.x; k < nel; k += SMC_ACC_BLOCK) dst[k] = rows[k];
  }
  if (a.npub > 1) __threadfence_system();
  if (smc_last_block(&a.ctl->ticket[4], gridDim.x) && threadIdx.x == 0) {
    const unsigned long long mine = a.ctl->accepted;
    a.ctl->accepted = 0;
    for (int q = 0; q < a.npub; ++q) {
      const int r = (a.npub == 1) ? a.rank : (a.rank + q) % a.nranks;
      smc_cnt(a, r, stage)[a.rank] = mine;
    }
    a.stats[(size_t)SMC_STATS * stage + 5] = (double)a.ctl->nclamp;
    a.ctl->stage = stage + 1;
    if (a.npub > 1) {
      __threadfence_system();
      const unsigned long long ep = a.ctl->epoch_base + (unsigned long long)stage + 1ULL;
      for (int q = 0; q < a.npub; ++q) st_release_sys_u64(smc_pub(a, (a.rank + q) % a.nranks) + a.rank, ep);
    }
  }
}

// acceptance rate of the last completed stage (the next stage's first kernel does the same): ends dyb_smc_run
__global__ void smc_finalize_kernel(SmcStageArgs a) {
  const int last = a.ctl->stage - 1;
  if (last < 1) return;
  smc_wait_peers(a, a.ctl->epoch_base + (unsigned long long)last + 1ULL);
  if (threadIdx.x == 0) smc_acpt_from_counts(a, last);
}

// initial particles -> rows [first, first + n) of payload buffer 0 of the own slab (smc.jl:96: logpost = phi[1] * loglh + prior)
__global__ void smc_init_pack_kernel(SmcStageArgs a, const double* __restrict__ para, const double* __restrict__ loglh,
                                     const double* __restrict__ prior, double phi0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int np = a.np, PS = np + 2;
  double* row = reinterpret_cast<double*>(a.slab[a.rank]) + (a.first + i) * PS;
  for (int k = 0; k < np; ++k) row[k] = para[i * np + k];
  row[np] = loglh[i];
  row[np + 1] = phi0 * loglh[i] + prior[i];
}

// start of a run with peer stores: the own rows of payload buffer 0 go to every peer, then the epoch flag (stage 0 complete)
__global__ void __launch_bounds__(256) smc_publish_init_kernel(SmcStageArgs a) {
  const int PS = a.np + 2;
  const long long nel = a.n * PS, off = a.first * PS;
  const double* src = reinterpret_cast<const double*>(a.slab[a.rank]) + off;
  for (int q = 1; q < a.npub; ++q) {
    double* dst = reinterpret_cast<double*>(a.slab[(a.rank + q) % a.nranks]) + off;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nel; k += (long long)gridDim.x * blockDim.x) dst[k] = src[k];
  }
  __threadfence_system();
  if (smc_last_block(&a.ctl->ticket[5], gridDim.x) && threadIdx.x == 0) {
    __threadfence_system();
    const unsigned long long ep = a.ctl->epoch_base + 1ULL;
    for (int q = 0; q < a.npub; ++q) st_release_sys_u64(smc_pub(a, (a.rank + q) % a.nranks) + a.rank, ep);
  }
}

__global__ void smc_set_epoch_kernel(SmcCtl* ctl, unsigned long long base) { ctl->epoch_base = base; }

// local rows of the current payload -> contiguous para / loglh / logpost (dyb_smc_get_particles)
__global__ void smc_unpack_rows_kernel(const double* __restrict__ pay, long long first, long long n, int np,
                                       double* __restrict__ para, double* __restrict__ loglh, double* __restrict__ logpost) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* row = pay + (first + i) * (np + 2);
  if (para) for (int k = 0; k < np; ++k) para[i * np + k] = row[k];
  if (loglh) loglh[i] = row[np];
  if (logpost) logpost[i] = row[np + 1];
}

}  // namespace dyb
