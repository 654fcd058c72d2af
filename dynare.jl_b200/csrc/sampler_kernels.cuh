// Device side of the samplers that drive the batched likelihood: counter-based random numbers, prior
// log-densities, parameter transformations, the SMC mutation step and random-walk Metropolis.
//
// Reference semantics restated here:
//   prior log-densities        src/estimation/estimation.jl:455-462, src/estimation/priors.jl:147-178,
//                              src/distributions/inversegamma1.jl:134-141
//   transformations to R       src/estimation/estimation.jl:465-487 (TransformVariables as𝕀 / asℝ₊ / asℝ / as(Real,a,b))
//   SMC selection / mutation   src/estimation/smc.jl:126-233, proposal densities :326-338
//   RWMH                       src/estimation/estimation.jl:964-989 (AdvancedMH.RWMH with a MvNormal proposal)
//
// Random numbers are Philox4x32-10 keyed by the sampler seed with the counter
//   (particle or chain index lo, hi, step, stream << 24 | block),
// so a draw depends only on (seed, global index, step, stream): results are independent of launch geometry
// and of how particles / chains are sharded over GPUs.  oracle/rng.py restates the generator in numpy.
#pragma once
#include "model_traits.cuh"
#include "smc_kernels.cuh"

namespace dyb {

// ------------------------------------------------------------------------------------------------
// Philox4x32-10
// ------------------------------------------------------------------------------------------------
enum RngStream : unsigned { RNG_RESAMPLE = 0, RNG_MIX = 1, RNG_ACCEPT = 2, RNG_NORMAL = 3, RNG_USER = 4 };

DYB_HD void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                          unsigned out[4]) {
  constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)M0 * c0;
    const unsigned long long p1 = (unsigned long long)M1 * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0;
    const unsigned n1 = (unsigned)p1;
    const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
    const unsigned n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

DYB_HD void rng_words(unsigned long long seed, unsigned long long idx, unsigned step, unsigned stream, unsigned block,
                      unsigned long long& x, unsigned long long& y) {
  unsigned o[4];
  philox4x32_10((unsigned)idx, (unsigned)(idx >> 32), step, (stream << 24) | (block & 0xFFFFFFu),
                (unsigned)seed, (unsigned)(seed >> 32), o);
  x = ((unsigned long long)o[1] << 32) | o[0];
  y = ((unsigned long long)o[3] << 32) | o[2];
}

// uniform on [0, 1) with 53 random bits
DYB_HD double rng_uniform(unsigned long long seed, unsigned long long idx, unsigned step, unsigned stream,
                          unsigned block = 0) {
  unsigned long long x, y;
  rng_words(seed, idx, step, stream, block, x, y);
  return (double)(x >> 11) * 1.1102230246251565e-16;      // 2^-53
}

// pair `block` of standard normals (Box-Muller): z0 = r cos(2 pi u2), z1 = r sin(2 pi u2), r = sqrt(-2 log u1), u1 in (0,1]
DYB_D void rng_normal_pair(unsigned long long seed, unsigned long long idx, unsigned step, unsigned stream,
                           unsigned block, double& z0, double& z1) {
  unsigned long long x, y;
  rng_words(seed, idx, step, stream, block, x, y);
  const double u1 = ((double)(x >> 11) + 1.0) * 1.1102230246251565e-16;
  const double u2 = (double)(y >> 11) * 1.1102230246251565e-16;
  const double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// fill z[0..np) with the normals of (idx, step)
DYB_D void rng_normal_vector(unsigned long long seed, unsigned long long idx, unsigned step, int np, double* z) {
  for (int j = 0; 2 * j < np; ++j) {
    double a, b;
    rng_normal_pair(seed, idx, step, RNG_NORMAL, (unsigned)j, a, b);
    z[2 * j] = a;
    if (2 * j + 1 < np) z[2 * j + 1] = b;
  }
}

// test hook: out[count x n]: kind 0 uniforms of `stream` (block = 0..count-1), kind 1 the normal vector
__global__ void rng_fill_kernel(unsigned long long seed, long long first, long long n, unsigned step, unsigned stream,
                                int kind, int count, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* o = out + i * count;
  if (kind == 0) {
    for (int b = 0; b < count; ++b) o[b] = rng_uniform(seed, (unsigned long long)(first + i), step, stream, (unsigned)b);
  } else {
    for (int j = 0; 2 * j < count; ++j) {
      double a, b;
      rng_normal_pair(seed, (unsigned long long)(first + i), step, RNG_NORMAL, (unsigned)j, a, b);
      o[2 * j] = a;
      if (2 * j + 1 < count) o[2 * j + 1] = b;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// priors and transformations
// ------------------------------------------------------------------------------------------------
struct PriorDevK {          // kernel-side copy of PriorDev (internal.h)
  int shape, transform;
  double a, b, cst;
};
struct PriorSetDev {
  int np;
  PriorDevK p[MAX_EST];
};

// shape codes of the reference (src/estimation/priors.jl:482-529): 1 Beta(a,b), 2 Gamma(shape a, scale b),
// 3 Normal(mean a, std b), 4 InverseGamma1(alpha a, theta b), 5 Uniform(a,b), 6 InverseGamma(alpha a, theta b),
// 8 Weibull(shape a, scale b).  `cst` is the x-independent part, computed once on the host.
DYB_D double prior_logpdf(const PriorDevK& p, double x) {
  const double NINF = -INFINITY;
  switch (p.shape) {
    case 1: return (x > 0.0 && x < 1.0) ? (p.a - 1.0) * log(x) + (p.b - 1.0) * log1p(-x) + p.cst : NINF;
    case 2: return (x > 0.0) ? (p.a - 1.0) * log(x) - x / p.b + p.cst : NINF;
    case 3: { const double z = (x - p.a) / p.b; return -0.5 * z * z + p.cst; }
    case 4: return (x > 0.0) ? p.cst - (2.0 * p.a + 1.0) * log(x) - p.b / (x * x) : NINF;
    case 5: return (x >= p.a && x <= p.b) ? p.cst : NINF;
    case 6: return (x > 0.0) ? p.cst - (p.a + 1.0) * log(x) - p.b / x : NINF;
    case 8: { if (!(x > 0.0)) return NINF; const double z = x / p.b; return p.cst + (p.a - 1.0) * log(z) - pow(z, p.a); }
    default: return NINF;
  }
}

DYB_D double log1pexp(double x) {          // log(1 + exp(x)) without overflow
  return x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x));
}
// y in R -> parameter x, and log |dx/dy|
DYB_D double transform_to_param(const PriorDevK& p, double y, double& logjac) {
  switch (p.transform) {
    case 1: logjac = y; return exp(y);
    case 2: { logjac = -(log1pexp(-y) + log1pexp(y)); return 1.0 / (1.0 + exp(-y)); }
    case 3: { logjac = log(p.b - p.a) - (log1pexp(-y) + log1pexp(y)); return (p.b - p.a) / (1.0 + exp(-y)) + p.a; }
    default: logjac = 0.0; return y;
  }
}
DYB_D double transform_from_param(const PriorDevK& p, double x) {
  switch (p.transform) {
    case 1: return log(x);
    case 2: return log(x / (1.0 - x));
    case 3: { const double u = (x - p.a) / (p.b - p.a); return log(u / (1.0 - u)); }
    default: return x;
  }
}

// logprior[i] = sum_k logpdf(prior_k, theta[k, i])   (estimation.jl:455-462; left-to-right sum)
__global__ void logprior_kernel(PriorSetDev ps, const double* __restrict__ theta, long long n,
                                double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* x = theta + i * ps.np;
  double s = 0.0;
  for (int k = 0; k < ps.np; ++k) s += prior_logpdf(ps.p[k], x[k]);
  out[i] = s;
}

// logpost = logprior + loglik, -Inf for failed draws (estimation.jl:503-515)
__global__ void logposterior_kernel(const double* __restrict__ logprior, const double* __restrict__ loglik,
                                    const int* __restrict__ status, long long n, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = logprior[i] + loglik[i];
  out[i] = (status[i] != 0 || !isfinite(v)) ? -INFINITY : v;
}

// theta = transform(y), logjac = sum_k log|d theta_k / d y_k|; direction 1: y = inverse(theta)
__global__ void transform_kernel(PriorSetDev ps, const double* __restrict__ in, long long n, int inverse,
                                 double* __restrict__ out, double* __restrict__ logjac) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double lj = 0.0;
  for (int k = 0; k < ps.np; ++k) {
    const double v = in[i * ps.np + k];
    if (inverse) out[i * ps.np + k] = transform_from_param(ps.p[k], v);
    else { double l; out[i * ps.np + k] = transform_to_param(ps.p[k], v, l); lj += l; }
  }
  if (logjac && !inverse) logjac[i] = lj;
}

// ------------------------------------------------------------------------------------------------
// SMC stage kernels (single-thread / single-warp kernels operate on O(np^2) or O(N/1024) data)
// ------------------------------------------------------------------------------------------------
struct SmcTune {            // device-resident scalars of `Tune` (smc.jl:13-62) that change between stages
  double c, acpt;
};
constexpr int SMC_STATS = 6;   // per stage: zhat, ESS, resampled, c, acpt, clamped

__global__ void smc_fill_kernel(double* __restrict__ x, long long n, double v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}

// logpost = phi[1] * loglh + logprior for the initial particles (smc.jl:96)
__global__ void smc_init_logpost_kernel(const double* __restrict__ loglh, const double* __restrict__ prior, double phi0,
                                        long long n, double* __restrict__ logpost) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) logpost[i] = phi0 * loglh[i] + prior[i];
}

// ESS (smc.jl:129), the selection flag (smc.jl:130) and the scale update (smc.jl:150-151) of one stage
__global__ void smc_stage_scalars_kernel(const double* __restrict__ scal /* zhat, sum w^2 */, long long npart,
                                         double trgt, SmcTune* __restrict__ tune, int* __restrict__ flag,
                                         double* __restrict__ stats) {
  if (blockIdx.x || threadIdx.x) return;
  const double ess = 1.0 / scal[1];
  const int f = ess < (double)npart / 2.0;
  *flag = f;
  const double e = exp(16.0 * (tune->acpt - trgt));
  tune->c = tune->c * (0.95 + 0.10 * e / (1.0 + e));
  stats[0] = scal[0]; stats[1] = ess; stats[2] = (double)f; stats[3] = tune->c;
}

// source index of each local output slot (global indices), only when *flag
__global__ void smc_select_kernel(const double* __restrict__ cw, long long npart, unsigned long long seed, unsigned stage,
                                  int systematic, long long first, long long n_local, const int* __restrict__ flag,
                                  long long* __restrict__ idx, unsigned long long* __restrict__ n_clamped) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_local) return;
  const long long g = first + i;
  if (!*flag) { idx[i] = g; return; }
  double ui;
  if (systematic) ui = ((double)g + rng_uniform(seed, 0ULL, stage, RNG_RESAMPLE)) / (double)npart;
  else ui = rng_uniform(seed, (unsigned long long)g, stage, RNG_RESAMPLE);
  long long a = 0, b = npart;
  while (a < b) {
    const long long mid = (a + b) >> 1;
    if (cw[mid] <= ui) a = mid + 1; else b = mid;
  }
  if (a >= npart) { a = npart - 1; atomicAdd(n_clamped, 1ULL); }
  idx[i] = a;
}

// para/loglh/logpost <- gathered rows of the all-gathered copies (only when *flag)
__global__ void smc_take_kernel(const double* __restrict__ para_g, const double* __restrict__ loglh_g,
                                const double* __restrict__ logpost_g, const long long* __restrict__ idx,
                                long long n_local, int np, const int* __restrict__ flag,
                                double* __restrict__ para, double* __restrict__ loglh, double* __restrict__ logpost) {
  if (!*flag) return;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_local * (np + 2)) return;
  const long long i = e / (np + 2);
  const int c = (int)(e % (np + 2));
  const long long s = idx[i];
  if (c < np) para[i * np + c] = para_g[s * np + c];
  else if (c == np) loglh[i] = loglh_g[s];
  else logpost[i] = logpost_g[s];
}

__global__ void smc_equal_weights_kernel(double* __restrict__ w, long long npart, const int* __restrict__ flag) {
  if (!*flag) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npart) w[i] = 1.0 / (double)npart;
}

// Cholesky factor of the particle covariance (lower triangle of R is used), its diagonal, standard errors and
// the log-determinants needed by the proposal densities (smc.jl:161-165, 326-338).  One warp.
//   out: L np x np column-major (upper part zero), rdiag np (variances), aux[0] = sum_k log L_kk,
//        aux[1] = sum_k log rdiag_k;  *err != 0 when R is not positive definite (the reference's cholesky throws)
__global__ void smc_chol_kernel(const double* __restrict__ R, int np, double* __restrict__ L, double* __restrict__ rdiag,
                                double* __restrict__ aux, int* __restrict__ err) {
  __shared__ double A[MAX_EST * MAX_EST];
  __shared__ double piv;
  const int lane = threadIdx.x;
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
    if (lane == 0) {
      const double d = A[j + np * j];
      if (!(d > 0.0) || !isfinite(d)) bad = true;
      piv = sqrt(d);
    }
    __syncwarp();
    const double ljj = piv;
    for (int i = j + lane; i < np; i += 32) A[i + np * j] = (i == j) ? ljj : A[i + np * j] / ljj;
    __syncwarp();
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

struct SmcProposeArgs {
  const double* para; const double* loglh; const double* logpost;     // current particles (local)
  const double* mu; const double* L; const double* rdiag; const double* aux; const SmcTune* tune;
  double alp_mix;                 // Tune.alp: mixture weight of the random-walk proposal
  unsigned long long seed; unsigned stage; long long first, n_local; int np;
  double* prop; double* qx; double* q0;
};

// one mutation proposal per particle (smc.jl:176-200) with the proposal log-densities (smc.jl:326-338).
// `MvNormal(mu, c^2 * Rdiag)` with a VECTOR second argument is Distributions' (deprecated) standard-deviation
// constructor, so the diagonal component's density uses sigma_k = c^2 Rdiag_k; reproduced as coded.
__global__ void __launch_bounds__(128) smc_propose_kernel(SmcProposeArgs a) {
  extern __shared__ double sh[];            // L (np*np), mu (np), rdiag (np)
  const int np = a.np;
  double* Ls = sh; double* mus = Ls + np * np; double* rds = mus + np;
  for (int e = threadIdx.x; e < np * np; e += blockDim.x) Ls[e] = a.L[e];
  for (int e = threadIdx.x; e < np; e += blockDim.x) { mus[e] = a.mu[e]; rds[e] = a.rdiag[e]; }
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_local) return;
  const unsigned long long g = (unsigned long long)(a.first + i);
  const double c = a.tune->c;
  const double LOG2PI = 1.8378770664093453;
  double z[MAX_EST], px[MAX_EST], dx[MAX_EST];
  rng_normal_vector(a.seed, g, a.stage, np, z);
  const double mix = rng_uniform(a.seed, g, a.stage, RNG_MIX);
  const int mixsel = mix < a.alp_mix ? 1 : (mix < a.alp_mix + (1.0 - a.alp_mix) / 2.0 ? 2 : 3);
  const double* p0 = a.para + i * np;
  for (int k = 0; k < np; ++k) {
    double s;
    if (mixsel == 2) s = (c * sqrt(rds[k])) * z[k];
    else { s = 0.0; for (int j = 0; j <= k; ++j) s += (c * Ls[k + np * j]) * z[j]; }
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
    // whitened residuals by forward substitution with c L
    double m1 = 0.0, m0 = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int k = 0; k < np; ++k) {
        const double ctr = (mixsel == 3) ? mus[k] : (pass == 0 ? p0[k] : px[k]);
        double s = ((pass == 0 ? px[k] : p0[k]) - ctr);
        for (int j = 0; j < k; ++j) s -= (c * Ls[k + np * j]) * dx[j];
        dx[k] = s / (c * Ls[k + np * k]);
        if (pass == 0) m1 += dx[k] * dx[k]; else m0 += dx[k] * dx[k];
      }
    }
    qx = c0 - 0.5 * m1; q0 = c0 - 0.5 * m0;
  }
  a.qx[i] = qx; a.q0[i] = q0;
}

struct SmcAcceptArgs {
  double* para; double* loglh; double* logpost;
  const double* prop; const double* lx; const int* status; const double* priorx; const double* qx; const double* q0;
  double phi, dphi; unsigned long long seed; unsigned stage; long long first, n_local; int np;
  unsigned long long* accepted;
};

// accept / reject (smc.jl:202-232)
__global__ void smc_accept_kernel(SmcAcceptArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool acc = false;
  if (i < a.n_local) {
    const double l0 = a.loglh[i];
    const double post0 = a.logpost[i] + a.dphi * l0;
    const double lx = a.status[i] != 0 ? -INFINITY : a.lx[i];
    const double postx = a.phi * lx + a.priorx[i];
    const double alp = exp((postx - a.qx[i]) - (post0 - a.q0[i]));
    const double u = rng_uniform(a.seed, (unsigned long long)(a.first + i), a.stage, RNG_ACCEPT);
    acc = u < alp;
    if (acc) {
      for (int k = 0; k < a.np; ++k) a.para[i * a.np + k] = a.prop[i * a.np + k];
      a.loglh[i] = lx;
      a.logpost[i] = postx;
    } else {
      a.logpost[i] = post0;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(a.accepted, (unsigned long long)__popc(m));
}

// acpt = mean(accepted) over ALL particles (smc.jl:236); counts: one entry per rank
__global__ void smc_acpt_kernel(const unsigned long long* __restrict__ counts, int nranks, long long npart,
                                SmcTune* __restrict__ tune, double* __restrict__ stats) {
  if (blockIdx.x || threadIdx.x) return;
  unsigned long long s = 0;
  for (int r = 0; r < nranks; ++r) s += counts[r];
  tune->acpt = (double)s / (double)npart;
  stats[4] = tune->acpt;
}

__global__ void smc_store_clamped_kernel(const unsigned long long* __restrict__ n_clamped, double* __restrict__ stats) {
  if (blockIdx.x || threadIdx.x) return;
  stats[5] = (double)*n_clamped;
}

// ------------------------------------------------------------------------------------------------
// random-walk Metropolis on many chains (estimation.jl:964-989)
// ------------------------------------------------------------------------------------------------
// proposal y' = y + L z (L = lower Cholesky factor of the proposal covariance), theta' = transform(y')
__global__ void __launch_bounds__(128)
rwmh_propose_kernel(PriorSetDev ps, const double* __restrict__ y, const double* __restrict__ L, int transformed,
                    unsigned long long seed, unsigned step, long long first, long long n,
                    double* __restrict__ yprop, double* __restrict__ theta, double* __restrict__ logjac) {
  extern __shared__ double sh[];
  const int np = ps.np;
  for (int e = threadIdx.x; e < np * np; e += blockDim.x) sh[e] = L[e];
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double z[MAX_EST];
  rng_normal_vector(seed, (unsigned long long)(first + i), step, np, z);
  double lj = 0.0;
  for (int k = 0; k < np; ++k) {
    double s = 0.0;
    for (int j = 0; j <= k; ++j) s += sh[k + np * j] * z[j];
    const double yk = y[i * np + k] + s;
    yprop[i * np + k] = yk;
    double l = 0.0;
    theta[i * np + k] = transformed ? transform_to_param(ps.p[k], yk, l) : yk;
    lj += l;
  }
  logjac[i] = lj;
}

// log target of the proposals and Metropolis accept: log u < logp' - logp
__global__ void rwmh_accept_kernel(double* __restrict__ y, double* __restrict__ logp, const double* __restrict__ yprop,
                                   const double* __restrict__ ll, const int* __restrict__ status,
                                   const double* __restrict__ prior, const double* __restrict__ logjac,
                                   unsigned long long seed, unsigned step, long long first, long long n, int np,
                                   unsigned long long* __restrict__ accepted /* per chain */,
                                   double* __restrict__ hist_y, double* __restrict__ hist_logp /* may be null */) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double lpx = prior[i] + ll[i];
  if (status[i] != 0 || !isfinite(lpx)) lpx = -INFINITY; else lpx += logjac[i];
  const double u = rng_uniform(seed, (unsigned long long)(first + i), step, RNG_ACCEPT);
  const bool acc = log(u) < lpx - logp[i];
  if (acc) {
    for (int k = 0; k < np; ++k) y[i * np + k] = yprop[i * np + k];
    logp[i] = lpx;
    accepted[i] += 1ULL;
  }
  if (hist_y) for (int k = 0; k < np; ++k) hist_y[i * np + k] = y[i * np + k];
  if (hist_logp) hist_logp[i] = logp[i];
}

// initial log target of the chains
__global__ void rwmh_init_kernel(const double* __restrict__ ll, const int* __restrict__ status,
                                 const double* __restrict__ prior, const double* __restrict__ logjac, long long n,
                                 double* __restrict__ logp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = prior[i] + ll[i];
  if (status[i] != 0 || !isfinite(v)) v = -INFINITY; else v += logjac[i];
  logp[i] = v;
}

}  // namespace dyb
