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
enum RngStream : unsigned { RNG_RESAMPLE = 0, RNG_MIX = 1, RNG_ACCEPT = 2, RNG_NORMAL = 3, RNG_USER = 4, RNG_PRIOR = 5 };

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

// ---- prior draws (prior_draw!, estimation.jl:329-335: pdraw[i] = rand(prior_i)) -------------------------------------------
// Counter layout of stream RNG_PRIOR: (draw index, round, block) with block = k << 12 | attempt << 2 | sub for parameter k:
// attempt t < 512 of a rejection sampler uses the normal pair of sub 0 / 2 and the uniform of sub 1 / 3 (first / second
// Gamma variate of the parameter), attempt 512 + g holds the uniform of the shape-below-one boost of variate g.
// Gamma(a, 1) by Marsaglia & Tsang (2000): d = a - 1/3, c = 1 / sqrt(9 d), v = (1 + c z)^3, accept when
// log u < z^2 / 2 + d - d v + d log v; a < 1: Gamma(a + 1, 1) u^(1/a).  oracle/sampler.py restates it in numpy.
DYB_D double rng_gamma(unsigned long long seed, unsigned long long idx, unsigned round, unsigned kbase, int g, double a) {
  double boost = 1.0;
  if (a < 1.0) {
    const double u = rng_uniform(seed, idx, round, RNG_PRIOR, kbase | ((512u + (unsigned)g) << 2));
    boost = pow(1.0 - u, 1.0 / a);                     // 1 - u in (0, 1]
    a += 1.0;
  }
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (unsigned t = 0; t < 512u; ++t) {
    double z, z1;
    rng_normal_pair(seed, idx, round, RNG_PRIOR, kbase | (t << 2) | (unsigned)(2 * g), z, z1);
    double v = 1.0 + c * z;
    if (v <= 0.0) continue;
    v = v * v * v;
    const double u = 1.0 - rng_uniform(seed, idx, round, RNG_PRIOR, kbase | (t << 2) | (unsigned)(2 * g + 1));
    if (log(u) < 0.5 * z * z + d - d * v + d * log(v)) return d * v * boost;
  }
  return d * boost;
}

DYB_D double prior_draw(const PriorDevK& p, unsigned long long seed, unsigned long long idx, unsigned round, int k) {
  const unsigned kbase = (unsigned)k << 12;
  switch (p.shape) {
    case 1: { const double x = rng_gamma(seed, idx, round, kbase, 0, p.a), y = rng_gamma(seed, idx, round, kbase, 1, p.b);
              return x / (x + y); }
    case 2: return p.b * rng_gamma(seed, idx, round, kbase, 0, p.a);
    case 3: { double z, z1; rng_normal_pair(seed, idx, round, RNG_PRIOR, kbase, z, z1); return p.a + p.b * z; }
    case 4: return sqrt(p.b / rng_gamma(seed, idx, round, kbase, 0, p.a));       // sqrt of InverseGamma(a, b) (inversegamma1.jl:167)
    case 5: return p.a + (p.b - p.a) * rng_uniform(seed, idx, round, RNG_PRIOR, kbase);
    case 6: return p.b / rng_gamma(seed, idx, round, kbase, 0, p.a);
    case 8: return p.b * pow(-log(1.0 - rng_uniform(seed, idx, round, RNG_PRIOR, kbase)), 1.0 / p.a);
    default: return NAN;
  }
}

// para[:, i] = a draw from the priors for every i < n whose status is non-zero (status null: every i); the draw of particle
// first + i in round r depends on (seed, first + i, r) only
__global__ void prior_draw_kernel(PriorSetDev ps, unsigned long long seed, long long first, long long n, unsigned round,
                                  const int* __restrict__ status, double* __restrict__ para, unsigned long long* __restrict__ n_drawn) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (status && status[i] == 0) return;
  for (int k = 0; k < ps.np; ++k) para[i * ps.np + k] = prior_draw(ps.p[k], seed, (unsigned long long)(first + i), round, k);
  if (n_drawn) atomicAdd(n_drawn, 1ULL);
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

// The stage kernels themselves are in smc_stage.cuh.

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
