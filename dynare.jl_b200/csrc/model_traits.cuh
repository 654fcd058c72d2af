// Common definitions for generated model headers and the kernels that consume them.
#pragma once
#include <cstdint>
#include <math.h>

#if defined(__CUDACC__)
#define DYB_HD __host__ __device__ __forceinline__
#define DYB_D __device__ __forceinline__
#else
#define DYB_HD inline
#define DYB_D inline
#endif

namespace dyb {

// estimated-parameter kinds (reference src/dynare_containers.jl:16 `EstimatedParameterType`,
// scattered by set_estimated_parameters!, src/estimation/estimation.jl:517-601)
enum EstKind : int {
  EST_PARAM = 0, EST_SD_SHOCK = 1, EST_VAR_SHOCK = 2, EST_CORR_SHOCK = 3,
  EST_SD_MEAS = 4, EST_VAR_MEAS = 5, EST_CORR_MEAS = 6
};

// per-draw status (include/dynare_b200.h DYB_ST_*)
enum DrawStatus : int {
  ST_OK = 0, ST_STEADY = 1, ST_INDETERMINATE = 2, ST_UNSTABLE = 3, ST_SOLVER = 4,
  ST_NONSTATIONARY = 5, ST_F_NOT_PD = 6,
  ST_REDO = 64          // internal to the solver: the fast pass could not verify this draw, the Householder pass takes it
};

constexpr int MAX_EST = 64;

DYB_HD constexpr int tri(int i, int j) {            // packed lower-triangular index, i >= j
  return i * (i + 1) / 2 + j;
}
DYB_HD constexpr int sym(int i, int j) { return i >= j ? tri(i, j) : tri(j, i); }

#if defined(__CUDACC__)
// cyclic reduction: relative change of ||A0(j)|| below which a non-vanishing A0 sequence is read as a unit root
constexpr double CR_UNIT_ROOT_EPS = 1e-6;

// doubling Lyapunov solve: the model counts as non-stationary when a root lies within 1e-6 of the unit circle (the
// reference splits at 1 - 1e-6, src/filters/kalman/kalman.jl:172-180): after LYAP_UNIT_ROOT_IT + 1 doublings the
// iterate is A^(2^20), whose norm is (1 - 1e-6)^(2^20) = 0.35 at that boundary
constexpr int LYAP_UNIT_ROOT_IT = 19;
constexpr double LYAP_UNIT_ROOT_NORM = 0.35;

// Newton's method on the static model of M (generated static_fj / static_resid_ss), for models without a closed-form
// steady state: full steps, halved while they do not reduce the sum of squared residuals; y holds the starting point
// on entry and the steady state on exit, NaN when the iteration fails (the caller's residual check then gives status 1).
// The reference solves the same system with a trust-region solver from the same point
// (src/steady_state/SteadyState.jl:183-252); both stop far below its acceptance threshold cbrt(eps).
constexpr int NEWTON_MAXIT = 50;
constexpr double NEWTON_TOL = 1e-13;
constexpr int NEWTON_HALVINGS = 30;
constexpr double NEWTON_STEP_TOL = 1e-6;
constexpr double NEWTON_ESCAPE = 1e8;       // iterates larger than this multiple of the starting point: a root at infinity
template <class M>
DYB_HD void newton_steady_state(const double* __restrict__ p, double* __restrict__ y) {
  constexpr int N = M::N;
  double f[N], J[N * N], yt[N];
  bool done = false;
  double y0max = 1.0, last_step = 0.0;
  for (int i = 0; i < N; ++i) y0max = y0max > fabs(y[i]) ? y0max : fabs(y[i]);
  for (int it = 0; it < NEWTON_MAXIT && !done; ++it) {
    for (int e = 0; e < N * N; ++e) J[e] = 0.0;
    M::static_fj(p, y, f, J);
    double fmax = 0.0, ymax = 0.0, g0 = 0.0;
    bool finite = true;
    for (int i = 0; i < N; ++i) {
      finite = finite && (f[i] - f[i] == 0.0);
      fmax = fmax > fabs(f[i]) ? fmax : fabs(f[i]);
      ymax = ymax > fabs(y[i]) ? ymax : fabs(y[i]);
      g0 += f[i] * f[i];
    }
    if (!finite || ymax > NEWTON_ESCAPE * y0max) break;
    // small residuals AND a small last step: along a root at infinity (e.g. 1/c -> 0) the residuals shrink while the
    // iterates keep doubling
    if (fmax <= NEWTON_TOL * (1.0 + ymax) && last_step <= NEWTON_STEP_TOL * (1.0 + ymax)) { done = true; break; }
    // solve J dy = -f by LU with partial pivoting (in place; f becomes dy)
    bool singular = false;
    for (int c = 0; c < N; ++c) {
      int piv = c;
      double best = fabs(J[c + N * c]);
      for (int r = c + 1; r < N; ++r) if (fabs(J[r + N * c]) > best) { best = fabs(J[r + N * c]); piv = r; }
      if (!(best > 0.0)) { singular = true; break; }
      if (piv != c) {
        for (int q = 0; q < N; ++q) { const double t = J[c + N * q]; J[c + N * q] = J[piv + N * q]; J[piv + N * q] = t; }
        const double t = f[c]; f[c] = f[piv]; f[piv] = t;
      }
      const double inv = 1.0 / J[c + N * c];
      for (int r = c + 1; r < N; ++r) {
        const double l = J[r + N * c] * inv;
        if (l != 0.0) {
          for (int q = c + 1; q < N; ++q) J[r + N * q] -= l * J[c + N * q];
          f[r] -= l * f[c];
        }
      }
    }
    if (singular) break;
    for (int r = N - 1; r >= 0; --r) {
      double x = -f[r];
      for (int q = r + 1; q < N; ++q) x -= J[r + N * q] * f[q];
      f[r] = x / J[r + N * r];
    }
    double t = 1.0;
    bool accepted = false;
    for (int h = 0; h < NEWTON_HALVINGS; ++h) {
      for (int i = 0; i < N; ++i) yt[i] = y[i] + t * f[i];
      const double g = M::static_resid_ss(p, yt);
      if (g - g == 0.0 && g < g0) { accepted = true; break; }
      t *= 0.5;
    }
    if (!accepted) break;
    last_step = 0.0;
    for (int i = 0; i < N; ++i) { last_step = last_step > fabs(yt[i] - y[i]) ? last_step : fabs(yt[i] - y[i]); y[i] = yt[i]; }
  }
  if (!done) for (int i = 0; i < N; ++i) y[i] = NAN;
}

// fast fp64 reciprocal square root: hardware seed (MUFU.RSQ64H, ~20 bits) + two Newton steps (<= 1 ulp for
// normal inputs); callers guard non-positive / non-finite arguments separately
DYB_D double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}

// Running product of the innovation determinants with its binary exponent split off after every factor, so that the
// T-step recursion takes no logarithm: prod = m 2^e with m in [1, 2), log(prod) = log(m) + e log 2 at the end.  A zero,
// denormal or non-finite product (never for a positive-definite F) is folded by a logarithm as it always was, which keeps
// the -Inf / NaN outcome of such a draw.  (The threshold-triggered log this replaces ran whenever ANY lane of the warp
// left [1e-200, 1e200]: 36 % of the periods of fs2000, 158 instructions each.)
struct LogDetAcc {
  double m = 1.0, extra = 0.0;
  int e = 0;
  DYB_D void mul(double det) {
    m *= det;
    const int hi = __double2hiint(m);
    const int ef = (hi >> 20) & 0x7ff;
    if (ef == 0 || ef == 0x7ff) { extra += log(m); m = 1.0; }
    else { e += ef - 1023; m = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(m)); }
  }
  DYB_D double value() const { return extra + (log(m) + (double)e * 0.6931471805599453); }
};

// fast fp64 reciprocal: hardware seed (MUFU.RCP64H) + two Newton steps; callers guard zero / non-finite arguments
DYB_D double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  y = fma(y, fma(-x, y, 1.0), y);
  y = fma(y, fma(-x, y, 1.0), y);
  return y;
}
#endif

}  // namespace dyb
