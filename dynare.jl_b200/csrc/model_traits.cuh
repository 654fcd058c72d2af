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
  ST_NONSTATIONARY = 5, ST_F_NOT_PD = 6
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
