/*
 * dynare_b200.h -- C ABI of libdynare_b200.so, the B200-native batched first-order DSGE
 * log-likelihood engine that replaces Dynare.jl's per-draw likelihood call.
 *
 * Every entry point cites the reference interface (file:line under the Dynare.jl tree) it
 * replaces.  Conventions:
 *   - plain pointers and sizes only; matrices are column-major (Julia layout); indices 0-based;
 *   - batches are "draw-major": the arrays of draw d start at  base + d * (size of one draw);
 *   - every function returns 0 on success or a DYB_ERR_* code (dyb_last_error() has the text);
 *     a failure of one draw is NOT an error: it is reported in status[d] with loglik[d] = -Inf,
 *     which is what the reference's try/catch does (src/estimation/estimation.jl:503-515);
 *   - a handle owns one CUDA stream, its device buffers and pinned staging; it is not thread-safe
 *     (one handle per host thread and GPU).  There is no CPU fallback.
 */
#ifndef DYNARE_B200_H
#define DYNARE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYB_VERSION 100

/* API return codes */
#define DYB_OK 0
#define DYB_ERR_ARG 1        /* bad argument (null pointer, size mismatch, unknown model) */
#define DYB_ERR_CUDA 2       /* CUDA runtime error */
#define DYB_ERR_STATE 3      /* call sequence error (e.g. no observations set) */
#define DYB_ERR_NCCL 4
#define DYB_ERR_UNSUPPORTED 5

/* per-draw status codes; loglik = -Inf whenever status != 0.
 * They name the exceptions the reference turns into -Inf (src/estimation/estimation.jl:16, 503-515). */
#define DYB_ST_OK 0
#define DYB_ST_STEADY 1          /* steady state not finite / static residuals > tol (SteadyState.jl:476-482), DomainError */
#define DYB_ST_INDETERMINATE 2   /* UndeterminateSystemException */
#define DYB_ST_UNSTABLE 3        /* UnstableSystemException */
#define DYB_ST_SOLVER 4          /* singular pivot / NaN in the solver (LAPACKException analogue) */
#define DYB_ST_NONSTATIONARY 5   /* Lyapunov doubling diverged: no finite unconditional variance */
#define DYB_ST_F_NOT_PD 6        /* innovation covariance not positive definite */

/* estimated-parameter kinds (src/dynare_containers.jl:16; scatter: src/estimation/estimation.jl:517-601) */
#define DYB_EST_PARAM 0
#define DYB_EST_SD_SHOCK 1
#define DYB_EST_VAR_SHOCK 2
#define DYB_EST_CORR_SHOCK 3
#define DYB_EST_SD_MEAS 4
#define DYB_EST_VAR_MEAS 5
#define DYB_EST_CORR_MEAS 6

typedef struct dyb_handle dyb_handle;

/* Sizes of a compiled model: what `Model` and `SSWs` carry
 * (src/dynare_containers.jl:94-188, src/estimation/estimation.jl:358-429). */
typedef struct dyb_model_dims {
  int32_t n;        /* endogenous incl. auxiliaries */
  int32_t nx;       /* exogenous shocks */
  int32_t npar;     /* model parameters */
  int32_t k;        /* n_states = n_bkwrd + n_both */
  int32_t nf;       /* forward variables (n_fwrd + n_both) */
  int32_t s;        /* static variables */
  int32_t ncol;     /* columns of the compressed Jacobian */
  int32_t ns;       /* Kalman states */
  int32_t ny;       /* observables */
  int32_t np_default; /* estimated parameters declared in the model file */
  int32_t ss_doubles; /* doubles per draw handed from the solve stage to the filter stage */
} dyb_model_dims;

/* Numerical options (reference: StochSimulOptions src/perturbations.jl:1-45, SteadyOptions tolf
 * src/steady_state/SteadyState.jl:183, kalman presample src/estimation/estimation.jl:659-661). */
typedef struct dyb_options {
  double cr_tol;            /* cyclic-reduction stopping tolerance on ||A0||_1, ||A2||_1   (default 1e-8)  */
  int32_t cr_maxit;         /* (default 100) */
  double steady_tol;        /* 2-norm of static residuals (default cbrt(eps)) */
  double lyap_tol;          /* relative increment at which doubling stops (default 1e-16) */
  int32_t lyap_maxit;       /* (default 60) */
  int32_t presample;        /* periods excluded from the likelihood sum (default 0) */
  int32_t solver;           /* 0 auto; 1 one thread per draw (any size); 2 cooperative register-resident (m <= 12) */
} dyb_options;

/* ---- library / registry ------------------------------------------------------------------- */
int dyb_version(void);
const char* dyb_last_error(void);
int dyb_device_count(int* count);
int dyb_model_count(void);
const char* dyb_model_name(int index);
int dyb_model_get_dims(const char* model, dyb_model_dims* dims);

/* ---- handle ----------------------------------------------------------------------------------
 * dyb_create replaces the construction of `SSWs(context, observations, varobs)` and of the cached LRE
 * workspaces (src/estimation/estimation.jl:376-428, src/dynare_containers.jl:964-1013). */
int dyb_create(const char* model, int device, dyb_handle** out);
int dyb_destroy(dyb_handle* h);
int dyb_get_dims(const dyb_handle* h, dyb_model_dims* dims);
int dyb_get_options(const dyb_handle* h, dyb_options* opt);
int dyb_set_options(dyb_handle* h, const dyb_options* opt);

/* calibrated values kept by non-estimated entries: context.work.params, model.Sigma_e, work.Sigma_m
 * (src/dynare_containers.jl:846-934).  Any pointer may be NULL to keep the current values. */
int dyb_set_calibration(dyb_handle* h, const double* params /* npar */,
                        const double* sigma_e /* nx*nx */, const double* sigma_m /* ny*ny */);
/* the estimated-parameter map: EstimatedParameters.index / .parametertype (src/dynare_containers.jl:800-841) */
int dyb_set_estimated_params(dyb_handle* h, int32_t np, const int32_t* kind,
                             const int32_t* index1, const int32_t* index2);
/* observations: ny x nobs column-major, NaN = missing (the Matrix{Union{Missing,Float64}} of
 * src/estimation/estimation.jl:603-608 after get_transposed_data!, src/data.jl:107-140) */
int dyb_set_observations(dyb_handle* h, const double* Y, int32_t ny, int32_t nobs);

/* ---- the hot path: loglikelihood(theta, context, observations, ssws) for N draws ---------------
 * (src/estimation/estimation.jl:603-702).  theta: np x N column-major.  Host pointers; pinned host
 * memory (dyb_host_alloc) is copied asynchronously without staging. */
int dyb_loglik_batch(dyb_handle* h, const double* theta, int64_t n_draws,
                     double* loglik /* N */, int32_t* status /* N */);
/* same, all pointers are device pointers on the handle's device; asynchronous on the handle's stream */
int dyb_loglik_batch_device(dyb_handle* h, const double* d_theta, int64_t n_draws,
                            double* d_loglik, int32_t* d_status);
int dyb_sync(dyb_handle* h);
/* run the handle's work on a caller-owned CUDA stream (cudaStream_t passed as void*; NULL restores the
 * handle's private stream) -- lets a host framework order and time the engine with its own events */
int dyb_set_stream(dyb_handle* h, void* cuda_stream);
/* number of kernels launched by this handle so far (for the bench's gpu_launches claim) */
int64_t dyb_kernel_launches(const dyb_handle* h);
/* elapsed device time (ms) of the last dyb_loglik_batch[_device] call's solve and filter kernels,
 * from CUDA events on the handle's stream; valid after dyb_sync */
int dyb_last_kernel_ms(dyb_handle* h, float* solve_ms, float* filter_ms);
/* accumulated CUDA-event timings of the solve / filter kernels over every likelihood batch since the last
 * reset (up to 4096 batches); read synchronises the stream */
int dyb_kernel_timers_reset(dyb_handle* h);
int dyb_kernel_timers_read(dyb_handle* h, int32_t* n_batches, float* solve_ms_total, float* filter_ms_total);
/* measured fp64 FMA throughput of the device (TFLOP/s) from a register-resident DFMA loop on all SMs:
 * the roofline denominator for this path, which MEASURED_PEAKS.json does not carry */
int dyb_fp64_peak(int device, double* tflops);

/* ---- stage-level entry points (host pointers) ---------------------------------------------------- */
/* LRE.first_order_solver! (src/perturbations.jl:663-664): g1 = [g1_1 | g1_2], n x (k+nx) per draw,
 * rows in declaration order (src/perturbations.jl:82-107); ybar n per draw; cr_iters / lyap_iters (iterations
 * taken by cyclic reduction and by the doubling Lyapunov solve) may be NULL. */
int dyb_first_order_batch(dyb_handle* h, const double* theta, int64_t n_draws,
                          double* g1, double* ybar, int32_t* cr_iters, int32_t* lyap_iters,
                          int32_t* status);
/* state space handed to the filter (src/estimation/estimation.jl:640-658), per draw:
 *   T ns x ns, RQR ns x ns (= R Q R'), P0 ns x ns, ybar_obs ny, H ny x ny; any output may be NULL */
int dyb_state_space_batch(dyb_handle* h, const double* theta, int64_t n_draws,
                          double* T, double* RQR, double* P0, double* ybar_obs, double* H,
                          int32_t* status);

/* generic batched Kalman log-likelihood for arbitrary (dense) state spaces -- the
 * kalman_likelihood(Y,Z,H,T,R,Q,a0,P,start,last,presample,ws) call (src/estimation/estimation.jl:687-700)
 * without a model: per draw T ns x ns, RQR ns x ns, P0 ns x ns, a0 ns (NULL = 0), H ny x ny (NULL = 0);
 * z_idx[ny] = state observed by each row of the selection matrix Z; Y ny x nobs shared by all draws. */
int dyb_kalman_batch(int device, int32_t ns, int32_t ny, int32_t nobs, int32_t presample,
                     const int32_t* z_idx, const double* Y,
                     const double* T, const double* RQR, const double* P0, const double* a0,
                     const double* H, int64_t n_draws, double* loglik, int32_t* status);

/* ---- SMC primitives on device (src/estimation/smc.jl) -------------------------------------------- */
/* correction step (smc.jl:118-123) + ESS (smc.jl:129): w_out = w * exp(dphi * loglh) / zhat */
int dyb_smc_reweight(int device, const double* w, const double* loglh, double dphi, int64_t n,
                     double* w_out, double* zhat, double* ess);
/* multinomial_resampling (smc.jl:340-361) with caller-provided uniforms: idx[i] = first j with
 * u[i] < cw[j] (0-based), cw = canonical blocked cumulative sum (1024-weight blocks, sequential
 * inside a block, sequential over block totals); clamped to n-1, *n_clamped counts clamps. */
int dyb_smc_resample_multinomial(int device, const double* w, const double* u, int64_t n,
                                 int64_t* idx, int64_t* n_clamped);
/* systematic resampling (the commented call smc.jl:133): u_i = (i + u0) / n */
int dyb_smc_resample_systematic(int device, const double* w, double u0, int64_t n,
                                int64_t* idx, int64_t* n_clamped);
/* weighted mean and covariance of the particles (smc.jl:154-161): para n x np ROW-major particles
 * (np x n column-major), mu np, R np x np */
int dyb_smc_moments(int device, const double* para, const double* w, int64_t n, int32_t np,
                    double* mu, double* R);

/* ---- pinned host memory ---------------------------------------------------------------------------- */
int dyb_host_alloc(void** ptr, int64_t bytes);
int dyb_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* DYNARE_B200_H */
