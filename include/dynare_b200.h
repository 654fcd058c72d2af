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
  int32_t mid_doubles; /* doubles per draw exchanged between the kernels of the solve stage */
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
  int32_t solver;           /* 0 auto (m <= 12: cooperative register-resident solver whose fast pass -- Gaussian elimination
                               without pivoting on equations pre-ordered at the calibration, result verified by its residual --
                               is followed by a Householder pass on the draws it could not verify); 1 one thread per draw (any
                               size; for run-time-size and medium models: the shared-memory kernel instead of the register-
                               blocked one); 2 cooperative solver, Householder pass only */
  int32_t filter_lanes;     /* lanes per draw of the model-specific Kalman kernel: 0 auto (16 up to 2 304 draws per call, else 1), 1 = one thread per draw (throughput kernel), 4, 8 = shared-tile kernel with the time
                             * update dealt over the lanes, 16, 32 = every stage of a period dealt over the lanes; the small-batch
                             * kernels are for calls that leave schedulers idle (an SMC stage on a shard); same sums in the same
                             * order: every variant returns the same bits */
  int32_t univariate_fallback; /* what a period whose innovation covariance F fails its Cholesky factorisation does:
                             * 0 (default) the draw is rejected (status 6, loglik -Inf); 1 the period is processed one observation
                             * at a time, observations with conditional variance <= 1e-10 skipped -- the behaviour of
                             * KalmanFilterTools' kalman_likelihood, the package behind the reference call
                             * src/estimation/estimation.jl:671-700 (restated from the published package: its source is not in
                             * the reference tree); needs diagonal measurement-error covariances; 2 every period of every draw
                             * univariate (diagnostic: must reproduce the multivariate likelihood).  Flagged draws are
                             * re-evaluated by a slow dense kernel; the production kernels are untouched. */
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
 * src/estimation/estimation.jl:603-608 after get_transposed_data!, src/data.jl:107-140).  Only the steady state of the
 * observed variables is removed on the device (estimation.jl:660-663); the reference's other detrending variants -- linear
 * trends, `prefilter`, log / first-difference transformations of the data file (src/data.jl:169-208, 332-353) -- are data
 * preparation done ONCE per estimation and stay with the caller: pass the prepared Y. */
int dyb_set_observations(dyb_handle* h, const double* Y, int32_t ny, int32_t nobs);

/* ---- the hot path: loglikelihood(theta, context, observations, ssws) for N draws ---------------
 * (src/estimation/estimation.jl:603-702).  theta: np x N column-major.  Host pointers; pinned host
 * memory (dyb_host_alloc) is copied asynchronously without staging. */
int dyb_loglik_batch(dyb_handle* h, const double* theta, int64_t n_draws,
                     double* loglik /* N */, int32_t* status /* N */);
/* same without the final synchronisation (results valid after dyb_sync); buffers must be pinned or device memory.
 * Two handles used alternately overlap one batch's PCIe copies with the other's kernels. */
int dyb_loglik_batch_async(dyb_handle* h, const double* theta, int64_t n_draws,
                           double* loglik /* N */, int32_t* status /* N */);
/* same, all pointers are device pointers on the handle's device; asynchronous on the handle's stream */
int dyb_loglik_batch_device(dyb_handle* h, const double* d_theta, int64_t n_draws,
                            double* d_loglik, int32_t* d_status);
int dyb_sync(dyb_handle* h);
/* non-blocking completion query for the asynchronous calls: *done = 1 once everything enqueued on the handle finished */
int dyb_query(dyb_handle* h, int32_t* done);
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
/* split of the armed batches' solve stage into its three kernels (model set-up, cyclic reduction, assembly);
 * call before dyb_kernel_timers_read; zeros on the one-kernel solver path */
int dyb_kernel_timers_read_solve_split(dyb_handle* h, float* model_ms_total, float* cr_ms_total,
                                       float* assemble_ms_total);
/* draws that the cooperative solver's fast pass (dyb_options.solver = 0) could not verify and its Householder pass solved
 * again, since the handle was created: the fast pass eliminates without pivoting and checks the residuals of
 * A0 + A1 G + A2 G[fw,:] G[bk,:] = 0 and of the g_u system against 1e-12 of their own scale (csrc/solve_coop.cuh) */
int dyb_solver_redo_count(dyb_handle* h, int64_t* n_redone);
/* measured fp64 FMA throughput of the device (TFLOP/s) from a register-resident DFMA loop on all SMs:
 * the roofline denominator for this path, which MEASURED_PEAKS.json does not carry */
int dyb_fp64_peak(int device, double* tflops);
/* same for the fp64 tensor-core path (mma.sync ... f64): shape 0 = m8n8k4, 1 = m16n8k16 */
int dyb_fp64_peak_dmma(int device, int32_t shape, double* tflops);

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

/* prior-predictive quantities for N draws (run_simulations, src/estimation/priorprediction.jl:64-98): steady state,
 * stdev = robustsqrt(diag(endogenous_variance)) of every endogenous variable (src/perturbations.jl:195-205) and the
 * impulse responses of irfs! (src/perturbations.jl:670-694) to a one-standard-deviation shock.  Per draw: ybar n,
 * stdev n, irfs n x periods x nx (column-major); status as in dyb_first_order_batch; any output but status may be NULL */
int dyb_prior_predictive_batch(dyb_handle* h, const double* theta, int64_t n_draws, int32_t periods,
                               double* ybar, double* stdev, double* irfs, int32_t* status);

/* fixed-interval Kalman smoother for N dense state spaces at once (the `kalman_smoother!` call of calibsmoother!,
 * src/filters/kalman/kalman.jl:126-150): smoothed states alphah ns x nobs, smoothed shocks etah nx x nobs, smoothed
 * measurement errors epsilonh ny x nobs, updated states att ns x nobs, predicted states a_pred ns x (nobs+1) per draw
 * (any may be NULL), log-likelihood and status.  R ns x nx and Q nx x nx per draw; z_idx / Y / H / a0 as in
 * dyb_kalman_batch. */
int dyb_kalman_smoother_batch(int device, int32_t ns, int32_t ny, int32_t nx, int32_t nobs, const int32_t* z_idx,
                              const double* Y, const double* T, const double* R, const double* Q, const double* P0,
                              const double* a0, const double* H, int64_t n_draws, double* alphah, double* etah,
                              double* epsilonh, double* att, double* a_pred, double* loglik, int32_t* status);
/* exact-diffuse Kalman filter and smoother for N dense state spaces with unit roots (the `diffuse_kalman_smoother!`
 * branch of calibsmoother!, src/filters/kalman/kalman.jl:163-223), in the ORIGINAL basis: with [Z1 Z2] the ordered
 * real Schur vectors of T (moduli > 1 - 1e-6 first, kalman.jl:166-180) and P22 the Lyapunov solution on the stable
 * block, the caller passes Pinf0 = Z1 Z1' and Pstar0 = Z2 P22 Z2' (the reference rotates the state space instead and
 * rotates the results back, kalman.jl:215-216 -- the filter is equivariant).  Univariate treatment of the observations,
 * so Finf may be rank deficient; Hdiag ny per draw (NULL = 0) is the diagonal of H.  A period is diffuse while
 * max diag(Pinf) > tol (the reference passes 1e-8, kalman.jl:211).  loglik is the diffuse log-likelihood
 * (Durbin & Koopman 2012, eq. 7.3), n_diffuse the number of diffuse periods; other arguments and outputs as in
 * dyb_kalman_smoother_batch; ybar_obs ny per draw (NULL = 0) is subtracted from the observations; with Pinf0 = 0 the
 * results are those of dyb_kalman_smoother_batch.
 * NOT batched on the device: the ordered real Schur decomposition that yields Z1, Z2 (LAPACK gees per draw, as the
 * reference calls it) runs on the HOST, one draw after another (the Python mirror's diffuse_initial_covariances, the
 * Julia shim's schur!), and the stable block's Lyapunov equation goes through dyb_lyapunov_batch; this smoother path
 * is a post-estimation call on a handful of draws, not part of the per-draw likelihood. */
int dyb_kalman_diffuse_smoother_batch(int device, int32_t ns, int32_t ny, int32_t nx, int32_t nobs, const int32_t* z_idx,
                                      const double* Y, const double* T, const double* R, const double* Q,
                                      const double* Pinf0, const double* Pstar0, const double* a0, const double* Hdiag,
                                      const double* ybar_obs, double tol, int64_t n_draws, double* alphah, double* etah, double* epsilonh,
                                      double* att, double* a_pred, double* loglik, int32_t* n_diffuse, int32_t* status);
/* Sigma_e (nx x nx) and Sigma_m (ny x ny) of every draw as set_estimated_parameters! builds them
 * (src/estimation/estimation.jl:517-601: stderr -> variance, corr -> covariance); either output may be NULL */
int dyb_shock_covariances_batch(dyb_handle* h, const double* theta, int64_t n_draws, double* sigma_e, double* sigma_m);
/* calibsmoother! (src/filters/kalman/kalman.jl:52-264) for N parameter draws: every endogenous variable is a state
 * (T[:, i_bkwrd_b] = g1_1, R = g1_2, P0 = endogenous_variance), observations detrended by the steady state; outputs in
 * LEVELS (steady state added back, kalman.jl:240-243): alphah n x nobs, etah nx x nobs, epsilonh ny x nobs per draw.
 * H is the measurement-error covariance of the draw (the reference's calibsmoother! uses H = 0, which is the same thing
 * for models without measurement errors).  Stationary models only (status 5 otherwise). */
int dyb_smoother_batch(dyb_handle* h, const double* theta, int64_t n_draws, double* alphah, double* etah,
                       double* epsilonh, double* loglik, int32_t* status);

/* unconditional forecast for N draws (forecasting_ -> forecast_!, src/forecast.jl:54-159): y0 n per draw in levels
 * (last smoothed state in calibsmoother mode, histval otherwise), Y n x (periods+1) per draw, Y[:,0] = y0 */
int dyb_forecast_batch(dyb_handle* h, const double* theta, int64_t n_draws, int32_t periods, const double* y0,
                       double* Y, int32_t* status);

/* generic batched Kalman log-likelihood for arbitrary (dense) state spaces -- the
 * kalman_likelihood(Y,Z,H,T,R,Q,a0,P,start,last,presample,ws) call (src/estimation/estimation.jl:687-700)
 * without a model: per draw T ns x ns, RQR ns x ns, P0 ns x ns, a0 ns (NULL = 0), H ny x ny (NULL = 0);
 * z_idx[ny] = state observed by each row of the selection matrix Z; Y ny x nobs shared by all draws. */
int dyb_kalman_batch(int device, int32_t ns, int32_t ny, int32_t nobs, int32_t presample,
                     const int32_t* z_idx, const double* Y,
                     const double* T, const double* RQR, const double* P0, const double* a0,
                     const double* H, int64_t n_draws, double* loglik, int32_t* status);

/* dyb_kalman_batch followed by the univariate fallback (see dyb_options.univariate_fallback; mode 1 or 2) on the draws
 * whose innovation covariance was not positive definite */
int dyb_kalman_batch_fallback(int device, int32_t ns, int32_t ny, int32_t nobs, int32_t presample,
                              const int32_t* z_idx, const double* Y,
                              const double* T, const double* RQR, const double* P0, const double* a0,
                              const double* H, int64_t n_draws, int32_t mode, double* loglik, int32_t* status);

/* which dense filter kernel dyb_kalman_batch / dyb_loglik_from_jacobian_batch use: 0 auto (four / eight lanes per draw up to
 * ns = 16 and ny = 8, tensor-core kernels up to ns = 256, scalar kernel beyond), 1 scalar, 2 tensor-core (fp64 mma.sync),
 * 3 tensor-core with the large-ns kernel's cp.async.bulk panel staging switched off, 4 lanes-per-draw kernel (ns <= 16,
 * ny <= 8 only), 5 tensor-core with the earlier ns <= 88 kernel (rank-ny downdate before the products, kalman_dmma.cuh)
 * instead of the folded one (kalman_dmma_fold.cuh); process-wide, for A/B tests */
int dyb_set_kalman_variant(int32_t variant);

/* ---- models without a generated device function -----------------------------------------------------------
 * The host evaluates steady state and Jacobian per draw (compute_steady_state!, get_dynamic_jacobian!,
 * set_compressed_jacobian!: src/steady_state/SteadyState.jl:183-252, src/model.jl:120-139,
 * src/perturbations.jl:616-635) -- e.g. on Julia threads -- and uploads them; everything after that runs here.
 * dyb_model_desc carries what `Model` / `LRE.Indices` / `SSWs` carry (src/dynare_containers.jl:110-131, 602;
 * src/estimation/estimation.jl:384-390): strictly increasing 0-based variable indices. */
typedef struct dyb_model_desc {
  int32_t n, nx, ny;
  int32_t n_static, n_bkwrd_b, n_fwrd_b;
  const int32_t* i_static;    /* variables appearing only at the current period */
  const int32_t* i_bkwrd_b;   /* variables with a lag (backward and both) = the solution's state columns */
  const int32_t* i_fwrd_b;    /* variables with a lead (forward and both) */
  const int32_t* obs_idx;     /* ny: observed variable of each row of Y, in varobs order */
} dyb_model_desc;
/* a handle with run-time sizes; dyb_set_observations / dyb_set_options / dyb_get_dims / dyb_destroy apply, the
 * theta-based entry points return DYB_ERR_UNSUPPORTED */
int dyb_create_generic(const dyb_model_desc* desc, int device, dyb_handle** out);
/* The two *_from_jacobian entry points below also accept a handle made by dyb_create(name) for a "SIZES-ONLY" model: a
 * translation unit that instantiates the register-resident kernels of the generated-model path for the model's sizes and
 * index sets, without a model function (host side: modelgen.emit_cuda_jacobian_fed + build.build_sizes_library, i.e.
 * engine.GenericSSWs(..., compile = True); a dynamic block of at most 12 variables).  Same arguments, same results to
 * solver accuracy; ls2003 with device-resident Jacobians: 30 M evaluations/s against 5.9 M for the run-time-size kernels
 * (fs2000 22.9 M, hs_nk 80 M; from pinned host memory the 1.2 - 3.7 KB per draw of PCIe traffic bound it at 9 - 26 M). */
/* LRE.first_order_solver! + compute_variance! + state-space assembly from compressed Jacobians
 * (n x ncol per draw, column-major, columns [lags of i_bkwrd_b | all n current | leads of i_fwrd_b | nx shocks]).
 * sigma_e: nx x nx, one per draw (sigma_e_per_draw != 0) or shared.  Outputs per draw (any may be NULL):
 * g1 n x (k+nx), T / RQR / P0 ns x ns, iterations taken. */
int dyb_first_order_from_jacobian_batch(dyb_handle* h, const double* jac, const double* sigma_e,
                                        int32_t sigma_e_per_draw, int64_t n_draws, double* g1, double* T,
                                        double* RQR, double* P0, int32_t* cr_iters, int32_t* lyap_iters,
                                        int32_t* status);
/* loglikelihood (src/estimation/estimation.jl:603-702) from host-evaluated Jacobians: ybar n per draw (NULL = 0)
 * for the detrending of the observables, sigma_m ny x ny per draw / shared / NULL (= 0), status_in (NULL = all ok)
 * marks draws whose steady state already failed on the host (passed through with loglik = -Inf) */
int dyb_loglik_from_jacobian_batch(dyb_handle* h, const double* jac, const double* ybar, const double* sigma_e,
                                   int32_t sigma_e_per_draw, const double* sigma_m, int32_t sigma_m_per_draw,
                                   const int32_t* status_in, int64_t n_draws, double* loglik, int32_t* status);
/* batched doubling solve of X = A X A' + B (k x k per draw): the solver behind compute_variance!
 * (src/estimation/estimation.jl:630); tol = relative increment at which doubling stops, status 5 when it diverges */
int dyb_lyapunov_batch(int device, int32_t k, const double* A, const double* B, int64_t n_draws, double tol,
                       int32_t maxit, double* X, int32_t* iters, int32_t* status);

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
/* covariance of the particles -> what the mutation step uses (smc.jl:161-165): `prox!(RPSD, IndPSD(), R)` (projection on
 * the positive semi-definite cone: symmetric eigen-decomposition, negative eigenvalues set to zero), `Rdiag = diag(RPSD)`,
 * `Rchol = cholesky(RPSD).L`.  R np x np symmetric (lower triangle read); Rpsd, L np x np column-major.  The projection is
 * the identity on a positive-definite matrix and is only carried out when the factorisation of R itself fails:
 * *projected = 0 (R positive definite, Rpsd = R), 1 (projected, then factorised), 2 (still not positive definite after
 * the projection: the reference's `cholesky` throws PosDefException; L is not usable).  The device routine is the one the
 * SMC stage runs (csrc/smc_stage.cuh). */
int dyb_smc_covariance_factor(int device, const double* R, int32_t np, double* Rpsd, double* L, int32_t* projected);

/* ---- samplers around the batched likelihood ---------------------------------------------------------------
 * Priors: the `prior` vector of EstimatedParameters (src/dynare_containers.jl:800-841).  shape[k] uses the
 * reference's mod-file codes (src/estimation/priors.jl:482-529): 1 Beta(a,b), 2 Gamma(shape a, scale b),
 * 3 Normal(mean a, std b), 4 InverseGamma1(alpha a, theta b) (src/distributions/inversegamma1.jl:134-141),
 * 5 Uniform(a,b), 6 InverseGamma(alpha a, theta b), 8 Weibull(shape a, scale b); (a, b) are the distribution's
 * own hyper-parameters, i.e. what src/distributions/distribution_parameters.jl:4-76 returns. */
int dyb_set_priors(dyb_handle* h, int32_t np, const int32_t* shape, const double* a, const double* b);
/* logpriordensity (src/estimation/estimation.jl:455-462) for N draws; host or device pointers */
int dyb_logprior_batch(dyb_handle* h, const double* theta, int64_t n_draws, double* logprior);
/* DSGELogPosteriorDensity callable (src/estimation/estimation.jl:503-515): logprior + loglik, -Inf for a
 * failed draw; loglik / status may be NULL */
int dyb_logposterior_batch(dyb_handle* h, const double* theta, int64_t n_draws, double* logpost,
                           double* loglik, int32_t* status);
/* DSGETransformation (src/estimation/estimation.jl:465-487): inverse = 0 maps R^np -> parameters and returns the
 * log-Jacobian (TransformedLogDensity); inverse = 1 maps parameters -> R^np (logjac unused) */
int dyb_transform_batch(dyb_handle* h, const double* in, int64_t n_draws, int32_t inverse, double* out,
                        double* logjac);
/* the samplers' counter-based generator (Philox4x32-10; counter = (index, step, stream, block), key = seed):
 * kind 0: out[count x n] uniforms on [0,1) of `stream` (1 mixture component, 2 accept, 0 resampling);
 * kind 1: the first `count` standard normals of (index, step).  For tests and for reproducing a run. */
int dyb_rng_fill(int device, uint64_t seed, int64_t first_index, int64_t n, uint32_t step, uint32_t stream,
                 int32_t kind, int32_t count, double* out);

/* SMC with the particles resident on the device(s): `smc(pdraw!, logprior, loglikelihood, names, np)`
 * (src/estimation/smc.jl:64-313).  Fields default to `Tune` (smc.jl:13-62). */
typedef struct dyb_smc dyb_smc;
typedef struct dyb_smc_options {
  int64_t npart;      /* particles in total, over all ranks (Tune.npart = 1000) */
  int32_t nphi;       /* tempering stages (400) */
  double lam;         /* bending coefficient of the schedule phi_i = ((i-1)/(nphi-1))^lam (2) */
  double c0;          /* initial scale of the proposal covariance (0.5) */
  double acpt0;       /* initial acceptance rate (0.25) */
  double trgt;        /* target acceptance rate (0.25) */
  double alp;         /* mixture weight of the random-walk proposal (0.9) */
  int32_t resampling; /* 0 multinomial (smc.jl:135), 1 systematic (the commented call smc.jl:133) */
  uint64_t seed;
} dyb_smc_options;
int dyb_smc_default_options(dyb_smc_options* opt);
/* phi: nphi user-supplied tempering values or NULL for the schedule above.  With a communicator attached
 * (dyb_comm_init) rank r owns particles [r*npart/nranks, (r+1)*npart/nranks). */
int dyb_smc_create(dyb_handle* h, const dyb_smc_options* opt, const double* phi, dyb_smc** out);
int dyb_smc_destroy(dyb_smc* s);
int dyb_smc_local_range(const dyb_smc* s, int64_t* first, int64_t* n_local);
/* how the stages run: *peer_stores = 1 when the mutated particles travel by stores into the peers' CUDA-IPC mapped
 * buffers over NVLink (the default with several ranks on one node), 0 for the in-place ncclAllGather (environment
 * DYB_SMC_TRANSPORT=nccl, or no peer access) or a single GPU; *cuda_graph = 1 when stages replay a captured CUDA graph
 * (DYB_SMC_GRAPH=0 and profiling switch it off).  Either output may be NULL. */
int dyb_smc_transport(const dyb_smc* s, int32_t* peer_stores, int32_t* cuda_graph);
/* initial particles (smc.jl:84-101): para_local np x n_local drawn from the prior by the caller; evaluates the
 * likelihood and the prior; status[i] != 0 marks draws the reference rejects (replace them and call again).  Local to
 * the rank: ranks may call it a different number of times before the (collective) dyb_smc_run. */
int dyb_smc_init(dyb_smc* s, const double* para_local, int32_t* status);
/* the same with the prior draws made on the device: `prior_draw!` (src/estimation/estimation.jl:329-335, pdraw[i] =
 * rand(prior_i)) inside the rejection loop of smc.jl:87-101 -- every local particle is drawn, the ones whose likelihood is
 * not finite are re-drawn, until none is left (at most max_rounds rounds, else DYB_ERR_STATE).  Philox stream keyed by
 * (seed, global particle index, round): the initial cloud does not depend on the number of GPUs.  *n_draws (may be NULL):
 * prior draws used by this rank.  Gamma variates by Marsaglia-Tsang, Beta as a ratio of Gammas, InverseGamma(1) as
 * (square roots of) reciprocal Gammas, Weibull / Uniform by inversion, Normal by Box-Muller. */
int dyb_smc_init_from_prior(dyb_smc* s, uint64_t seed, int32_t max_rounds, int64_t* n_draws);
/* prior draws on their own: para np x n_draws column-major (host or device), draw i = particle first + i of round `round`
 * of the stream above; needs dyb_set_priors */
int dyb_prior_draw(dyb_handle* h, int64_t n_draws, int64_t first, uint64_t seed, int32_t round, double* para);
/* the next n_stages tempering stages: correction, selection, mutation (smc.jl:110-243); asynchronous.  With several
 * ranks this is a collective call: every rank runs the same number of stages.  A stage is seven fused kernels around the
 * likelihood batch (csrc/smc_stage.cuh), replayed as a CUDA graph; its one exchange -- the mutated particles -- is
 * written into every peer's memory by the accept kernel itself. */
int dyb_smc_run(dyb_smc* s, int32_t n_stages);
int dyb_smc_stages_done(const dyb_smc* s);
/* per-stage time split from CUDA events on the handle's stream: when profiling is on, every stage records 7 events;
 * ms[6] = totals of (correction + weight all-gather, selection + particle all-gather, moments + Cholesky, proposal,
 * likelihood of the proposals, prior + accept + acceptance all-gather) over the profiled stages.  Synchronises. */
int dyb_smc_set_profiling(dyb_smc* s, int32_t on);
int dyb_smc_get_timing(dyb_smc* s, int32_t* n_stages, double* ms);
/* local slices of the particle system (host or device destinations, any may be NULL); synchronises */
int dyb_smc_get_particles(dyb_smc* s, double* para, double* w, double* loglh, double* logpost);
/* stats nphi x 6 row-major (zhat, ESS, resampled, c, acpt, clamped indices so far); mu np, R np x np of the last
 * mutation step; logmdd = sum_i log zhat_i (smc.jl:301).  Any output may be NULL; synchronises. */
int dyb_smc_get_stats(dyb_smc* s, double* stats, double* mu, double* R, double* logmdd);

/* random-walk Metropolis on n_chains chains at once (run_mcmc, src/estimation/estimation.jl:964-989): proposal
 * y' = y + chol_cov z; transformed != 0: y lives in R^np and the target is the TransformedLogDensity
 * (estimation.jl:926-929).  y np x n_chains in/out, logp / accepted n_chains out, history np x n_chains x n_iter and
 * history_logp n_chains x n_iter optional.  first_chain offsets the chain index used by the generator (sharding
 * over ranks); first_step lets a run be continued with fresh random numbers. */
int dyb_rwmh_run(dyb_handle* h, int64_t n_chains, int64_t first_chain, int32_t n_iter, uint32_t first_step,
                 int32_t transformed, const double* chol_cov, uint64_t seed, double* y, double* logp,
                 int64_t* accepted, double* history, double* history_logp);

/* ---- several GPUs: one process per GPU.  The NCCL communicator bootstraps the SMC stage coupling (exchange of CUDA-IPC
 * handles, barriers) and carries the particle all-gather when peer stores are unavailable.
 * Rank 0 calls dyb_comm_unique_id and ships the 128 bytes to the other ranks by any host channel. */
#define DYB_COMM_ID_BYTES 128
int dyb_comm_unique_id(void* id /* DYB_COMM_ID_BYTES out */);
int dyb_comm_init(dyb_handle* h, int32_t rank, int32_t nranks, const void* id);
int dyb_comm_destroy(dyb_handle* h);
int dyb_comm_info(const dyb_handle* h, int32_t* rank, int32_t* nranks);
/* in-place sum over ranks of a device buffer of doubles (is_int64 = 0) or int64 (1) on the handle's stream */
int dyb_comm_allreduce_sum(dyb_handle* h, void* d_buf, int64_t count, int32_t is_int64);

/* ---- pinned host memory ---------------------------------------------------------------------------- */
int dyb_host_alloc(void** ptr, int64_t bytes);
int dyb_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* DYNARE_B200_H */
