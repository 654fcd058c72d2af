// C ABI of libdynare_b200.so (declared in include/dynare_b200.h).
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "internal.h"

namespace dyb {

static std::vector<const ModelOps*>& reg_mut() {
  static std::vector<const ModelOps*> r;
  return r;
}
const std::vector<const ModelOps*>& registry() { return reg_mut(); }
void register_model(const ModelOps* m) { reg_mut().push_back(m); }

static thread_local std::string g_err;
std::string& last_error_ref() { return g_err; }

cudaStream_t device_stream(int device) {
  static std::mutex mu;
  static std::vector<cudaStream_t> streams;
  std::lock_guard<std::mutex> lk(mu);
  if ((int)streams.size() <= device) streams.resize(device + 1, nullptr);
  if (!streams[device]) cudaStreamCreateWithFlags(&streams[device], cudaStreamNonBlocking);
  return streams[device];
}

}  // namespace dyb

using namespace dyb;

static void default_options(dyb_options* o) {
  o->cr_tol = 1e-8;
  o->cr_maxit = 100;
  o->steady_tol = std::cbrt(2.220446049250313e-16);
  o->lyap_tol = 1e-16;
  o->lyap_maxit = 60;
  o->presample = 0;
  o->solver = 0;
  o->filter_lanes = 0;
  o->univariate_fallback = 0;
}

namespace dyb {
// fp64 tensor-core microbenchmark: SHAPE 0 = mma.m8n8k4 (256 FMA / warp instruction), 1 = mma.m16n8k16 (2048 FMA);
// 4 independent accumulator tiles per warp, operands constant in registers
template <int SHAPE>
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double seed) {
  const double a0 = seed + 1e-3 * threadIdx.x, b0 = 1.0 - 1e-3 * threadIdx.x;
  if (SHAPE == 0) {
    double c[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) { c[t][0] = t; c[t][1] = -t; }
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a0), "d"(b0));
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 4; ++t) s += c[t][0] + c[t][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {
    double c[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t) { c[t][0] = t; c[t][1] = -t; c[t][2] = 2 * t; c[t][3] = 1; }
    const double a[8] = {a0, a0 + 1, a0 + 2, a0 + 3, a0 + 4, a0 + 5, a0 + 6, a0 + 7};
    const double b[4] = {b0, b0 + 1, b0 + 2, b0 + 3};
#pragma unroll 2
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                       "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 4; ++t) s += c[t][0] + c[t][1] + c[t][2] + c[t][3];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}
}  // namespace dyb

extern "C" {

int dyb_version(void) { return DYB_VERSION; }
const char* dyb_last_error(void) { return g_err.c_str(); }

int dyb_device_count(int* count) {
  if (!count) return fail(DYB_ERR_ARG, "count is null");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { cudaGetLastError(); *count = 0; return fail(DYB_ERR_CUDA, cudaGetErrorString(e)); }
  *count = c;
  return DYB_OK;
}

int dyb_model_count(void) { return (int)registry().size(); }
const char* dyb_model_name(int index) {
  if (index < 0 || index >= (int)registry().size()) return nullptr;
  return registry()[index]->name;
}
int dyb_model_get_dims(const char* model, dyb_model_dims* dims) {
  const ModelOps* m = find_model(model);
  if (!m || !dims) return fail(DYB_ERR_ARG, "unknown model or null dims");
  *dims = m->dims;
  return DYB_OK;
}

int dyb_create(const char* model, int device, dyb_handle** out) {
  if (!out) return fail(DYB_ERR_ARG, "out is null");
  *out = nullptr;
  const ModelOps* m = find_model(model);
  if (!m) return fail(DYB_ERR_ARG, std::string("unknown model: ") + (model ? model : "(null)"));
  int ndev = 0;
  DYB_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(DYB_ERR_ARG, "no such CUDA device");
  DYB_CUDA(cudaSetDevice(device));
  dyb_handle* h = new dyb_handle();
  h->ops = m;
  h->dims = m->dims;
  h->device = device;
  m->defaults(h->hs);
  default_options(&h->hs.opt);
  cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  h->stream = h->own_stream;
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&h->ev[i]);
  if (e != cudaSuccess) { delete h; return fail(DYB_ERR_CUDA, cudaGetErrorString(e)); }
  *out = h;
  return DYB_OK;
}

int dyb_destroy(dyb_handle* h) {
  if (!h) return DYB_OK;
  if (h->n_smc > 0) return fail(DYB_ERR_STATE, "dyb_destroy: a dyb_smc object is still attached to this handle (dyb_smc_destroy first)");
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  comm_release(h);
  generic_release(h);
  h->hs.crp.release();
  h->d_jfJ.release(); h->d_jfyb.release(); h->d_jfse.release(); h->d_jfsm.release(); h->d_jfsi.release();
  h->d_uT.release(); h->d_uQ.release(); h->d_uP.release(); h->d_uyb.release(); h->d_uH.release(); h->d_uz.release(); h->d_uscr.release();
  h->d_theta.release(); h->d_ss.release(); h->d_st1.release(); h->d_ll.release(); h->d_st2.release(); h->d_Y.release(); h->d_mid.release();
  h->h_in.release(); h->h_out.release();
  for (int i = 0; i < 4; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (cudaEvent_t e : h->timer_ev) cudaEventDestroy(e);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return DYB_OK;
}

int dyb_get_dims(const dyb_handle* h, dyb_model_dims* dims) {
  if (!h || !dims) return fail(DYB_ERR_ARG, "null argument");
  *dims = h->dims;
  return DYB_OK;
}
int dyb_get_options(const dyb_handle* h, dyb_options* opt) {
  if (!h || !opt) return fail(DYB_ERR_ARG, "null argument");
  *opt = h->hs.opt;
  return DYB_OK;
}
int dyb_set_options(dyb_handle* h, const dyb_options* opt) {
  if (!h || !opt) return fail(DYB_ERR_ARG, "null argument");
  if (!(opt->cr_tol > 0) || opt->cr_maxit < 1 || !(opt->steady_tol > 0) || !(opt->lyap_tol >= 0) ||
      opt->lyap_maxit < 1 || opt->presample < 0 || opt->solver < 0 || opt->solver > 2 ||
      !(opt->filter_lanes == 0 || opt->filter_lanes == 1 || opt->filter_lanes == 4 || opt->filter_lanes == 8 ||
        opt->filter_lanes == 16 || opt->filter_lanes == 32) ||
      opt->univariate_fallback < 0 || opt->univariate_fallback > 2)
    return fail(DYB_ERR_ARG, "invalid option value");
  h->hs.opt = *opt;
  h->config_epoch += 1;
  return DYB_OK;
}

int dyb_set_calibration(dyb_handle* h, const double* params, const double* sigma_e, const double* sigma_m) {
  if (h && !h->ops) return fail(DYB_ERR_UNSUPPORTED, "this handle has no device model function (dyb_create_generic): use the *_from_jacobian entry points");
  if (!h) return fail(DYB_ERR_ARG, "null handle");
  const dyb_model_dims& d = h->dims;
  if (params) h->hs.params.assign(params, params + d.npar);
  if (sigma_e) h->hs.sigma_e.assign(sigma_e, sigma_e + d.nx * d.nx);
  if (sigma_m) h->hs.sigma_m.assign(sigma_m, sigma_m + d.ny * d.ny);
  h->config_epoch += 1;
  return DYB_OK;
}

int dyb_set_estimated_params(dyb_handle* h, int32_t np, const int32_t* kind, const int32_t* i1, const int32_t* i2) {
  if (h && !h->ops) return fail(DYB_ERR_UNSUPPORTED, "this handle has no device model function (dyb_create_generic): use the *_from_jacobian entry points");
  if (!h || np < 0 || np > MAX_EST || (np > 0 && (!kind || !i1 || !i2))) return fail(DYB_ERR_ARG, "bad estimated-parameter map");
  const dyb_model_dims& d = h->dims;
  for (int q = 0; q < np; ++q) {
    int lim1, lim2;
    switch (kind[q]) {
      case DYB_EST_PARAM: lim1 = d.npar; lim2 = 1 << 30; break;
      case DYB_EST_SD_SHOCK: case DYB_EST_VAR_SHOCK: case DYB_EST_CORR_SHOCK: lim1 = lim2 = d.nx; break;
      case DYB_EST_SD_MEAS: case DYB_EST_VAR_MEAS: case DYB_EST_CORR_MEAS: lim1 = lim2 = d.ny; break;
      default: return fail(DYB_ERR_ARG, "unknown estimated-parameter kind");
    }
    if (i1[q] < 0 || i1[q] >= lim1 || i2[q] < 0 || i2[q] >= lim2) return fail(DYB_ERR_ARG, "estimated-parameter index out of range");
  }
  h->hs.np = np;
  for (int q = 0; q < np; ++q) { h->hs.est_type[q] = kind[q]; h->hs.est_i[q] = i1[q]; h->hs.est_j[q] = i2[q]; }
  h->config_epoch += 1;
  return DYB_OK;
}

int dyb_set_observations(dyb_handle* h, const double* Y, int32_t ny, int32_t nobs) {
  if (!h || !Y) return fail(DYB_ERR_ARG, "null argument");
  if (ny != h->dims.ny) return fail(DYB_ERR_ARG, "observations must have ny rows");
  if (nobs < 0) return fail(DYB_ERR_ARG, "nobs < 0");
  if ((size_t)ny * nobs * sizeof(double) > 200 * 1024) return fail(DYB_ERR_UNSUPPORTED, "observation matrix larger than 200 KB");
  DYB_CUDA(cudaSetDevice(h->device));
  DYB_CUDA(h->d_Y.reserve(sizeof(double) * (size_t)ny * (nobs > 0 ? nobs : 1)));
  if (nobs > 0) DYB_CUDA(cudaMemcpyAsync(h->d_Y.p, Y, sizeof(double) * (size_t)ny * nobs,
                                         is_device_ptr(Y) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  h->ny_obs = ny; h->nobs = nobs;
  h->config_epoch += 1;
  return DYB_OK;
}

static cudaEvent_t* timer_slot(dyb_handle* h);
static int reserve_batch(dyb_handle* h, int64_t n) {
  const dyb_model_dims& d = h->dims;
  const size_t N = (size_t)(n > 0 ? n : 1);
  const void* before[3] = {h->d_ss.p, h->d_st1.p, h->d_mid.p};
  DYB_CUDA(h->d_ss.reserve(sizeof(double) * d.ss_doubles * N));
  DYB_CUDA(h->d_st1.reserve(sizeof(int) * N));
  DYB_CUDA(h->d_mid.reserve(sizeof(double) * (size_t)h->ops->mid_doubles * N));
  if (before[0] != h->d_ss.p || before[1] != h->d_st1.p || before[2] != h->d_mid.p) h->config_epoch += 1;   // captured graphs hold the old addresses
  return DYB_OK;
}

int dyb_loglik_batch_device(dyb_handle* h, const double* d_theta, int64_t n, double* d_loglik, int32_t* d_status) {
  if (h && !h->ops) return fail(DYB_ERR_UNSUPPORTED, "this handle has no device model function (dyb_create_generic): use the *_from_jacobian entry points");
  if (!h || n < 0 || (n > 0 && (!d_theta || !d_loglik || !d_status))) return fail(DYB_ERR_ARG, "null argument");
  if (h->nobs <= 0 && h->d_Y.p == nullptr) return fail(DYB_ERR_STATE, "dyb_set_observations has not been called");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  if (!h->ops->register_solver) {
    // a model beyond the register-resident kernels: generated model function -> run-time-size solver -> dense filter
    h->ev_valid = false;
    return hybrid_loglik(h, d_theta, n, d_loglik, d_status);
  }
  int rc = reserve_batch(h, n);
  if (rc) return rc;
  SolveOut so{h->d_ss.as<double>(), h->d_st1.as<int>(), nullptr, nullptr, nullptr, nullptr, h->d_mid.as<double>()};
  // inside a stream capture (the SMC stage graph) the timing events are left out: a captured record is no timestamp
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  DYB_CUDA(cudaStreamIsCapturing(h->stream, &cap));
  const bool timed = cap == cudaStreamCaptureStatusNone;
  cudaEvent_t* ev = (timed && h->timers_armed) ? timer_slot(h) : nullptr;
  if (ev) { h->timer_n += 1; so.ev_after_model = ev[3]; so.ev_after_cr = ev[4]; } else ev = h->ev;
  if (timed) DYB_CUDA(cudaEventRecord(ev[0], h->stream));
  DYB_CUDA(h->ops->solve(h->hs, d_theta, n, so, h->stream));
  if (timed) DYB_CUDA(cudaEventRecord(ev[1], h->stream));
  DYB_CUDA(h->ops->kalman(h->d_ss.as<double>(), n, h->d_Y.as<double>(), h->nobs, h->hs.opt.presample, h->hs.opt.filter_lanes,
                          h->d_st1.as<int>(), d_loglik, d_status, h->stream));
  if (timed) DYB_CUDA(cudaEventRecord(ev[2], h->stream));
  if (h->hs.opt.univariate_fallback) {
    rc = univariate_fallback_from_handoff(h, n, h->d_st1.as<int>(), d_loglik, d_status);
    if (rc) return rc;
  }
  h->ev_valid = timed && (ev == h->ev);
  h->launches += 1 + h->ops->solve_launches(h->hs);
  return DYB_OK;
}

// timed variant used when the accumulating timers are armed: records into the timer ring instead
static cudaEvent_t* timer_slot(dyb_handle* h) {
  if (h->timer_n >= 4096) return nullptr;
  while ((int)h->timer_ev.size() < 5 * (h->timer_n + 1)) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    h->timer_ev.push_back(e);
  }
  return &h->timer_ev[5 * h->timer_n];
}

static int loglik_batch_impl(dyb_handle* h, const double* theta, int64_t n, double* loglik, int32_t* status, bool async) {
  if (h && !h->ops) return fail(DYB_ERR_UNSUPPORTED, "this handle has no device model function (dyb_create_generic): use the *_from_jacobian entry points");
  if (!h || n < 0 || (n > 0 && (!theta || !loglik || !status))) return fail(DYB_ERR_ARG, "null argument");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  const int np = h->hs.np;
  const size_t N = (size_t)n;
  const bool dev_in = is_device_ptr(theta);
  const bool pin_in = dev_in || is_pinned_ptr(theta);
  const bool pin_ll = is_pinned_ptr(loglik) || is_device_ptr(loglik);
  const bool pin_st = is_pinned_ptr(status) || is_device_ptr(status);
  if (async && !(pin_in && pin_ll && pin_st))
    return fail(DYB_ERR_ARG, "the asynchronous call needs pinned (dyb_host_alloc) or device buffers");
  DYB_CUDA(h->d_theta.reserve(sizeof(double) * np * N + 8));
  DYB_CUDA(h->d_ll.reserve(sizeof(double) * N));
  DYB_CUDA(h->d_st2.reserve(sizeof(int) * N));
  const void* src = theta;
  if (!pin_in) {
    DYB_CUDA(h->h_in.reserve(sizeof(double) * np * N + 8));
    std::memcpy(h->h_in.p, theta, sizeof(double) * np * N);
    src = h->h_in.p;
  }
  DYB_CUDA(cudaMemcpyAsync(h->d_theta.p, src, sizeof(double) * np * N,
                           dev_in ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  int rc = dyb_loglik_batch_device(h, h->d_theta.as<double>(), n, h->d_ll.as<double>(), h->d_st2.as<int>());
  if (rc) return rc;
  DYB_CUDA(h->h_out.reserve((sizeof(double) + sizeof(int)) * N + 16));
  double* ll_stage = h->h_out.as<double>();
  int* st_stage = reinterpret_cast<int*>(ll_stage + N);
  DYB_CUDA(cudaMemcpyAsync(pin_ll ? (void*)loglik : (void*)ll_stage, h->d_ll.p, sizeof(double) * N, cudaMemcpyDefault, h->stream));
  DYB_CUDA(cudaMemcpyAsync(pin_st ? (void*)status : (void*)st_stage, h->d_st2.p, sizeof(int) * N, cudaMemcpyDefault, h->stream));
  if (async) return DYB_OK;
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  if (!pin_ll) std::memcpy(loglik, ll_stage, sizeof(double) * N);
  if (!pin_st) std::memcpy(status, st_stage, sizeof(int) * N);
  return DYB_OK;
}

int dyb_loglik_batch(dyb_handle* h, const double* theta, int64_t n, double* loglik, int32_t* status) {
  return loglik_batch_impl(h, theta, n, loglik, status, false);
}

// same call without the final synchronisation: H2D, kernels and D2H are enqueued on the handle's stream and the call
// returns; results are valid after dyb_sync.  Buffers must be pinned (dyb_host_alloc) or device memory and stay
// untouched until then.  Two handles used alternately overlap one batch's copies with the other's kernels.
int dyb_loglik_batch_async(dyb_handle* h, const double* theta, int64_t n, double* loglik, int32_t* status) {
  return loglik_batch_impl(h, theta, n, loglik, status, true);
}

int dyb_sync(dyb_handle* h) {
  if (!h) return fail(DYB_ERR_ARG, "null handle");
  DYB_CUDA(cudaSetDevice(h->device));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  return DYB_OK;
}

// non-blocking completion query for the asynchronous entry points: *done = 1 when the handle's stream is idle
int dyb_query(dyb_handle* h, int32_t* done) {
  if (!h || !done) return fail(DYB_ERR_ARG, "null argument");
  DYB_CUDA(cudaSetDevice(h->device));
  const cudaError_t e = cudaStreamQuery(h->stream);
  if (e == cudaSuccess) { *done = 1; return DYB_OK; }
  if (e == cudaErrorNotReady) { cudaGetLastError(); *done = 0; return DYB_OK; }
  return fail(DYB_ERR_CUDA, cudaGetErrorString(e));
}

int64_t dyb_kernel_launches(const dyb_handle* h) { return h ? h->launches : 0; }

int dyb_last_kernel_ms(dyb_handle* h, float* solve_ms, float* filter_ms) {
  if (!h) return fail(DYB_ERR_ARG, "null handle");
  if (!h->ev_valid) return fail(DYB_ERR_STATE, "no likelihood batch has run");
  DYB_CUDA(cudaSetDevice(h->device));
  DYB_CUDA(cudaEventSynchronize(h->ev[2]));
  float a = 0, b = 0;
  DYB_CUDA(cudaEventElapsedTime(&a, h->ev[0], h->ev[1]));
  DYB_CUDA(cudaEventElapsedTime(&b, h->ev[1], h->ev[2]));
  if (solve_ms) *solve_ms = a;
  if (filter_ms) *filter_ms = b;
  return DYB_OK;
}

int dyb_set_stream(dyb_handle* h, void* cuda_stream) {
  if (!h) return fail(DYB_ERR_ARG, "null handle");
  DYB_CUDA(cudaSetDevice(h->device));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  h->config_epoch += 1;
  return DYB_OK;
}

int dyb_kernel_timers_reset(dyb_handle* h) {
  if (!h) return fail(DYB_ERR_ARG, "null handle");
  DYB_CUDA(cudaSetDevice(h->device));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  h->timer_n = 0;
  h->timers_armed = true;
  return DYB_OK;
}

int dyb_kernel_timers_read(dyb_handle* h, int32_t* n_batches, float* solve_ms_total, float* filter_ms_total) {
  if (!h) return fail(DYB_ERR_ARG, "null handle");
  DYB_CUDA(cudaSetDevice(h->device));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  float a = 0, b = 0;
  for (int i = 0; i < h->timer_n; ++i) {
    float x = 0, y = 0;
    DYB_CUDA(cudaEventElapsedTime(&x, h->timer_ev[5 * i], h->timer_ev[5 * i + 1]));
    DYB_CUDA(cudaEventElapsedTime(&y, h->timer_ev[5 * i + 1], h->timer_ev[5 * i + 2]));
    a += x; b += y;
  }
  if (n_batches) *n_batches = h->timer_n;
  if (solve_ms_total) *solve_ms_total = a;
  if (filter_ms_total) *filter_ms_total = b;
  h->timers_armed = false;
  return DYB_OK;
}

// split of the solve stage of the armed batches: model set-up (steady state, Jacobian, static elimination), cyclic
// reduction (+ g_u), assembly (static rows, Lyapunov, state space); all zero on the one-kernel solver path
int dyb_kernel_timers_read_solve_split(dyb_handle* h, float* model_ms_total, float* cr_ms_total, float* assemble_ms_total) {
  if (!h) return fail(DYB_ERR_ARG, "null handle");
  DYB_CUDA(cudaSetDevice(h->device));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  float m = 0, c = 0, s = 0;
  if (h->ops && h->ops->solve_launches(h->hs) >= 3) {
    for (int i = 0; i < h->timer_n; ++i) {
      float x = 0, y = 0, z = 0;
      DYB_CUDA(cudaEventElapsedTime(&x, h->timer_ev[5 * i], h->timer_ev[5 * i + 3]));
      DYB_CUDA(cudaEventElapsedTime(&y, h->timer_ev[5 * i + 3], h->timer_ev[5 * i + 4]));
      DYB_CUDA(cudaEventElapsedTime(&z, h->timer_ev[5 * i + 4], h->timer_ev[5 * i + 1]));
      m += x; c += y; s += z;
    }
  }
  if (model_ms_total) *model_ms_total = m;
  if (cr_ms_total) *cr_ms_total = c;
  if (assemble_ms_total) *assemble_ms_total = s;
  return DYB_OK;
}

}  // extern "C"

namespace dyb {
// host or device inputs of a sizes-only model -> device (persistent buffers; device pointers are used in place)
struct JfInputs { const double *jac, *ybar, *se, *sm; const int* si; };
static int jacfed_stage(dyb_handle* h, const double* jac, const double* ybar, const double* sigma_e, int se_pd,
                        const double* sigma_m, int sm_pd, const int* status_in, long long n, JfInputs* o) {
  const dyb_model_dims& d = h->dims;
  const size_t N = (size_t)n;
  cudaError_t e = cudaSuccess;
  auto stage = [&](DevBuf& b, const void* p, size_t bytes) -> const void* {
    if (!p || is_device_ptr(p)) return p;
    if (e == cudaSuccess) e = b.reserve(bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b.p, p, bytes, cudaMemcpyHostToDevice, h->stream);
    return b.p;
  };
  o->jac = static_cast<const double*>(stage(h->d_jfJ, jac, sizeof(double) * d.n * d.ncol * N));
  o->ybar = static_cast<const double*>(stage(h->d_jfyb, ybar, sizeof(double) * d.n * N));
  o->se = static_cast<const double*>(stage(h->d_jfse, sigma_e, sizeof(double) * d.nx * d.nx * (se_pd ? N : 1)));
  o->sm = static_cast<const double*>(stage(h->d_jfsm, sigma_m, sizeof(double) * d.ny * d.ny * (sm_pd ? N : 1)));
  o->si = static_cast<const int*>(stage(h->d_jfsi, status_in, sizeof(int) * N));
  if (e != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(e));
  return DYB_OK;
}

int jacfed_loglik(dyb_handle* h, const double* jac, const double* ybar, const double* sigma_e, int se_pd,
                  const double* sigma_m, int sm_pd, const int* status_in, long long n, double* loglik, int* status) {
  if (n < 0 || (n > 0 && (!jac || !sigma_e || !loglik || !status))) return fail(DYB_ERR_ARG, "null argument");
  if (h->d_Y.p == nullptr) return fail(DYB_ERR_STATE, "dyb_set_observations has not been called");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  int rc = reserve_batch(h, n);
  if (rc) return rc;
  JfInputs in{};
  rc = jacfed_stage(h, jac, ybar, sigma_e, se_pd, sigma_m, sm_pd, status_in, n, &in);
  if (rc) return rc;
  Stager sg(h->stream);
  double* dll = sg.out(loglik, (size_t)n);
  int* dst = sg.out(status, (size_t)n);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  SolveOut so{h->d_ss.as<double>(), h->d_st1.as<int>(), nullptr, nullptr, nullptr, nullptr, h->d_mid.as<double>()};
  cudaEvent_t* ev = h->timers_armed ? timer_slot(h) : nullptr;
  if (ev) { h->timer_n += 1; so.ev_after_model = ev[3]; so.ev_after_cr = ev[4]; } else ev = h->ev;
  DYB_CUDA(cudaEventRecord(ev[0], h->stream));
  DYB_CUDA(h->ops->solve_from_jacobian(h->hs, in.jac, in.ybar, in.se, se_pd, in.sm, sm_pd, in.si, n, so, h->stream));
  DYB_CUDA(cudaEventRecord(ev[1], h->stream));
  DYB_CUDA(h->ops->kalman(h->d_ss.as<double>(), n, h->d_Y.as<double>(), h->nobs, h->hs.opt.presample, h->hs.opt.filter_lanes,
                          h->d_st1.as<int>(), dll, dst, h->stream));
  DYB_CUDA(cudaEventRecord(ev[2], h->stream));
  h->ev_valid = (ev == h->ev);
  h->launches += 4;
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

int jacfed_first_order(dyb_handle* h, const double* jac, const double* sigma_e, int se_pd, long long n, double* g1,
                       double* T, double* RQR, double* P0, int* cr_iters, int* lyap_iters, int* status) {
  if (n < 0 || (n > 0 && (!jac || !sigma_e || !status))) return fail(DYB_ERR_ARG, "null argument");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  const dyb_model_dims& d = h->dims;
  int rc = reserve_batch(h, n);
  if (rc) return rc;
  JfInputs in{};
  rc = jacfed_stage(h, jac, nullptr, sigma_e, se_pd, nullptr, 0, nullptr, n, &in);
  if (rc) return rc;
  Stager sg(h->stream);
  const size_t N = (size_t)n, ns2 = (size_t)d.ns * d.ns;
  SolveOut so;
  so.ss = h->d_ss.as<double>();
  so.status = sg.out(status, N);
  so.g1 = sg.out(g1, (size_t)d.n * (d.k + d.nx) * N);
  so.ybar = nullptr;
  so.cr_iters = sg.out(cr_iters, N);
  so.lyap_iters = sg.out(lyap_iters, N);
  so.mid = h->d_mid.as<double>();
  double* dT = sg.out(T, ns2 * N); double* dQ = sg.out(RQR, ns2 * N); double* dP = sg.out(P0, ns2 * N);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  if (so.g1) DYB_CUDA(cudaMemsetAsync(so.g1, 0, sizeof(double) * (size_t)d.n * (d.k + d.nx) * N, h->stream));
  if (so.cr_iters) DYB_CUDA(cudaMemsetAsync(so.cr_iters, 0, sizeof(int) * N, h->stream));
  if (so.lyap_iters) DYB_CUDA(cudaMemsetAsync(so.lyap_iters, 0, sizeof(int) * N, h->stream));
  if (dT) DYB_CUDA(cudaMemsetAsync(dT, 0, sizeof(double) * ns2 * N, h->stream));
  if (dQ) DYB_CUDA(cudaMemsetAsync(dQ, 0, sizeof(double) * ns2 * N, h->stream));
  if (dP) DYB_CUDA(cudaMemsetAsync(dP, 0, sizeof(double) * ns2 * N, h->stream));
  DYB_CUDA(h->ops->solve_from_jacobian(h->hs, in.jac, nullptr, in.se, se_pd, nullptr, 0, nullptr, n, so, h->stream));
  if (dT || dQ || dP) DYB_CUDA(h->ops->unpack(h->d_ss.as<double>(), n, dT, dQ, dP, nullptr, nullptr, h->stream));
  h->launches += 4;
  DYB_CUDA(sg.finish());
  return DYB_OK;
}
}  // namespace dyb

extern "C" {
int dyb_solver_redo_count(dyb_handle* h, int64_t* n_redone) {
  if (!h || !n_redone) return fail(DYB_ERR_ARG, "null argument");
  DYB_CUDA(cudaSetDevice(h->device));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  unsigned long long v = 0;
  if (h->hs.crp.d_redo) DYB_CUDA(cudaMemcpy(&v, h->hs.crp.d_redo, sizeof(v), cudaMemcpyDeviceToHost));
  *n_redone = (int64_t)v;
  return DYB_OK;
}

namespace dyb {
// 8 independent DFMA chains per thread, 4096 iterations: register-resident, saturates the fp64 pipe
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
}  // namespace dyb

// measured fp64 TENSOR-core throughput (TFLOP/s): shape 0 = m8n8k4, 1 = m16n8k16
int dyb_fp64_peak_dmma(int device, int32_t shape, double* tflops) {
  if (!tflops || shape < 0 || shape > 1) return fail(DYB_ERR_ARG, "bad argument");
  DYB_CUDA(cudaSetDevice(device));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  DYB_CUDA(cudaMalloc(&out, sizeof(double) * blocks * threads));
  cudaStream_t st = device_stream(device);
  cudaEvent_t e0, e1;
  DYB_CUDA(cudaEventCreate(&e0)); DYB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    DYB_CUDA(cudaEventRecord(e0, st));
    if (shape == 0) dmma_peak_kernel<0><<<blocks, threads, 0, st>>>(out, iters, 1.0 + rep);
    else dmma_peak_kernel<1><<<blocks, threads, 0, st>>>(out, iters, 1.0 + rep);
    DYB_CUDA(cudaEventRecord(e1, st));
    DYB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    DYB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double fma_per_instr = shape == 0 ? 256.0 : 2048.0;
    const double fl = 2.0 * fma_per_instr * 4.0 * iters * (double)blocks * (threads / 32);
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  *tflops = best;
  return DYB_OK;
}

int dyb_fp64_peak(int device, double* tflops) {
  if (!tflops) return fail(DYB_ERR_ARG, "null argument");
  DYB_CUDA(cudaSetDevice(device));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int blocks = sms * 8, threads = 256, iters = 8192;
  double* out = nullptr;
  DYB_CUDA(cudaMalloc(&out, sizeof(double) * blocks * threads));
  cudaStream_t st = device_stream(device);
  cudaEvent_t e0, e1;
  DYB_CUDA(cudaEventCreate(&e0)); DYB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    DYB_CUDA(cudaEventRecord(e0, st));
    fp64_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 1.0 + rep);
    DYB_CUDA(cudaEventRecord(e1, st));
    DYB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    DYB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8.0 * iters * (double)blocks * threads;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  *tflops = best;
  return DYB_OK;
}

int dyb_first_order_batch(dyb_handle* h, const double* theta, int64_t n, double* g1, double* ybar,
                          int32_t* cr_iters, int32_t* lyap_iters, int32_t* status) {
  if (h && !h->ops) return fail(DYB_ERR_UNSUPPORTED, "this handle has no device model function (dyb_create_generic): use the *_from_jacobian entry points");
  if (!h || n < 0 || (n > 0 && (!theta || !status))) return fail(DYB_ERR_ARG, "null argument");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  const dyb_model_dims& d = h->dims;
  if (!h->ops->register_solver) {
    Stager sgh(h->stream);
    const size_t Nh = (size_t)n;
    const double* dth = sgh.in(theta, (size_t)h->hs.np * Nh);
    double* dg1 = sgh.out(g1, (size_t)d.n * (d.k + d.nx) * Nh);
    double* dyb = sgh.out(ybar, (size_t)d.n * Nh);
    int* dci = sgh.out(cr_iters, Nh); int* dli = sgh.out(lyap_iters, Nh); int* dst = sgh.out(status, Nh);
    if (sgh.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sgh.err));
    int rch = hybrid_first_order(h, dth, n, dg1, dyb, dci, dli, dst);
    if (rch) return rch;
    DYB_CUDA(sgh.finish());
    return DYB_OK;
  }
  int rc = reserve_batch(h, n);
  if (rc) return rc;
  Stager sg(h->stream);
  const size_t N = (size_t)n;
  const double* dth = sg.in(theta, (size_t)h->hs.np * N);
  SolveOut so;
  so.ss = h->d_ss.as<double>();
  so.status = sg.out(status, N);
  so.g1 = sg.out(g1, (size_t)d.n * (d.k + d.nx) * N);
  so.ybar = sg.out(ybar, (size_t)d.n * N);
  so.cr_iters = sg.out(cr_iters, N);
  so.lyap_iters = sg.out(lyap_iters, N);
  so.mid = h->d_mid.as<double>();
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  if (so.g1) DYB_CUDA(cudaMemsetAsync(so.g1, 0, sizeof(double) * (size_t)d.n * (d.k + d.nx) * N, h->stream));
  if (so.cr_iters) DYB_CUDA(cudaMemsetAsync(so.cr_iters, 0, sizeof(int) * N, h->stream));
  if (so.lyap_iters) DYB_CUDA(cudaMemsetAsync(so.lyap_iters, 0, sizeof(int) * N, h->stream));
  DYB_CUDA(h->ops->solve(h->hs, dth, n, so, h->stream));
  h->launches += h->ops->solve_launches(h->hs);
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

int dyb_state_space_batch(dyb_handle* h, const double* theta, int64_t n, double* T, double* RQR, double* P0,
                          double* ybar_obs, double* H, int32_t* status) {
  if (h && !h->ops) return fail(DYB_ERR_UNSUPPORTED, "this handle has no device model function (dyb_create_generic): use the *_from_jacobian entry points");
  if (!h || n < 0 || (n > 0 && (!theta || !status))) return fail(DYB_ERR_ARG, "null argument");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  const dyb_model_dims& d = h->dims;
  if (!h->ops->register_solver)
    return fail(DYB_ERR_UNSUPPORTED, "dyb_state_space_batch: not available for models beyond the register-resident kernels (use dyb_first_order_batch / dyb_loglik_batch)");
  int rc = reserve_batch(h, n);
  if (rc) return rc;
  Stager sg(h->stream);
  const size_t N = (size_t)n, ns2 = (size_t)d.ns * d.ns;
  const double* dth = sg.in(theta, (size_t)h->hs.np * N);
  SolveOut so{h->d_ss.as<double>(), sg.out(status, N), nullptr, nullptr, nullptr, nullptr, h->d_mid.as<double>()};
  double* dT = sg.out(T, ns2 * N);
  double* dQ = sg.out(RQR, ns2 * N);
  double* dP = sg.out(P0, ns2 * N);
  double* dy = sg.out(ybar_obs, (size_t)d.ny * N);
  double* dH = sg.out(H, (size_t)d.ny * d.ny * N);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  DYB_CUDA(cudaMemsetAsync(h->d_ss.p, 0, sizeof(double) * d.ss_doubles * N, h->stream));
  DYB_CUDA(h->ops->solve(h->hs, dth, n, so, h->stream));
  DYB_CUDA(h->ops->unpack(h->d_ss.as<double>(), n, dT, dQ, dP, dy, dH, h->stream));
  h->launches += 1 + h->ops->solve_launches(h->hs);
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

int dyb_host_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return fail(DYB_ERR_ARG, "bad argument");
  DYB_CUDA(cudaMallocHost(ptr, (size_t)bytes));
  return DYB_OK;
}
int dyb_host_free(void* ptr) {
  if (!ptr) return DYB_OK;
  DYB_CUDA(cudaFreeHost(ptr));
  return DYB_OK;
}

}  // extern "C"
