// Internal definitions shared by the translation units of libdynare_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "registry.h"

namespace dyb {

std::string& last_error_ref();
inline int fail(int code, const std::string& msg) { last_error_ref() = msg; return code; }

#define DYB_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail(DYB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));        \
  } while (0)

inline const ModelOps* find_model(const char* name) {
  if (!name) return nullptr;
  for (auto* m : registry()) if (std::strcmp(m->name, name) == 0) return m;
  return nullptr;
}

inline bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
inline bool is_pinned_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// growable device / pinned buffers
struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return static_cast<T*>(p); }
};
struct PinBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

// Per-call staging of arguments that may live on the host or on the device.
struct Stager {
  cudaStream_t st;
  struct Back { void* host; void* dev; size_t bytes; };
  std::vector<void*> temps;
  std::vector<Back> backs;
  cudaError_t err = cudaSuccess;
  explicit Stager(cudaStream_t s) : st(s) {}
  template <class T> const T* in(const T* p, size_t count) {
    if (!p || is_device_ptr(p)) return p;
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, count * sizeof(T) + 8);
    if (e != cudaSuccess) { err = e; return nullptr; }
    temps.push_back(d);
    e = cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) err = e;
    return static_cast<const T*>(d);
  }
  template <class T> T* out(T* p, size_t count) {
    if (!p || is_device_ptr(p)) return p;
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, count * sizeof(T) + 8);
    if (e != cudaSuccess) { err = e; return nullptr; }
    temps.push_back(d);
    backs.push_back({p, d, count * sizeof(T)});
    return static_cast<T*>(d);
  }
  template <class T> T* temp(size_t count) {
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, count * sizeof(T) + 8);
    if (e != cudaSuccess) { err = e; return nullptr; }
    temps.push_back(d);
    return static_cast<T*>(d);
  }
  cudaError_t finish() {
    cudaError_t e = err;
    for (auto& b : backs) {
      cudaError_t e2 = cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = e2;
    }
    cudaError_t e3 = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = e3;
    for (void* t : temps) cudaFree(t);
    temps.clear(); backs.clear();
    return e;
  }
  ~Stager() { for (void* t : temps) cudaFree(t); }
};

cudaStream_t device_stream(int device);

}  // namespace dyb

struct dyb_handle;
namespace dyb { void comm_release(dyb_handle* h); }

// prior of one estimated parameter with its constants hoisted (sampler_kernels.cuh)
namespace dyb {
struct PriorDev {
  int shape;        // reference mod-file codes (src/estimation/priors.jl:482-529)
  int transform;    // 0 identity, 1 exp, 2 logistic, 3 scaled logistic on [a, b]
  double a, b;      // hyper-parameters
  double cst;       // normalising constant of the log-density
};
struct Comm;        // NCCL communicator wrapper (sampler.cu)
struct GenericHost; // run-time model description (generic.cu)
void generic_release(dyb_handle* h);
// dyb_options.univariate_fallback for the model-specific path: re-evaluates flagged draws from the hand-off buffer (generic.cu)
// generated models beyond the register-resident solver (ModelOps::register_solver == false): generic.cu
int hybrid_loglik(dyb_handle* h, const double* d_theta, long long n, double* d_loglik, int* d_status);
// sizes-only models (ModelOps::solve_from_jacobian): the register-resident kernels on host-evaluated Jacobians (api.cu)
int jacfed_loglik(dyb_handle* h, const double* jac, const double* ybar, const double* sigma_e, int sigma_e_per_draw,
                  const double* sigma_m, int sigma_m_per_draw, const int* status_in, long long n, double* loglik, int* status);
int jacfed_first_order(dyb_handle* h, const double* jac, const double* sigma_e, int sigma_e_per_draw, long long n, double* g1,
                       double* T, double* RQR, double* P0, int* cr_iters, int* lyap_iters, int* status);
int hybrid_first_order(dyb_handle* h, const double* d_theta, long long n, double* d_g1, double* d_ybar, int* d_cr_iters,
                       int* d_lyap_iters, int* d_status);
int univariate_fallback_from_handoff(dyb_handle* h, long long n, const int* d_status_solve, double* d_loglik, int* d_status);
}  // namespace dyb

struct dyb_handle {
  const dyb::ModelOps* ops = nullptr;     // null for handles made by dyb_create_generic
  dyb_model_dims dims{};
  dyb::GenericHost* generic = nullptr;    // generic.cu
  int device = 0;
  cudaStream_t stream = nullptr;        // stream in use
  cudaStream_t own_stream = nullptr;    // the handle's private stream
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool ev_valid = false;
  std::vector<cudaEvent_t> timer_ev;    // triples (start, after solve, after filter)
  int timer_n = 0;
  bool timers_armed = false;
  dyb::HostState hs;
  dyb::DevBuf d_theta, d_ss, d_st1, d_ll, d_st2, d_Y, d_mid;
  dyb::DevBuf d_uT, d_uQ, d_uP, d_uyb, d_uH, d_uz, d_uscr;     // dense state spaces of the univariate fallback (generic.cu)
  dyb::DevBuf d_jfJ, d_jfyb, d_jfse, d_jfsm, d_jfsi;           // staged inputs of a sizes-only model (jacfed_*, api.cu)
  dyb::PinBuf h_in, h_out;
  int ny_obs = 0, nobs = 0;
  int64_t launches = 0;
  // samplers (sampler.cu)
  std::vector<dyb::PriorDev> priors;    // one per estimated parameter once dyb_set_priors has been called
  dyb::Comm* comm = nullptr;
  int n_smc = 0;                        // dyb_smc objects attached (dyb_destroy refuses while > 0)
  long long config_epoch = 0;           // bumped by every setter: invalidates captured CUDA graphs
};
