// Samplers around the batched likelihood (include/dynare_b200.h, "samplers" section): priors and the log-posterior,
// the SMC tempering recursion with its particles resident on the device(s), random-walk Metropolis on many chains,
// and the NCCL communicator that couples the particle shards of several GPUs.
#include <dlfcn.h>
#include <nccl.h>

#include "internal.h"
#include "sampler_kernels.cuh"

namespace dyb {

// ---- NCCL, bound at run time so that single-GPU users need no libnccl ------------------------------
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // reuse a copy already mapped into the process (e.g. the one PyTorch ships), else the system one
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    api.lib = h;
#define DYB_SYM(field, name)                                                     \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));           \
    if (!api.field) { api.why = std::string("libnccl lacks ") + name; api.lib = nullptr; return; }
    DYB_SYM(GetUniqueId, "ncclGetUniqueId")
    DYB_SYM(CommInitRank, "ncclCommInitRank")
    DYB_SYM(CommDestroy, "ncclCommDestroy")
    DYB_SYM(AllGather, "ncclAllGather")
    DYB_SYM(AllReduce, "ncclAllReduce")
    DYB_SYM(GroupStart, "ncclGroupStart")
    DYB_SYM(GroupEnd, "ncclGroupEnd")
    DYB_SYM(GetErrorString, "ncclGetErrorString")
#undef DYB_SYM
  });
  return api.lib ? &api : nullptr;
}

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
};

void comm_release(dyb_handle* h) {
  if (h && h->comm) {
    NcclApi* api = nccl_api();
    if (api && h->comm->comm) api->CommDestroy(h->comm->comm);
    delete h->comm;
    h->comm = nullptr;
  }
}

#define DYB_NCCL(expr)                                                                          \
  do {                                                                                          \
    ncclResult_t r_ = (expr);                                                                   \
    if (r_ != ncclSuccess)                                                                      \
      return fail(DYB_ERR_NCCL, std::string(#expr) + ": " + nccl_api()->GetErrorString(r_));    \
  } while (0)

static inline unsigned nblk(long long n, int b) { return (unsigned)((n + b - 1) / b); }

static int prior_set(const dyb_handle* h, PriorSetDev* ps) {
  if ((int)h->priors.size() != h->hs.np) return fail(DYB_ERR_STATE, "dyb_set_priors has not been called for the current estimated parameters");
  ps->np = h->hs.np;
  for (int k = 0; k < ps->np; ++k) {
    const PriorDev& p = h->priors[k];
    ps->p[k] = PriorDevK{p.shape, p.transform, p.a, p.b, p.cst};
  }
  return DYB_OK;
}

// all-gather `count` elements per rank; in place when send == recv + rank*count.  Single rank: copy if needed.
static int all_gather(dyb_handle* h, const void* send, void* recv, size_t count, ncclDataType_t type, size_t elsize) {
  if (!h->comm || h->comm->nranks == 1) {
    if (send != recv) DYB_CUDA(cudaMemcpyAsync(recv, send, count * elsize, cudaMemcpyDeviceToDevice, h->stream));
    return DYB_OK;
  }
  DYB_NCCL(nccl_api()->AllGather(send, recv, count, type, h->comm->comm, h->stream));
  return DYB_OK;
}

}  // namespace dyb

using namespace dyb;

struct dyb_smc {
  dyb_handle* h = nullptr;
  long long npart = 0, first = 0, n_local = 0;
  int np = 0, nphi = 0, rank = 0, nranks = 1;
  std::vector<double> phi;
  double trgt = 0.25, alp = 0.9;
  int systematic = 0;
  unsigned long long seed = 0;
  int stage = 0;                 // stages completed (0 = not initialised, 1 = particles initialised, ...)
  bool profiling = false;
  std::vector<cudaEvent_t> tev;  // SMC_TSEG + 1 events per profiled stage
  DevBuf para, loglh, logpost, w_g, wt_g, cw_g, bt, offs, scal, flag, tune, idx, para_g, loglh_g, logpost_g,
      part_g, mu, R, L, rdiag, aux, err, prop, qx, q0, lx, st, priorx, acc_g, nclamp, stats;
  void release() {
    DevBuf* all[] = {&para, &loglh, &logpost, &w_g, &wt_g, &cw_g, &bt, &offs, &scal, &flag, &tune, &idx, &para_g,
                     &loglh_g, &logpost_g, &part_g, &mu, &R, &L, &rdiag, &aux, &err, &prop, &qx, &q0, &lx, &st,
                     &priorx, &acc_g, &nclamp, &stats};
    for (DevBuf* b : all) b->release();
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
    tev.clear();
  }
};
// segments of a stage timed by dyb_smc_get_timing: correction, selection, moments + Cholesky, proposal, likelihood, accept
constexpr int SMC_TSEG = 6;

extern "C" {

// ---- SMC primitives -------------------------------------------------------------------------------------

int dyb_smc_reweight(int device, const double* w, const double* loglh, double dphi, int64_t n,
                     double* w_out, double* zhat, double* ess) {
  if (n <= 0 || !w || !loglh || !w_out) return fail(DYB_ERR_ARG, "bad argument");
  DYB_CUDA(cudaSetDevice(device));
  cudaStream_t st = device_stream(device);
  Stager sg(st);
  const size_t N = (size_t)n;
  const long long nb = (n + SMC_BLOCK - 1) / SMC_BLOCK;
  const double* dw = sg.in(w, N);
  const double* dl = sg.in(loglh, N);
  double* dwo = sg.out(w_out, N);
  double* wt = sg.temp<double>(N);
  double* bt = sg.temp<double>((size_t)nb);
  double* scal = sg.temp<double>(2);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  smc_incweight_kernel<<<nblk(n, 256), 256, 0, st>>>(dw, dl, dphi, n, wt);
  smc_block_sum_kernel<<<(unsigned)nb, 32, 0, st>>>(wt, n, 0, bt);
  smc_total_kernel<<<1, 32, 0, st>>>(bt, nb, scal);
  smc_normalise_kernel<<<nblk(n, 256), 256, 0, st>>>(wt, scal, n, dwo);
  smc_block_sum_kernel<<<(unsigned)nb, 32, 0, st>>>(dwo, n, 1, bt);
  smc_total_kernel<<<1, 32, 0, st>>>(bt, nb, scal + 1);
  DYB_CUDA(cudaGetLastError());
  double hs[2];
  DYB_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
  DYB_CUDA(sg.finish());
  if (zhat) *zhat = hs[0];
  if (ess) *ess = 1.0 / hs[1];
  return DYB_OK;
}

static int resample_impl(int device, const double* w, const double* u, double u0, int systematic, int64_t n,
                         int64_t* idx, int64_t* n_clamped) {
  if (n <= 0 || !w || !idx || (!systematic && !u)) return fail(DYB_ERR_ARG, "bad argument");
  DYB_CUDA(cudaSetDevice(device));
  cudaStream_t st = device_stream(device);
  Stager sg(st);
  const size_t N = (size_t)n;
  const long long nb = (n + SMC_BLOCK - 1) / SMC_BLOCK;
  const double* dw = sg.in(w, N);
  const double* du = systematic ? nullptr : sg.in(u, N);
  long long* didx = reinterpret_cast<long long*>(sg.out(idx, N));
  double* cw = sg.temp<double>(N);
  double* bt = sg.temp<double>((size_t)nb);
  double* offs = sg.temp<double>((size_t)nb);
  unsigned long long* dcl = sg.temp<unsigned long long>(1);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  DYB_CUDA(cudaMemsetAsync(dcl, 0, sizeof(unsigned long long), st));
  smc_block_scan_kernel<<<(unsigned)nb, 32, 0, st>>>(dw, n, cw, bt);
  smc_offsets_kernel<<<1, 32, 0, st>>>(bt, nb, offs);
  smc_add_offsets_kernel<<<nblk(n, 256), 256, 0, st>>>(cw, offs, n);
  smc_search_kernel<<<nblk(n, 256), 256, 0, st>>>(cw, n, du, u0, systematic, 0, n, didx, dcl);
  DYB_CUDA(cudaGetLastError());
  unsigned long long hcl = 0;
  DYB_CUDA(cudaMemcpyAsync(&hcl, dcl, sizeof(hcl), cudaMemcpyDeviceToHost, st));
  DYB_CUDA(sg.finish());
  if (n_clamped) *n_clamped = (int64_t)hcl;
  return DYB_OK;
}

int dyb_smc_resample_multinomial(int device, const double* w, const double* u, int64_t n, int64_t* idx, int64_t* n_clamped) {
  return resample_impl(device, w, u, 0.0, 0, n, idx, n_clamped);
}
int dyb_smc_resample_systematic(int device, const double* w, double u0, int64_t n, int64_t* idx, int64_t* n_clamped) {
  return resample_impl(device, w, nullptr, u0, 1, n, idx, n_clamped);
}

// second central moment partials in the canonical order: staged through shared memory when a block of particles fits
static cudaError_t launch_moment2(const double* para, const double* w, const double* mu, long long n, int np, long long nb,
                                  double* part, int device, cudaStream_t st) {
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  const size_t smem = smc_moment2_smem_bytes(np);
  if (nb > 0 && smem + 1024 <= (size_t)max_smem) {
    cudaError_t e = cudaFuncSetAttribute(smc_moment2_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;                   // per device: set on every call (host-side, microseconds)
    unsigned groups = (unsigned)((296 + nb - 1) / nb);                       // about two CTAs per SM in total
    const unsigned maxg = (unsigned)((np * np + SMC_MOM_TILE - 1) / SMC_MOM_TILE);
    if (groups > maxg) groups = maxg;
    if (groups < 1) groups = 1;
    smc_moment2_staged_kernel<<<dim3((unsigned)nb, groups), dim3(32, SMC_MOM_TILE), smem, st>>>(para, w, mu, n, np, part);
  } else if (nb > 0) {
    smc_moment_kernel<<<dim3((unsigned)nb, nblk(np * np, SMC_MOM_TILE)), dim3(32, SMC_MOM_TILE), 0, st>>>(para, w, mu, n, np, part);
  }
  return cudaGetLastError();
}

int dyb_smc_moments(int device, const double* para, const double* w, int64_t n, int32_t np, double* mu, double* R) {
  if (n <= 0 || np <= 0 || np > MAX_EST || !para || !w || !mu || !R) return fail(DYB_ERR_ARG, "bad argument");
  DYB_CUDA(cudaSetDevice(device));
  cudaStream_t st = device_stream(device);
  Stager sg(st);
  const size_t N = (size_t)n;
  const long long nb = (n + SMC_BLOCK - 1) / SMC_BLOCK;
  const double* dp = sg.in(para, N * np);
  const double* dw = sg.in(w, N);
  double* dmu = sg.out(mu, (size_t)np);
  double* dR = sg.out(R, (size_t)np * np);
  double* part = sg.temp<double>((size_t)nb * np * np);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  smc_moment_kernel<<<dim3((unsigned)nb, nblk(np, SMC_MOM_TILE)), dim3(32, SMC_MOM_TILE), 0, st>>>(dp, dw, nullptr, n, np, part);
  smc_reduce_partials_kernel<<<nblk(np, 64), 64, 0, st>>>(part, nb, np, dmu);
  DYB_CUDA(launch_moment2(dp, dw, dmu, n, np, nb, part, device, st));
  smc_reduce_partials_kernel<<<nblk(np * np, 64), 64, 0, st>>>(part, nb, np * np, dR);
  DYB_CUDA(cudaGetLastError());
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

// ---- communicator ------------------------------------------------------------------------------------
int dyb_comm_unique_id(void* id) {
  if (!id) return fail(DYB_ERR_ARG, "null argument");
  NcclApi* api = nccl_api();
  if (!api) return fail(DYB_ERR_NCCL, "NCCL is not available");
  ncclUniqueId u;
  DYB_NCCL(api->GetUniqueId(&u));
  static_assert(sizeof(ncclUniqueId) == DYB_COMM_ID_BYTES, "ncclUniqueId size");
  std::memcpy(id, &u, sizeof(u));
  return DYB_OK;
}

int dyb_comm_init(dyb_handle* h, int32_t rank, int32_t nranks, const void* id) {
  if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(DYB_ERR_ARG, "bad argument");
  NcclApi* api = nccl_api();
  if (!api) return fail(DYB_ERR_NCCL, "NCCL is not available");
  DYB_CUDA(cudaSetDevice(h->device));
  comm_release(h);
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  Comm* c = new Comm();
  c->rank = rank; c->nranks = nranks;
  ncclResult_t r = api->CommInitRank(&c->comm, nranks, u, rank);
  if (r != ncclSuccess) { delete c; return fail(DYB_ERR_NCCL, std::string("ncclCommInitRank: ") + api->GetErrorString(r)); }
  h->comm = c;
  // first collective: NCCL builds its channels and registers buffers lazily, keep that out of the samplers' timings
  double* warm = nullptr;
  DYB_CUDA(cudaMalloc(&warm, 8 * (size_t)nranks));
  DYB_CUDA(cudaMemsetAsync(warm, 0, 8 * (size_t)nranks, h->stream));
  DYB_NCCL(api->AllGather(warm + rank, warm, 1, ncclDouble, c->comm, h->stream));
  DYB_NCCL(api->AllReduce(warm, warm, 1, ncclDouble, ncclSum, c->comm, h->stream));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  DYB_CUDA(cudaFree(warm));
  return DYB_OK;
}

int dyb_comm_destroy(dyb_handle* h) {
  if (!h) return fail(DYB_ERR_ARG, "null handle");
  if (h->stream) cudaStreamSynchronize(h->stream);
  comm_release(h);
  return DYB_OK;
}

int dyb_comm_info(const dyb_handle* h, int32_t* rank, int32_t* nranks) {
  if (!h) return fail(DYB_ERR_ARG, "null handle");
  if (rank) *rank = h->comm ? h->comm->rank : 0;
  if (nranks) *nranks = h->comm ? h->comm->nranks : 1;
  return DYB_OK;
}

// sum over ranks of `count` doubles / int64 held on the device (in place); a no-op without a communicator
int dyb_comm_allreduce_sum(dyb_handle* h, void* d_buf, int64_t count, int32_t is_int64) {
  if (!h || !d_buf || count < 0) return fail(DYB_ERR_ARG, "bad argument");
  if (!h->comm || h->comm->nranks == 1 || count == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  DYB_NCCL(nccl_api()->AllReduce(d_buf, d_buf, (size_t)count, is_int64 ? ncclInt64 : ncclDouble, ncclSum, h->comm->comm, h->stream));
  return DYB_OK;
}

// ---- priors / posterior --------------------------------------------------------------------------------
int dyb_set_priors(dyb_handle* h, int32_t np, const int32_t* shape, const double* a, const double* b) {
  if (!h || np < 0 || np > MAX_EST || (np > 0 && (!shape || !a || !b))) return fail(DYB_ERR_ARG, "bad argument");
  if (np != h->hs.np) return fail(DYB_ERR_ARG, "one prior per estimated parameter is required");
  std::vector<PriorDev> pr((size_t)np);
  for (int k = 0; k < np; ++k) {
    PriorDev p;
    p.shape = shape[k]; p.a = a[k]; p.b = b[k];
    switch (shape[k]) {
      case 1: if (!(a[k] > 0 && b[k] > 0)) return fail(DYB_ERR_ARG, "Beta prior needs a, b > 0");
        p.cst = -(std::lgamma(a[k]) + std::lgamma(b[k]) - std::lgamma(a[k] + b[k])); p.transform = 2; break;
      case 2: if (!(a[k] > 0 && b[k] > 0)) return fail(DYB_ERR_ARG, "Gamma prior needs shape, scale > 0");
        p.cst = -std::lgamma(a[k]) - a[k] * std::log(b[k]); p.transform = 1; break;
      case 3: if (!(b[k] > 0)) return fail(DYB_ERR_ARG, "Normal prior needs std > 0");
        p.cst = -std::log(b[k]) - 0.5 * std::log(2.0 * M_PI); p.transform = 0; break;
      case 4: if (!(a[k] > 0 && b[k] > 0)) return fail(DYB_ERR_ARG, "InverseGamma1 prior needs alpha, theta > 0");
        p.cst = std::log(2.0) + a[k] * std::log(b[k]) - std::lgamma(a[k]); p.transform = 1; break;
      case 5: if (!(b[k] > a[k])) return fail(DYB_ERR_ARG, "Uniform prior needs a < b");
        p.cst = -std::log(b[k] - a[k]); p.transform = 3; break;
      case 6: if (!(a[k] > 0 && b[k] > 0)) return fail(DYB_ERR_ARG, "InverseGamma prior needs alpha, theta > 0");
        p.cst = a[k] * std::log(b[k]) - std::lgamma(a[k]); p.transform = 1; break;
      case 8: if (!(a[k] > 0 && b[k] > 0)) return fail(DYB_ERR_ARG, "Weibull prior needs shape, scale > 0");
        p.cst = std::log(a[k] / b[k]); p.transform = 1; break;
      default: return fail(DYB_ERR_ARG, "unknown prior shape code");
    }
    pr[k] = p;
  }
  h->priors = pr;
  return DYB_OK;
}

int dyb_logprior_batch(dyb_handle* h, const double* theta, int64_t n, double* logprior) {
  if (!h || n < 0 || (n > 0 && (!theta || !logprior))) return fail(DYB_ERR_ARG, "null argument");
  if (n == 0) return DYB_OK;
  PriorSetDev ps;
  int rc = prior_set(h, &ps);
  if (rc) return rc;
  DYB_CUDA(cudaSetDevice(h->device));
  Stager sg(h->stream);
  const double* dth = sg.in(theta, (size_t)ps.np * n);
  double* dlp = sg.out(logprior, (size_t)n);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  logprior_kernel<<<nblk(n, 128), 128, 0, h->stream>>>(ps, dth, n, dlp);
  DYB_CUDA(cudaGetLastError());
  h->launches += 1;
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

int dyb_transform_batch(dyb_handle* h, const double* in, int64_t n, int32_t inverse, double* out, double* logjac) {
  if (!h || n < 0 || (n > 0 && (!in || !out))) return fail(DYB_ERR_ARG, "null argument");
  if (n == 0) return DYB_OK;
  PriorSetDev ps;
  int rc = prior_set(h, &ps);
  if (rc) return rc;
  DYB_CUDA(cudaSetDevice(h->device));
  Stager sg(h->stream);
  const double* din = sg.in(in, (size_t)ps.np * n);
  double* dout = sg.out(out, (size_t)ps.np * n);
  double* dlj = sg.out(logjac, (size_t)n);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  transform_kernel<<<nblk(n, 128), 128, 0, h->stream>>>(ps, din, n, inverse, dout, dlj);
  DYB_CUDA(cudaGetLastError());
  h->launches += 1;
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

int dyb_logposterior_batch(dyb_handle* h, const double* theta, int64_t n, double* logpost, double* loglik, int32_t* status) {
  if (!h || n < 0 || (n > 0 && (!theta || !logpost))) return fail(DYB_ERR_ARG, "null argument");
  if (n == 0) return DYB_OK;
  PriorSetDev ps;
  int rc = prior_set(h, &ps);
  if (rc) return rc;
  DYB_CUDA(cudaSetDevice(h->device));
  Stager sg(h->stream);
  const size_t N = (size_t)n;
  const double* dth = sg.in(theta, (size_t)ps.np * N);
  double* dpost = sg.out(logpost, N);
  double* dll = loglik ? sg.out(loglik, N) : sg.temp<double>(N);
  int* dst = status ? sg.out(status, N) : sg.temp<int>(N);
  double* dpr = sg.temp<double>(N);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  rc = dyb_loglik_batch_device(h, dth, n, dll, dst);
  if (rc) return rc;
  logprior_kernel<<<nblk(n, 128), 128, 0, h->stream>>>(ps, dth, n, dpr);
  logposterior_kernel<<<nblk(n, 256), 256, 0, h->stream>>>(dpr, dll, dst, n, dpost);
  DYB_CUDA(cudaGetLastError());
  h->launches += 2;
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

int dyb_rng_fill(int device, uint64_t seed, int64_t first, int64_t n, uint32_t step, uint32_t stream, int32_t kind,
                 int32_t count, double* out) {
  if (n <= 0 || count <= 0 || !out || kind < 0 || kind > 1) return fail(DYB_ERR_ARG, "bad argument");
  DYB_CUDA(cudaSetDevice(device));
  cudaStream_t st = device_stream(device);
  Stager sg(st);
  double* d = sg.out(out, (size_t)n * count);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  rng_fill_kernel<<<nblk(n, 128), 128, 0, st>>>(seed, first, n, step, stream, kind, count, d);
  DYB_CUDA(cudaGetLastError());
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

// ---- SMC ---------------------------------------------------------------------------------------------
int dyb_smc_default_options(dyb_smc_options* o) {
  if (!o) return fail(DYB_ERR_ARG, "null argument");
  o->npart = 1000; o->nphi = 400; o->lam = 2.0; o->c0 = 0.5; o->acpt0 = 0.25; o->trgt = 0.25; o->alp = 0.9;
  o->resampling = 0; o->seed = 0;
  return DYB_OK;
}

int dyb_smc_create(dyb_handle* h, const dyb_smc_options* o, const double* phi, dyb_smc** out) {
  if (!h || !o || !out) return fail(DYB_ERR_ARG, "null argument");
  *out = nullptr;
  if (o->npart < 1 || o->nphi < 2 || !(o->alp >= 0 && o->alp <= 1) || !(o->c0 > 0) || o->resampling < 0 || o->resampling > 1)
    return fail(DYB_ERR_ARG, "invalid SMC option");
  if ((int)h->priors.size() != h->hs.np) return fail(DYB_ERR_STATE, "dyb_set_priors must be called before dyb_smc_create");
  const int nranks = h->comm ? h->comm->nranks : 1, rank = h->comm ? h->comm->rank : 0;
  if (nranks > 1 && o->npart % ((long long)SMC_BLOCK * nranks) != 0)
    return fail(DYB_ERR_ARG, "with several ranks the particle count must be a multiple of 1024 * nranks");
  DYB_CUDA(cudaSetDevice(h->device));
  dyb_smc* s = new dyb_smc();
  s->h = h; s->npart = o->npart; s->nranks = nranks; s->rank = rank;
  s->n_local = o->npart / nranks; s->first = s->n_local * rank;
  s->np = h->hs.np; s->nphi = o->nphi; s->trgt = o->trgt; s->alp = o->alp; s->systematic = o->resampling; s->seed = o->seed;
  s->phi.resize((size_t)o->nphi);
  for (int i = 0; i < o->nphi; ++i) s->phi[i] = phi ? phi[i] : std::pow((double)i / (double)(o->nphi - 1), o->lam);   // smc.jl:48
  const size_t N = (size_t)s->npart, n = (size_t)s->n_local, np = (size_t)s->np;
  const size_t nb = (N + SMC_BLOCK - 1) / SMC_BLOCK;
  cudaError_t e = cudaSuccess;
  auto R = [&](DevBuf& b, size_t bytes) { if (e == cudaSuccess) e = b.reserve(bytes + 16); };
  R(s->para, 8 * np * n); R(s->loglh, 8 * n); R(s->logpost, 8 * n); R(s->w_g, 8 * N); R(s->wt_g, 8 * N); R(s->cw_g, 8 * N);
  R(s->bt, 8 * nb); R(s->offs, 8 * nb); R(s->scal, 16); R(s->flag, 4); R(s->tune, sizeof(SmcTune)); R(s->idx, 8 * n);
  R(s->para_g, 8 * np * N); R(s->loglh_g, 8 * N); R(s->logpost_g, 8 * N); R(s->part_g, 8 * nb * np * np);
  R(s->mu, 8 * np); R(s->R, 8 * np * np); R(s->L, 8 * np * np); R(s->rdiag, 8 * np); R(s->aux, 16); R(s->err, 4);
  R(s->prop, 8 * np * n); R(s->qx, 8 * n); R(s->q0, 8 * n); R(s->lx, 8 * n); R(s->st, 4 * n); R(s->priorx, 8 * n);
  R(s->acc_g, 8 * (size_t)nranks); R(s->nclamp, 8); R(s->stats, 8 * SMC_STATS * (size_t)o->nphi);
  if (e != cudaSuccess) { s->release(); delete s; return fail(DYB_ERR_CUDA, cudaGetErrorString(e)); }
  const SmcTune t0{o->c0, o->acpt0};
  DYB_CUDA(cudaMemcpyAsync(s->tune.p, &t0, sizeof(t0), cudaMemcpyHostToDevice, h->stream));
  DYB_CUDA(cudaMemsetAsync(s->err.p, 0, 4, h->stream));
  DYB_CUDA(cudaMemsetAsync(s->nclamp.p, 0, 8, h->stream));
  DYB_CUDA(cudaMemsetAsync(s->stats.p, 0, 8 * SMC_STATS * (size_t)o->nphi, h->stream));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  *out = s;
  return DYB_OK;
}

int dyb_smc_destroy(dyb_smc* s) {
  if (!s) return DYB_OK;
  cudaSetDevice(s->h->device);
  cudaStreamSynchronize(s->h->stream);
  s->release();
  delete s;
  return DYB_OK;
}

int dyb_smc_local_range(const dyb_smc* s, int64_t* first, int64_t* n_local) {
  if (!s) return fail(DYB_ERR_ARG, "null argument");
  if (first) *first = s->first;
  if (n_local) *n_local = s->n_local;
  return DYB_OK;
}

// initial particles (smc.jl:84-101): the caller draws them from the prior; the likelihood, the prior density and
// logpost = phi[1] * loglh + logprior are evaluated here.  status[i] != 0 marks draws the reference would reject
// (obj = -Inf): replace them and call again.
int dyb_smc_init(dyb_smc* s, const double* para_local, int32_t* status) {
  if (!s || !para_local) return fail(DYB_ERR_ARG, "null argument");
  dyb_handle* h = s->h;
  DYB_CUDA(cudaSetDevice(h->device));
  PriorSetDev ps;
  int rc = prior_set(h, &ps);
  if (rc) return rc;
  const long long n = s->n_local;
  DYB_CUDA(cudaMemcpyAsync(s->para.p, para_local, 8 * (size_t)s->np * n, cudaMemcpyDefault, h->stream));
  rc = dyb_loglik_batch_device(h, s->para.as<double>(), n, s->loglh.as<double>(), s->st.as<int>());
  if (rc) return rc;
  logprior_kernel<<<nblk(n, 128), 128, 0, h->stream>>>(ps, s->para.as<double>(), n, s->priorx.as<double>());
  // logpost = phi[0] * loglh + prior, written with the accept kernel's arithmetic: phi * l + prior
  smc_init_logpost_kernel<<<nblk(n, 256), 256, 0, h->stream>>>(s->loglh.as<double>(), s->priorx.as<double>(), s->phi[0], n,
                                                              s->logpost.as<double>());
  smc_fill_kernel<<<nblk(s->npart, 256), 256, 0, h->stream>>>(s->w_g.as<double>(), s->npart, 1.0 / (double)s->npart);
  const double st0[SMC_STATS] = {1.0, (double)s->npart, 0.0, 0.0, 0.0, 0.0};
  DYB_CUDA(cudaMemcpyAsync(s->stats.p, st0, sizeof(st0), cudaMemcpyHostToDevice, h->stream));
  DYB_CUDA(cudaGetLastError());
  h->launches += 3;
  if (status) DYB_CUDA(cudaMemcpyAsync(status, s->st.p, 4 * (size_t)n, cudaMemcpyDefault, h->stream));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  s->stage = 1;
  return DYB_OK;
}

// one tempering stage i (0-based, i >= 1): correction, selection, mutation (smc.jl:110-243); asynchronous
static int smc_stage(dyb_smc* s, int i) {
  dyb_handle* h = s->h;
  cudaStream_t st = h->stream;
  const long long N = s->npart, n = s->n_local, first = s->first;
  const int np = s->np;
  const long long nb = (N + SMC_BLOCK - 1) / SMC_BLOCK, nb_l = (n + SMC_BLOCK - 1) / SMC_BLOCK;
  const double phi = s->phi[i], dphi = s->phi[i] - s->phi[i - 1];
  double* stats = s->stats.as<double>() + (size_t)SMC_STATS * i;
  double* w_g = s->w_g.as<double>();
  double* wt_g = s->wt_g.as<double>();
  int* flag = s->flag.as<int>();
  SmcTune* tune = s->tune.as<SmcTune>();
  PriorSetDev ps;
  int rc = prior_set(h, &ps);
  if (rc) return rc;

  cudaEvent_t* tev = nullptr;
  if (s->profiling) {
    const size_t base = s->tev.size();
    for (int k = 0; k <= SMC_TSEG; ++k) { cudaEvent_t e; DYB_CUDA(cudaEventCreate(&e)); s->tev.push_back(e); }
    tev = &s->tev[base];
  }
#define DYB_MARK(k) do { if (tev) DYB_CUDA(cudaEventRecord(tev[k], st)); } while (0)
  DYB_MARK(0);
  // 1. correction (smc.jl:118-123) and ESS (:129); the weights of ALL particles are replicated on every rank
  smc_incweight_kernel<<<nblk(n, 256), 256, 0, st>>>(w_g + first, s->loglh.as<double>(), dphi, n, wt_g + first);
  rc = all_gather(h, wt_g + first, wt_g, (size_t)n, ncclDouble, 8);
  if (rc) return rc;
  smc_block_sum_kernel<<<(unsigned)nb, 32, 0, st>>>(wt_g, N, 0, s->bt.as<double>());
  smc_total_kernel<<<1, 32, 0, st>>>(s->bt.as<double>(), nb, s->scal.as<double>());
  smc_normalise_kernel<<<nblk(N, 256), 256, 0, st>>>(wt_g, s->scal.as<double>(), N, w_g);
  smc_block_sum_kernel<<<(unsigned)nb, 32, 0, st>>>(w_g, N, 1, s->bt.as<double>());
  smc_total_kernel<<<1, 32, 0, st>>>(s->bt.as<double>(), nb, s->scal.as<double>() + 1);
  smc_stage_scalars_kernel<<<1, 32, 0, st>>>(s->scal.as<double>(), N, s->trgt, tune, flag, stats);

  DYB_MARK(1);
  // 2. selection (smc.jl:130-145): every kernel is predicated on the device flag, the collectives are unconditional
  smc_block_scan_kernel<<<(unsigned)nb, 32, 0, st>>>(w_g, N, s->cw_g.as<double>(), s->bt.as<double>());
  smc_offsets_kernel<<<1, 32, 0, st>>>(s->bt.as<double>(), nb, s->offs.as<double>());
  smc_add_offsets_kernel<<<nblk(N, 256), 256, 0, st>>>(s->cw_g.as<double>(), s->offs.as<double>(), N);
  smc_select_kernel<<<nblk(n, 256), 256, 0, st>>>(s->cw_g.as<double>(), N, s->seed, (unsigned)i, s->systematic, first, n, flag,
                                                  s->idx.as<long long>(), s->nclamp.as<unsigned long long>());
  if (s->nranks > 1) DYB_NCCL(nccl_api()->GroupStart());
  rc = all_gather(h, s->para.p, s->para_g.p, (size_t)np * n, ncclDouble, 8);
  if (!rc) rc = all_gather(h, s->loglh.p, s->loglh_g.p, (size_t)n, ncclDouble, 8);
  if (!rc) rc = all_gather(h, s->logpost.p, s->logpost_g.p, (size_t)n, ncclDouble, 8);
  if (s->nranks > 1) DYB_NCCL(nccl_api()->GroupEnd());
  if (rc) return rc;
  smc_take_kernel<<<nblk(n * (np + 2), 256), 256, 0, st>>>(s->para_g.as<double>(), s->loglh_g.as<double>(), s->logpost_g.as<double>(),
                                                           s->idx.as<long long>(), n, np, flag, s->para.as<double>(),
                                                           s->loglh.as<double>(), s->logpost.as<double>());
  smc_equal_weights_kernel<<<nblk(N, 256), 256, 0, st>>>(w_g, N, flag);

  DYB_MARK(2);
  // 3. mutation: particle mean and covariance (smc.jl:154-165) in the canonical blocked order
  double* part_g = s->part_g.as<double>();
  smc_moment_kernel<<<dim3((unsigned)nb_l, nblk(np, SMC_MOM_TILE)), dim3(32, SMC_MOM_TILE), 0, st>>>(
      s->para.as<double>(), w_g + first, nullptr, n, np, part_g + (size_t)s->rank * nb_l * np);
  rc = all_gather(h, part_g + (size_t)s->rank * nb_l * np, part_g, (size_t)nb_l * np, ncclDouble, 8);
  if (rc) return rc;
  smc_reduce_partials_kernel<<<nblk(np, 64), 64, 0, st>>>(part_g, nb, np, s->mu.as<double>());
  DYB_CUDA(launch_moment2(s->para.as<double>(), w_g + first, s->mu.as<double>(), n, np, nb_l,
                          part_g + (size_t)s->rank * nb_l * np * np, h->device, st));
  rc = all_gather(h, part_g + (size_t)s->rank * nb_l * np * np, part_g, (size_t)nb_l * np * np, ncclDouble, 8);
  if (rc) return rc;
  smc_reduce_partials_kernel<<<nblk(np * np, 64), 64, 0, st>>>(part_g, nb, np * np, s->R.as<double>());
  smc_chol_kernel<<<1, 32, 0, st>>>(s->R.as<double>(), np, s->L.as<double>(), s->rdiag.as<double>(), s->aux.as<double>(), s->err.as<int>());

  DYB_MARK(3);
  SmcProposeArgs pa;
  pa.para = s->para.as<double>(); pa.loglh = s->loglh.as<double>(); pa.logpost = s->logpost.as<double>();
  pa.mu = s->mu.as<double>(); pa.L = s->L.as<double>(); pa.rdiag = s->rdiag.as<double>(); pa.aux = s->aux.as<double>();
  pa.tune = tune; pa.alp_mix = s->alp; pa.seed = s->seed; pa.stage = (unsigned)i; pa.first = first; pa.n_local = n; pa.np = np;
  pa.prop = s->prop.as<double>(); pa.qx = s->qx.as<double>(); pa.q0 = s->q0.as<double>();
  smc_propose_kernel<<<nblk(n, 128), 128, sizeof(double) * (np * np + 2 * np), st>>>(pa);
  DYB_CUDA(cudaGetLastError());
  DYB_MARK(4);
  rc = dyb_loglik_batch_device(h, s->prop.as<double>(), n, s->lx.as<double>(), s->st.as<int>());
  if (rc) return rc;
  DYB_MARK(5);
  logprior_kernel<<<nblk(n, 128), 128, 0, st>>>(ps, s->prop.as<double>(), n, s->priorx.as<double>());
  unsigned long long* acc = s->acc_g.as<unsigned long long>();
  DYB_CUDA(cudaMemsetAsync(acc + s->rank, 0, 8, st));
  SmcAcceptArgs aa;
  aa.para = s->para.as<double>(); aa.loglh = s->loglh.as<double>(); aa.logpost = s->logpost.as<double>();
  aa.prop = s->prop.as<double>(); aa.lx = s->lx.as<double>(); aa.status = s->st.as<int>(); aa.priorx = s->priorx.as<double>();
  aa.qx = s->qx.as<double>(); aa.q0 = s->q0.as<double>(); aa.phi = phi; aa.dphi = dphi; aa.seed = s->seed; aa.stage = (unsigned)i;
  aa.first = first; aa.n_local = n; aa.np = np; aa.accepted = acc + s->rank;
  smc_accept_kernel<<<nblk(n, 256), 256, 0, st>>>(aa);
  rc = all_gather(h, acc + s->rank, acc, 1, ncclUint64, 8);
  if (rc) return rc;
  smc_acpt_kernel<<<1, 32, 0, st>>>(acc, s->nranks, N, tune, stats);
  smc_store_clamped_kernel<<<1, 32, 0, st>>>(s->nclamp.as<unsigned long long>(), stats);
  DYB_CUDA(cudaGetLastError());
  DYB_MARK(6);
#undef DYB_MARK
  h->launches += 23;
  return DYB_OK;
}

int dyb_smc_run(dyb_smc* s, int32_t n_stages) {
  if (!s || n_stages < 0) return fail(DYB_ERR_ARG, "bad argument");
  if (s->stage < 1) return fail(DYB_ERR_STATE, "dyb_smc_init has not been called");
  DYB_CUDA(cudaSetDevice(s->h->device));
  for (int k = 0; k < n_stages && s->stage < s->nphi; ++k) {
    int rc = smc_stage(s, s->stage);
    if (rc) return rc;
    s->stage += 1;
  }
  return DYB_OK;
}

int dyb_smc_stages_done(const dyb_smc* s) { return s ? s->stage : 0; }

int dyb_smc_set_profiling(dyb_smc* s, int32_t on) {
  if (!s) return fail(DYB_ERR_ARG, "null argument");
  s->profiling = on != 0;
  return DYB_OK;
}

int dyb_smc_get_timing(dyb_smc* s, int32_t* n_stages, double* ms) {
  if (!s || !ms) return fail(DYB_ERR_ARG, "null argument");
  DYB_CUDA(cudaSetDevice(s->h->device));
  DYB_CUDA(cudaStreamSynchronize(s->h->stream));
  for (int k = 0; k < SMC_TSEG; ++k) ms[k] = 0.0;
  const size_t ns = s->tev.size() / (SMC_TSEG + 1);
  for (size_t i = 0; i < ns; ++i)
    for (int k = 0; k < SMC_TSEG; ++k) {
      float x = 0;
      DYB_CUDA(cudaEventElapsedTime(&x, s->tev[i * (SMC_TSEG + 1) + k], s->tev[i * (SMC_TSEG + 1) + k + 1]));
      ms[k] += x;
    }
  if (n_stages) *n_stages = (int32_t)ns;
  return DYB_OK;
}

int dyb_smc_get_particles(dyb_smc* s, double* para, double* w, double* loglh, double* logpost) {
  if (!s) return fail(DYB_ERR_ARG, "null argument");
  dyb_handle* h = s->h;
  DYB_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)s->n_local;
  if (para) DYB_CUDA(cudaMemcpyAsync(para, s->para.p, 8 * n * s->np, cudaMemcpyDefault, h->stream));
  if (w) DYB_CUDA(cudaMemcpyAsync(w, s->w_g.as<double>() + s->first, 8 * n, cudaMemcpyDefault, h->stream));
  if (loglh) DYB_CUDA(cudaMemcpyAsync(loglh, s->loglh.p, 8 * n, cudaMemcpyDefault, h->stream));
  if (logpost) DYB_CUDA(cudaMemcpyAsync(logpost, s->logpost.p, 8 * n, cudaMemcpyDefault, h->stream));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  return DYB_OK;
}

// per-stage records (nphi x 6, row-major: zhat, ESS, resampled, c, acpt, clamped so far), the particle mean and
// covariance of the last mutation step, and logmdd = sum_i log zhat_i (smc.jl:301).  Synchronises.
int dyb_smc_get_stats(dyb_smc* s, double* stats, double* mu, double* R, double* logmdd) {
  if (!s) return fail(DYB_ERR_ARG, "null argument");
  dyb_handle* h = s->h;
  DYB_CUDA(cudaSetDevice(h->device));
  std::vector<double> hs((size_t)SMC_STATS * s->nphi);
  int err = 0;
  DYB_CUDA(cudaMemcpyAsync(hs.data(), s->stats.p, 8 * hs.size(), cudaMemcpyDeviceToHost, h->stream));
  DYB_CUDA(cudaMemcpyAsync(&err, s->err.p, 4, cudaMemcpyDeviceToHost, h->stream));
  if (mu) DYB_CUDA(cudaMemcpyAsync(mu, s->mu.p, 8 * (size_t)s->np, cudaMemcpyDefault, h->stream));
  if (R) DYB_CUDA(cudaMemcpyAsync(R, s->R.p, 8 * (size_t)s->np * s->np, cudaMemcpyDefault, h->stream));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  if (stats) std::memcpy(stats, hs.data(), 8 * hs.size());
  if (logmdd) {
    double a = 0.0;
    for (int i = 1; i < s->stage; ++i) a += std::log(hs[(size_t)SMC_STATS * i]);
    *logmdd = a;
  }
  if (err) return fail(DYB_ERR_STATE, "the particle covariance was not positive definite at some stage (the reference's cholesky throws)");
  return DYB_OK;
}

// ---- random-walk Metropolis ------------------------------------------------------------------------------
int dyb_rwmh_run(dyb_handle* h, int64_t n_chains, int64_t first_chain, int32_t n_iter, uint32_t first_step,
                 int32_t transformed, const double* chol_cov, uint64_t seed, double* y, double* logp,
                 int64_t* accepted, double* history, double* history_logp) {
  if (!h || n_chains < 0 || n_iter < 0 || first_chain < 0 || (n_chains > 0 && (!chol_cov || !y))) return fail(DYB_ERR_ARG, "bad argument");
  if (n_chains == 0) return DYB_OK;
  PriorSetDev ps;
  int rc = prior_set(h, &ps);
  if (rc) return rc;
  DYB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int np = ps.np;
  const size_t n = (size_t)n_chains;
  Stager sg(st);
  const double* dL = sg.in(chol_cov, (size_t)np * np);
  double* dy;
  if (is_device_ptr(y)) dy = y;
  else {
    dy = sg.out(y, np * n);
    if (dy) DYB_CUDA(cudaMemcpyAsync(dy, y, 8 * np * n, cudaMemcpyHostToDevice, st));
  }
  double* dlogp = logp ? sg.out(logp, n) : sg.temp<double>(n);
  unsigned long long* dacc = accepted ? reinterpret_cast<unsigned long long*>(sg.out(accepted, n)) : sg.temp<unsigned long long>(n);
  double* dhist = history ? sg.out(history, np * n * (size_t)n_iter) : nullptr;
  double* dhlp = history_logp ? sg.out(history_logp, n * (size_t)n_iter) : nullptr;
  double* yprop = sg.temp<double>(np * n);
  double* theta = sg.temp<double>(np * n);
  double* lj = sg.temp<double>(n);
  double* ll = sg.temp<double>(n);
  double* pr = sg.temp<double>(n);
  int* stt = sg.temp<int>(n);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  DYB_CUDA(cudaMemsetAsync(dacc, 0, 8 * n, st));
  // log target at the initial point
  if (transformed) {
    transform_kernel<<<nblk(n_chains, 128), 128, 0, st>>>(ps, dy, n_chains, 0, theta, lj);
  } else {
    DYB_CUDA(cudaMemcpyAsync(theta, dy, 8 * np * n, cudaMemcpyDeviceToDevice, st));
    DYB_CUDA(cudaMemsetAsync(lj, 0, 8 * n, st));
  }
  rc = dyb_loglik_batch_device(h, theta, n_chains, ll, stt);
  if (rc) return rc;
  logprior_kernel<<<nblk(n_chains, 128), 128, 0, st>>>(ps, theta, n_chains, pr);
  rwmh_init_kernel<<<nblk(n_chains, 256), 256, 0, st>>>(ll, stt, pr, lj, n_chains, dlogp);
  h->launches += 3;
  for (int it = 0; it < n_iter; ++it) {
    rwmh_propose_kernel<<<nblk(n_chains, 128), 128, sizeof(double) * np * np, st>>>(ps, dy, dL, transformed, seed, first_step + (unsigned)it,
                                                                                   first_chain, n_chains, yprop, theta, lj);
    rc = dyb_loglik_batch_device(h, theta, n_chains, ll, stt);
    if (rc) return rc;
    logprior_kernel<<<nblk(n_chains, 128), 128, 0, st>>>(ps, theta, n_chains, pr);
    rwmh_accept_kernel<<<nblk(n_chains, 256), 256, 0, st>>>(dy, dlogp, yprop, ll, stt, pr, lj, seed, first_step + (unsigned)it, first_chain,
                                                            n_chains, np, dacc, dhist ? dhist + np * n * (size_t)it : nullptr,
                                                            dhlp ? dhlp + n * (size_t)it : nullptr);
    h->launches += 3;
  }
  DYB_CUDA(cudaGetLastError());
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

}  // extern "C"
