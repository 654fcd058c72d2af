// Samplers around the batched likelihood (include/dynare_b200.h, "samplers" section): priors and the log-posterior,
// the SMC tempering recursion with its particles resident on the device(s), random-walk Metropolis on many chains,
// and the NCCL communicator that couples the particle shards of several GPUs.
#include <dlfcn.h>
#include <cstdlib>
#include <nccl.h>

#include "internal.h"
#include "sampler_kernels.cuh"
#include "smc_stage.cuh"

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
  double trgt = 0.25, alp = 0.9, c0 = 0.5, acpt0 = 0.25;
  int systematic = 0;
  unsigned long long seed = 0;
  int stage = 0;                 // stages completed (0 = not initialised, 1 = particles initialised, ...)
  bool profiling = false;
  std::vector<cudaEvent_t> tev;  // SMC_TSEG + 1 events per profiled stage
  // exchange of the mutated particles (smc_stage.cuh): peer stores into CUDA-IPC mapped slabs, or one ncclAllGather
  bool p2p = false;              // peers' slabs are mapped into this process
  bool published = false;        // the initial payload has reached every rank
  unsigned long long epoch = 0;  // epoch base of the flags, advanced at every (re)start of the recursion
  void* peer_slab[SMC_MAX_RANKS] = {};
  size_t slab_bytes = 0, off_pay1 = 0, off_pub = 0, off_cnt = 0;
  // one stage as a CUDA graph per payload parity (the launch sequence does not depend on the stage)
  bool use_graph = true;
  cudaGraphExec_t graph[2] = {nullptr, nullptr};
  long long graph_cfg[2] = {-1, -1};
  DevBuf slab, ctl, phi_d, w_g, wt_g, cwl, bt, bt2, btw, offs, idx, xsel, part1, part2, mu, R, L, rdiag, aux, prop, qx, q0,
      lx, st, priorx, stats, tmp_para, tmp_ll, tmp_lp;
  void release() {
    DevBuf* all[] = {&slab, &ctl, &phi_d, &w_g, &wt_g, &cwl, &bt, &bt2, &btw, &offs, &idx, &xsel, &part1, &part2, &mu, &R, &L,
                     &rdiag, &aux, &prop, &qx, &q0, &lx, &st, &priorx, &stats, &tmp_para, &tmp_ll, &tmp_lp};
    for (DevBuf* b : all) b->release();
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
    tev.clear();
    for (int k = 0; k < 2; ++k) if (graph[k]) { cudaGraphExecDestroy(graph[k]); graph[k] = nullptr; }
  }
};
// segments of a stage timed by dyb_smc_get_timing: correction, selection, moments + Cholesky, proposal, likelihood, accept
constexpr int SMC_TSEG = 6;

static SmcStageArgs stage_args(const dyb_smc* s) {
  SmcStageArgs a{};
  a.ctl = s->ctl.as<SmcCtl>(); a.phi = s->phi_d.as<double>();
  for (int r = 0; r < SMC_MAX_RANKS; ++r) a.slab[r] = nullptr;
  if (s->p2p) for (int r = 0; r < s->nranks; ++r) a.slab[r] = static_cast<char*>(s->peer_slab[r]);
  a.slab[s->rank] = s->slab.as<char>();
  a.off_pay1 = s->off_pay1; a.off_pub = s->off_pub; a.off_cnt = s->off_cnt;
  a.nranks = s->nranks; a.rank = s->rank;
  a.npub = s->p2p ? s->nranks : 1; a.nwait = s->p2p ? s->nranks : 0;
  a.N = s->npart; a.n = s->n_local; a.first = s->first; a.np = s->np;
  a.w = s->w_g.as<double>(); a.wt = s->wt_g.as<double>(); a.cwl = s->cwl.as<double>(); a.bt = s->bt.as<double>();
  a.bt2 = s->bt2.as<double>(); a.btw = s->btw.as<double>(); a.offs = s->offs.as<double>(); a.idx = s->idx.as<long long>(); a.xsel = s->xsel.as<double>();
  a.part1 = s->part1.as<double>(); a.part2 = s->part2.as<double>(); a.mu = s->mu.as<double>(); a.R = s->R.as<double>();
  a.L = s->L.as<double>(); a.rdiag = s->rdiag.as<double>(); a.aux = s->aux.as<double>(); a.stats = s->stats.as<double>();
  a.prop = s->prop.as<double>(); a.qx = s->qx.as<double>(); a.q0 = s->q0.as<double>();
  a.lx = s->lx.as<double>(); a.status = s->st.as<int>();
  a.seed = s->seed; a.systematic = s->systematic; a.trgt = s->trgt; a.alp = s->alp;
  return a;
}

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

int dyb_smc_covariance_factor(int device, const double* R, int32_t np, double* Rpsd, double* L, int32_t* projected) {
  if (np <= 0 || np > MAX_EST || !R || !Rpsd || !L || !projected) return fail(DYB_ERR_ARG, "bad argument");
  DYB_CUDA(cudaSetDevice(device));
  cudaStream_t st = device_stream(device);
  Stager sg(st);
  const size_t n2 = (size_t)np * np;
  const double* dRin = sg.in(R, n2);
  double* dR = sg.out(Rpsd, n2);
  double* dL = sg.out(L, n2);
  double* drd = sg.temp<double>((size_t)np);
  double* daux = sg.temp<double>(4);
  double* dwork = sg.temp<double>(2 * n2);
  int* derr = sg.temp<int>(1);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  DYB_CUDA(cudaMemcpyAsync(dR, dRin, 8 * n2, cudaMemcpyDeviceToDevice, st));
  DYB_CUDA(cudaMemsetAsync(daux, 0, 32, st));
  DYB_CUDA(cudaMemsetAsync(derr, 0, 4, st));
  smc_covariance_factor_kernel<<<1, 32, 8 * n2, st>>>(dR, np, dL, drd, daux, dwork, derr);
  DYB_CUDA(cudaGetLastError());
  double aux1[4] = {0, 0, 0, 0};
  int err1 = 0;
  DYB_CUDA(cudaMemcpyAsync(aux1, daux, 32, cudaMemcpyDeviceToHost, st));
  DYB_CUDA(cudaMemcpyAsync(&err1, derr, 4, cudaMemcpyDeviceToHost, st));
  DYB_CUDA(sg.finish());
  *projected = err1 ? 2 : (aux1[2] > 0.0 ? 1 : 0);
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
  h->config_epoch += 1;
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
  double* dlj = inverse ? nullptr : sg.out(logjac, (size_t)n);      // the inverse map writes no log-Jacobian
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

// peer mappings of the payload slabs: every rank all-gathers its cudaIpcMemHandle_t through NCCL and opens the others'
static int smc_map_peers(dyb_smc* s) {
  dyb_handle* h = s->h;
  NcclApi* api = nccl_api();
  const int G = s->nranks;
  struct Rec { cudaIpcMemHandle_t hnd; int device; int ok; };
  static_assert(sizeof(Rec) % 8 == 0, "record size");
  Rec mine{};
  mine.device = h->device;
  mine.ok = cudaIpcGetMemHandle(&mine.hnd, s->slab.p) == cudaSuccess ? 1 : 0;
  if (!mine.ok) cudaGetLastError();
  Rec* d_all = nullptr;
  DYB_CUDA(cudaMalloc(&d_all, sizeof(Rec) * G));
  DYB_CUDA(cudaMemcpyAsync(d_all + s->rank, &mine, sizeof(Rec), cudaMemcpyHostToDevice, h->stream));
  DYB_NCCL(api->AllGather(d_all + s->rank, d_all, sizeof(Rec), ncclChar, h->comm->comm, h->stream));
  std::vector<Rec> all((size_t)G);
  DYB_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(Rec) * G, cudaMemcpyDeviceToHost, h->stream));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  int ok = 1;
  for (int r = 0; r < G; ++r) ok = ok && all[r].ok;
  for (int r = 0; r < G && ok; ++r) {
    if (r == s->rank) { s->peer_slab[r] = s->slab.p; continue; }
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, h->device, all[r].device) != cudaSuccess || !can) { cudaGetLastError(); ok = 0; break; }
    if (cudaIpcOpenMemHandle(&s->peer_slab[r], all[r].hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError(); s->peer_slab[r] = nullptr; ok = 0;
    }
  }
  // every rank must take the same decision
  int* d_ok = reinterpret_cast<int*>(d_all);
  DYB_CUDA(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  DYB_NCCL(api->AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, h->comm->comm, h->stream));
  DYB_CUDA(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  DYB_CUDA(cudaFree(d_all));
  if (!ok) {
    for (int r = 0; r < G; ++r) if (r != s->rank && s->peer_slab[r]) { cudaIpcCloseMemHandle(s->peer_slab[r]); s->peer_slab[r] = nullptr; }
    cudaGetLastError();
  }
  s->p2p = ok != 0;
  return DYB_OK;
}

// all ranks have finished everything enqueued so far (used before slabs are rewritten from scratch or unmapped)
static int smc_rank_barrier(dyb_smc* s) {
  dyb_handle* h = s->h;
  if (s->nranks == 1 || !h->comm) return DYB_OK;
  DYB_NCCL(nccl_api()->AllReduce(s->aux.p, s->aux.p, 1, ncclDouble, ncclSum, h->comm->comm, h->stream));
  return DYB_OK;
}

int dyb_smc_create(dyb_handle* h, const dyb_smc_options* o, const double* phi, dyb_smc** out) {
  if (!h || !o || !out) return fail(DYB_ERR_ARG, "null argument");
  *out = nullptr;
  if (o->npart < 1 || o->nphi < 2 || !(o->alp >= 0 && o->alp <= 1) || !(o->c0 > 0) || o->resampling < 0 || o->resampling > 1)
    return fail(DYB_ERR_ARG, "invalid SMC option");
  if ((int)h->priors.size() != h->hs.np) return fail(DYB_ERR_STATE, "dyb_set_priors must be called before dyb_smc_create");
  const int nranks = h->comm ? h->comm->nranks : 1, rank = h->comm ? h->comm->rank : 0;
  if (nranks > SMC_MAX_RANKS) return fail(DYB_ERR_ARG, "at most 16 ranks");
  if (nranks > 1 && o->npart % ((long long)SMC_BLOCK * nranks) != 0)
    return fail(DYB_ERR_ARG, "with several ranks the particle count must be a multiple of 1024 * nranks");
  DYB_CUDA(cudaSetDevice(h->device));
  dyb_smc* s = new dyb_smc();
  s->h = h; s->npart = o->npart; s->nranks = nranks; s->rank = rank;
  s->n_local = o->npart / nranks; s->first = s->n_local * rank;
  s->np = h->hs.np; s->nphi = o->nphi; s->trgt = o->trgt; s->alp = o->alp; s->systematic = o->resampling; s->seed = o->seed;
  s->c0 = o->c0; s->acpt0 = o->acpt0;
  s->phi.resize((size_t)o->nphi);
  for (int i = 0; i < o->nphi; ++i) s->phi[i] = phi ? phi[i] : std::pow((double)i / (double)(o->nphi - 1), o->lam);   // smc.jl:48
  const char* eg = std::getenv("DYB_SMC_GRAPH");
  s->use_graph = !(eg && eg[0] == '0');
  const size_t N = (size_t)s->npart, n = (size_t)s->n_local, np = (size_t)s->np;
  const size_t nb = (N + SMC_BLOCK - 1) / SMC_BLOCK;
  const size_t pay = 8 * N * (np + 2);
  s->off_pay1 = (pay + 255) / 256 * 256;
  s->off_pub = 2 * s->off_pay1;
  s->off_cnt = s->off_pub + 8 * SMC_MAX_RANKS;
  s->slab_bytes = s->off_cnt + 8 * 2 * SMC_MAX_RANKS;
  cudaError_t e = cudaSuccess;
  auto R = [&](DevBuf& b, size_t bytes) { if (e == cudaSuccess) e = b.reserve(bytes + 16); };
  // the slab is a cudaMalloc allocation of its own (cudaIpcGetMemHandle covers whole allocations)
  if (cudaMalloc(&s->slab.p, s->slab_bytes) == cudaSuccess) s->slab.cap = s->slab_bytes; else e = cudaGetLastError();
  R(s->ctl, sizeof(SmcCtl)); R(s->phi_d, 8 * (size_t)o->nphi);
  R(s->w_g, 8 * N); R(s->wt_g, 8 * N); R(s->cwl, 8 * N); R(s->bt, 8 * nb); R(s->bt2, 8 * nb); R(s->btw, 8 * nb); R(s->offs, 8 * nb);
  R(s->idx, 8 * N); R(s->xsel, 8 * N * (np + 2)); R(s->part1, 8 * nb * np); R(s->part2, 8 * nb * np * np);
  R(s->mu, 8 * np); R(s->R, 8 * np * np); R(s->L, 8 * np * np); R(s->rdiag, 8 * np); R(s->aux, 32 + 16 * np * np);     // aux: 4 scalars + the 2 np^2 workspace of the PSD projection
  R(s->prop, 8 * np * n); R(s->qx, 8 * n); R(s->q0, 8 * n); R(s->lx, 8 * n); R(s->st, 4 * n); R(s->priorx, 8 * n);
  R(s->stats, 8 * SMC_STATS * (size_t)o->nphi);
  R(s->tmp_para, 8 * np * n); R(s->tmp_ll, 8 * n); R(s->tmp_lp, 8 * n);
  int rc = DYB_OK;
  if (e != cudaSuccess) rc = fail(DYB_ERR_CUDA, cudaGetErrorString(e));
  auto CK = [&](cudaError_t x) { if (rc == DYB_OK && x != cudaSuccess) rc = fail(DYB_ERR_CUDA, cudaGetErrorString(x)); };
  if (rc == DYB_OK) {
    CK(cudaMemsetAsync(s->slab.p, 0, s->slab_bytes, h->stream));
    CK(cudaMemsetAsync(s->ctl.p, 0, sizeof(SmcCtl), h->stream));
    CK(cudaMemsetAsync(s->stats.p, 0, 8 * SMC_STATS * (size_t)o->nphi, h->stream));
    CK(cudaMemsetAsync(s->aux.p, 0, 32, h->stream));
    CK(cudaMemcpyAsync(s->phi_d.p, s->phi.data(), 8 * (size_t)o->nphi, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  if (rc == DYB_OK && nranks > 1) {
    const char* et = std::getenv("DYB_SMC_TRANSPORT");
    if (!(et && std::strcmp(et, "nccl") == 0)) rc = smc_map_peers(s);
  }
  if (rc != DYB_OK) { s->release(); delete s; return rc; }
  h->n_smc += 1;
  *out = s;
  return DYB_OK;
}

int dyb_smc_destroy(dyb_smc* s) {
  if (!s) return DYB_OK;
  dyb_handle* h = s->h;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (s->p2p) {
    // nobody may still be storing into a slab that is about to go away
    if (smc_rank_barrier(s) == DYB_OK) cudaStreamSynchronize(h->stream);
    for (int r = 0; r < s->nranks; ++r) if (r != s->rank && s->peer_slab[r]) cudaIpcCloseMemHandle(s->peer_slab[r]);
  }
  s->release();
  h->n_smc -= 1;
  delete s;
  return DYB_OK;
}

int dyb_smc_local_range(const dyb_smc* s, int64_t* first, int64_t* n_local) {
  if (!s) return fail(DYB_ERR_ARG, "null argument");
  if (first) *first = s->first;
  if (n_local) *n_local = s->n_local;
  return DYB_OK;
}

// 1 when the mutated particles travel by peer stores (CUDA IPC over NVLink), 0 for the NCCL all-gather / single GPU
int dyb_smc_transport(const dyb_smc* s, int32_t* peer_stores, int32_t* cuda_graph) {
  if (!s) return fail(DYB_ERR_ARG, "null argument");
  if (peer_stores) *peer_stores = s->p2p ? 1 : 0;
  if (cuda_graph) *cuda_graph = (s->use_graph && !s->profiling) ? 1 : 0;
  return DYB_OK;
}

// prior density, packing into the payload, initial weights / control block / records: what follows the likelihood of the
// initial particles (which is in s->lx / s->st for the draws in s->prop)
static int smc_init_finish(dyb_smc* s, const PriorSetDev& ps) {
  dyb_handle* h = s->h;
  const long long n = s->n_local;
  cudaStream_t st = h->stream;
  logprior_kernel<<<nblk(n, 128), 128, 0, st>>>(ps, s->prop.as<double>(), n, s->priorx.as<double>());
  const SmcStageArgs a = stage_args(s);
  smc_init_pack_kernel<<<nblk(n, 256), 256, 0, st>>>(a, s->prop.as<double>(), s->lx.as<double>(), s->priorx.as<double>(), s->phi[0]);
  smc_fill_kernel<<<nblk(s->npart, 256), 256, 0, st>>>(s->w_g.as<double>(), s->npart, 1.0 / (double)s->npart);
  SmcCtl c0{};
  c0.stage = 1; c0.tune = SmcTune{s->c0, s->acpt0};
  DYB_CUDA(cudaMemcpyAsync(s->ctl.p, &c0, sizeof(c0), cudaMemcpyHostToDevice, st));
  const double st0[SMC_STATS] = {1.0, (double)s->npart, 0.0, 0.0, 0.0, 0.0};
  DYB_CUDA(cudaMemcpyAsync(s->stats.p, st0, sizeof(st0), cudaMemcpyHostToDevice, st));
  DYB_CUDA(cudaGetLastError());
  h->launches += 3;
  return DYB_OK;
}

// initial particles (smc.jl:84-101): the caller draws them from the prior; the likelihood, the prior density and
// logpost = phi[1] * loglh + logprior are evaluated here.  status[i] != 0 marks draws the reference would reject
// (obj = -Inf): replace them and call again.  Local to the rank (no collective).
int dyb_smc_init(dyb_smc* s, const double* para_local, int32_t* status) {
  if (!s || !para_local) return fail(DYB_ERR_ARG, "null argument");
  dyb_handle* h = s->h;
  DYB_CUDA(cudaSetDevice(h->device));
  PriorSetDev ps;
  int rc = prior_set(h, &ps);
  if (rc) return rc;
  const long long n = s->n_local;
  cudaStream_t st = h->stream;
  DYB_CUDA(cudaMemcpyAsync(s->prop.p, para_local, 8 * (size_t)s->np * n, cudaMemcpyDefault, st));
  rc = dyb_loglik_batch_device(h, s->prop.as<double>(), n, s->lx.as<double>(), s->st.as<int>());
  if (rc) return rc;
  rc = smc_init_finish(s, ps);
  if (rc) return rc;
  if (status) DYB_CUDA(cudaMemcpyAsync(status, s->st.p, 4 * (size_t)n, cudaMemcpyDefault, st));
  DYB_CUDA(cudaStreamSynchronize(st));
  s->stage = 1;
  s->published = false;
  return DYB_OK;
}

// the same with the prior draws made on the device (prior_draw!, estimation.jl:329-335, inside the rejection loop of
// smc.jl:87-101): round 0 draws every local particle, round r > 0 re-draws the ones whose likelihood was not finite; the host
// only reads one counter per round.  The draw of global particle j in round r depends on (seed, j, r) alone, so the initial
// cloud does not depend on the number of GPUs.
int dyb_smc_init_from_prior(dyb_smc* s, uint64_t seed, int32_t max_rounds, int64_t* n_draws) {
  if (!s || max_rounds < 1) return fail(DYB_ERR_ARG, "bad argument");
  dyb_handle* h = s->h;
  DYB_CUDA(cudaSetDevice(h->device));
  PriorSetDev ps;
  int rc = prior_set(h, &ps);
  if (rc) return rc;
  const long long n = s->n_local;
  cudaStream_t st = h->stream;
  unsigned long long* cnt = s->idx.as<unsigned long long>();          // free until the first selection step
  DYB_CUDA(cudaMemsetAsync(cnt, 0, 8, st));
  unsigned long long total = 0;
  bool done = false;
  for (int r = 0; r < max_rounds && !done; ++r) {
    prior_draw_kernel<<<nblk(n, 128), 128, 0, st>>>(ps, seed, s->first, n, (unsigned)r, r == 0 ? nullptr : s->st.as<int>(),
                                                     s->prop.as<double>(), cnt);
    DYB_CUDA(cudaGetLastError());
    h->launches += 1;
    unsigned long long now = 0;
    DYB_CUDA(cudaMemcpyAsync(&now, cnt, 8, cudaMemcpyDeviceToHost, st));
    DYB_CUDA(cudaStreamSynchronize(st));
    if (now == total) { done = true; break; }                         // nothing was re-drawn: every particle is finite
    total = now;
    rc = dyb_loglik_batch_device(h, s->prop.as<double>(), n, s->lx.as<double>(), s->st.as<int>());
    if (rc) return rc;
  }
  if (n_draws) *n_draws = (int64_t)total;
  if (!done) return fail(DYB_ERR_STATE, "could not draw initial particles with a finite likelihood within max_rounds");
  rc = smc_init_finish(s, ps);
  if (rc) return rc;
  DYB_CUDA(cudaStreamSynchronize(st));
  s->stage = 1;
  s->published = false;
  return DYB_OK;
}

// prior draws on their own: para np x n (draw-major), draw i = particle first + i of round `round`
int dyb_prior_draw(dyb_handle* h, int64_t n, int64_t first, uint64_t seed, int32_t round, double* para) {
  if (!h || n < 0 || round < 0 || (n > 0 && !para)) return fail(DYB_ERR_ARG, "bad argument");
  if (n == 0) return DYB_OK;
  DYB_CUDA(cudaSetDevice(h->device));
  PriorSetDev ps;
  int rc = prior_set(h, &ps);
  if (rc) return rc;
  Stager sg(h->stream);
  double* d = sg.out(para, (size_t)n * ps.np);
  if (sg.err != cudaSuccess) return fail(DYB_ERR_CUDA, cudaGetErrorString(sg.err));
  prior_draw_kernel<<<nblk(n, 128), 128, 0, h->stream>>>(ps, seed, first, n, (unsigned)round, nullptr, d, nullptr);
  DYB_CUDA(cudaGetLastError());
  h->launches += 1;
  DYB_CUDA(sg.finish());
  return DYB_OK;
}

// in-place all-gather of the rows each rank wrote into payload buffer `par` of its own slab, and of the acceptance counts
static int smc_nccl_exchange(dyb_smc* s, int par, bool counts) {
  dyb_handle* h = s->h;
  NcclApi* api = nccl_api();
  double* pay = reinterpret_cast<double*>(s->slab.as<char>() + (par ? s->off_pay1 : 0));
  const size_t chunk = (size_t)s->n_local * (s->np + 2);
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(s->slab.as<char>() + s->off_cnt) + par * SMC_MAX_RANKS;
  int rc = DYB_OK;
  ncclResult_t r0 = api->GroupStart();
  ncclResult_t r1 = api->AllGather(pay + chunk * s->rank, pay, chunk, ncclDouble, h->comm->comm, h->stream);
  ncclResult_t r2 = counts ? api->AllGather(cnt + s->rank, cnt, 1, ncclUint64, h->comm->comm, h->stream) : ncclSuccess;
  ncclResult_t r3 = api->GroupEnd();                   // always closed, whatever happened in between
  for (ncclResult_t r : {r0, r1, r2, r3})
    if (r != ncclSuccess && rc == DYB_OK) rc = fail(DYB_ERR_NCCL, std::string("SMC particle exchange: ") + api->GetErrorString(r));
  return rc;
}

// the launch sequence of ONE tempering stage (smc.jl:110-243); every kernel reads the stage from the control block.
// `par` = parity of the stage (only the NCCL exchange needs it on the host).  tev: 7 events or null.
static int smc_stage_launch(dyb_smc* s, int par, cudaEvent_t* tev) {
  dyb_handle* h = s->h;
  cudaStream_t st = h->stream;
  const long long N = s->npart, n = s->n_local;
  const int np = s->np;
  const unsigned nb = (unsigned)((N + SMC_BLOCK - 1) / SMC_BLOCK);
  PriorSetDev ps;
  int rc = prior_set(h, &ps);
  if (rc) return rc;
  const SmcStageArgs a = stage_args(s);
#define DYB_MARK(k) do { if (tev) DYB_CUDA(cudaEventRecord(tev[k], st)); } while (0)
  DYB_MARK(0);
  smc_w1_kernel<<<nb, 256, 0, st>>>(a);
  smc_w2_kernel<<<nb, 256, 0, st>>>(a);
  DYB_MARK(1);
  smc_select_take_kernel<<<nblk(N, 256), 256, 0, st>>>(a);
  DYB_MARK(2);
  {
    // particle mean, then covariance + Cholesky: one wave of CTAs each, a block of 1024 rows staged through shared memory
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    const size_t staged = smc_moment2_smem_bytes(np);
    const bool fits = staged + 1024 <= (size_t)max_smem;
    const size_t smem = fits ? staged : sizeof(double) * np * np;
    const dim3 blk(32, SMC_MOM_WARPS);
    auto groups_for = [&](int ncomp) {
      unsigned g = 148 / nb, maxg = (unsigned)((ncomp + SMC_MOM_WARPS - 1) / SMC_MOM_WARPS);
      if (!fits) g = maxg;
      if (g > maxg) g = maxg;
      return g < 1 ? 1u : g;
    };
    if (fits) {
      DYB_CUDA(cudaFuncSetAttribute(smc_moments_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      DYB_CUDA(cudaFuncSetAttribute(smc_moments_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smc_moments_kernel<false, true><<<dim3(nb, groups_for(np)), blk, smem, st>>>(a);
      smc_moments_kernel<true, true><<<dim3(nb, groups_for(np * np)), blk, smem, st>>>(a);
    } else {
      smc_moments_kernel<false, false><<<dim3(nb, groups_for(np)), blk, smem, st>>>(a);
      smc_moments_kernel<true, false><<<dim3(nb, groups_for(np * np)), blk, smem, st>>>(a);
    }
  }
  DYB_MARK(3);
  smc_propose2_kernel<<<nblk(n, 128), 128, sizeof(double) * (np * np + 2 * np), st>>>(a);
  DYB_CUDA(cudaGetLastError());
  DYB_MARK(4);
  rc = dyb_loglik_batch_device(h, s->prop.as<double>(), n, s->lx.as<double>(), s->st.as<int>());
  if (rc) return rc;
  DYB_MARK(5);
  smc_accept_publish_kernel<<<nblk(n, SMC_ACC_BLOCK), SMC_ACC_BLOCK, sizeof(double) * SMC_ACC_BLOCK * (np + 2), st>>>(a, ps);
  DYB_CUDA(cudaGetLastError());
  if (s->nranks > 1 && !s->p2p) { rc = smc_nccl_exchange(s, par, true); if (rc) return rc; }
  DYB_MARK(6);
#undef DYB_MARK
  h->launches += 7;
  return DYB_OK;
}

// first run after dyb_smc_init: the initial payload reaches every rank
static int smc_publish_initial(dyb_smc* s) {
  dyb_handle* h = s->h;
  s->epoch += (unsigned long long)s->nphi + 2ULL;
  smc_set_epoch_kernel<<<1, 1, 0, h->stream>>>(s->ctl.as<SmcCtl>(), s->epoch);
  if (s->nranks > 1) {
    int rc = smc_rank_barrier(s);            // every rank is done with the previous contents of the slabs
    if (rc) return rc;
    if (s->p2p) {
      const SmcStageArgs a = stage_args(s);
      const long long nel = s->n_local * (s->np + 2);
      unsigned g = nblk(nel, 256); if (g > 296) g = 296;
      smc_publish_init_kernel<<<g, 256, 0, h->stream>>>(a);
    } else {
      rc = smc_nccl_exchange(s, 0, false);
      if (rc) return rc;
    }
  }
  DYB_CUDA(cudaGetLastError());
  s->published = true;
  return DYB_OK;
}

int dyb_smc_run(dyb_smc* s, int32_t n_stages) {
  if (!s || n_stages < 0) return fail(DYB_ERR_ARG, "bad argument");
  if (s->stage < 1) return fail(DYB_ERR_STATE, "dyb_smc_init has not been called");
  dyb_handle* h = s->h;
  DYB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  int rc = DYB_OK;
  if (!s->published) { rc = smc_publish_initial(s); if (rc) return rc; }
  bool ran = false;
  for (int k = 0; k < n_stages && s->stage < s->nphi; ++k) {
    const int par = s->stage & 1;
    if (s->profiling) {
      const size_t base = s->tev.size();
      for (int q = 0; q <= SMC_TSEG; ++q) { cudaEvent_t e; DYB_CUDA(cudaEventCreate(&e)); s->tev.push_back(e); }
      rc = smc_stage_launch(s, par, &s->tev[base]);
      if (rc) return rc;
    } else if (s->use_graph && k > 0) {
      // stage 1 of a call runs as plain launches (it also sizes the likelihood buffers); later ones replay a graph
      if (!s->graph[par] || s->graph_cfg[par] != h->config_epoch) {
        if (s->graph[par]) { cudaGraphExecDestroy(s->graph[par]); s->graph[par] = nullptr; }
        cudaGraph_t g = nullptr;
        const int64_t launches0 = h->launches;
        DYB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        rc = smc_stage_launch(s, par, nullptr);
        const cudaError_t ce = cudaStreamEndCapture(st, &g);
        h->launches = launches0;
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (ce != cudaSuccess || !g) { cudaGetLastError(); s->use_graph = false; }
        else {
          const cudaError_t ie = cudaGraphInstantiate(&s->graph[par], g, 0);
          cudaGraphDestroy(g);
          if (ie != cudaSuccess) { cudaGetLastError(); s->graph[par] = nullptr; s->use_graph = false; }
          else s->graph_cfg[par] = h->config_epoch;
        }
      }
      if (s->use_graph) {
        DYB_CUDA(cudaGraphLaunch(s->graph[par], st));
        h->launches += 7 + 1 + h->ops->solve_launches(h->hs);
      } else {
        rc = smc_stage_launch(s, par, nullptr);
        if (rc) return rc;
      }
    } else {
      rc = smc_stage_launch(s, par, nullptr);
      if (rc) return rc;
    }
    s->stage += 1;
    ran = true;
  }
  if (ran) {
    smc_finalize_kernel<<<1, 32, 0, st>>>(stage_args(s));
    DYB_CUDA(cudaGetLastError());
    h->launches += 1;
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
  if (s->stage < 1) return fail(DYB_ERR_STATE, "dyb_smc_init has not been called");
  dyb_handle* h = s->h;
  DYB_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)s->n_local;
  const double* pay = reinterpret_cast<const double*>(s->slab.as<char>() + (((s->stage - 1) & 1) ? s->off_pay1 : 0));
  smc_unpack_rows_kernel<<<nblk(s->n_local, 256), 256, 0, h->stream>>>(pay, s->first, s->n_local, s->np, s->tmp_para.as<double>(),
                                                                     s->tmp_ll.as<double>(), s->tmp_lp.as<double>());
  DYB_CUDA(cudaGetLastError());
  if (para) DYB_CUDA(cudaMemcpyAsync(para, s->tmp_para.p, 8 * n * s->np, cudaMemcpyDefault, h->stream));
  if (w) DYB_CUDA(cudaMemcpyAsync(w, s->w_g.as<double>() + s->first, 8 * n, cudaMemcpyDefault, h->stream));
  if (loglh) DYB_CUDA(cudaMemcpyAsync(loglh, s->tmp_ll.p, 8 * n, cudaMemcpyDefault, h->stream));
  if (logpost) DYB_CUDA(cudaMemcpyAsync(logpost, s->tmp_lp.p, 8 * n, cudaMemcpyDefault, h->stream));
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
  SmcCtl ctl{};
  DYB_CUDA(cudaMemcpyAsync(hs.data(), s->stats.p, 8 * hs.size(), cudaMemcpyDeviceToHost, h->stream));
  DYB_CUDA(cudaMemcpyAsync(&ctl, s->ctl.p, sizeof(ctl), cudaMemcpyDeviceToHost, h->stream));
  if (mu) DYB_CUDA(cudaMemcpyAsync(mu, s->mu.p, 8 * (size_t)s->np, cudaMemcpyDefault, h->stream));
  if (R) DYB_CUDA(cudaMemcpyAsync(R, s->R.p, 8 * (size_t)s->np * s->np, cudaMemcpyDefault, h->stream));
  DYB_CUDA(cudaStreamSynchronize(h->stream));
  if (stats) std::memcpy(stats, hs.data(), 8 * hs.size());
  if (logmdd) {
    double a = 0.0;
    for (int i = 1; i < s->stage; ++i) a += std::log(hs[(size_t)SMC_STATS * i]);
    *logmdd = a;
  }
  if (ctl.err) return fail(DYB_ERR_STATE, "the particle covariance was not positive definite at some stage (the reference's cholesky throws)");
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
