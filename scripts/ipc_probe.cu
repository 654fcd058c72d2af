// Probe: can two processes (one per GPU) on this box exchange data through CUDA IPC peer mappings with a
// flag-based hand-shake from inside kernels?  (What the fused SMC exchange in csrc/peer_exchange.cuh relies on.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o ipc_probe scripts/ipc_probe.cu && ./ipc_probe [nranks]
#include <cuda_runtime.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) { fprintf(stderr, "[%d] %s: %s\n", g_rank, #x, cudaGetErrorString(e_)); exit(2); } \
  } while (0)
static int g_rank = -1;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

struct Peers { double* data[8]; unsigned long long* flag[8]; };

// every rank writes its chunk into all peers' buffers, then raises flag[rank] on every peer (last CTA), then a second
// kernel waits for all flags
__global__ void publish(Peers P, int rank, int nranks, int n, unsigned long long epoch, unsigned int* ticket) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double v = (double)(rank * 1000000 + i) + (double)epoch;
    for (int r = 0; r < nranks; ++r) P.data[r][(size_t)rank * n + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      *ticket = 0;
      __threadfence_system();
      for (int r = 0; r < nranks; ++r) st_release_sys(P.flag[r] + rank, epoch);
    }
  }
}
__global__ void wait_all(const unsigned long long* flags, int nranks, unsigned long long epoch) {
  if (threadIdx.x < nranks) while (ld_acquire_sys(flags + threadIdx.x) < epoch) {}
}
__global__ void check(const double* data, int nranks, int n, unsigned long long epoch, int* bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nranks * n) return;
  const int r = i / n, k = i % n;
  if (data[i] != (double)(r * 1000000 + k) + (double)epoch) atomicAdd(bad, 1);
}

int main(int argc, char** argv) {
  const int nranks = argc > 1 ? atoi(argv[1]) : 2;
  const int n = 4096;
  // pipes: to[r][s] carries rank r's handles to rank s
  std::vector<int> rd(nranks * nranks), wr(nranks * nranks);
  for (int i = 0; i < nranks * nranks; ++i) { int p[2]; if (pipe(p)) return 3; rd[i] = p[0]; wr[i] = p[1]; }
  std::vector<pid_t> kids;
  for (int r = 0; r < nranks; ++r) {
    pid_t pid = fork();
    if (pid == 0) { g_rank = r; break; }
    kids.push_back(pid);
  }
  if (g_rank < 0) {
    int rc = 0;
    for (pid_t k : kids) { int st; waitpid(k, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st)) rc = 1; }
    printf(rc ? "ipc_probe FAILED\n" : "ipc_probe ok\n");
    return rc;
  }
  const int rank = g_rank;
  CK(cudaSetDevice(rank));
  double* data; unsigned long long* flag; unsigned int* ticket; int* bad;
  CK(cudaMalloc(&data, sizeof(double) * n * nranks));
  CK(cudaMalloc(&flag, 8 * 8));
  CK(cudaMalloc(&ticket, 4)); CK(cudaMalloc(&bad, 4));
  CK(cudaMemset(flag, 0, 64)); CK(cudaMemset(ticket, 0, 4)); CK(cudaMemset(bad, 0, 4));
  CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t hd, hf;
  CK(cudaIpcGetMemHandle(&hd, data)); CK(cudaIpcGetMemHandle(&hf, flag));
  for (int s = 0; s < nranks; ++s) {
    if (s == rank) continue;
    if (write(wr[rank * nranks + s], &hd, sizeof(hd)) != sizeof(hd)) return 4;
    if (write(wr[rank * nranks + s], &hf, sizeof(hf)) != sizeof(hf)) return 4;
  }
  Peers P;
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) { P.data[r] = data; P.flag[r] = flag; continue; }
    cudaIpcMemHandle_t a, b;
    if (read(rd[r * nranks + rank], &a, sizeof(a)) != sizeof(a)) return 5;
    if (read(rd[r * nranks + rank], &b, sizeof(b)) != sizeof(b)) return 5;
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, rank, r));
    if (!can) { fprintf(stderr, "[%d] no peer access to %d\n", rank, r); return 6; }
    CK(cudaIpcOpenMemHandle((void**)&P.data[r], a, cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle((void**)&P.flag[r], b, cudaIpcMemLazyEnablePeerAccess));
  }
  cudaStream_t st; CK(cudaStreamCreate(&st));
  const int iters = 2000;
  // plain launches
  auto t0 = std::chrono::steady_clock::now();
  for (int e = 1; e <= iters; ++e) {
    publish<<<(n + 255) / 256, 256, 0, st>>>(P, rank, nranks, n, e, ticket);
    wait_all<<<1, 32, 0, st>>>(flag, nranks, e);
  }
  CK(cudaStreamSynchronize(st));
  auto t1 = std::chrono::steady_clock::now();
  check<<<(n * nranks + 255) / 256, 256, 0, st>>>(data, nranks, n, iters, bad);
  int hb = -1;
  CK(cudaMemcpyAsync(&hb, bad, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const double us = std::chrono::duration<double, std::micro>(t1 - t0).count() / iters;
  // same inside a CUDA graph (100 exchanges per graph)
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int e = 1; e <= 100; ++e) {
    publish<<<(n + 255) / 256, 256, 0, st>>>(P, rank, nranks, n, iters + e, ticket);
    wait_all<<<1, 32, 0, st>>>(flag, nranks, iters + e);
  }
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  auto t2 = std::chrono::steady_clock::now();
  CK(cudaGraphLaunch(ge, st));
  CK(cudaStreamSynchronize(st));
  auto t3 = std::chrono::steady_clock::now();
  const double usg = std::chrono::duration<double, std::micro>(t3 - t2).count() / 100;
  printf("[rank %d/%d] bad=%d  exchange of %d doubles to every peer + flag wait: %.2f us (launches), %.2f us (graph)\n",
         rank, nranks, hb, n, us, usg);
  // leave peers mapped until everyone is done: crude barrier through the pipes
  char c = 1;
  for (int s = 0; s < nranks; ++s) if (s != rank && write(wr[rank * nranks + s], &c, 1) != 1) return 7;
  for (int r = 0; r < nranks; ++r) if (r != rank && read(rd[r * nranks + rank], &c, 1) != 1) return 7;
  return hb == 0 ? 0 : 1;
}
