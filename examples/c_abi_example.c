/* Plain-C use of libdynare_b200.so -- the same calls a Julia `ccall` shim makes (julia/DynareB200.jl).
 *   gcc -std=c11 -I include examples/c_abi_example.c -L dynare.jl_b200 -l:libdynare_b200.so -Wl,-rpath,$PWD/dynare.jl_b200 -lm -o c_abi_example
 *   ./c_abi_example            (needs a CUDA device)
 * Evaluates the fs2000 log-likelihood at the reference's estimated_params_init point
 * (test/models/estimation/fs2000.mod:81-91) and at a perturbed draw, on synthetic observations. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "dynare_b200.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (call);                                                                \
    if (rc_ != DYB_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, dyb_last_error()); return 1; } \
  } while (0)

int main(void) {
  int ndev = 0;
  if (dyb_device_count(&ndev) != DYB_OK || ndev == 0) { printf("no CUDA device: nothing to run\n"); return 0; }
  printf("libdynare_b200 version %d, %d model(s):", dyb_version(), dyb_model_count());
  for (int i = 0; i < dyb_model_count(); ++i) printf(" %s", dyb_model_name(i));
  printf("\n");

  dyb_handle* h = NULL;
  CHECK(dyb_create("fs2000", 0, &h));
  dyb_model_dims d;
  CHECK(dyb_get_dims(h, &d));
  const int nobs = 192;
  double* Y = malloc(sizeof(double) * d.ny * nobs);
  unsigned s = 12345u;                              /* any data will do for the example */
  for (int i = 0; i < d.ny * nobs; ++i) { s = s * 1664525u + 1013904223u; Y[i] = 0.005 + 0.01 * ((s >> 8) / 16777216.0 - 0.5); }
  CHECK(dyb_set_observations(h, Y, d.ny, nobs));

  enum { N = 4 };
  const double mode[9] = {0.4186, 0.9901, 0.0038, 1.0141, 0.8623, 0.6837, 0.0020, 0.0133, 0.0029};
  double theta[9 * N], ll[N];
  int32_t st[N];
  for (int j = 0; j < N; ++j)
    for (int k = 0; k < 9; ++k) theta[k + 9 * j] = mode[k] * (1.0 + 0.01 * j);
  theta[1 + 9 * 3] = 1.5;                           /* bet > 1: no steady state -> status 1, loglik = -Inf */
  CHECK(dyb_loglik_batch(h, theta, N, ll, st));
  for (int j = 0; j < N; ++j) printf("draw %d: status %d loglik %.10g\n", j, st[j], ll[j]);
  if (st[0] != DYB_ST_OK || !isfinite(ll[0]) || st[3] == DYB_ST_OK || !isinf(ll[3])) { fprintf(stderr, "unexpected result\n"); return 1; }

  /* the scalar call of the reference is the batch of one */
  double ll1; int32_t st1;
  CHECK(dyb_loglik_batch(h, theta, 1, &ll1, &st1));
  if (ll1 != ll[0]) { fprintf(stderr, "batch-size dependence\n"); return 1; }
  printf("kernels launched: %lld\n", (long long)dyb_kernel_launches(h));
  CHECK(dyb_destroy(h));
  free(Y);
  printf("ok\n");
  return 0;
}
