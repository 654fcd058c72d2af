// Host-side registry of compiled models: one ModelOps per generated model header.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>
#include "../../include/dynare_b200.h"
#include "model_traits.cuh"

namespace dyb {

// everything the solve kernel needs from the host besides theta
// device-side cache of the cooperative solver's fast pass (solve_coop.cuh): the pivot order of the equations found at the
// calibration point, re-derived when params / sigma_e / sigma_m change (key), and the count of draws the fast pass handed
// to the Householder pass.  Owned by the handle (released in dyb_destroy); never copied.
struct CrPermCache {
  int* d_perm = nullptr;                // MM ints
  double* d_mid1 = nullptr;             // mid buffer of the one calibration draw
  int* d_st1 = nullptr;
  unsigned long long* d_redo = nullptr; // [0] draws redone by the Householder pass since creation, [1] those of the current call
  long long* d_list = nullptr;          // the current call's draws for the Householder pass, densely packed (list_cap entries)
  long long list_cap = 0;
  unsigned long long key = 0;
  bool valid = false;
  void release() {
    if (d_perm) cudaFree(d_perm); if (d_mid1) cudaFree(d_mid1); if (d_st1) cudaFree(d_st1); if (d_redo) cudaFree(d_redo);
    if (d_list) cudaFree(d_list);
    d_perm = nullptr; d_mid1 = nullptr; d_st1 = nullptr; d_redo = nullptr; d_list = nullptr; list_cap = 0; valid = false;
  }
};

struct HostState {
  std::vector<double> params, sigma_e, sigma_m;
  int np = 0;
  int est_type[MAX_EST], est_i[MAX_EST], est_j[MAX_EST];
  dyb_options opt;
  mutable CrPermCache crp;
  HostState() = default;
  HostState(const HostState&) = delete;
  HostState& operator=(const HostState&) = delete;
};

struct SolveOut {
  double* ss;        // SoA hand-off buffer, ss_doubles x N
  int* status;
  double* g1;        // optional AoS outputs (device), may be null
  double* ybar;
  int* cr_iters;
  int* lyap_iters;
  double* mid;       // scratch for the cooperative solver, mid_doubles x N (device)
  cudaEvent_t ev_after_model = nullptr, ev_after_cr = nullptr;   // optional timing marks between the solver's kernels
};

struct ModelOps {
  const char* name;
  dyb_model_dims dims;
  int mid_doubles;   // doubles per draw of scratch between the kernels of the cooperative solver
  void (*defaults)(HostState&);
  int (*solve_launches)(const HostState&);   // kernels one solve() call launches
  cudaError_t (*solve)(const HostState&, const double* d_theta, long long n, SolveOut out, cudaStream_t);
  // lanes: dyb_options.filter_lanes (0 auto, 1, 4, 8)
  cudaError_t (*kalman)(const double* d_ss, long long n, const double* d_Y, int nobs, int presample, int lanes,
                        const int* d_status_in, double* d_loglik, int* d_status_out, cudaStream_t);
  // expand the SoA hand-off buffer into dense per-draw matrices (any output may be null)
  cudaError_t (*unpack)(const double* d_ss, long long n, double* T, double* RQR, double* P0,
                        double* ybar_obs, double* H, cudaStream_t);
  void (*bkwrd_indices)(int* i_bkwrd_b /* dims.k */);   // variables with a lag, 0-based (Model.i_bkwrd_b)
  void (*obs_indices)(int* obs_idx /* dims.ny */);      // observed variable of each row of Y, 0-based
  void (*obs_state_positions)(int* z_idx /* dims.ny */);   // position of each observable in the Kalman state (SSWs.obs_idx_state)
  // ---- models whose dynamic block does not fit the register-resident solver (m > 12): the generated model function
  // feeds the run-time-size solver and the dense filter (generic.cu) on the device, no host hop
  bool register_solver;                                 // the three-kernel register-resident path exists for this model
  void (*static_indices)(int* i_static /* dims.s */);
  void (*fwrd_indices)(int* i_fwrd_b /* dims.nf */);
  // theta -> per draw: compressed Jacobian n x ncol (column-major), steady state n, Sigma_e nx x nx, Sigma_m ny x ny, status
  cudaError_t (*jacobian_batch)(const HostState&, const double* d_theta, long long n, double* d_jac, double* d_ybar,
                                double* d_sigma_e, double* d_sigma_m, int* d_status, cudaStream_t);
  // ---- "sizes-only" models (modelgen.emit_cuda_jacobian_fed): no model function, the HOST evaluates steady state and
  // compressed Jacobian per draw and the register-resident kernels, compiled for the model's sizes and index sets, do the
  // rest.  Null for models with a generated model function.  Inputs as dyb_loglik_from_jacobian_batch (device pointers).
  cudaError_t (*solve_from_jacobian)(const HostState&, const double* d_jac, const double* d_ybar, const double* d_sigma_e,
                                     int sigma_e_per_draw, const double* d_sigma_m, int sigma_m_per_draw,
                                     const int* d_status_in, long long n, SolveOut out, cudaStream_t);
};

const std::vector<const ModelOps*>& registry();
void register_model(const ModelOps*);

}  // namespace dyb
