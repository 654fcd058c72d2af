/* TEST INFRASTRUCTURE (CPU oracle port) -- never linked into the product library. */
#ifndef ORACLE_MODEL_H
#define ORACLE_MODEL_H
typedef struct oracle_model {
  const char* name;
  int n, nx, npar, k, nf, s, ncol, ns, ny, nl, np_default;
  const int *i_bkwrd_b, *i_fwrd_b, *i_static, *i_dyn, *state_ids, *obs_idx_state, *lagged_state_ids;
  const int *est_type, *est_i, *est_j;
  const double *params0, *sigma_e0, *sigma_m0;
  void (*steady_state)(const double* p, double* y);
  double (*static_resid_ss)(const double* p, const double* y);
  void (*jacobian)(const double* p, const double* y, double* J);   /* compressed, n x ncol col-major, zeroed by caller */
} oracle_model;
#endif
