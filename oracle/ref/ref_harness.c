/* ref_harness.c -- storage for the module variables / grid(:) / gbuffers(:) of the translated reference routines (one copy
 * per thread = per emulated rank), setters for the Python driver, and the unit-100 history.  TEST INFRASTRUCTURE (oracle/). */
#include "ref_prelude.h"
#include "ref_types.h"

__thread grid_type GRID_[REF_MAXLEV];
__thread mpi_buffers GBUFFERS_[REF_MAXLEV];
__thread int g_myrank = 0, g_nprocs = 1, g_bmask = 0, g_netcdf_output = 0, g_aggressive = 0, g_nlevs = 0, g_ns_pre = 3, g_ns_post = 2,
             g_ns_coarsest = 40, g_nsmall = 8, g_solver_maxiter = 50;
__thread double g_hlim = 0.0, g_nhtheta_b = 0.0, g_nhtheta_s = 0.0, g_solver_prec = 1e-12;
__thread const char *g_cmatrix = "real", *g_relax_method = "RB", *g_interp_type = "linear";

static __thread double hist_res[256], hist_conv[256];
static __thread int hist_n = 0;
void cb_history(double res, double conv) { if (hist_n < 256) { hist_res[hist_n] = res; hist_conv[hist_n] = conv; ++hist_n; } }
void ref_history_reset(void) { hist_n = 0; }
int ref_history(double* res, double* conv, int cap) { int n = hist_n < cap ? hist_n : cap; for (int q = 0; q < n; ++q) { res[q] = hist_res[q]; conv[q] = hist_conv[q]; } return hist_n; }

grid_type* ref_grid(int lev) { return &GRID_[lev]; }
void ref_reset_state(void) { memset(GRID_, 0, sizeof(GRID_)); memset(GBUFFERS_, 0, sizeof(GBUFFERS_)); hist_n = 0; }   /* (arrays of a previous case are leaked: test process) */
int ref_set_int(const char* name, int v) {
#define S(x) if (!strcmp(name, #x)) { g_##x = v; return 0; }
  S(myrank) S(nprocs) S(bmask) S(netcdf_output) S(aggressive) S(nlevs) S(ns_pre) S(ns_post) S(ns_coarsest) S(nsmall) S(solver_maxiter)
#undef S
  return 1;
}
int ref_get_int(const char* name) {
#define S(x) if (!strcmp(name, #x)) return g_##x;
  S(myrank) S(nprocs) S(bmask) S(nlevs) S(nsmall)
#undef S
  return -999;
}
int ref_set_real(const char* name, double v) {
#define S(x) if (!strcmp(name, #x)) { g_##x = v; return 0; }
  S(hlim) S(nhtheta_b) S(nhtheta_s) S(solver_prec)
#undef S
  return 1;
}
int ref_set_str(const char* name, const char* value) {
  char* v = strdup(value);
  if (!strcmp(name, "cmatrix")) g_cmatrix = v; else if (!strcmp(name, "relax_method")) g_relax_method = v;
  else if (!strcmp(name, "interp_type")) g_interp_type = v; else return 1;
  return 0;
}
