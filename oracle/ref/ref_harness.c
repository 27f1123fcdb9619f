/* ref_harness.c -- storage for the module variables / grid(:) of the translated reference routines and the call-back
 * trampolines (the halo fills are supplied by oracle/refdriver.py).  TEST INFRASTRUCTURE (oracle/). */
#include "ref_prelude.h"

grid_t GRID_[REF_MAXLEV];
int g_myrank = 0, g_nprocs = 1, g_bmask = 0, g_netcdf_output = 0, g_aggressive = 0, g_nlevs = 0, g_ns_pre = 3, g_ns_post = 2, g_ns_coarsest = 40,
    g_mpi_proc_null = -2, g_nsmall = 8;
double g_hlim = 0.0, g_nhtheta_b = 0.0, g_nhtheta_s = 0.0;
const char *g_cmatrix = "real", *g_relax_method = "RB", *g_interp_type = "linear";

typedef void (*halo_relax_fn)(int, double*, int, int, int);
typedef void (*halo_fn)(int, farr, const char*);
static halo_relax_fn h_relax = 0;
static halo_fn h_any = 0;
void ref_set_callbacks(halo_relax_fn a, halo_fn b) { h_relax = a; h_any = b; }
void cb_fill_halo_relax(int lev, double* p, int nx, int ny, int nz) { if (!h_relax) abort(); h_relax(lev, p, nx, ny, nz); }
void cb_fill_halo(int lev, farr a, const char* lbc_null) { if (!h_any) abort(); h_any(lev, a, lbc_null); }

/* one rank: the global sum is the local one (src/mg_mpi_exchange.f90:1555-1571), nothing is ever gathered */
void cb_global_sum(int lev, double local, double* global) { (void)lev; *global = local; }
void cb_gather(int lev, farr x, farr y) { (void)lev; (void)x; (void)y; abort(); }
void cb_split(int lev, farr x, farr y) { (void)lev; (void)x; (void)y; abort(); }
static double hist_res[256], hist_conv[256]; static int hist_n = 0;
void cb_history(double res, double conv) { if (hist_n < 256) { hist_res[hist_n] = res; hist_conv[hist_n] = conv; ++hist_n; } }
void ref_history_reset(void) { hist_n = 0; }
int ref_history(double* res, double* conv, int cap) { int n = hist_n < cap ? hist_n : cap; for (int q = 0; q < n; ++q) { res[q] = hist_res[q]; conv[q] = hist_conv[q]; } return hist_n; }
grid_t* ref_grid(int lev) { return &GRID_[lev]; }
void ref_set_str(const char* name, const char* value) {
  char* v = strdup(value);
  if (!strcmp(name, "cmatrix")) g_cmatrix = v; else if (!strcmp(name, "relax_method")) g_relax_method = v;
  else if (!strcmp(name, "interp_type")) g_interp_type = v; else abort();
}
