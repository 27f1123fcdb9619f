/* ref_prelude.h -- run-time support for the C text that oracle/f90toc.py generates from the reference's Fortran sources.
 * TEST INFRASTRUCTURE (oracle/): only tests/, tests/golden/make_ref_golden.py and __graft_entry__.build() use it.
 * Arrays are descriptors (base pointer, lower bounds, extents) so that every Fortran array keeps its own index origin
 * (src/mg_grids.f90:170-260: p(nz,0:ny+1,0:nx+1), zr(nz,-1:ny+2,-1:nx+2), ...); column-major like Fortran. */
#ifndef MGB_REF_PRELUDE_H
#define MGB_REF_PRELUDE_H
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double* a; int lb[4]; int n[4]; } farr;

#define FA1(d, i1) (d).a[(long)((i1) - (d).lb[0])]
#define FA2(d, i1, i2) (d).a[(long)((i1) - (d).lb[0]) + (long)(d).n[0] * (long)((i2) - (d).lb[1])]
#define FA3(d, i1, i2, i3) (d).a[(long)((i1) - (d).lb[0]) + (long)(d).n[0] * ((long)((i2) - (d).lb[1]) + (long)(d).n[1] * (long)((i3) - (d).lb[2]))]
#define FA4(d, i1, i2, i3, i4) \
  (d).a[(long)((i1) - (d).lb[0]) + (long)(d).n[0] * ((long)((i2) - (d).lb[1]) + (long)(d).n[1] * ((long)((i3) - (d).lb[2]) + (long)(d).n[2] * (long)((i4) - (d).lb[3])))]

static inline farr mk1(double* a, int l1, int u1) { farr d = {a, {l1, 0, 0, 0}, {u1 - l1 + 1, 0, 0, 0}}; return d; }
static inline farr mk2(double* a, int l1, int u1, int l2, int u2) { farr d = {a, {l1, l2, 0, 0}, {u1 - l1 + 1, u2 - l2 + 1, 0, 0}}; return d; }
static inline farr mk3(double* a, int l1, int u1, int l2, int u2, int l3, int u3) {
  farr d = {a, {l1, l2, l3, 0}, {u1 - l1 + 1, u2 - l2 + 1, u3 - l3 + 1, 0}}; return d; }
static inline farr mk4(double* a, int l1, int u1, int l2, int u2, int l3, int u3, int l4, int u4) {
  farr d = {a, {l1, l2, l3, l4}, {u1 - l1 + 1, u2 - l2 + 1, u3 - l3 + 1, u4 - l4 + 1}}; return d; }
static inline double* fa_mem(long n) { return (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }
static inline farr fa_alloc1(int l1, int u1) { return mk1(fa_mem((long)(u1 - l1 + 1)), l1, u1); }
static inline farr fa_alloc2(int l1, int u1, int l2, int u2) { return mk2(fa_mem((long)(u1 - l1 + 1) * (u2 - l2 + 1)), l1, u1, l2, u2); }
static inline farr fa_alloc3(int l1, int u1, int l2, int u2, int l3, int u3) {
  return mk3(fa_mem((long)(u1 - l1 + 1) * (u2 - l2 + 1) * (u3 - l3 + 1)), l1, u1, l2, u2, l3, u3); }
static inline void fa_free(farr d) { free(d.a); }

/* x**2 is x*x in every Fortran compiler */
static inline double SQD(double x) { return x * x; }
static inline int SQI(int x) { return x * x; }
static inline double MAXD(double a, double b) { return a > b ? a : b; }
static inline double MIND(double a, double b) { return a < b ? a : b; }
static inline int MAXI(int a, int b) { return a > b ? a : b; }
static inline int MINI(int a, int b) { return a < b ? a : b; }
/* number of elements of the section lo:hi:st */
static inline int SECCOUNT(int lo, int hi, int st) { int n = (hi - lo + st) / st; return n > 0 ? n : 0; }
/* Fortran character comparison: trailing blanks do not count */
static inline int fstr_eq(const char* a, const char* b) {
  size_t la = strlen(a), lb = strlen(b);
  while (la > 0 && a[la - 1] == ' ') --la;
  while (lb > 0 && b[lb - 1] == ' ') --lb;
  return la == lb && memcmp(a, b, la) == 0;
}

/* the reference's grid_type (src/mg_grids.f90:24-65), the members the translated routines touch */
typedef struct {
  int nx, ny, nz, npx, npy, incx, incy, gather, ngx, ngy, key, color, nd;
  int neighb[9];
  farr dx, dy, zeta, h, p, b, r, zr, zw, cw, dzw, zxdy, zydx, dummy3dnz, dummy3dnzp, dummy3, ca;
} grid_t;
#define REF_MAXLEV 40
extern grid_t GRID_[REF_MAXLEV];

/* module variables (src/mg_namelist.f90:11-50, mg_mpi.f90:7-8, mg_grids.f90:113-119) */
extern int g_myrank, g_nprocs, g_bmask, g_netcdf_output, g_aggressive, g_nlevs, g_ns_pre, g_ns_post, g_ns_coarsest, g_mpi_proc_null, g_nsmall;
extern double g_hlim, g_nhtheta_b, g_nhtheta_s;
extern const char *g_cmatrix, *g_relax_method, *g_interp_type;

/* call-backs implemented by the harness (single-rank halo fills, global_sum, gather / split are never reached on one rank) */
void cb_fill_halo_relax(int lev, double* p, int nx, int ny, int nz);
void cb_fill_halo(int lev, farr a, const char* lbc_null);
void cb_global_sum(int lev, double local, double* global);
void cb_gather(int lev, farr x, farr y);
void cb_split(int lev, farr x, farr y);
void cb_history(double res, double conv);
#endif
