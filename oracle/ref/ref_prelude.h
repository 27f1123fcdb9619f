/* ref_prelude.h -- run-time support for the C text that oracle/f90toc.py generates from the reference's Fortran sources.
 * TEST INFRASTRUCTURE (oracle/): only tests/, tests/golden/make_ref_golden.py and __graft_entry__.build() use it.
 * Arrays are descriptors (base pointer, lower bounds, extents) so that every Fortran array keeps its own index origin
 * (src/mg_grids.f90:170-260: p(nz,0:ny+1,0:nx+1), zr(nz,-1:ny+2,-1:nx+2), ...); column-major like Fortran.
 * Module variables are thread-local: one thread = one emulated MPI rank (oracle/ref/ref_mpi.c). */
#ifndef MGB_REF_PRELUDE_H
#define MGB_REF_PRELUDE_H
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double* a; int lb[5]; int n[5]; } farr;
typedef struct { int* a; int lb[5]; int n[5]; } iarr;

#define FA1(d, i1) (d).a[(long)((i1) - (d).lb[0])]
#define FA2(d, i1, i2) (d).a[(long)((i1) - (d).lb[0]) + (long)(d).n[0] * (long)((i2) - (d).lb[1])]
#define FA3(d, i1, i2, i3) (d).a[(long)((i1) - (d).lb[0]) + (long)(d).n[0] * ((long)((i2) - (d).lb[1]) + (long)(d).n[1] * (long)((i3) - (d).lb[2]))]
#define FA4(d, i1, i2, i3, i4) \
  (d).a[(long)((i1) - (d).lb[0]) + (long)(d).n[0] * ((long)((i2) - (d).lb[1]) + (long)(d).n[1] * ((long)((i3) - (d).lb[2]) + (long)(d).n[2] * (long)((i4) - (d).lb[3])))]
#define FA5(d, i1, i2, i3, i4, i5)                                                                                                     \
  (d).a[(long)((i1) - (d).lb[0]) + (long)(d).n[0] * ((long)((i2) - (d).lb[1]) + (long)(d).n[1] * ((long)((i3) - (d).lb[2]) +            \
        (long)(d).n[2] * ((long)((i4) - (d).lb[3]) + (long)(d).n[3] * (long)((i5) - (d).lb[4]))))]

#define REF_MK(T, E, name)                                                                                                        \
  static inline T name##1(E* a, int l1, int u1) { T d = {a, {l1, 0, 0, 0, 0}, {u1 - l1 + 1, 0, 0, 0, 0}}; return d; }             \
  static inline T name##2(E* a, int l1, int u1, int l2, int u2) { T d = {a, {l1, l2, 0, 0, 0}, {u1 - l1 + 1, u2 - l2 + 1, 0, 0, 0}}; return d; } \
  static inline T name##3(E* a, int l1, int u1, int l2, int u2, int l3, int u3) {                                                \
    T d = {a, {l1, l2, l3, 0, 0}, {u1 - l1 + 1, u2 - l2 + 1, u3 - l3 + 1, 0, 0}}; return d; }                                     \
  static inline T name##4(E* a, int l1, int u1, int l2, int u2, int l3, int u3, int l4, int u4) {                                \
    T d = {a, {l1, l2, l3, l4, 0}, {u1 - l1 + 1, u2 - l2 + 1, u3 - l3 + 1, u4 - l4 + 1, 0}}; return d; }                          \
  static inline T name##5(E* a, int l1, int u1, int l2, int u2, int l3, int u3, int l4, int u4, int l5, int u5) {                \
    T d = {a, {l1, l2, l3, l4, l5}, {u1 - l1 + 1, u2 - l2 + 1, u3 - l3 + 1, u4 - l4 + 1, u5 - l5 + 1}}; return d; }
REF_MK(farr, double, mk)
REF_MK(iarr, int, imk)
static inline long ref_ext(int l, int u) { return u >= l ? (long)(u - l + 1) : 0; }
static inline double* fa_mem(long n) { return (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }
static inline int* ia_mem(long n) { return (int*)calloc((size_t)(n > 0 ? n : 1), sizeof(int)); }
static inline farr fa_alloc1(int l1, int u1) { return mk1(fa_mem(ref_ext(l1, u1)), l1, u1); }
static inline farr fa_alloc2(int l1, int u1, int l2, int u2) { return mk2(fa_mem(ref_ext(l1, u1) * ref_ext(l2, u2)), l1, u1, l2, u2); }
static inline farr fa_alloc3(int l1, int u1, int l2, int u2, int l3, int u3) {
  return mk3(fa_mem(ref_ext(l1, u1) * ref_ext(l2, u2) * ref_ext(l3, u3)), l1, u1, l2, u2, l3, u3); }
static inline farr fa_alloc4(int l1, int u1, int l2, int u2, int l3, int u3, int l4, int u4) {
  return mk4(fa_mem(ref_ext(l1, u1) * ref_ext(l2, u2) * ref_ext(l3, u3) * ref_ext(l4, u4)), l1, u1, l2, u2, l3, u3, l4, u4); }
static inline farr fa_alloc5(int l1, int u1, int l2, int u2, int l3, int u3, int l4, int u4, int l5, int u5) {
  return mk5(fa_mem(ref_ext(l1, u1) * ref_ext(l2, u2) * ref_ext(l3, u3) * ref_ext(l4, u4) * ref_ext(l5, u5)), l1, u1, l2, u2, l3, u3, l4, u4, l5, u5); }
static inline iarr ia_alloc1(int l1, int u1) { return imk1(ia_mem(ref_ext(l1, u1)), l1, u1); }
static inline void fa_free(farr d) { free(d.a); }
static inline void ia_free(iarr d) { free(d.a); }

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

#define REF_MAXLEV 40

/* module variables (src/mg_namelist.f90:11-50, mg_mpi.f90:7-8, mg_grids.f90:113-119), one copy per emulated rank */
extern __thread int g_myrank, g_nprocs, g_bmask, g_netcdf_output, g_aggressive, g_nlevs, g_ns_pre, g_ns_post, g_ns_coarsest, g_nsmall, g_solver_maxiter;
extern __thread double g_hlim, g_nhtheta_b, g_nhtheta_s, g_solver_prec;
extern __thread const char *g_cmatrix, *g_relax_method, *g_interp_type;

/* MPI constants and the thread-emulated MPI (oracle/ref/ref_mpi.c) */
enum { g_mpi_proc_null = -2, g_mpi_comm_world = 0, g_mpi_request_null = -1, g_mpi_double_precision = 0, g_mpi_sum = 0, g_mpi_status_size = 4 };
void rmpi_irecv(double* buf, int count, int src, int tag, int comm, int* req);
void rmpi_isend(double* buf, int count, int dst, int tag, int comm, int* req);
void rmpi_waitany(int count, int* reqs, int* index);
void rmpi_allreduce_sum(double* in, double* out, int count, int comm);
void rmpi_allgather(double* in, int n, double* out, int comm);
void rmpi_comm_split(int comm, int color, int key, int* newcomm);
void rmpi_barrier(int comm);

/* unit 100 of the reference (the residual history) */
void cb_history(double res, double conv);
#endif
