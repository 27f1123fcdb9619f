// mgb_internal.h -- internal declarations of libmgb (B200-native multigrid pressure solver).
// Not part of the public ABI (that is include/mgb.h).
#pragma once
#include <string>
#include <vector>
#include <cstddef>

namespace mgb {

// ---- namelist /nhparam/ (reference: src/mg_namelist.f90:11-50) -------------------------------------
struct Params {
  int nsmall = 8, ns_coarsest = 40, ns_pre = 3, ns_post = 2;
  double solver_prec = 1e-6;
  int solver_maxiter = 50;
  std::string cmatrix = "real", relax_method = "RB", interp_type = "linear", restrict_type = "avg";
  bool aggressive = false, netcdf_output = false, bmask = false;
};
// returns 0 ok / 1 file absent (defaults kept) / <0 parse error (msg filled)
int namelist_read(const char* path, Params& P, std::string& msg);
int namelist_set(Params& P, const char* key, const char* value, std::string& msg);

// ---- level hierarchy (reference: src/mg_grids.f90:468-738) ------------------------------------------
struct LevelDims {
  int nx = 0, ny = 0, nz = 0;      // local sizes without halo
  int npx = 1, npy = 1;            // process grid at this level
  int incx = 1, incy = 1;          // rank stride to the neighbours
  int gather = 0, ngx = 1, ngy = 1, key = 0, color = 0;
  int neighb[8] = {-1, -1, -1, -1, -1, -1, -1, -1};   // S,E,N,W,SW,SE,NE,NW ; -1 = MPI_PROC_NULL
  int members[4] = {-1, -1, -1, -1}; int nmembers = 0; // gather family ordered by key
};
int plan_levels(const Params& P, int nx, int ny, int nz, int npx, int npy, int rank,
                std::vector<LevelDims>& out, std::string& msg);

// ---- device-side view of one level -------------------------------------------------------------------
// 3-D fields are (k,j,i), k fastest, j and i carry one halo cell: index k + nz*(j + (ny+2)*i), k 0-based.
// The stencil coefficients are held as eight separate planes (SoA) instead of the reference's
// cA(8,nz,ny,nx) (AoS); mgb_get_field("cA") converts back.
struct LevelView {
  int nx, ny, nz;
  double *p, *b, *r;
  double* cA[8];
  double *bet;        // reciprocal Thomas pivots of the column matrix (depend on cA(1),cA(2) only)
  double *g;          // g = T^-1 e1 (rank-1 correction vector of the red-black bottom coupling)
  double *c5b, *c8b;  // 2-D copies of cA(5,1,:,:), cA(8,1,:,:) (bottom diagonal couplings)
  double *g1, *q1, *x1, *s;   // 2-D work planes of the rank-1 correction
  double *x1b;                // second k=1 plane (double buffer of the fused B+C kernel)
  int rb_chunk, rb_W;         // the bottom recurrence runs in independent chunks of rb_chunk columns warmed up over rb_W columns
  // fused relaxation kernel (kernels_relax.cu): the k=1 plane of q as tagged 16-byte words {lo, launch, hi, launch}, the launch
  // counter {number, exit count}, the mapped host error word + device abort flag raised by a wait that times out, and that timeout
  void* q1t;
  unsigned* tctl;
  int *err, *abort_flag;
  unsigned timeout_ms;
};
// warm-up length of the chunked bottom recurrence: smallest w with rho^w <= 1e-20 (rho = contraction factor)
inline int rb_warmup(double rho) { int w = 0; double acc = 1.0; while (acc > 1e-20 && w < 64) { acc *= rho; ++w; } return w; }
constexpr int RB_CHUNK = 8, RB_WMAX = 24;

struct Halo1Flags { int south, east, north, west, sw, se, ne, nw; };  // 1 = neighbour exists

}  // namespace mgb
