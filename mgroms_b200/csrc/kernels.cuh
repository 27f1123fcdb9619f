// kernels.cuh -- launch wrappers of the sm_100a kernels (definitions in kernels_*.cu).
#pragma once
#include <cuda_runtime.h>
#include "mgb_internal.h"

namespace mgb {

extern long long g_launches;   // number of kernel launches issued by this library
void kernels_init();           // one-time function attributes (dynamic shared memory limits), occupancy-derived grid caps
void note_launch_error(cudaError_t e);   // remembers the first failing launch ...
cudaError_t take_launch_error();         // ... for the next API boundary (solver.cu: finish)
int device_sm_count();
int thomas_variant();                    // MGB_THOMAS (experiments): 0 plain loops, 1 prefetching form                   // SMs of the current device (grid caps are derived from it, not hard-coded)

// ---- halo (single-GPU part: physical edges; src/mg_mpi_exchange.f90) ---------------------------------
// nh = 1 array a(nzl, 0:ny+1, 0:nx+1): homogeneous-Neumann mirror on edges without neighbour, physical corners.
void launch_halo1_phys(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f);
// corners that depend on just-received halos ("flag_*" logic, src/mg_mpi_exchange.f90:720-743)
void launch_halo1_flagcorners(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f);
// two-launch multi-GPU variant for nh = 1: `pre` = physical mirrors + pack of the S/N rows; `post` = unpack + flag corners (+ x1 upkeep)
void launch_halo1_pre(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f, double* packS, double* packN);
void launch_halo1_post(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f, const double* recvS, const double* recvN, double* x1);
// One-launch multi-GPU variant over NVLink peer memory (no NCCL, no fences): every rank owns a mailbox per level that its
// neighbours have mapped (CUDA IPC).  A double travels as one 16-byte store {lo, n, hi, n} where n is the number of the
// exchange: the receiver polls the word until both tags equal n, so the payload is its own arrival flag (each 8-byte half
// is single-copy atomic).  The kernel (1) mirrors the physical edges and stores its boundary faces into the neighbours'
// mailboxes, (2) polls its own mailbox and unpacks it into the halo ring (+ flag corners, + x1 upkeep).
// Two mailbox parities make the write-after-read hazard between consecutive exchanges impossible (a neighbour can be
// at most one exchange ahead).  The grid is capped so that all blocks are co-resident (step 2 spins).
struct P2PBox {
  uint4* peer[8];                     // mailbox (both parities) of the neighbour in direction q (S,E,N,W,SW,SE,NE,NW)
  const uint4* mine;                  // my mailbox
  unsigned long long* seq;            // exchanges completed at this level (local)
  unsigned int* done;                 // blocks that have read *seq; the last one advances it (local)
  int* err;                           // mapped host word: set to 1 when a poll times out
  int* abort_flag;                    // device word raised with it: later waits return at once instead of timing out again
  unsigned timeout_ms;                // wall-clock limit of one wait
  int nzmax;                          // slot sizing: rows are nzmax*nx, columns nzmax*ny, corners nzmax words
};
void launch_halo1_p2p(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f, const P2PBox& B, double* x1);
// nh = 2 array a(nzl, -1:ny+2, -1:nx+2) (zr, zw): ring 1 mirror, ring 2 = 2A-B, corners point mirror
void launch_halo2_phys(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f);
void launch_halo2_flagcorners(cudaStream_t st, double* a, int nzl, int ny, int nx, Halo1Flags f);
// pack / unpack of strided S/N faces and nh-wide E/W faces for the inter-GPU exchange
void launch_pack_rows(cudaStream_t st, const double* a, double* buf, int nzl, int ny, int nx, int nh, int j0, int nrows, int i0, int ncols);
void launch_unpack_rows(cudaStream_t st, double* a, const double* buf, int nzl, int ny, int nx, int nh, int j0, int nrows, int i0, int ncols);

// ---- relaxation (src/mg_relax.f90:151-334) -------------------------------------------------------------
struct ColourSet {     // columns visited by one colour of a sweep
  int istart, istep;   // i = istart, istart+istep, ... <= nx
  int jrule;           // 0: j = 1 + mod(i+rb,2), step 2 (red-black);  1: j = jfix, step 2 (four-colour);
                       // 2: the single column j = jfix - 2 i (Gauss-Seidel wavefront jfix = j + 2 i)
  int rb, jfix;
  int bottom_mode;     // 0: no k=1 diagonal terms ('simple');  1: rank-1 form (two (i+1) terms from the x1 snapshot);
                       // 2: all four terms read from p (no same-colour coupling, four-colour)
  int fuse_halo;       // 1: level without neighbour ranks -> the kernels write the Neumann halo themselves and keep x1 = p(1,:,:)
                       // 2: level with neighbours -> kernel C keeps the interior of x1 current, the halo exchange does the ring
};
void launch_snapshot_x1(cudaStream_t st, const LevelView& L);
void launch_relax_A(cudaStream_t st, const LevelView& L, const ColourSet& cs);
void launch_relax_B(cudaStream_t st, const LevelView& L, const ColourSet& cs);
void launch_relax_C(cudaStream_t st, const LevelView& L, const ColourSet& cs);
// B and C fused, for neighbour-free levels with a chunked recurrence; reads the k=1 plane x1old, writes x1new completely
void launch_relax_BC(cudaStream_t st, const LevelView& L, const ColourSet& cs, const double* x1old, double* x1new);
// one colour of the rank-1 red-black sweep in ONE persistent launch (A + B + C fused, kernels_relax.cu); big levels only.
// mirror = 1 on a level without neighbour ranks (the kernel writes the Neumann halo of p and of the k=1 plane itself).
// x1old is read, x1new is written on the whole interior (the halo exchange completes its ring when there are neighbours).
bool relax_fused_supported(const LevelView& L);
size_t relax_fused_num_tiles(int nx, int ny, int nz);     // flags to allocate for the level (0: never fused)
void relax_fused_init();                                  // occupancy query / shared-memory attributes (outside any capture)
void launch_relax_fused(cudaStream_t st, const LevelView& L, int rb, int mirror, const double* x1old, double* x1new);
// whole relax(lev, nsweeps) of a small, neighbour-free level in one single-CTA launch (red-black only)
bool relax_small_supported(const LevelView& L);
// hasnb = 1: the level has neighbour ranks and the halo travels over peer memory: the kernel runs the exchange itself after every colour
void launch_relax_small(cudaStream_t st, const LevelView& L, int nsweeps, int rank1, int hasnb, const Halo1Flags& hf, const P2PBox& B);
// whole relax(lev, nsweeps) of a mid-size, neighbour-free level in one cooperative launch with software grid barriers
// (red-black only); `bar` = two zeroed device words owned by the level
bool relax_mid_supported(const LevelView& L);
void launch_relax_mid(cudaStream_t st, const LevelView& L, int nsweeps, int rank1, unsigned* bar, int hasnb, const Halo1Flags& hf, const P2PBox& B);

// ---- residual + norm (src/mg_relax.f90:337-515) ----------------------------------------------------------
// r = b - A p on the interior; partial sums of r^2 per block into `partials`, then reduced in fixed order to *sum_out.
int  residual_num_partials(const LevelView& L);
void launch_residual(cudaStream_t st, const LevelView& L, int real_matrix, double* partials, double* sum_out);
void launch_sumsq(cudaStream_t st, const LevelView& L, const double* a, double* partials, double* sum_out);

// ---- intergrid (src/mg_intergrids.f90) -------------------------------------------------------------------
// y_c(k2,j2,i2) = sum of the 8 fine cells of x_f ; y has local dims (nxc,nyc,nzc) with 1 halo
void launch_restrict8(cudaStream_t st, const double* xf, int nzf, int nyf, int nxf, double* yc, int nzc, int nyc, int nxc);
// the same on a neighbour-free coarse level, fused with halo(y_c), p_c = 0 (whole array) and, when rc != nullptr, r_c = y_c
void launch_restrict8_fused(cudaStream_t st, const double* xf, int nzf, int nyf, int nxf, double* yc, double* pc, double* rc, int nzc, int nyc, int nxc);
// trilinear/nearest prolongation of xc (nzc,nyc,nxc) onto the fine interior.
// mode 0: rf = interp           (reference's first step; halo(rf) and p+=rf done by the caller)
// mode 1: pf += interp          (fused; caller refreshes halo(pf))
// mode 2: as mode 1 on a neighbour-free fine level: the kernel writes the Neumann halo of pf itself (linear only)
void launch_prolong(cudaStream_t st, const double* xc, int nzc, int nyc, int nxc, double* rf, double* pf, int linear, int mode);
// pf += P xc (linear) on the interior of a big level, the fine field staged through shared memory by the bulk-copy engine
// (kernels_bulk.cu); the caller refreshes the halo of pf
bool prolong_bulk_supported(int nzc, int nyc, int nxc);
void prolong_bulk_init();
void launch_prolong_bulk(cudaStream_t st, const double* xc, int nzc, int nyc, int nxc, double* pf);
void launch_axpy_all(cudaStream_t st, double* y, const double* x, size_t n);          // y += x
void launch_add_halo_ring(cudaStream_t st, double* p, const double* r, int nz, int ny, int nx);  // p += r on the halo ring only
void launch_fill(cudaStream_t st, double* a, double v, size_t n);
void launch_copy(cudaStream_t st, double* dst, const double* src, size_t n);

// ---- gather / split (src/mg_gather.f90) -------------------------------------------------------------------
// copy the interior of sub-block x (nz, nys, nxs, halo 1) into y (nz, ny, nx, halo 1) at block offset (l,m)
void launch_gather_place(cudaStream_t st, const double* x, int nzl, int nys, int nxs, double* y, int ny, int nx, int l, int m);
// y(k,j,i) = x(k, j+m*nys, i+l*nxs), i in 0..nxs+1, j in 0..nys+1
void launch_split(cudaStream_t st, const double* x, int nzl, int ny, int nx, double* y, int nys, int nxs, int l, int m);

// ---- matrices (src/mg_define_matrix.f90, src/mg_zr_zw.f90) ---------------------------------------------------
void launch_coarsen2d(cudaStream_t st, const double* f, int nyf, int nxf, double* c, int nyc, int nxc, double w);
// tab: 4*(nz+1) doubles on the device: cs_r[k], cff_r[k], cs_w[k], cff_w[k]
void launch_zrzw(cudaStream_t st, const double* h, const double* zeta, const double* tab, double hlim, double* zr, double* zw, int nz, int ny, int nx);
struct MatrixView {
  int nx, ny, nz, lev1;
  const double *dx, *dy, *zr, *zw;
  const double *rmask; int bmask;     // level mask (0:ny+1,0:nx+1) and namelist bmask: umask/vmask = 1 unless bmask
  double *cw, *dzw, *zxdy, *zydx;
  double* cA[8];
  double *bet, *g, *c5b, *c8b, *g1;
};
void launch_define_matrix(cudaStream_t st, const MatrixView& M);        // coefficients + diagonal + derived (no cA exchange in between)
void launch_define_matrix_coef(cudaStream_t st, const MatrixView& M);   // cw, metrics, cA(2..8)
void launch_define_matrix_diag(cudaStream_t st, const MatrixView& M);   // cA(1), bet, g, g1  (after the cA halo exchange that bmask adds)
// rmask of a coarse level: ones, halo zeroed on sides without neighbour when bmask (src/mg_define_matrix.f90:155-161)
void launch_level_rmask(cudaStream_t st, double* rm, int ny, int nx, Halo1Flags f, int bmask);
void launch_refresh_derived(cudaStream_t st, const MatrixView& M);   // c5b, c8b, bet, g, g1 from the cA planes
void launch_rho(cudaStream_t st, const MatrixView& M, double* out_dev);   // contraction factor of the bottom recurrence
// max(|c5b|,|c8b|) > 0 ?  result (0/1) written to *flag_dev
void launch_any_nonzero2(cudaStream_t st, const double* a, const double* b, size_t n, int* flag_dev);

// ---- rhs / correction (src/mg_compute_rhs.f90, src/mg_correct_uvw.f90) ------------------------------------------
// model layout (i fastest) <-> solver layout (k fastest); src dims: (ni, nj, nk); dst index k + nk*(j + nj*i)
void launch_transpose_ijk_to_kji(cudaStream_t st, const double* src, double* dst, int ni, int nj, int nk);
void launch_transpose_kji_to_ijk(cudaStream_t st, const double* src, double* dst, int ni, int nj, int nk);
struct RhsView {
  int nx, ny, nz;
  const double *dx, *dy, *zw, *dzw, *zxdy, *zydx, *cw, *rmask;
  const double *ut, *vt, *wt;   // k-fastest: ut(nz, ny+2, nx+1) i=1..nx+1 ; vt(nz, ny+1, nx+2) j=1..ny+1 ; wt(nz+1, ny+2, nx+2)
  double *flux;                 // scratch (nz+1, ny+2, nx+2)
  double *b;
  int west_phys, east_phys, south_phys, north_phys;   // wall faces to zero (lbc_null); all 0 with bmask (the masks close the walls)
  int bmask;
};
void launch_rhs_uf(cudaStream_t st, const RhsView& R);       // flux(k,j,i) = uf, i=1..nx+1, j=1..ny (+ wall zero)
void launch_rhs_vf(cudaStream_t st, const RhsView& R);
void launch_rhs_wf(cudaStream_t st, const RhsView& R);
void launch_rhs_div(cudaStream_t st, const RhsView& R, int dir);   // dir 0: b = d_i uf ; 1: b += d_j vf ; 2: b += d_k wf
struct CorrView {
  int nx, ny, nz;
  const double *dx, *dy, *zr, *zw, *p;
  double *ut, *vt, *wt;
  const double* rmask; int bmask;
};
void launch_correct_uvw(cudaStream_t st, const CorrView& C);

// ---- layout conversion for accessors -------------------------------------------------------------------------
void launch_soa_to_aos8(cudaStream_t st, double* const* planes_dev, double* aos, size_t npts);
void launch_aos8_to_soa(cudaStream_t st, const double* aos, double* const* planes_dev, size_t npts);

}  // namespace mgb
