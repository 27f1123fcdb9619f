/* mgb.h -- C ABI of the B200-native multigrid pressure solver (libmgb.so).
 *
 * Drop-in boundary for the hot path of CESR-lab/mgroms 0.3.9: the entry points below are what a thin
 * ISO_C_BINDING shim of the reference's `module nhydro` (src/nhydro.f90) binds, one C function per
 * Fortran procedure, plus accessors for the module-global state other code reaches into
 * (grid(lev)%p, %b, %r, %cA -- src/mg_grids.f90:24-65) and fine-grained operator calls used by tests.
 * See INTEGRATION.md for the Fortran-side stub.
 *
 * Conventions
 *   - one solver instance per process (the reference's state is module-global, src/mg_grids.f90:113-119);
 *     one process drives one GPU; ranks of an npx*npy process grid map 1:1 to GPUs.
 *   - every function returns 0 on success, non-zero on error; mgb_last_error() gives the message
 *     (the reference `stop`s with a message instead, e.g. src/mg_grids.f90:529-530).
 *   - all reals are double (rp=8), all integers int (ip=4)  -- src/mg_namelist.f90:8.
 *   - levels are 1-based, 1 = finest (src/mg_grids.f90:117).
 *   - host arrays are Fortran column-major exactly as the reference declares them.
 *   - there is NO CPU fallback: compute entry points fail if no CUDA device is usable.
 */
#ifndef MGB_H
#define MGB_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- life cycle ------------------------------------------------------------------------------ */

/* nhydro_init(nx,ny,nz,npxg,npyg)  src/nhydro.f90:18-34  (+ mg_mpi_init: rank/size, src/mg_mpi.f90:12-36;
 * read_nhnamelist, src/mg_namelist.f90:55-127).  nx,ny,nz are LOCAL sizes; rank in [0,npx*npy).
 * namelist_path NULL or "" means "nh_namelist in the cwd if present, else defaults".
 * device < 0 selects cudaGetDevice()'s current device. */
int mgb_init(int nx, int ny, int nz, int npx, int npy, int rank, const char* namelist_path, int device);

/* Programmatic override of one namelist key (same 13 keys as namelist /nhparam/, src/mg_namelist.f90:37-50).
 * Must be called before mgb_init; mgb_reset_params() restores the defaults of src/mg_namelist.f90:11-35. */
int mgb_set_param(const char* key, const char* value);
int mgb_reset_params(void);

/* Host-only: parse `namelist_path` (NULL/"" = ./nh_namelist) like mgb_init would, apply the mgb_set_param overrides and
 * write "key=value;..." for the 13 keys (+ file=0|1: whether a file was found) to out.  No device needed. */
int mgb_namelist_probe(const char* namelist_path, char* out, int cap);

/* Inter-GPU transport for npx*npy > 1 (replaces MPI_COMM_WORLD of src/mg_mpi_exchange.f90 and the
 * MPI_COMM_SPLIT sub-communicators of src/mg_grids.f90:717).  Rank 0 fills a 128-byte id with
 * mgb_comm_unique_id, the caller broadcasts it (any side channel), every rank calls mgb_comm_init. */
int mgb_comm_unique_id(void* id128);
int mgb_comm_init(const void* id128);

/* nhydro_clean()  src/nhydro.f90:137-141 */
int mgb_clean(void);

/* ---- matrices -------------------------------------------------------------------------------- */

/* nhydro_matrices(dx,dy,zeta,h,rmask,hc,theta_b,theta_s)  src/nhydro.f90:36-50 -> define_matrices_topo
 * src/mg_define_matrix.f90:28-208.  dx,dy,zeta,h,rmask: (0:ny+1,0:nx+1), j fastest.  rmask may be NULL (=1). */
int mgb_matrices(const double* dx, const double* dy, const double* zeta, const double* h, const double* rmask,
                 double hc, double theta_b, double theta_s);

/* ---- solve ----------------------------------------------------------------------------------- */

/* nhydro_solve(nx,ny,nz,rmaska,ua,va,wa)  src/nhydro.f90:53-102: compute_rhs -> solve_p -> correct_uvw.
 * u(1:nx+1,0:ny+1,1:nz), v(0:nx+1,1:ny+1,1:nz), w(0:nx+1,0:ny+1,0:nz), i fastest, corrected in place.
 * Host buffers; the H2D/D2H copies are part of the call. */
int mgb_solve(const double* rmask, double* u, double* v, double* w);

/* nhydro_check_nondivergence(...)  src/nhydro.f90:105-134: recomputes grid(1)%b only. */
int mgb_check_nondivergence(const double* rmask, const double* u, const double* v, const double* w);

/* The three steps of nhydro_solve, separately (compute_rhs src/mg_compute_rhs.f90:14, solve_p
 * src/mg_solvers.f90:17, correct_uvw src/mg_correct_uvw.f90:15).  mgb_solve_p works on the device-resident
 * grid(1)%b and leaves grid(1)%p on the device. */
int mgb_compute_rhs(const double* rmask, const double* u, const double* v, const double* w);
int mgb_solve_p(void);
int mgb_correct_uvw(double* u, double* v, double* w);

/* Residual history of the last solve: res[0] = rnorm0/bnorm (=1 for p=0), then rnorm/bnorm per F-cycle --
 * the lines the reference writes to unit 100 (src/mg_solvers.f90:59,72).  Returns the number of entries
 * (may exceed cap; only cap are written); conv may be NULL. */
int mgb_get_history(double* res, double* conv, int cap);
/* Write the history in the reference's fort.100 format (list-directed: value, conv) to `path`. */
int mgb_write_fort100(const char* path);
/* Device time (CUDA events) of the last mgb_solve_p loop in ms, and the number of F-cycles done. */
double mgb_last_solve_ms(void);
int mgb_last_solve_iters(void);

/* ---- operators, one call per reference routine (tests, tuning) ------------------------------- */
int mgb_relax(int lev, int nsweeps);          /* relax            src/mg_relax.f90:16-47          */
int mgb_residual(int lev, double* res);       /* compute_residual src/mg_relax.f90:337-383        */
int mgb_fine2coarse(int lev);                 /* fine2coarse      src/mg_intergrids.f90:16-72     */
int mgb_coarse2fine(int lev);                 /* coarse2fine      src/mg_intergrids.f90:167-228   */
int mgb_fcycle(void);                         /* Fcycle           src/mg_solvers.f90:104-126      */
int mgb_vcycle(int lev1);                     /* Vcycle           src/mg_solvers.f90:129-151      */
int mgb_fill_halo(int lev, const char* name); /* fill_halo        src/mg_mpi_exchange.f90:10-16   */

/* ---- state access ---------------------------------------------------------------------------- */
int mgb_nlevs(void);
/* info[0..11] = nx,ny,nz,npx,npy,incx,incy,gather,ngx,ngy,key,color ; info[12..19] = neighbours S,E,N,W,SW,SE,NE,NW
 * (-1 = MPI_PROC_NULL)  -- src/mg_grids.f90:503-738.  Works without a GPU after mgb_plan_levels(). */
int mgb_level_info(int lev, int* info20);
/* Host-only planning of the level hierarchy (no device touched): same logic mgb_init uses. */
int mgb_plan_levels(int nx, int ny, int nz, int npx, int npy, int rank);
/* Members (level-1 ranks, ordered by key) of this rank's gather family at `lev`; returns the count. */
int mgb_gather_members(int lev, int* out4);

/* Number of doubles of field `name` at `lev` in the reference's layout:
 *   "p","b","r": (nz,0:ny+1,0:nx+1)   "cA": (8,nz,0:ny+1,0:nx+1)   "dx","dy","zeta","h": (0:ny+1,0:nx+1)
 *   "zr": (nz,-1:ny+2,-1:nx+2)  "zw": (nz+1,-1:ny+2,-1:nx+2)  "cw": (nz+1,0:ny+1,0:nx+1)
 *   level 1 only: "dzw": (nz+1,..)  "zxdy","zydx": (nz,..)      returns -1 for an unknown name. */
long long mgb_field_size(int lev, const char* name);
int mgb_get_field(int lev, const char* name, double* out);
int mgb_set_field(int lev, const char* name, const double* in);

/* ---- diagnostics ----------------------------------------------------------------------------- */
const char* mgb_last_error(void);
/* Transport of the nh = 1 halo exchange (replaces the Isend/Irecv pairs of src/mg_mpi_exchange.f90:396-745):
 * 0 = single rank (no communicator), 1 = NVLink peer-memory mailboxes, 2 = NCCL send/recv (fallback, or MGB_NO_P2P=1). */
int mgb_halo_transport(void);
/* Kernel launches issued by this library since the last mgb_reset_launch_count() (bench.py's gpu_launches). */
long long mgb_launch_count(void);
void mgb_reset_launch_count(void);
/* Which kernels relax(lev, .) runs for this level and configuration: "small" (single CTA), "mid" (cooperative launch),
 * "fused" (one persistent launch per colour), "a+bc" / "a+b+c" (separate launches), "a" (no k=1 diagonal coupling),
 * suffixed "+halo" / "+exchange" when the halo fill is a separate launch.  Diagnostic for tests and bench.py. */
int mgb_relax_path(int lev, char* out, int cap);
/* Per-kernel-class device time accumulated by mgb_profile_level1(): runs `reps` launches of each level-1
 * hot kernel on the current state and reports average ms per launch.
 * out[0]=relax colour (all launches of one colour), [1]=residual+norm, [2]=restrict, [3]=prolong+add. */
int mgb_profile_level1(int reps, double* out4);
/* Device time (us) of one fill_halo(grid(lev)%p) -- src/mg_mpi_exchange.f90:396 -- averaged over `reps` back-to-back
 * exchanges on the solver's stream (collective: every rank must call it with the same arguments). */
int mgb_profile_halo(int lev, int reps, double* us_per_exchange);
/* mg_tictoc (src/mg_tictoc.f90): the reference's tic/toc sites -- nhydro_solve, compute_rhs (src/nhydro.f90:67-100), solve,
 * Fcycle (src/mg_solvers.f90:45-124), relax_3D_8_{RB,GS,FC}, residual_3D_8 (src/mg_relax.f90:128-370) -- timed per level with
 * CUDA events.  Enabling it replaces the CUDA-graph replay of the F-cycle by plain launches (a profiling mode);
 * MGB_TICTOC=1 in the environment enables it at mgb_init.  mgb_write_tictoc writes the table of print_tictoc
 * (src/mg_tictoc.f90:114-153: seconds in E9.3 and call counts in I9, "Total" + one column per level) to `path`,
 * NULL or "" = "fort.<10+rank>" like the reference. */
int mgb_tictoc_enable(int on);
int mgb_write_tictoc(const char* path);
const char* mgb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MGB_H */
