/* mg_testseamount_mp.c -- the reference's seamount driver on npx x npy ranks, ONE PROCESS PER GPU, in plain C over the C ABI.
 *
 * What `mpirun -np N ./mg_testseamount` does for the reference (src/mg_testseamount.f90 with MPI ranks), done without MPI:
 * the parent forks N-1 children before anything touches CUDA, rank 0 creates the communicator id (mgb_comm_unique_id) and
 * hands it to the others through pipes -- the side channel a Fortran host would replace by one MPI_Bcast of 128 bytes
 *  -- and every rank calls mgb_init(..., rank, ..., device = rank), mgb_comm_init, then the usual
 * nhydro_matrices / nhydro_solve / nhydro_check_nondivergence / nhydro_clean sequence on its own sub-domain.
 * Rank 0 prints the residual history (identical on all ranks) and writes fort.100.
 *
 *   gcc -O2 -Iinclude examples/mg_testseamount_mp.c -Lmgroms_b200 -lmgb -lm -Wl,-rpath,'$ORIGIN/../mgroms_b200' -o examples/mg_testseamount_mp
 *   examples/mg_testseamount_mp npx npy [nx ny nz]      (nx ny nz = LOCAL sizes, default 32 32 16; needs npx*npy GPUs)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>
#include "mgb.h"

#define CHECK(call) do { if ((call) != 0) { fprintf(stderr, "[rank %d] %s failed: %s\n", rank, #call, mgb_last_error()); return 1; } } while (0)

static int run_rank(int rank, int npx, int npy, int nx, int ny, int nz, const int* id_rd, int (*id_wr)[2]) {
  const int N = npx * npy, pi = rank % npx, pj = rank / npx;            /* src/mg_grids.f90:593-594 */
  const double Lx = 1.0e4, Ly = 1.0e4, Htot = 4.0e3, hc = 4.0e3, theta_b = 0.0, theta_s = 0.0;   /* src/mg_testseamount.f90:76-82 */
  const size_t n2 = (size_t)(nx + 2) * (ny + 2);
  double *dx = malloc(n2 * 8), *dy = malloc(n2 * 8), *zeta = malloc(n2 * 8), *h = malloc(n2 * 8), *rmask = malloc(n2 * 8);
  const size_t nu = (size_t)(nx + 1) * (ny + 2) * nz, nv = (size_t)(nx + 2) * (ny + 1) * nz, nw = (size_t)(nx + 2) * (ny + 2) * (nz + 1);
  double *u = calloc(nu, 8), *v = calloc(nv, 8), *w = calloc(nw, 8);
  if (!dx || !dy || !zeta || !h || !rmask || !u || !v || !w) { fprintf(stderr, "out of memory\n"); return 1; }
  CHECK(mgb_reset_params());
  CHECK(mgb_set_param("solver_prec", "1.d-10"));
  CHECK(mgb_init(nx, ny, nz, npx, npy, rank, NULL, rank));
  unsigned char id[128];
  if (rank == 0) {
    CHECK(mgb_comm_unique_id(id));
    for (int r = 1; r < N; ++r) if (write(id_wr[r][1], id, sizeof id) != (ssize_t)sizeof id) { perror("write"); return 1; }
  } else if (read(*id_rd, id, sizeof id) != (ssize_t)sizeof id) { perror("read"); return 1; }
  CHECK(mgb_comm_init(id));
  /* seamount of the GLOBAL domain, evaluated on this rank's (0:ny+1, 0:nx+1) window (src/mg_setup_tests.f90:130-148) */
  const double ddx = Lx / (npx * nx), ddy = Ly / (npy * ny), x0 = 0.5 * Lx, y0 = 0.5 * Ly;
  for (int i = 0; i <= nx + 1; ++i)
    for (int j = 0; j <= ny + 1; ++j) {
      const size_t o = (size_t)j + (size_t)(ny + 2) * i;
      const double x = ((double)(i + pi * nx) - 0.5) * ddx, y = ((double)(j + pj * ny) - 0.5) * ddy;
      dx[o] = ddx; dy[o] = ddy; zeta[o] = 0.0; rmask[o] = 1.0;
      h[o] = Htot * (1.0 - 0.5 * exp(-(x - x0) * (x - x0) / ((Lx / 5.0) * (Lx / 5.0)) - (y - y0) * (y - y0) / ((Ly / 5.0) * (Ly / 5.0))));
    }
  CHECK(mgb_matrices(dx, dy, zeta, h, rmask, hc, theta_b, theta_s));
  for (size_t q = (size_t)(nx + 2) * (ny + 2); q < nw; ++q) w[q] = -1.0;      /* u = v = 0, w(:,:,0) = 0, w(:,:,1:nz) = -1 */
  CHECK(mgb_solve(rmask, u, v, w));
  if (rank == 0) {
    double res[512], conv[512];
    const int n = mgb_get_history(res, conv, 512);
    printf("mgb %s\n%d x %d ranks, local %d x %d x %d, levels %d, F-cycles %d, device time of the solve loop %.3f ms\n", mgb_version(), npx, npy, nx, ny, nz,
           mgb_nlevs(), mgb_last_solve_iters(), mgb_last_solve_ms());
    for (int q = 0; q < n && q < 512; ++q) printf("  ite %2d  res/bnorm = %.16e  conv = %.3e\n", q, res[q], conv[q]);
    CHECK(mgb_write_fort100("fort.100"));
  }
  CHECK(mgb_check_nondivergence(rmask, u, v, w));
  CHECK(mgb_clean());
  free(dx); free(dy); free(zeta); free(h); free(rmask); free(u); free(v); free(w);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s npx npy [nx ny nz]\n", argv[0]); return 2; }
  const int npx = atoi(argv[1]), npy = atoi(argv[2]);
  const int nx = argc > 5 ? atoi(argv[3]) : 32, ny = argc > 5 ? atoi(argv[4]) : 32, nz = argc > 5 ? atoi(argv[5]) : 16;
  const int N = npx * npy;
  if (N < 1 || N > 64) { fprintf(stderr, "npx*npy must be 1..64\n"); return 2; }
  int pipes[64][2];
  pid_t pids[64];
  for (int r = 1; r < N; ++r) if (pipe(pipes[r]) != 0) { perror("pipe"); return 1; }
  for (int r = 1; r < N; ++r) {                       /* fork BEFORE the first CUDA call: a CUDA context does not survive fork */
    pids[r] = fork();
    if (pids[r] < 0) { perror("fork"); return 1; }
    if (pids[r] == 0) {
      close(pipes[r][1]);
      _exit(run_rank(r, npx, npy, nx, ny, nz, &pipes[r][0], NULL));
    }
    close(pipes[r][0]);
  }
  int rc = run_rank(0, npx, npy, nx, ny, nz, NULL, pipes);
  for (int r = 1; r < N; ++r) close(pipes[r][1]);     /* a rank still waiting for the id (rank 0 failed early) sees EOF and stops */
  for (int r = 1; r < N; ++r) {
    int st = 0;
    waitpid(pids[r], &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) { fprintf(stderr, "rank %d failed (status %d)\n", r, st); rc = 1; }
  }
  return rc;
}
