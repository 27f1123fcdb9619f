/* mg_testseamount.c -- a plain-C host driver over the C ABI of libmgb.so (include/mgb.h).
 *
 * Same flow as the reference's driver src/mg_testseamount.f90 (one rank): nhydro_init (:84) -> seamount geometry
 * (src/mg_setup_tests.f90:130-148) -> nhydro_matrices (:104) -> u = v = 0, w = -1 (:119-123) -> nhydro_solve (:185)
 * -> nhydro_check_nondivergence (:198) -> nhydro_clean.  It writes the residual history to fort.100 like the reference.
 * Nothing here knows about CUDA, Python or torch: this is what a Fortran/C host model links against.
 *
 *   gcc -O2 -Iinclude examples/mg_testseamount.c -Lmgroms_b200 -lmgb -lm -Wl,-rpath,'$ORIGIN/../mgroms_b200' -o examples/mg_testseamount
 *   examples/mg_testseamount [nx ny nz]          (default 64 64 64, the reference's case)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mgb.h"

#define CHECK(call) do { if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, mgb_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
  const int nx = argc > 3 ? atoi(argv[1]) : 64, ny = argc > 3 ? atoi(argv[2]) : 64, nz = argc > 3 ? atoi(argv[3]) : 64;
  const double Lx = 1.0e4, Ly = 1.0e4, Htot = 4.0e3;          /* src/mg_testseamount.f90:76-78 */
  const double hc = 4.0e3, theta_b = 0.0, theta_s = 0.0;      /* :80-82 */
  const size_t n2 = (size_t)(nx + 2) * (ny + 2);
  double *dx = malloc(n2 * 8), *dy = malloc(n2 * 8), *zeta = malloc(n2 * 8), *h = malloc(n2 * 8), *rmask = malloc(n2 * 8);
  const size_t nu = (size_t)(nx + 1) * (ny + 2) * nz, nv = (size_t)(nx + 2) * (ny + 1) * nz, nw = (size_t)(nx + 2) * (ny + 2) * (nz + 1);
  double *u = calloc(nu, 8), *v = calloc(nv, 8), *w = calloc(nw, 8);
  if (!dx || !dy || !zeta || !h || !rmask || !u || !v || !w) { fprintf(stderr, "out of memory\n"); return 1; }

  CHECK(mgb_reset_params());
  CHECK(mgb_set_param("solver_prec", "1.d-10"));
  CHECK(mgb_init(nx, ny, nz, 1, 1, 0, NULL, -1));              /* nh_namelist of the cwd is read if present */

  /* arrays are (0:ny+1, 0:nx+1), j fastest */
  for (int i = 0; i <= nx + 1; ++i)
    for (int j = 0; j <= ny + 1; ++j) {
      const size_t o = (size_t)j + (size_t)(ny + 2) * i;
      const double ddx = Lx / nx, ddy = Ly / ny;
      const double x = (i - 0.5) * ddx, y = (j - 0.5) * ddy, x0 = 0.5 * Lx, y0 = 0.5 * Ly;
      dx[o] = ddx; dy[o] = ddy; zeta[o] = 0.0; rmask[o] = 1.0;
      h[o] = Htot * (1.0 - 0.5 * exp(-(x - x0) * (x - x0) / ((Lx / 5.0) * (Lx / 5.0)) - (y - y0) * (y - y0) / ((Ly / 5.0) * (Ly / 5.0))));
    }
  CHECK(mgb_matrices(dx, dy, zeta, h, rmask, hc, theta_b, theta_s));

  /* w(0:nx+1, 0:ny+1, 0:nz), i fastest: w(:,:,0) = 0, w(:,:,1:nz) = -1 */
  for (size_t q = (size_t)(nx + 2) * (ny + 2); q < nw; ++q) w[q] = -1.0;
  CHECK(mgb_solve(rmask, u, v, w));

  double res[512], conv[512];
  const int n = mgb_get_history(res, conv, 512);
  printf("mgb %s\nlevels %d, F-cycles %d, device time of the solve loop %.3f ms\n", mgb_version(), mgb_nlevs(), mgb_last_solve_iters(), mgb_last_solve_ms());
  for (int q = 0; q < n && q < 512; ++q) printf("  ite %2d  res/bnorm = %.16e  conv = %.3e\n", q, res[q], conv[q]);
  CHECK(mgb_write_fort100("fort.100"));

  CHECK(mgb_check_nondivergence(rmask, u, v, w));             /* recomputes grid(1)%b from the corrected velocities */
  const long long nb = mgb_field_size(1, "b");
  double* b = malloc((size_t)nb * 8);
  if (!b) return 1;
  CHECK(mgb_get_field(1, "b", b));
  double bmax = 0.0;
  for (int i = 1; i <= nx; ++i) for (int j = 1; j <= ny; ++j) for (int k = 0; k < nz; ++k) {
    const double a = fabs(b[(size_t)k + (size_t)nz * ((size_t)j + (size_t)(ny + 2) * i)]);
    if (a > bmax) bmax = a;
  }
  printf("max |div| after the correction: %.6e\n", bmax);
  CHECK(mgb_clean());
  free(dx); free(dy); free(zeta); free(h); free(rmask); free(u); free(v); free(w); free(b);
  return 0;
}
