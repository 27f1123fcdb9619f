"""Quick performance probe of the single-GPU solve (development aid, not the bench)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from mgroms_b200 import nhydro as nh, testcases as tc

kind = sys.argv[1] if len(sys.argv) > 1 else "seamount"
nx, ny, nz = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (512, 512, 64)
nh.set_params(solver_prec=1e-10, solver_maxiter=30)
nh.nhydro_init(nx, ny, nz)
(dx, dy, zeta, h), vert = tc.bench_geometry(kind, nx, ny, nz, 1, 1, 0)
t = time.time(); nh.nhydro_matrices(dx, dy, zeta, h, None, **vert); print("matrices s", time.time() - t)
u, v, w = tc.uvw_wminus1(nx, ny, nz)
t = time.time(); nh.compute_rhs(None, u, v, w); print("rhs s (incl H2D)", time.time() - t)
for rep in range(2):
    nh.reset_launch_count()
    t = time.time(); nh.solve_p(); wall = time.time() - t
    hist, conv = nh.history()
    print(f"solve: {nh.last_solve_ms():.2f} ms device, {wall*1e3:.2f} ms wall, iters {nh.last_solve_iters()}, launches {nh.launch_count()}")
print("history", hist)
npts = nx * ny * nz
prof = nh.profile_level1(10)
bytes_per = dict(relax_colour_ms=68, residual_ms=88, restrict_ms=10, prolong_ms=17)
for k, ms in prof.items():
    print(f"{k}: {ms:.4f} ms  -> {bytes_per[k]*npts/ms/1e6:.0f} GB/s algorithmic")
