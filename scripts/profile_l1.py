"""Runs only the level-1 hot kernels (for ncu --set full captures)."""
import sys
sys.path.insert(0, ".")
from mgroms_b200 import nhydro as nh, testcases as tc
kind = sys.argv[1] if len(sys.argv) > 1 else "seamount"
nx, ny, nz = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (512, 512, 64)
nh.set_params(solver_prec=1e-10, solver_maxiter=30)
nh.nhydro_init(nx, ny, nz)
(dx, dy, zeta, h), vert = tc.bench_geometry(kind, nx, ny, nz, 1, 1, 0)
nh.nhydro_matrices(dx, dy, zeta, h, None, **vert)
u, v, w = tc.uvw_wminus1(nx, ny, nz)
nh.compute_rhs(None, u, v, w)
print(nh.profile_level1(int(sys.argv[5]) if len(sys.argv) > 5 else 4))
