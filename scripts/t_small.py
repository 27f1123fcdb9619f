import sys; sys.path.insert(0, ".")
from mgroms_b200 import nhydro as nh, testcases as tc
nx, ny, nz = 512, 512, 64
nh.set_params(solver_prec=1e-10)
nh.nhydro_init(nx, ny, nz)
(dx, dy, zeta, h), vert = tc.bench_geometry("seamount", nx, ny, nz, 1, 1, 0)
nh.nhydro_matrices(dx, dy, zeta, h, None, **vert)
nh.compute_rhs(None, *tc.uvw_wminus1(nx, ny, nz))
nh.relax(6, 40); nh.relax(6, 40)
