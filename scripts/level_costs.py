"""In-situ cost of the operators per level (wall clock around synchronous API calls, many sweeps to amortise)."""
import sys, time
sys.path.insert(0, ".")
from mgroms_b200 import nhydro as nh, testcases as tc
nx, ny, nz = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (512, 512, 64)
nh.set_params(solver_prec=1e-10, solver_maxiter=30)
nh.nhydro_init(nx, ny, nz)
(dx, dy, zeta, h), vert = tc.bench_geometry("seamount", nx, ny, nz, 1, 1, 0)
nh.nhydro_matrices(dx, dy, zeta, h, None, **vert)
nh.compute_rhs(None, *tc.uvw_wminus1(nx, ny, nz)); nh.solve_p()
L = nh.nlevs()
plan = nh.plan_levels(nx, ny, nz)
def t(f, n=3):
    f(); best = 1e9
    for _ in range(n):
        t0 = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0)
    return best * 1e6
tot = 0
for lev in range(1, L + 1):
    d = plan[lev - 1]
    ns = 20
    tr = (t(lambda: nh.relax(lev, ns)) - 15) / (2 * ns)
    tres = t(lambda: nh.compute_residual(lev)) - 25
    line = f"lev {lev} {d['nx']}x{d['ny']}x{d['nz']}: relax colour {tr:8.1f} us  residual {tres:8.1f} us"
    visits = lev if lev < L else L
    sweeps = 5 if lev < L else 40
    cost = visits * sweeps * 2 * tr + (visits * tres if lev < L else 0)
    if lev < L:
        tf = t(lambda: nh.fine2coarse(lev)) - 15; tc2 = t(lambda: nh.coarse2fine(lev)) - 15
        line += f"  restrict {tf:7.1f}  prolong(exact mode) {tc2:7.1f}"
        cost += (1 + visits) * (tf + tc2)
    tot += cost
    print(line + f"   -> ~{cost/1e3:.2f} ms per F-cycle")
print(f"sum ~{tot/1e3:.2f} ms per F-cycle (API-call estimates)")
