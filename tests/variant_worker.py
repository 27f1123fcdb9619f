"""Runs one relaxation / solve parity check in a fresh process, so that kernel-selection knobs (environment variables the
library reads once per process: MGB_NO_SMALL_KERNEL, MGB_NO_MID_KERNEL, MGB_NO_FUSED_HALO, MGB_FUSED_RELAX,
MGB_FUSED_MIN_TILES, MGB_F_GRID_CAP, MGB_F_VARIANT ...) can force a code path that the grid size would not select.

    python tests/variant_worker.py KIND NX NY NZ [solve] [bench]     (bench: dx = dy = 156.25 m as in bench.py)

Prints "OK ..." and exits 0 when the CUDA path matches the CPU oracle (bit-exact against its rank-1 formulation)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)

from common import history_close, make_oracle, make_product, neumann_halo, random_interior, relerr, velocities  # noqa: E402


def main():
    kind, nx, ny, nz = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    do_solve = "solve" in sys.argv[5:]
    bench_dx = "bench" in sys.argv[5:]
    from mgroms_b200 import nhydro as nh
    nh.load(); nh.reset_params()
    prm = dict(solver_prec=1e-10, solver_maxiter=20)
    o1 = make_oracle(kind, nx, ny, nz, relax_order=1, bench_dx=bench_dx, **prm)
    make_product(nh, kind, nx, ny, nz, bench_dx=bench_dx, **prm)
    worst = 0.0
    paths = [nh.relax_path(lev) for lev in range(1, o1.nlevs + 1)]
    want = os.environ.get("MGB_EXPECT_PATH")
    if want and not paths[0].startswith(want):
        print(f"FAIL level 1 takes the path '{paths[0]}', expected '{want}'"); sys.exit(1)
    for lev in range(1, o1.nlevs + 1):
        shp = o1.get("p", lev).shape
        p = neumann_halo(random_interior(shp, 30 + lev)); b = random_interior(shp, 40 + lev)
        o1.set("p", p, lev); o1.set("b", b, lev); o1.relax(lev, 2)
        nh.set_field("p", p, lev); nh.set_field("b", b, lev); nh.relax(lev, 2)
        g = nh.get_field("p", lev); ref = o1.get("p", lev)
        if not np.array_equal(g, ref):
            bad = np.argwhere(g != ref)
            print(f"FAIL relax lev {lev}: {len(bad)} of {g.size} differ, rel {relerr(g, ref):.3e}, first (i,j,k) {bad[:5].tolist()}")
            sys.exit(1)
        worst = max(worst, relerr(g, ref))
    msg = f"relax bit-exact on {o1.nlevs} levels"
    if do_solve:
        o1.compute_rhs(); o1.solve_p()
        nh.compute_rhs(None, *velocities("wminus1", nx, ny, nz)); nh.solve_p()
        hg, _ = nh.history(); h1, _ = o1.history()
        history_close(hg, h1, rtol=1e-12, floor=1e-16)
        e = relerr(nh.get_field("p"), o1.get("p"))
        assert e < 1e-12, e
        msg += f"; solve {len(hg) - 1} cycles res {hg[-1]:.2e} p_rel {e:.1e}"
    nh.nhydro_clean()
    print("OK", msg, "paths", paths, "launches", nh.launch_count())


if __name__ == "__main__":
    main()
