"""The CUDA path (through the C ABI) against REFERENCE output: the committed vectors of tests/golden/ref_golden.* were
produced by the reference's own Fortran routines (translated statement by statement, oracle/f90toc.py) on one rank.

Bit-exact: stencil coefficients of every level, right-hand side, residual, restriction, prolongation.
North-star tolerances: per-cycle residual norms 1e-10 relative (+ the FP64 floor of a reordered sweep), final pressure 1e-9
relative, corrected velocities 1e-9 -- the smoother runs in the rank-1 form, so these are not bit-equal.
"""
import numpy as np
import pytest

from common import history_close, relerr
from refgolden import digest, load, load_ops, seeded_field, sub

pytestmark = pytest.mark.gpu
G = load()


def product_case(nh, c):
    from mgroms_b200 import testcases as tc
    nx, ny, nz = c["dims"]
    nh.reset_params()
    nh.set_params(solver_prec=c.get("tol", 1e-10), solver_maxiter=c.get("maxite", 20), **c["params"])
    nh.nhydro_init(nx, ny, nz, 1, 1)
    if c.get("bench_dx"):
        geo, vert = tc.bench_geometry(c["kind"], nx, ny, nz, 1, 1, 0)
    else:
        geo = {"seamount": tc.seamount, "rndtopo": tc.rndtopo, "cuc": tc.cuc}[c["kind"]](nx, ny, nz, 1, 1, 0)
        vert = tc.CUC_VERT if c["kind"] == "cuc" else tc.SEAMOUNT_VERT
    rm = tc.bmask_rmask(nx, ny) if c["params"].get("bmask") else None
    nh.nhydro_matrices(*geo, rm, vert["hc"], vert["theta_b"], vert["theta_s"])
    return rm


@pytest.mark.parametrize("c", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_cuda_path_against_reference_vectors(nh, c):
    from mgroms_b200 import testcases as tc
    nx, ny, nz = c["dims"]
    rm = product_case(nh, c)
    assert nh.nlevs() == c["nlevs"]
    for lev in range(1, c["nlevs"] + 1):
        assert digest(nh.get_field("cA", lev)) == c["sha_cA"][lev - 1], ("cA", lev)
    u, v, w = tc.uvw_wminus1(nx, ny, nz) if c["uvw"] == "wminus1" else tc.uvw_random(nx, ny, nz, 1, 1, 0)
    nh.nhydro_solve(nx, ny, nz, rm, u, v, w)          # compute_rhs + solve_p + correct_uvw, host arrays in and out
    assert digest(nh.get_field("b")[1:-1, 1:-1]) == c["sha_b_interior"]
    history_close(nh.history()[0], np.array(c["hist_res"]))
    st, arr = c["sub_stride"], load_ops()
    assert np.abs(sub(nh.get_field("p"), st) - arr[c["arrays"] + "/p_sub"]).max() <= 1e-9 * c["p_max"]
    for got, ref in ((u, arr[c["arrays"] + "/u_sub"]), (w, arr[c["arrays"] + "/w_sub"])):
        assert np.abs(sub(got, st) - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("oc", G["ops_cases"], ids=[c["name"] for c in G["ops_cases"]])
def test_cuda_operators_against_reference_vectors(nh, oc):
    from mgroms_b200 import testcases as tc
    from oracle.refdriver import neumann_nh1
    ops = load_ops(); key = oc["key"]
    nx, ny, nz = oc["dims"]
    rm = product_case(nh, dict(oc, tol=1e-10, maxite=20))
    for lev in range(1, oc["nlevs"] + 1):
        assert np.array_equal(nh.get_field("cA", lev), ops[f"{key}/cA{lev}"]), lev
    u, v, w = tc.uvw_random(nx, ny, nz, 1, 1, 0)
    nh.compute_rhs(rm, u, v, w)
    assert np.array_equal(nh.get_field("b")[1:-1, 1:-1], ops[f"{key}/b"][1:-1, 1:-1])
    p0 = seeded_field((nx + 2, ny + 2, nz), oc["p0_seed"]); neumann_nh1(p0)
    nh.set_field("b", ops[f"{key}/b"]); nh.set_field("p", p0)
    nh.relax(1, 1)                                      # rank-1 form of the sweep: round-off level, not bit-equal
    assert relerr(nh.get_field("p"), ops[f"{key}/p_after_relax1"]) < 1e-12
    nh.set_field("p", ops[f"{key}/p_after_relax1"])
    res = nh.compute_residual(1)
    assert np.array_equal(nh.get_field("r"), ops[f"{key}/r"])
    assert abs(res - float(ops[f"{key}/res_after_relax1"][0])) <= 1e-13 * res
    nh.fine2coarse(1)
    assert np.array_equal(nh.get_field("b", 2), ops[f"{key}/b2"])
    p2 = seeded_field(nh.get_field("p", 2).shape, oc["p2_seed"]); neumann_nh1(p2)
    nh.set_field("p", p2, 2)
    nh.coarse2fine(1)
    assert np.array_equal(nh.get_field("p"), ops[f"{key}/p_after_prolong"])
    nh.set_field("p", p0); nh.set_field("b", ops[f"{key}/b"])
    nh.Vcycle(1)
    assert relerr(nh.get_field("p"), ops[f"{key}/p_after_vcycle"]) < 1e-11


def _ref_available():
    from oracle import refdriver
    return refdriver.available()


@pytest.mark.skipif(not _ref_available(), reason="needs the prebuilt oracle/_ref/libmgref.so (oracle/build_ref.py, built where /root/reference exists)")
@pytest.mark.parametrize("kind,nx,ny,nz,prm", [("rndtopo", 128, 64, 32, {}), ("cuc", 64, 128, 64, {}), ("seamount", 128, 128, 16, dict(bmask=True))])
def test_cuda_path_against_the_live_translated_reference(nh, kind, nx, ny, nz, prm):
    """Bigger grids than the committed vectors (level 1 leaves the small / cooperative kernels): the translated reference
    runs beside the CUDA path.  Coefficients, rhs bit-exact; history, p and u,v,w within the north-star tolerances."""
    from mgroms_b200 import testcases as tc
    from oracle.refdriver import RefSolver
    c = dict(kind=kind, dims=[nx, ny, nz], params=prm, tol=1e-10, maxite=12)
    rm = product_case(nh, c)
    geo = {"seamount": tc.seamount, "rndtopo": tc.rndtopo, "cuc": tc.cuc}[kind](nx, ny, nz, 1, 1, 0)
    vert = tc.CUC_VERT if kind == "cuc" else tc.SEAMOUNT_VERT
    R = RefSolver(nx, ny, nz, **prm)
    R.matrices(*geo, rm, vert["hc"], vert["theta_b"], vert["theta_s"])
    assert nh.nlevs() == R.nlevs
    for lev in range(1, R.nlevs + 1):
        for name in ("zr", "zw", "cw", "cA"):
            assert np.array_equal(nh.get_field(name, lev), R.get(name, lev)), (name, lev)
    u, v, w = tc.uvw_random(nx, ny, nz, 1, 1, 0)
    ur, vr, wr = u.copy(), v.copy(), w.copy()
    R.compute_rhs(rm, ur, vr, wr)
    hr, _ = R.solve_p(1e-10, 12)
    R.correct_uvw(ur, vr, wr)
    nh.nhydro_solve(nx, ny, nz, rm, u, v, w)
    assert np.array_equal(nh.get_field("b")[1:-1, 1:-1], R.get("b")[1:-1, 1:-1])
    history_close(nh.history()[0], hr)
    assert relerr(nh.get_field("p"), R.get("p")) < 1e-9
    for a, b in ((u, ur), (v, vr), (w, wr)):
        assert np.abs(a - b).max() <= 1e-9 * max(1.0, np.abs(b).max())
