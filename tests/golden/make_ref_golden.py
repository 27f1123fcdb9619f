"""Generates tests/golden/ref_golden.json + ref_golden_ops.npz by RUNNING THE REFERENCE'S OWN SOURCE on one rank.

The image has no Fortran compiler, so the reference's Fortran routines are translated statement by statement to C by
oracle/f90toc.py (read in place from /root/reference/src, output only under oracle/_ref/, never committed), compiled with
`gcc -O2 -ffp-contract=off` and driven by oracle/refdriver.py.  The whole solver path executes from the reference's text:
nhydro_init (define_grids, define_neighbours, define_gather_informations), nhydro_matrices (define_matrices_topo,
setup_zr_zw, define_matrix), compute_rhs, solve_p / Fcycle / Vcycle, relax / relax_3D_8_heart / tridiag, compute_residual,
fine2coarse, coarse2fine, correct_uvw, and the MPI layer -- fill_halo_2D/3D/3D_relax/4D, gather, split, global_sum -- with
the MPI library calls emulated between threads (oracle/ref/ref_mpi.c; one thread = one rank).  These vectors are therefore
REFERENCE output (of a mechanical translation), not oracle output: they pin oracle/mgoracle.cpp and the CUDA path, on one
rank AND for the 2x1 / 2x2 / 4x2 decompositions the multi-GPU tests and bench.py's in-run parity check use.

Needs /root/reference (this container only).  Run from the repo root:  python tests/golden/make_ref_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mgroms_b200 import testcases as tc  # noqa: E402
from oracle.refdriver import RefSolver, RefWorld  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# the reference's own CPU-runnable case first (BASELINE.json configs[0]: mg_testseamount 64^3, literal Lx = Ly = 10 km,
# u = v = 0, w = -1), then the geometries / namelist switches of the other drivers at sizes that run in milliseconds
CASES = [
    dict(name="c1 seamount 64^3 (mg_testseamount)", kind="seamount", dims=[64, 64, 64], uvw="wminus1", tol=1e-10, maxite=20, params={}),
    dict(name="seamount 32x32x16 dx=156.25", kind="seamount", dims=[32, 32, 16], bench_dx=True, uvw="random", tol=1e-10, maxite=20, params={}),
    dict(name="rndtopo 32x16x16", kind="rndtopo", dims=[32, 16, 16], uvw="random", tol=1e-10, maxite=20, params={}),
    dict(name="cuc 16x32x8", kind="cuc", dims=[16, 32, 8], uvw="random", tol=1e-10, maxite=20, params={}),
    dict(name="cuc 32x32x32 w=-1", kind="cuc", dims=[32, 32, 32], uvw="wminus1", tol=1e-10, maxite=20, params={}),
    dict(name="rndtopo 16^3 bmask", kind="rndtopo", dims=[16, 16, 16], uvw="random", tol=1e-10, maxite=20, params=dict(bmask=True)),
    dict(name="cuc 32x32x16 bmask", kind="cuc", dims=[32, 32, 16], uvw="random", tol=1e-10, maxite=20, params=dict(bmask=True)),
    dict(name="seamount 16^3 FC", kind="seamount", dims=[16, 16, 16], uvw="random", tol=1e-10, maxite=20, params=dict(relax_method="FC")),
    dict(name="rndtopo 16^3 GS", kind="rndtopo", dims=[16, 16, 16], uvw="random", tol=1e-10, maxite=20, params=dict(relax_method="GS")),
    dict(name="cuc 16^3 nearest", kind="cuc", dims=[16, 16, 16], uvw="random", tol=1e-10, maxite=20, params=dict(interp_type="nearest")),
]
OPS_CASES = [   # operator by operator, full arrays (tiny grids)
    dict(name="ops rndtopo 8x16x8", kind="rndtopo", dims=[8, 16, 8], params={}),
    dict(name="ops cuc 16x8x16 bmask", kind="cuc", dims=[16, 8, 16], params=dict(bmask=True)),
]


# the decomposed cases: every configuration of tests/test_gpu_multi.py (without its kernel-forcing variants) and the two
# cases bench.py checks on its live communicator before a multi-GPU measurement, for each process grid it runs on
MULTI_CASES = [
    dict(kind="seamount", dims=[32, 32, 16], grid=[2, 1], uvw="wminus1", params=dict(solver_prec=1e-10, solver_maxiter=20)),
    dict(kind="rndtopo", dims=[16, 32, 16], grid=[2, 1], uvw="random", params=dict(solver_prec=1e-10, solver_maxiter=20, nsmall=16)),
    dict(kind="seamount", dims=[16, 16, 16], grid=[2, 2], uvw="random", params=dict(solver_prec=1e-10, solver_maxiter=20)),
    dict(kind="cuc", dims=[16, 32, 8], grid=[2, 2], uvw="wminus1", params=dict(solver_prec=1e-9, solver_maxiter=20, nsmall=16)),
    dict(kind="seamount", dims=[16, 32, 16], grid=[4, 2], uvw="random", params=dict(solver_prec=1e-10, solver_maxiter=20, nsmall=16)),
    dict(kind="rndtopo", dims=[16, 32, 16], grid=[2, 1], uvw="random", params=dict(solver_prec=1e-10, solver_maxiter=30, bmask=True)),
    dict(kind="seamount", dims=[16, 16, 16], grid=[2, 2], uvw="random", params=dict(solver_prec=1e-10, solver_maxiter=30, bmask=True, nsmall=16)),
    dict(kind="rndtopo", dims=[64, 64, 32], grid=[2, 1], uvw="random", params=dict(solver_prec=1e-10, solver_maxiter=20)),
    dict(kind="seamount", dims=[64, 64, 32], grid=[2, 2], uvw="wminus1", params=dict(solver_prec=1e-10, solver_maxiter=20)),
    dict(kind="rndtopo", dims=[64, 64, 32], grid=[4, 2], uvw="random", params=dict(solver_prec=1e-10, solver_maxiter=20, nsmall=32)),
    dict(kind="cuc", dims=[64, 128, 64], grid=[2, 1], uvw="wminus1", params=dict(solver_prec=1e-9, solver_maxiter=20)),
    dict(kind="cuc", dims=[64, 64, 64], grid=[2, 2], uvw="random", params=dict(solver_prec=1e-9, solver_maxiter=20)),
    dict(kind="cuc", dims=[64, 64, 64], grid=[4, 2], uvw="random", params=dict(solver_prec=1e-9, solver_maxiter=20)),
] + [dict(kind="seamount", dims=[16, 32, 16], grid=g, uvw="random", params=dict(solver_prec=1e-10, solver_maxiter=20, nsmall=16)) for g in ([2, 1], [2, 2])] \
  + [dict(kind="rndtopo", dims=[64, 64, 32], grid=g, uvw="random", params=dict(solver_prec=1e-10, solver_maxiter=20)) for g in ([2, 2], [4, 2])]


def multi_key(kind, dims, grid, params) -> str:
    """Identity of a decomposed case (how tests and bench.py look their vector up)."""
    extras = ",".join(f"{k}={params[k]}" for k in sorted(params) if k not in ("solver_prec", "solver_maxiter"))
    return f"{kind} {dims[0]}x{dims[1]}x{dims[2]} per rank on {grid[0]}x{grid[1]}" + (f" [{extras}]" if extras else "")


def run_multi(c):
    nx, ny, nz = c["dims"]; npx, npy = c["grid"]; N = npx * npy
    PR = RefWorld.PerRank
    prm = dict(c["params"])
    W = RefWorld(nx, ny, nz, npx, npy, **prm)
    gen = {"seamount": tc.seamount, "rndtopo": tc.rndtopo, "cuc": tc.cuc}[c["kind"]]
    vert = tc.CUC_VERT if c["kind"] == "cuc" else tc.SEAMOUNT_VERT
    geos = [gen(nx, ny, nz, npx, npy, r) for r in range(N)]
    rms = PR([tc.bmask_rmask(nx, ny, npx, npy, r) if prm.get("bmask") else None for r in range(N)])
    W.call("matrices", *[PR([g[q] for g in geos]) for q in range(4)], rms, vert["hc"], vert["theta_b"], vert["theta_s"])
    uvw = [tc.uvw_wminus1(nx, ny, nz) if c["uvw"] == "wminus1" else tc.uvw_random(nx, ny, nz, npx, npy, r) for r in range(N)]
    W.call("compute_rhs", rms, *[PR([x[q] for x in uvw]) for q in range(3)])
    b = W.get("b")
    hist = W.call("solve_p", prm["solver_prec"], prm["solver_maxiter"])[0]
    W.call("correct_uvw", *[PR([x[q] for x in uvw]) for q in range(3)])
    p = W.get("p")
    e = dict(c, key=multi_key(c["kind"], c["dims"], c["grid"], c["params"]), nlevs=W.nlevs,
             levels=[{k: v for k, v in W.call("level_info", lev)[0].items() if k != "neighb"} for lev in range(1, W.nlevs + 1)],
             hist_res=[float(x) for x in hist[0]], hist_conv=[float(x) for x in hist[1][1:]],
             sha_cA=[[digest(a) for a in W.get("cA", lev)] for lev in range(1, W.nlevs + 1)],
             sha_b_interior=[digest(x[1:-1, 1:-1]) for x in b], sha_p=[digest(x) for x in p], sha_u=[digest(x[0]) for x in uvw],
             p_max=float(max(np.abs(x).max() for x in p)), sub_stride=stride_of(c["dims"]))
    W.close()
    return e, [sub(x, stride_of(c["dims"])) for x in p]


def digest(a) -> str:
    """sha256 of the IEEE bit patterns (with -0.0 folded into +0.0)."""
    return hashlib.sha256(np.ascontiguousarray(np.asarray(a, dtype=np.float64) + 0.0).tobytes()).hexdigest()


def geometry(c):
    nx, ny, nz = c["dims"]
    if c.get("bench_dx"):
        return tc.bench_geometry(c["kind"], nx, ny, nz, 1, 1, 0)
    fn = {"seamount": tc.seamount, "rndtopo": tc.rndtopo, "cuc": tc.cuc}[c["kind"]]
    return fn(nx, ny, nz, 1, 1, 0), (tc.CUC_VERT if c["kind"] == "cuc" else tc.SEAMOUNT_VERT)


def build_ref(c):
    nx, ny, nz = c["dims"]
    R = RefSolver(nx, ny, nz, **c["params"])
    geo, vert = geometry(c)
    rm = tc.bmask_rmask(nx, ny) if c["params"].get("bmask") else None
    R.matrices(*geo, rm, vert["hc"], vert["theta_b"], vert["theta_s"])
    return R, rm


def stride_of(dims):
    """About 8 x 8 x 6 samples per field whatever its size."""
    nx, ny, nz = dims
    return [max(1, (nx + 2) // 8), max(1, (ny + 2) // 8), max(1, nz // 6)]


def sub(a, st):
    return np.ascontiguousarray(a[::st[0], ::st[1], ::st[2]])


def seeded_field(shape, seed):
    """Deterministic pseudo-random field from integer arithmetic only (bit-identical everywhere, no RNG library)."""
    n = int(np.prod(shape))
    k = (np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(1000003)) * np.uint64(0x9E3779B97F4A7C15)
    k ^= k >> np.uint64(29); k *= np.uint64(0xBF58476D1CE4E5B9); k ^= k >> np.uint64(32)
    return ((k >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5).reshape(shape)


def main():
    out = {"generator": "tests/golden/make_ref_golden.py: the reference's Fortran text translated by oracle/f90toc.py, gcc -O2 -ffp-contract=off, one rank",
           "reference": "CESR-lab/mgroms 0.3.9, /root/reference/src", "cases": []}
    ops = {}
    for c in CASES:
        nx, ny, nz = c["dims"]
        R, rm = build_ref(c)
        u, v, w = tc.uvw_wminus1(nx, ny, nz) if c["uvw"] == "wminus1" else tc.uvw_random(nx, ny, nz, 1, 1, 0)
        R.compute_rhs(rm, u, v, w)
        b = R.get("b").copy()
        res, conv = R.solve_p(c["tol"], c["maxite"])
        R.correct_uvw(u, v, w)
        p = R.get("p")
        e = dict(c)
        e.update(nlevs=R.nlevs, level_dims=[list(R.dims[l]) for l in range(1, R.nlevs + 1)],
                 hist_res=[float(x) for x in res], hist_conv=[float(x) for x in conv[1:]],
                 sha_cA=[digest(R.get("cA", l)) for l in range(1, R.nlevs + 1)], sha_b_interior=digest(b[1:-1, 1:-1]),
                 sha_p=digest(p), sha_u=digest(u), sha_v=digest(v), sha_w=digest(w),
                 p_max=float(np.abs(p).max()), sub_stride=stride_of(c["dims"]), arrays=f"solve{len(out['cases'])}")
        for nm, arr in (("p_sub", p), ("u_sub", u), ("w_sub", w)):
            ops[f"{e['arrays']}/{nm}"] = sub(arr, stride_of(c["dims"]))
        out["cases"].append(e)
        print(f"{c['name']}: {R.nlevs} levels, {len(res) - 1} F-cycles, res {res[-1]:.3e}")
    for n_, c in enumerate(OPS_CASES):
        nx, ny, nz = c["dims"]
        R, rm = build_ref(c)
        u, v, w = tc.uvw_random(nx, ny, nz, 1, 1, 0)
        R.compute_rhs(rm, u, v, w)
        key = f"ops{n_}"
        for lev in range(1, R.nlevs + 1):
            ops[f"{key}/cA{lev}"] = R.get("cA", lev).copy()
        ops[f"{key}/b"] = R.get("b").copy()
        # one operator at a time from a seeded state (p with a Neumann halo, as every caller leaves it)
        from oracle.refdriver import neumann_nh1
        p0 = seeded_field((nx + 2, ny + 2, nz), 11 + n_); neumann_nh1(p0)
        R.get("p")[...] = p0
        R.relax(1, 1); ops[f"{key}/p_after_relax1"] = R.get("p").copy()
        ops[f"{key}/res_after_relax1"] = np.array([R.residual(1)]); ops[f"{key}/r"] = R.get("r").copy()
        R.fine2coarse(1); ops[f"{key}/b2"] = R.get("b", 2).copy()
        p2 = seeded_field(R.get("p", 2).shape, 23 + n_); neumann_nh1(p2)
        R.get("p", 2)[...] = p2
        R.coarse2fine(1); ops[f"{key}/p_after_prolong"] = R.get("p").copy(); ops[f"{key}/r_interp"] = R.get("r").copy()
        R.get("p")[...] = p0
        R.get("b")[...] = ops[f"{key}/b"]
        R.vcycle(1); ops[f"{key}/p_after_vcycle"] = R.get("p").copy()
        out.setdefault("ops_cases", []).append(dict(c, key=key, nlevs=R.nlevs, p0_seed=11 + n_, p2_seed=23 + n_))
    seen = set()
    for c in MULTI_CASES:
        k = multi_key(c["kind"], c["dims"], c["grid"], c["params"])
        if k in seen:
            continue
        seen.add(k)
        e, psubs = run_multi(c)
        e["arrays"] = f"multi{len(out.get('multi_cases', []))}"
        for r, a in enumerate(psubs):
            ops[f"{e['arrays']}/p_sub_r{r}"] = a
        out.setdefault("multi_cases", []).append(e)
        print(f"{e['key']}: {e['nlevs']} levels (gathered: {[l['gather'] for l in e['levels']]}), {len(e['hist_res']) - 1} F-cycles, res {e['hist_res'][-1]:.3e}")
    with open(os.path.join(HERE, "ref_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    np.savez_compressed(os.path.join(HERE, "ref_golden_ops.npz"), **ops)
    print("wrote", len(out["cases"]), "solve cases and", len(OPS_CASES), "operator cases:",
          os.path.getsize(os.path.join(HERE, "ref_golden.json")) // 1024, "KiB json,", os.path.getsize(os.path.join(HERE, "ref_golden_ops.npz")) // 1024, "KiB npz")


if __name__ == "__main__":
    main()
