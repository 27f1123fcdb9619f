"""Generates tests/golden/ref_golden.json + ref_golden_ops.npz by RUNNING THE REFERENCE'S OWN SOURCE on one rank.

The image has no Fortran compiler, so the reference's Fortran routines are translated statement by statement to C by
oracle/f90toc.py (read in place from /root/reference/src, output only under oracle/_ref/, never committed), compiled with
`gcc -O2 -ffp-contract=off` and driven by oracle/refdriver.py.  Everything arithmetic -- define_matrix, setup_zr_zw,
compute_rhs, relax / relax_3D_8_heart / tridiag, compute_residual, fine2coarse, coarse2fine, solve_p / Fcycle / Vcycle,
correct_uvw -- executes from the reference's text; the one-rank halo fills and the level loop of define_matrices_topo are
restated in refdriver.py (copies, 2A - B, 2x2 averages).  These vectors are therefore REFERENCE output (of a mechanical
translation), not oracle output: they are what pins oracle/mgoracle.cpp and the CUDA path.

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
from oracle.refdriver import RefSolver  # noqa: E402

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


def sub(a):
    return np.ascontiguousarray(a[::5, ::5, ::3])


def seeded_field(shape, seed):
    """Deterministic pseudo-random field from integer arithmetic only (bit-identical everywhere, no RNG library)."""
    n = int(np.prod(shape))
    k = (np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(1000003)) * np.uint64(0x9E3779B97F4A7C15)
    k ^= k >> np.uint64(29); k *= np.uint64(0xBF58476D1CE4E5B9); k ^= k >> np.uint64(32)
    return ((k >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5).reshape(shape)


def main():
    out = {"generator": "tests/golden/make_ref_golden.py: the reference's Fortran text translated by oracle/f90toc.py, gcc -O2 -ffp-contract=off, one rank",
           "reference": "CESR-lab/mgroms 0.3.9, /root/reference/src", "cases": []}
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
                 p_max=float(np.abs(p).max()), p_sub=sub(p).ravel().tolist(), u_sub=sub(u).ravel().tolist(), w_sub=sub(w).ravel().tolist(),
                 sub_stride=[5, 5, 3])
        out["cases"].append(e)
        print(f"{c['name']}: {R.nlevs} levels, {len(res) - 1} F-cycles, res {res[-1]:.3e}")
    ops = {}
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
    with open(os.path.join(HERE, "ref_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    np.savez_compressed(os.path.join(HERE, "ref_golden_ops.npz"), **ops)
    print("wrote", len(out["cases"]), "solve cases and", len(OPS_CASES), "operator cases:",
          os.path.getsize(os.path.join(HERE, "ref_golden.json")) // 1024, "KiB json,", os.path.getsize(os.path.join(HERE, "ref_golden_ops.npz")) // 1024, "KiB npz")


if __name__ == "__main__":
    main()
