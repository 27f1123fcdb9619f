"""Generates tests/golden/oracle_histories.json from the CPU ORACLE.

These are regression fixtures of the oracle (so that an accidental change of the restatement is caught), NOT reference
output: the reference (Fortran 90 + MPI) cannot be compiled in this image and ships no golden vectors of its own.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import make_case  # noqa: E402

CASES = [
    dict(name="seamount32 1x1 RB", kind="seamount", local=[32, 32, 32], npx=1, npy=1, params=dict(solver_prec=1e-12)),
    dict(name="seamount32 1x1 rank1-form", kind="seamount", local=[32, 32, 32], npx=1, npy=1, params=dict(solver_prec=1e-12, relax_order=1)),
    dict(name="rndtopo 32x32x16 2x2", kind="rndtopo", local=[16, 16, 16], npx=2, npy=2, params=dict(solver_prec=1e-10)),
    dict(name="cuc 32x16x8 2x1 gathered", kind="cuc", local=[16, 16, 8], npx=2, npy=1, params=dict(solver_prec=1e-10, nsmall=16)),
    dict(name="seamount 16^3 FC nearest", kind="seamount", local=[16, 16, 16], npx=1, npy=1, params=dict(solver_prec=1e-8, relax_method="FC", interp_type="nearest")),
    dict(name="seamount 32x16x16 GS", kind="seamount", local=[32, 16, 16], npx=1, npy=1, params=dict(solver_prec=1e-9, relax_method="GS")),
    dict(name="seamount 32x32x16 bmask", kind="seamount", local=[32, 32, 16], npx=1, npy=1, params=dict(solver_prec=1e-10, bmask=True)),
    dict(name="rndtopo 32x32x16 2x2 bmask", kind="rndtopo", local=[16, 16, 16], npx=2, npy=2, params=dict(solver_prec=1e-10, bmask=True)),
]
out = {"generator": "tests/golden/make_golden.py (CPU oracle, -ffp-contract=off)", "cases": []}
for c in CASES:
    o = make_case(c["kind"], *c["local"], c["npx"], c["npy"], **c["params"])
    o.compute_rhs(); o.solve_p()
    h, _ = o.history()
    c = dict(c); c["history"] = [float(x) for x in h]; c["p_abs_sum"] = float(np.abs(o.get("p")).sum())
    out["cases"].append(c)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_histories.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out["cases"]), "cases")
