"""Access to the committed REFERENCE golden vectors (tests/golden/ref_golden.json + ref_golden_ops.npz, produced by
tests/golden/make_ref_golden.py from the reference's own Fortran text) and the digest they were hashed with."""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(np.asarray(a, dtype=np.float64) + 0.0).tobytes()).hexdigest()


def load():
    with open(os.path.join(HERE, "golden", "ref_golden.json")) as f:
        return json.load(f)


_ARR = None


def load_ops():
    """Operator-level arrays and the sub-sampled fields of the solve cases (keys `<case arrays>/<name>`)."""
    global _ARR
    if _ARR is None:
        _ARR = np.load(os.path.join(HERE, "golden", "ref_golden_ops.npz"))
    return _ARR


def multi_key(kind, dims, grid, params) -> str:
    """Same identity string as tests/golden/make_ref_golden.py::multi_key."""
    extras = ",".join(f"{k}={params[k]}" for k in sorted(params) if k not in ("solver_prec", "solver_maxiter"))
    return f"{kind} {dims[0]}x{dims[1]}x{dims[2]} per rank on {grid[0]}x{grid[1]}" + (f" [{extras}]" if extras else "")


def multi_case(kind, dims, grid, params):
    """The committed reference vector of a decomposed case, or None."""
    k = multi_key(kind, list(dims), list(grid), params)
    for c in load().get("multi_cases", []):
        if c["key"] == k:
            return c
    return None


def seeded_field(shape, seed):
    """Same integer-arithmetic field as tests/golden/make_ref_golden.py."""
    n = int(np.prod(shape))
    k = (np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(1000003)) * np.uint64(0x9E3779B97F4A7C15)
    k ^= k >> np.uint64(29); k *= np.uint64(0xBF58476D1CE4E5B9); k ^= k >> np.uint64(32)
    return ((k >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5).reshape(shape)


def sub(a, stride):
    return np.ascontiguousarray(a[::stride[0], ::stride[1], ::stride[2]])


def case_ids(cases):
    return [c["name"] for c in cases]
