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


def load_ops():
    return np.load(os.path.join(HERE, "golden", "ref_golden_ops.npz"))


def seeded_field(shape, seed):
    """Same integer-arithmetic field as tests/golden/make_ref_golden.py."""
    n = int(np.prod(shape))
    k = (np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(1000003)) * np.uint64(0x9E3779B97F4A7C15)
    k ^= k >> np.uint64(29); k *= np.uint64(0xBF58476D1CE4E5B9); k ^= k >> np.uint64(32)
    return ((k >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5).reshape(shape)


def sub(a, stride=(5, 5, 3)):
    return np.ascontiguousarray(a[::stride[0], ::stride[1], ::stride[2]]).ravel()


def case_ids(cases):
    return [c["name"] for c in cases]
