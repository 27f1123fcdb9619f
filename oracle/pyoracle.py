"""ctypes front-end of the CPU oracle (oracle/mgoracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package mgroms_b200 never does.  Parity: pinned against the translated reference, one rank and
decomposed (oracle/refdriver.py, tests/test_cpu_ref_pin.py) -- see mgoracle.cpp header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_NATIVE = False
_dp = C.POINTER(C.c_double)


def use_native_build(on: bool = True):
    """CPU-baseline legs of bench.py: compile the oracle -O3 -march=native ON THE MACHINE THAT RUNS IT
    instead of loading the portable libmgoracle.so that travels with the snapshot.  Must be called before the first lib()."""
    global _NATIVE
    assert _LIB is None or _NATIVE == on, "the oracle library is already loaded"
    _NATIVE = on


def _native_so() -> str:
    import hashlib
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("flags"):
                    flags = ln; break
    except OSError:
        pass
    src = os.path.join(_HERE, "mgoracle.cpp")
    tag = hashlib.sha1((flags + str(os.path.getmtime(src))).encode()).hexdigest()[:12]
    d = os.path.join(_HERE, "_native"); os.makedirs(d, exist_ok=True)
    so = os.path.join(d, f"libmgoracle_{tag}.so")
    if not os.path.exists(so):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        tmp = so + f".{os.getpid()}.tmp"
        subprocess.check_call([cxx, "-O3", "-march=native", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-w", "-shared", "-o", tmp, src])
        os.replace(tmp, so)
    return so


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libmgoracle.so")
    src = os.path.join(_HERE, "mgoracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(_native_so() if _NATIVE else build())
        L.mgo_create.restype = C.c_void_p
        L.mgo_create.argtypes = [C.c_int] * 5 + [C.c_char_p]
        L.mgo_last_error.restype = C.c_char_p
        L.mgo_last_error.argtypes = [C.c_void_p]
        L.mgo_field_size.restype = C.c_longlong
        L.mgo_solve_seconds.restype = C.c_double
        L.mgo_bnorm.restype = C.c_double
        for name in ("mgo_destroy", "mgo_nlevs", "mgo_matrices", "mgo_compute_rhs", "mgo_correct_uvw", "mgo_solve_p",
                     "mgo_nhydro_solve", "mgo_fcycle", "mgo_solve_seconds", "mgo_bnorm"):
            getattr(L, name).argtypes = [C.c_void_p] + ([C.c_double] * 3 if name == "mgo_matrices" else [])
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


class Oracle:
    """One emulated MPI world (npx*npy ranks in this process) running the restated reference path."""

    def __init__(self, nx, ny, nz, npx=1, npy=1, namelist: str | None = None, **params):
        L = lib()
        self.L = L
        self.nx, self.ny, self.nz, self.npx, self.npy = nx, ny, nz, npx, npy
        self.nranks = npx * npy
        self.h = C.c_void_p(L.mgo_create(nx, ny, nz, npx, npy, (namelist or "").encode()))
        for k, v in params.items():
            if isinstance(v, bool):
                v = ".true." if v else ".false."
            if L.mgo_set_param(self.h, k.encode(), str(v).encode()) != 0:
                raise KeyError(k)
        if L.mgo_init(self.h, nx, ny, nz) != 0:
            raise RuntimeError(L.mgo_last_error(self.h).decode())
        self.nlevs = L.mgo_nlevs(self.h)

    def close(self):
        if self.h:
            self.L.mgo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- topology
    def level_info(self, lev, rank=0):
        info = (C.c_int * 20)()
        if self.L.mgo_level_info(self.h, rank, lev, info) != 0:
            raise IndexError(lev)
        keys = ["nx", "ny", "nz", "npx", "npy", "incx", "incy", "gather", "ngx", "ngy", "key", "color"]
        d = {k: info[q] for q, k in enumerate(keys)}
        d["neighb"] = [info[12 + q] for q in range(8)]
        return d

    def gather_members(self, lev, rank=0):
        out = (C.c_int * 4)()
        n = self.L.mgo_gather_members(self.h, rank, lev, out)
        return [out[q] for q in range(n)]

    # ---- phases of the public API (nhydro.f90)
    def set_geometry(self, rank, dx, dy, zeta, h, rmask=None):
        self.L.mgo_set_geometry(self.h, rank, _p(dx), _p(dy), _p(zeta), _p(h), _p(rmask) if rmask is not None else None)

    def matrices(self, hc, theta_b, theta_s):
        self.L.mgo_matrices(self.h, float(hc), float(theta_b), float(theta_s))

    def set_uvw(self, rank, u, v, w):
        self.L.mgo_set_uvw(self.h, rank, _p(u), _p(v), _p(w))

    def get_uvw(self, rank=0):
        nx, ny, nz = self.nx, self.ny, self.nz
        u = np.empty((nz, ny + 2, nx + 1)); v = np.empty((nz, ny + 1, nx + 2)); w = np.empty((nz + 1, ny + 2, nx + 2))
        self.L.mgo_get_uvw(self.h, rank, _p(u), _p(v), _p(w))
        return u, v, w

    def compute_rhs(self): self.L.mgo_compute_rhs(self.h)
    def correct_uvw(self): self.L.mgo_correct_uvw(self.h)
    def solve_p(self): self.L.mgo_solve_p(self.h)
    def nhydro_solve(self): self.L.mgo_nhydro_solve(self.h)
    def fcycle(self): self.L.mgo_fcycle(self.h)
    def vcycle(self, lev1): self.L.mgo_vcycle(self.h, lev1)
    def relax(self, lev, nsweeps): self.L.mgo_relax(self.h, lev, nsweeps)
    def fine2coarse(self, lev): self.L.mgo_fine2coarse(self.h, lev)
    def coarse2fine(self, lev): self.L.mgo_coarse2fine(self.h, lev)
    def gather(self, lev): self.L.mgo_gather(self.h, lev)
    def split(self, lev): self.L.mgo_split(self.h, lev)

    def fill_halo(self, lev, name):
        if self.L.mgo_fill_halo(self.h, lev, name.encode()) != 0:
            raise KeyError(name)

    def residual(self, lev):
        r = C.c_double()
        self.L.mgo_residual(self.h, lev, C.byref(r))
        return r.value

    def set_param(self, key, val):
        if self.L.mgo_set_param(self.h, key.encode(), str(val).encode()) != 0:
            raise KeyError(key)

    # ---- fields: returned in C order, i slowest: p/b/r -> (nx+2, ny+2, nz); cA -> (nx+2, ny+2, nz, 8)
    def _shape(self, lev, name, rank):
        g = self.level_info(lev, rank)
        nx, ny, nz = g["nx"], g["ny"], g["nz"]
        if name == "cA": return (nx + 2, ny + 2, nz, 8)
        if name in ("p", "b", "r", "zxdy", "zydx"): return (nx + 2, ny + 2, nz)
        if name in ("cw", "dzw"): return (nx + 2, ny + 2, nz + 1)
        if name == "zr": return (nx + 4, ny + 4, nz)
        if name == "zw": return (nx + 4, ny + 4, nz + 1)
        if name == "dummy3": return (nx // g["ngx"] + 2, ny // g["ngy"] + 2, nz)
        if name in ("dx", "dy", "zeta", "h"): return (nx + 2, ny + 2)
        raise KeyError(name)

    def get(self, name, lev=1, rank=0):
        out = np.empty(self._shape(lev, name, rank))
        n = self.L.mgo_field_size(self.h, rank, lev, name.encode())
        assert n == out.size, (name, n, out.size)
        if self.L.mgo_get_field(self.h, rank, lev, name.encode(), _p(out)) != 0:
            raise KeyError(name)
        return out

    def set(self, name, arr, lev=1, rank=0):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.shape == self._shape(lev, name, rank), (arr.shape, self._shape(lev, name, rank))
        if self.L.mgo_set_field(self.h, rank, lev, name.encode(), _p(arr)) != 0:
            raise KeyError(name)

    def history(self):
        res = (C.c_double * 512)(); conv = (C.c_double * 512)()
        n = self.L.mgo_get_history(self.h, res, conv, 512)
        return np.array(res[:n]), np.array(conv[:n])

    @property
    def solve_seconds(self): return self.L.mgo_solve_seconds(self.h)
    @property
    def bnorm(self): return self.L.mgo_bnorm(self.h)


def make_case(kind, nx, ny, nz, npx=1, npy=1, uvw="wminus1", bench_dx=False, **params):
    """Oracle world with geometry + matrices + velocities of a named synthetic case installed."""
    from mgroms_b200 import testcases as tc   # input generators only (pure numpy), not the product path
    o = Oracle(nx, ny, nz, npx, npy, **params)
    vert = None
    for r in range(npx * npy):
        if bench_dx:
            (dx, dy, zeta, h), vert = tc.bench_geometry(kind, nx, ny, nz, npx, npy, r)
        elif kind == "seamount":
            dx, dy, zeta, h = tc.seamount(nx, ny, nz, npx, npy, r); vert = tc.SEAMOUNT_VERT
        elif kind == "rndtopo":
            dx, dy, zeta, h = tc.rndtopo(nx, ny, nz, npx, npy, r); vert = tc.SEAMOUNT_VERT
        elif kind == "cuc":
            dx, dy, zeta, h = tc.cuc(nx, ny, nz, npx, npy, r); vert = tc.CUC_VERT
        else:
            raise ValueError(kind)
        o.set_geometry(r, dx, dy, zeta, h, tc.bmask_rmask(nx, ny, npx, npy, r) if params.get("bmask") else None)
        if uvw == "wminus1":
            o.set_uvw(r, *tc.uvw_wminus1(nx, ny, nz))
        else:
            o.set_uvw(r, *tc.uvw_random(nx, ny, nz, npx, npy, r))
    o.matrices(vert["hc"], vert["theta_b"], vert["theta_s"])
    return o
