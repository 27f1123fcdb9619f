"""refdriver.py -- driver of the TRANSLATED reference (oracle/_ref/libmgref.so, built by oracle/build_ref.py).

TEST INFRASTRUCTURE (oracle/): used by tests/ and tests/golden/make_ref_golden.py only.

Everything on the solver path runs from the reference's own source text (translated statement by statement to C by
oracle/f90toc.py): nhydro_init -> define_grids / define_neighbours / define_gather_informations (src/mg_grids.f90),
nhydro_matrices -> define_matrices_topo / setup_zr_zw / define_matrix, compute_rhs, solve_p / Fcycle / Vcycle, relax +
relax_3D_8_{RB,FC,GS} + relax_3D_8_heart + tridiag, compute_residual, fine2coarse, coarse2fine, correct_uvw, AND the MPI layer:
fill_halo_2D / _3D / _3D_relax / _4D (src/mg_mpi_exchange.f90), gather / split (src/mg_gather.f90), global_sum.  The MPI calls
themselves (Isend / Irecv / Waitany / Allreduce / Allgather / Comm_split) are emulated between THREADS by oracle/ref/ref_mpi.c:
one Python thread = one rank, the translated module variables are thread-local.  Not translated: namelist reading (the
values are set through setters), NetCDF output, timers.

Arrays are NumPy C-order with reversed dimensions, e.g. Fortran p(nz,0:ny+1,0:nx+1) <-> p[i, j, k-1] of shape (nx+2, ny+2, nz).

    RefSolver(nx, ny, nz, **namelist)                     one rank, runs in the calling thread
    RefWorld(nx, ny, nz, npx, npy, **namelist)            npx*npy ranks = threads; every call is collective
"""
from __future__ import annotations

import ctypes as C
import json
import os
import queue
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class farr(C.Structure):
    _fields_ = [("a", C.POINTER(C.c_double)), ("lb", C.c_int * 5), ("n", C.c_int * 5)]


class iarr(C.Structure):
    _fields_ = [("a", C.POINTER(C.c_int)), ("lb", C.c_int * 5), ("n", C.c_int * 5)]


_lib = None
_types = {}
_FAST = False


def use_fast_build(on: bool = True):
    """bench.py's CPU arm: the -O3 / AVX2+FMA build of the same translated text (timing only, never used as a checker)."""
    global _FAST, _lib
    if on != _FAST:
        _FAST, _lib = on, None


def available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libmgref.so")) or os.path.isdir("/root/reference/src")


def _ctype_of(kind: str):
    if kind == "int":
        return C.c_int
    if kind == "double":
        return C.c_double
    if kind == "farr":
        return farr
    if kind == "iarr":
        return iarr
    if kind.startswith("int["):
        return C.c_int * int(kind[4:-1])
    raise ValueError(kind)


def lib():
    global _lib
    if _lib is None:
        import sys
        sys.path.insert(0, HERE)
        import build_ref
        so = os.path.join(HERE, "_ref", "libmgref.so")
        stale = os.path.exists(so) and os.path.getmtime(so) < max(os.path.getmtime(os.path.join(HERE, f)) for f in
                                                                   ("f90toc.py", "build_ref.py", "ref/ref_prelude.h", "ref/ref_harness.c", "ref/ref_mpi.c"))
        if not os.path.exists(so) or not os.path.exists(so.replace("libmgref.so", "libmgref_fast.so")) or os.environ.get("MGB_REF_REBUILD") \
                or (stale and os.path.isdir("/root/reference/src")):
            so = build_ref.build(quiet=True)
        with open(os.path.join(HERE, "_ref", "layout.json")) as f:
            lay = json.load(f)
        for tname, members in lay["types"].items():          # ctypes mirror of grid_type / mpi_buffers (generated layout)
            if tname not in _types:
                _types[tname] = type(tname, (C.Structure,), {"_fields_": [(n, _ctype_of(k)) for n, k in members]})
        L = C.CDLL(so.replace("libmgref.so", "libmgref_fast.so") if _FAST else so)
        D, I, PD = C.c_double, C.c_int, C.POINTER(C.c_double)
        L.ref_grid.restype = C.POINTER(_types["grid_type"]); L.ref_grid.argtypes = [I]
        L.ref_nhydro_init.argtypes = [I] * 5
        L.ref_nhydro_matrices.argtypes = [farr] * 5 + [D] * 3
        L.ref_nhydro_solve.argtypes = [I, I, I, PD, PD, PD, PD]
        L.ref_compute_rhs.argtypes = [farr] * 4
        L.ref_correct_uvw.argtypes = [farr] * 4
        L.ref_relax.argtypes = [I, I]
        L.ref_compute_residual.argtypes = [I, PD]
        L.ref_fine2coarse.argtypes = [I]
        L.ref_coarse2fine.argtypes = [I]
        L.ref_solve_p.argtypes = [D, I]
        L.ref_vcycle.argtypes = [I]
        L.ref_fill_halo_3d_relax.argtypes = [I, PD, C.POINTER(I), C.POINTER(I), C.POINTER(I)]
        L.ref_fill_halo_3d.argtypes = [I, farr, C.c_char_p]
        L.ref_fill_halo_2d.argtypes = [I, farr]
        L.ref_gather_3d.argtypes = [I, farr, farr]
        L.ref_split.argtypes = [I, farr, farr]
        L.ref_set_int.argtypes = [C.c_char_p, I]; L.ref_get_int.argtypes = [C.c_char_p]
        L.ref_set_real.argtypes = [C.c_char_p, D]
        L.ref_set_str.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_history.argtypes = [PD, PD, I]
        L.ref_setup_seamount.argtypes = [I] * 5 + [D] * 3 + [farr] * 4
        L.rmpi_world_init.argtypes = [I]; L.rmpi_attach.argtypes = [I]
        _lib = L
    return _lib


def desc(a: np.ndarray, lb) -> farr:
    """Descriptor of a C-order array whose REVERSED shape is the Fortran shape, with Fortran lower bounds `lb`."""
    assert a.flags.c_contiguous and a.dtype == np.float64
    d = farr()
    d.a = a.ctypes.data_as(C.POINTER(C.c_double))
    shp = a.shape[::-1]
    for q in range(5):
        d.lb[q] = lb[q] if q < len(shp) else 0
        d.n[q] = shp[q] if q < len(shp) else 0
    return d


def view(d: farr) -> np.ndarray:
    """NumPy view (no copy) of an array owned by the translated reference."""
    rank = sum(1 for q in range(5) if d.n[q] > 0)
    return np.ctypeslib.as_array(d.a, shape=tuple(d.n[q] for q in range(rank))[::-1])


def neumann_nh1(A):
    """Homogeneous-Neumann ring of a (nx+2, ny+2[, nz]) array: what fill_halo does on one rank (used by tests to prepare
    inputs; the solver path itself runs the translated fill_halo_*)."""
    nx, ny = A.shape[0] - 2, A.shape[1] - 2
    A[1:nx + 1, 0] = A[1:nx + 1, 1]; A[nx + 1, 1:ny + 1] = A[nx, 1:ny + 1]
    A[1:nx + 1, ny + 1] = A[1:nx + 1, ny]; A[0, 1:ny + 1] = A[1, 1:ny + 1]
    A[0, 0] = A[1, 1]; A[nx + 1, 0] = A[nx, 1]; A[nx + 1, ny + 1] = A[nx, ny]; A[0, ny + 1] = A[1, ny]


class _Rank:
    """The reference API on one emulated rank.  Every method must run in the thread that owns the rank."""

    def __init__(self, rank):
        self.rank, self.L, self.keep = rank, lib(), []

    def init(self, nx, ny, nz, npx, npy, prm):
        L = self.L
        L.rmpi_attach(self.rank)
        L.ref_reset_state()
        ints = dict(bmask=int(bool(prm.get("bmask", False))), ns_pre=prm.get("ns_pre", 3), ns_post=prm.get("ns_post", 2),
                    ns_coarsest=prm.get("ns_coarsest", 40), nsmall=prm.get("nsmall", 8), netcdf_output=0, aggressive=0,
                    solver_maxiter=prm.get("solver_maxiter", 50))
        for k, v in ints.items():
            assert L.ref_set_int(k.encode(), int(v)) == 0, k
        assert L.ref_set_real(b"solver_prec", float(prm.get("solver_prec", 1e-12))) == 0
        for k, dflt in (("cmatrix", "real"), ("relax_method", "RB"), ("interp_type", "linear")):
            assert L.ref_set_str(k.encode(), str(prm.get(k, dflt)).encode()) == 0
        self.bmask = bool(prm.get("bmask", False))
        self.nx, self.ny, self.nz = nx, ny, nz
        L.ref_nhydro_init(nx, ny, nz, npx, npy)           # define_grids + define_neighbours (+ gather informations)
        self.nlevs = L.ref_get_int(b"nlevs")
        self.dims = {lev: (L.ref_grid(lev).contents.nx, L.ref_grid(lev).contents.ny, L.ref_grid(lev).contents.nz) for lev in range(1, self.nlevs + 1)}
        return self.nlevs

    def setup_seamount(self, npx, npy, Lx=1.0e4, Ly=1.0e4, Htot=4.0e3):
        """dx, dy, zeta, h of this rank from the reference's own test set-up (src/mg_setup_tests.f90:108-157)."""
        out = [np.zeros((self.nx + 2, self.ny + 2)) for _ in range(4)]
        self.L.ref_setup_seamount(self.nx, self.ny, self.nz, npx, npy, Lx, Ly, Htot, *[desc(a, (0, 0)) for a in out])
        return out

    def level_info(self, lev):
        g = self.L.ref_grid(lev).contents
        return dict(nx=g.nx, ny=g.ny, nz=g.nz, npx=g.npx, npy=g.npy, incx=g.incx, incy=g.incy, gather=g.gather, ngx=g.ngx, ngy=g.ngy,
                    key=g.key, neighb=[g.neighb[q] for q in range(1, 9)])

    def get(self, name, lev=1):
        return view(getattr(self.L.ref_grid(lev).contents, "ca" if name == "cA" else name.lower()))

    def matrices(self, dx, dy, zeta, h, rmask, hc, theta_b, theta_s):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (dx, dy, zeta, h)]
        rm = np.ascontiguousarray(rmask, dtype=np.float64) if rmask is not None else np.ones((self.nx + 2, self.ny + 2))
        self.keep = arrs + [rm]
        self.L.ref_nhydro_matrices(*[desc(a, (0, 0)) for a in arrs], desc(rm, (0, 0)), hc, theta_b, theta_s)

    def _uvw(self, u, v, w):
        return desc(u, (1, 0, 1)), desc(v, (0, 1, 1)), desc(w, (0, 0, 0))

    def compute_rhs(self, rmask, u, v, w):
        self._rm = np.ascontiguousarray(rmask, dtype=np.float64) if rmask is not None else np.ones((self.nx + 2, self.ny + 2))
        self.L.ref_compute_rhs(desc(self._rm, (0, 0)), *self._uvw(u, v, w))

    def correct_uvw(self, u, v, w):
        self.L.ref_correct_uvw(desc(self._rm, (0, 0)), *self._uvw(u, v, w))

    def nhydro_solve(self, rmask, u, v, w):
        """The reference's public entry point (src/nhydro.f90:53-102): compute_rhs + solve_p(solver_prec, solver_maxiter) + correct_uvw."""
        rm = np.ascontiguousarray(rmask, dtype=np.float64) if rmask is not None else np.ones((self.nx + 2, self.ny + 2))
        PD = C.POINTER(C.c_double)
        self.L.ref_history_reset()
        self.L.ref_nhydro_solve(self.nx, self.ny, self.nz, rm.ctypes.data_as(PD), u.ctypes.data_as(PD), v.ctypes.data_as(PD), w.ctypes.data_as(PD))
        return self.history()

    def solve_p(self, tol, maxite):
        self.L.ref_history_reset()
        self.L.ref_solve_p(tol, maxite)
        return self.history()

    def history(self):
        res, conv = (C.c_double * 256)(), (C.c_double * 256)()
        n = self.L.ref_history(res, conv, 256)
        return np.array(res[:n]), np.array(conv[:n])

    def relax(self, lev, nsweeps): self.L.ref_relax(lev, nsweeps)
    def fine2coarse(self, lev): self.L.ref_fine2coarse(lev)
    def coarse2fine(self, lev): self.L.ref_coarse2fine(lev)
    def vcycle(self, lev): self.L.ref_vcycle(lev)
    def fcycle(self): self.L.ref_fcycle()

    def residual(self, lev):
        r = C.c_double(0.0)
        self.L.ref_compute_residual(lev, C.byref(r))
        return r.value

    def fill_halo(self, lev, name):
        """fill_halo(lev, grid(lev)%name): the generic's 2-D / 3-D specific, exactly as the reference's callers reach it."""
        d = getattr(self.L.ref_grid(lev).contents, name.lower())
        rank = sum(1 for q in range(5) if d.n[q] > 0)
        (self.L.ref_fill_halo_2d if rank == 2 else self.L.ref_fill_halo_3d)(lev, d, *(() if rank == 2 else (None,)))

    def fill_halo_relax(self, lev, name):
        g = self.L.ref_grid(lev).contents
        d = getattr(g, name.lower())
        nx, ny, nz = C.c_int(g.nx), C.c_int(g.ny), C.c_int(d.n[0])
        self.L.ref_fill_halo_3d_relax(lev, d.a, C.byref(nx), C.byref(ny), C.byref(nz))


class RefSolver(_Rank):
    """One rank, in the calling thread (the common case of the tests)."""

    def __init__(self, nx, ny, nz, **prm):
        super().__init__(0)
        self.L.rmpi_world_init(1)
        self.init(nx, ny, nz, 1, 1, prm)


class RefWorld:
    """npx*npy ranks as threads.  call(method, *args) runs the method on every rank concurrently (a collective), with
    per-rank arguments given as RefWorld.PerRank([...]); returns the list of per-rank results."""

    class PerRank(list):
        pass

    def __init__(self, nx, ny, nz, npx, npy, **prm):
        self.n = npx * npy
        lib().rmpi_world_init(self.n)
        self.q_in = [queue.Queue() for _ in range(self.n)]
        self.q_out = [queue.Queue() for _ in range(self.n)]
        self.threads = [threading.Thread(target=self._worker, args=(r,), daemon=True) for r in range(self.n)]
        for t in self.threads:
            t.start()
        self.nlevs = self.call("init", nx, ny, nz, npx, npy, prm)[0]

    def _worker(self, r):
        me = _Rank(r)
        while True:
            item = self.q_in[r].get()
            if item is None:
                return
            name, args = item
            try:
                self.q_out[r].put(("ok", getattr(me, name)(*args)))
            except BaseException as e:  # pragma: no cover
                self.q_out[r].put(("err", repr(e)))

    def call(self, name, *args, timeout=600):
        for r in range(self.n):
            self.q_in[r].put((name, tuple(a[r] if isinstance(a, RefWorld.PerRank) else a for a in args)))
        out = []
        for r in range(self.n):
            st, val = self.q_out[r].get(timeout=timeout)
            if st != "ok":
                raise RuntimeError(f"rank {r}: {val}")
            out.append(val)
        return out

    def get(self, name, lev=1):
        return [a.copy() for a in self.call("get", name, lev)]

    def close(self):
        for r in range(self.n):
            self.q_in[r].put(None)
