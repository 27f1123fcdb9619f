"""refdriver.py -- one-rank driver of the TRANSLATED reference (oracle/_ref/libmgref.so, built by oracle/build_ref.py).

TEST INFRASTRUCTURE (oracle/): used by tests/ and tests/golden/make_ref_golden.py only.

What runs from the reference's own source text (translated statement by statement, see oracle/f90toc.py):
  find_grid_levels, define_grid_dims (src/mg_grids.f90:468-577), setup_zr_zw (src/mg_zr_zw.f90), define_matrix
  (src/mg_define_matrix.f90:211-680), compute_rhs, correct_uvw, relax + relax_3D_8_{RB,FC,GS} + relax_3D_8_heart + tridiag,
  compute_residual(_3D_8), fine2coarse(_3D), coarse2fine(_3D_linear, _3D_nearest), solve_p, Fcycle, Vcycle,
  fill_halo_2D_bmask.
What this file restates for ONE rank (no arithmetic beyond copies, the nh = 2 extrapolation 2A - B and the 2x2 averages):
  * grid allocation with the reference's bounds (src/mg_grids.f90:170-260);
  * the level loop of define_matrices_topo (src/mg_define_matrix.f90:28-208): copy / 2x2 coarsening of dx, dy, zeta, h;
  * the physical-boundary branches of fill_halo_2D / _3D / _3D_relax / _4D (src/mg_mpi_exchange.f90:23-352, 396-745,
    750-1242, 1247-1550): on one rank every neighbour is MPI_PROC_NULL, so only those branches execute.
Arrays are NumPy C-order with reversed dimensions, e.g. Fortran p(nz,0:ny+1,0:nx+1) <-> p[i, j, k-1] of shape (nx+2, ny+2, nz).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class farr(C.Structure):
    _fields_ = [("a", C.POINTER(C.c_double)), ("lb", C.c_int * 4), ("n", C.c_int * 4)]


_FARR_MEMBERS = ["dx", "dy", "zeta", "h", "p", "b", "r", "zr", "zw", "cw", "dzw", "zxdy", "zydx", "dummy3dnz", "dummy3dnzp", "dummy3", "ca"]


class grid_t(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nx", "ny", "nz", "npx", "npy", "incx", "incy", "gather", "ngx", "ngy", "key", "color", "nd")] + \
               [("neighb", C.c_int * 9)] + [(n, farr) for n in _FARR_MEMBERS]


_lib = None


def available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libmgref.so")) or os.path.isdir("/root/reference/src")


def lib():
    global _lib
    if _lib is None:
        import sys
        sys.path.insert(0, HERE)
        import build_ref
        so = os.path.join(HERE, "_ref", "libmgref.so")
        if os.path.isdir("/root/reference/src") or not os.path.exists(so):
            so = build_ref.build(quiet=True)
        L = C.CDLL(so)
        L.ref_grid.restype = C.POINTER(grid_t)
        L.ref_grid.argtypes = [C.c_int]
        D, I, PD = C.c_double, C.c_int, C.POINTER(C.c_double)
        L.ref_find_grid_levels.argtypes = [I] * 5
        L.ref_setup_zr_zw_croco.argtypes = [D, D, D, farr, farr, farr, farr, C.c_char_p]
        L.ref_define_matrix.argtypes = [I, farr, farr, farr, farr, farr]
        L.ref_compute_rhs.argtypes = [farr] * 4
        L.ref_correct_uvw.argtypes = [farr] * 4
        L.ref_relax.argtypes = [I, I]
        L.ref_compute_residual.argtypes = [I, PD]
        L.ref_fine2coarse.argtypes = [I]
        L.ref_coarse2fine.argtypes = [I]
        L.ref_solve_p.argtypes = [D, I]
        L.ref_vcycle.argtypes = [I]
        L.ref_fill_halo_2d_bmask.argtypes = [I, farr]
        L.ref_set_str.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_history.argtypes = [PD, PD, I]
        _lib = L
    return _lib


def desc(a: np.ndarray, lb) -> farr:
    """Descriptor of a C-order array whose REVERSED shape is the Fortran shape, with Fortran lower bounds `lb`."""
    assert a.flags.c_contiguous and a.dtype == np.float64
    d = farr()
    d.a = a.ctypes.data_as(C.POINTER(C.c_double))
    shp = a.shape[::-1]
    for q in range(4):
        d.lb[q] = lb[q] if q < len(shp) else 0
        d.n[q] = shp[q] if q < len(shp) else 0
    return d


_HALO_RELAX = C.CFUNCTYPE(None, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int)
_HALO_ANY = C.CFUNCTYPE(None, C.c_int, farr, C.c_char_p)


def neumann_nh1(A):
    """Physical-boundary branches of fill_halo_2D / fill_halo_3D_relax / fill_halo_3D with nh = 1 (src/mg_mpi_exchange.f90:
    110-200, 507-598, 953-1095): A[i, j, ...] with the ring at index 0 and n+1."""
    nx, ny = A.shape[0] - 2, A.shape[1] - 2
    A[1:nx + 1, 0] = A[1:nx + 1, 1]            # south   a(0,1:nx)    = a(1,1:nx)
    A[nx + 1, 1:ny + 1] = A[nx, 1:ny + 1]      # east    a(1:ny,nx+1) = a(1:ny,nx)
    A[1:nx + 1, ny + 1] = A[1:nx + 1, ny]      # north
    A[0, 1:ny + 1] = A[1, 1:ny + 1]            # west
    A[0, 0] = A[1, 1]                          # south-west  a(0,0) = a(1,1)
    A[nx + 1, 0] = A[nx, 1]                    # south-east  a(0,nx+1) = a(1,nx)
    A[nx + 1, ny + 1] = A[nx, ny]              # north-east
    A[0, ny + 1] = A[1, ny]                    # north-west


def halo3d(A, nx, ny, lbc):
    """fill_halo_3D on one rank (src/mg_mpi_exchange.f90:947-1095): A[i + nh - 1 ... ] of shape (nx+2nh, ny+2nh, nz)."""
    nh = (A.shape[1] - ny) // 2
    o = nh - 1                                  # Fortran index (j, i) -> A[i + o, j + o]
    lbc = lbc.decode() if lbc else None

    def P(j, i):                                # slices in Fortran index space (inclusive ranges as tuples)
        def s(r):
            if isinstance(r, tuple):
                lo, hi, st = (r + (1,))[:3]
                return slice(lo + o, (hi + o + (1 if st > 0 else -1)) if (hi + o + (1 if st > 0 else -1)) >= 0 else None, st)
            return r + o
        return (s(i), s(j))
    full_j, full_i = (1 - nh, ny + nh), (1 - nh, nx + nh)
    # south
    if lbc == "v":
        A[P(1, full_i)] = 0.0
    else:
        A[P(0, (1, nx))] = A[P(1, (1, nx))]
        if nh == 2:
            A[P(-1, (1, nx))] = 2.0 * A[P(1, (1, nx))] - A[P(2, (1, nx))]
    # east
    if lbc == "u":
        A[P(full_j, nx + 1)] = 0.0
    else:
        A[P((1, ny), nx + 1)] = A[P((1, ny), nx)]
        if nh == 2:
            A[P((1, ny), nx + 2)] = 2.0 * A[P((1, ny), nx)] - A[P((1, ny), nx - 1)]
    # north
    if lbc == "v":
        A[P(ny + 1, full_i)] = 0.0
    else:
        A[P(ny + 1, (1, nx))] = A[P(ny, (1, nx))]
        if nh == 2:
            A[P(ny + 2, (1, nx))] = 2.0 * A[P(ny, (1, nx))] - A[P(ny - 1, (1, nx))]
    # west
    if lbc == "u":
        A[P(full_j, 1)] = 0.0
    else:
        A[P((1, ny), 0)] = A[P((1, ny), 1)]
        if nh == 2:
            A[P((1, ny), -1)] = 2.0 * A[P((1, ny), 1)] - A[P((1, ny), 2)]
    # corners: mirrors of the interior corner block (p(:,1-nh:0,1-nh:0) = p(:,nh:1:-1,nh:1:-1), ...), or zero with lbc_null
    lo, hi_j, hi_i = (1 - nh, 0), (ny + 1, ny + nh), (nx + 1, nx + nh)
    mj_lo, mi_lo, mj_hi, mi_hi = (nh, 1, -1), (nh, 1, -1), (ny, ny - nh + 1, -1), (nx, nx - nh + 1, -1)
    if lbc == "u":
        A[P(lo, lo)] = 0.0; A[P(lo, hi_i)] = 0.0; A[P(hi_j, hi_i)] = 0.0; A[P(hi_j, lo)] = 0.0
    elif lbc == "v":
        A[P(lo, lo)] = A[P(mj_lo, mi_lo)]; A[P(lo, hi_i)] = A[P(mj_lo, mi_hi)]
        A[P(hi_j, hi_i)] = 0.0; A[P(hi_j, lo)] = 0.0
    else:
        A[P(lo, lo)] = A[P(mj_lo, mi_lo)]; A[P(lo, hi_i)] = A[P(mj_lo, mi_hi)]
        A[P(hi_j, hi_i)] = A[P(mj_hi, mi_hi)]; A[P(hi_j, lo)] = A[P(mj_hi, mi_lo)]


class RefSolver:
    """nhydro_init / nhydro_matrices / nhydro_solve of the reference on one MPI rank (src/nhydro.f90:18-143)."""

    _active = None

    def __init__(self, nx, ny, nz, bmask=False, relax_method="RB", interp_type="linear", cmatrix="real",
                 ns_pre=3, ns_post=2, ns_coarsest=40, nsmall=8):
        L = self.L = lib()
        RefSolver._active = self
        for k, v in dict(bmask=int(bmask), ns_pre=ns_pre, ns_post=ns_post, ns_coarsest=ns_coarsest, nsmall=nsmall, myrank=0, nprocs=1,
                         netcdf_output=0, aggressive=0).items():
            C.c_int.in_dll(L, "g_" + k).value = v
        for k, v in dict(cmatrix=cmatrix, relax_method=relax_method, interp_type=interp_type).items():
            L.ref_set_str(k.encode(), v.encode())
        self.bmask = bool(bmask)
        self._cb = (_HALO_RELAX(self._halo_relax), _HALO_ANY(self._halo_any))
        L.ref_set_callbacks(*self._cb)
        null = C.c_int.in_dll(L, "g_mpi_proc_null").value
        # nhydro_init -> grids_init: levels and per-level sizes from the translated find_grid_levels / define_grid_dims
        g1 = L.ref_grid(1).contents
        g1.nx, g1.ny, g1.nz, g1.npx, g1.npy, g1.incx, g1.incy = nx, ny, nz, 1, 1, 1, 1
        L.ref_find_grid_levels(1, 1, nx, ny, nz)
        self.nlevs = C.c_int.in_dll(L, "g_nlevs").value
        L.ref_define_grid_dims()
        self.A = {}                      # (lev, name) -> numpy array (kept alive; the descriptors point into them)
        self.dims = {}
        for lev in range(1, self.nlevs + 1):
            g = L.ref_grid(lev).contents
            lx, ly, lz = g.nx, g.ny, g.nz
            self.dims[lev] = (lx, ly, lz)
            for q in range(9):
                g.neighb[q] = null
            g.nd = 8
            spec = {n: ((lx + 2, ly + 2), (0, 0)) for n in ("dx", "dy", "zeta", "h")}
            spec.update({n: ((lx + 2, ly + 2, lz), (1, 0, 0)) for n in ("p", "b", "r", "zxdy", "zydx", "dummy3dnz")})
            spec.update({n: ((lx + 2, ly + 2, lz + 1), (1, 0, 0)) for n in ("cw", "dzw", "dummy3dnzp")})
            spec["zr"] = ((lx + 4, ly + 4, lz), (1, -1, -1))
            spec["zw"] = ((lx + 4, ly + 4, lz + 1), (1, -1, -1))
            spec["ca"] = ((lx + 2, ly + 2, lz, 8), (1, 1, 0, 0))
            for n, (shape, lb) in spec.items():
                a = np.zeros(shape)
                self.A[(lev, n)] = a
                setattr(g, n, desc(a, lb))

    # ---- call-backs ------------------------------------------------------------------------------------------
    def _halo_relax(self, lev, ptr, nx, ny, nz):
        A = np.ctypeslib.as_array(ptr, shape=(nx + 2, ny + 2, nz))
        neumann_nh1(A)

    def _halo_any(self, lev, d, lbc):
        rank = sum(1 for q in range(4) if d.n[q] > 0)
        shape = tuple(d.n[q] for q in range(rank))[::-1]
        A = np.ctypeslib.as_array(d.a, shape=shape)
        nx, ny, _ = self.dims[lev]
        if rank == 2:
            neumann_nh1(A)
        elif rank == 3:
            halo3d(A, nx, ny, lbc)
        elif rank == 4:
            if not self.bmask:             # fill_halo_4D (src/mg_mpi_exchange.f90:1325-1410): zero ring unless bmask
                A[:, 0] = 0.0; A[:, -1] = 0.0; A[0, :] = 0.0; A[-1, :] = 0.0

    # ---- the reference API -----------------------------------------------------------------------------------
    def get(self, name, lev=1):
        return self.A[(lev, "ca" if name == "cA" else name)]

    def desc_of(self, name, lev):
        return getattr(self.L.ref_grid(lev).contents, name)

    def matrices(self, dx, dy, zeta, h, rmask, hc, theta_b, theta_s):
        """define_matrices_topo (src/mg_define_matrix.f90:28-208); inputs of shape (nx+2, ny+2) = Fortran (0:ny+1,0:nx+1)."""
        L = self.L
        C.c_double.in_dll(L, "g_hlim").value = hc
        C.c_double.in_dll(L, "g_nhtheta_b").value = theta_b
        C.c_double.in_dll(L, "g_nhtheta_s").value = theta_s
        for lev in range(1, self.nlevs + 1):
            nx, ny, nz = self.dims[lev]
            if lev == 1:
                for n, src in (("dx", dx), ("dy", dy), ("zeta", zeta), ("h", h)):
                    self.A[(1, n)][...] = src
                rmaski = np.ascontiguousarray(rmask, dtype=np.float64) if rmask is not None else np.ones((nx + 2, ny + 2))
            else:
                nxf, nyf, _ = self.dims[lev - 1]
                for n, w in (("dx", 0.5), ("dy", 0.5), ("zeta", 0.25), ("h", 0.25)):
                    f = self.A[(lev - 1, n)]
                    # c(1:nyc,1:nxc) = w * ( f(1:nyf:2,1:nxf:2) + f(2:nyf+1:2,1:nxf:2) + f(1:nyf:2,2:nxf+1:2) + f(2:nyf+1:2,2:nxf+1:2) )
                    self.A[(lev, n)][1:nx + 1, 1:ny + 1] = w * (((f[1:nxf:2, 1:nyf:2] + f[1:nxf:2, 2:nyf + 1:2]) + f[2:nxf + 1:2, 1:nyf:2])
                                                                + f[2:nxf + 1:2, 2:nyf + 1:2])
                rmaski = np.ones((nx + 2, ny + 2))
                if self.bmask:
                    L.ref_fill_halo_2d_bmask(lev, desc(rmaski, (0, 0)))
            for n in ("dx", "dy", "zeta", "h"):
                neumann_nh1(self.A[(lev, n)])
            L.ref_setup_zr_zw_croco(hc, theta_b, theta_s, self.desc_of("zeta", lev), self.desc_of("h", lev), self.desc_of("zr", lev),
                                    self.desc_of("zw", lev), b"new_s_coord")
            halo3d(self.A[(lev, "zr")], nx, ny, None)
            halo3d(self.A[(lev, "zw")], nx, ny, None)
            self._rmaski = rmaski
            L.ref_define_matrix(lev, self.desc_of("dx", lev), self.desc_of("dy", lev), self.desc_of("zr", lev), self.desc_of("zw", lev),
                                desc(rmaski, (0, 0)))

    def _uvw_desc(self, u, v, w):
        return desc(u, (1, 0, 1)), desc(v, (0, 1, 1)), desc(w, (0, 0, 0))

    def compute_rhs(self, rmask, u, v, w):
        nx, ny, _ = self.dims[1]
        self._rm = np.ascontiguousarray(rmask, dtype=np.float64) if rmask is not None else np.ones((nx + 2, ny + 2))
        # rmask is consumed as rmask(j,i) with shape (0:ny+1,0:nx+1) (declared (i,j), used (j,i))
        self.L.ref_compute_rhs(desc(self._rm, (0, 0)), *self._uvw_desc(u, v, w))

    def correct_uvw(self, u, v, w):
        self.L.ref_correct_uvw(desc(self._rm, (0, 0)), *self._uvw_desc(u, v, w))

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
