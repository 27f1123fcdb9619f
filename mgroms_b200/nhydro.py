"""Host-side mirror of the reference's `module nhydro` (src/nhydro.f90) over the C ABI of libmgb.so.

Same procedure names, argument order and meaning as the Fortran module:

    nhydro_init(nx, ny, nz, npxg, npyg)                      src/nhydro.f90:18
    nhydro_matrices(dx, dy, zeta, h, rmask, hc, theta_b, theta_s)   :36
    nhydro_solve(nx, ny, nz, rmask, u, v, w)                 :53   (u, v, w corrected in place)
    nhydro_check_nondivergence(nx, ny, nz, rmask, u, v, w)   :105
    nhydro_clean()                                           :137

Arrays are NumPy float64, C-contiguous, shaped so that their memory equals the Fortran declaration:
dx,dy,zeta,h,rmask -> (nx+2, ny+2) [Fortran (0:ny+1,0:nx+1)];  u -> (nz, ny+2, nx+1), v -> (nz, ny+1, nx+2),
w -> (nz+1, ny+2, nx+2) [Fortran (i,j,k)].  Errors raise MgbError (the reference `stop`s).

This module never imports anything under oracle/ and has no CPU compute path: every compute call goes to the
CUDA library and fails loudly if it (or a GPU) is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libmgb.so")
_lib = None
_dp = C.POINTER(C.c_double)


class MgbError(RuntimeError):
    pass


def load(build_if_missing: bool = True):
    """dlopen libmgb.so (building it with nvcc when absent) and declare the prototypes of include/mgb.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBPATH):
        if not build_if_missing:
            raise MgbError("libmgb.so is not built; run `python -m mgroms_b200.build`")
        from . import build as _b
        _b.build()
    L = C.CDLL(_LIBPATH)
    L.mgb_last_error.restype = C.c_char_p
    L.mgb_version.restype = C.c_char_p
    L.mgb_init.argtypes = [C.c_int] * 6 + [C.c_char_p, C.c_int]
    L.mgb_set_param.argtypes = [C.c_char_p, C.c_char_p]
    L.mgb_plan_levels.argtypes = [C.c_int] * 6
    L.mgb_namelist_probe.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.mgb_level_info.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.mgb_gather_members.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.mgb_comm_unique_id.argtypes = [C.c_void_p]
    L.mgb_comm_init.argtypes = [C.c_void_p]
    L.mgb_matrices.argtypes = [_dp] * 5 + [C.c_double] * 3
    L.mgb_solve.argtypes = [_dp] * 4
    L.mgb_check_nondivergence.argtypes = [_dp] * 4
    L.mgb_compute_rhs.argtypes = [_dp] * 4
    L.mgb_correct_uvw.argtypes = [_dp] * 3
    L.mgb_get_history.argtypes = [_dp, _dp, C.c_int]
    L.mgb_write_fort100.argtypes = [C.c_char_p]
    L.mgb_last_solve_ms.restype = C.c_double
    L.mgb_relax.argtypes = [C.c_int, C.c_int]
    L.mgb_residual.argtypes = [C.c_int, _dp]
    L.mgb_fine2coarse.argtypes = [C.c_int]
    L.mgb_coarse2fine.argtypes = [C.c_int]
    L.mgb_vcycle.argtypes = [C.c_int]
    L.mgb_fill_halo.argtypes = [C.c_int, C.c_char_p]
    L.mgb_field_size.argtypes = [C.c_int, C.c_char_p]
    L.mgb_field_size.restype = C.c_longlong
    L.mgb_get_field.argtypes = [C.c_int, C.c_char_p, _dp]
    L.mgb_set_field.argtypes = [C.c_int, C.c_char_p, _dp]
    L.mgb_relax_path.argtypes = [C.c_int, C.c_char_p, C.c_int]
    L.mgb_launch_count.restype = C.c_longlong
    L.mgb_reset_launch_count.restype = None
    L.mgb_profile_level1.argtypes = [C.c_int, _dp]
    L.mgb_profile_halo.argtypes = [C.c_int, C.c_int, _dp]
    L.mgb_tictoc_enable.argtypes = [C.c_int]
    L.mgb_write_tictoc.argtypes = [C.c_char_p]
    _lib = L
    return L


def _chk(rc):
    if rc != 0:
        raise MgbError(load().mgb_last_error().decode())


def _p(a, shape=None, allow_none=False):
    if a is None:
        if allow_none:
            return None
        raise MgbError("array argument is None")
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
        raise MgbError("arrays must be C-contiguous float64 NumPy arrays")
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise MgbError(f"array has shape {a.shape}, expected {shape}")
    return a.ctypes.data_as(_dp)


_dims = None   # (nx, ny, nz) of the live instance


# ---- namelist -------------------------------------------------------------------------------------------
def set_params(**kw):
    """Override namelist keys programmatically (applied on top of the nh_namelist file at nhydro_init)."""
    L = load()
    for k, v in kw.items():
        if isinstance(v, bool):
            v = ".true." if v else ".false."
        _chk(L.mgb_set_param(k.encode(), str(v).encode()))


def reset_params():
    _chk(load().mgb_reset_params())


def namelist_probe(path=None) -> dict:
    """Parse a namelist file exactly as nhydro_init would (host only) and return the 13 keys."""
    buf = C.create_string_buffer(1024)
    _chk(load().mgb_namelist_probe((path or "").encode(), buf, 1024))
    return dict(kv.split("=", 1) for kv in buf.value.decode().split(";"))


# ---- the five procedures of module nhydro -------------------------------------------------------------------
def nhydro_init(nx, ny, nz, npxg=1, npyg=1, rank=0, namelist=None, device=-1):
    global _dims
    _chk(load().mgb_init(nx, ny, nz, npxg, npyg, rank, (namelist or "").encode(), device))
    _dims = (nx, ny, nz)


def nhydro_matrices(dx, dy, zeta, h, rmask, hc, theta_b, theta_s):
    nx, ny, nz = _dims
    s2 = (nx + 2, ny + 2)
    _chk(load().mgb_matrices(_p(dx, s2), _p(dy, s2), _p(zeta, s2), _p(h, s2), _p(rmask, s2, True), float(hc), float(theta_b), float(theta_s)))


def _uvw_ptrs(nx, ny, nz, u, v, w):
    return _p(u, (nz, ny + 2, nx + 1)), _p(v, (nz, ny + 1, nx + 2)), _p(w, (nz + 1, ny + 2, nx + 2))


def nhydro_solve(nx, ny, nz, rmask, u, v, w):
    if (nx, ny, nz) != _dims:
        raise MgbError("nhydro_solve: sizes differ from nhydro_init")
    _chk(load().mgb_solve(_p(rmask, (nx + 2, ny + 2), True), *_uvw_ptrs(nx, ny, nz, u, v, w)))


def nhydro_check_nondivergence(nx, ny, nz, rmask, u, v, w):
    _chk(load().mgb_check_nondivergence(_p(rmask, (nx + 2, ny + 2), True), *_uvw_ptrs(nx, ny, nz, u, v, w)))


def nhydro_clean():
    global _dims
    _chk(load().mgb_clean())
    _dims = None


# ---- the three steps of nhydro_solve and the operators -----------------------------------------------------------
def compute_rhs(rmask, u, v, w):
    nx, ny, nz = _dims
    _chk(load().mgb_compute_rhs(_p(rmask, (nx + 2, ny + 2), True), *_uvw_ptrs(nx, ny, nz, u, v, w)))


def solve_p():
    _chk(load().mgb_solve_p())


def correct_uvw(u, v, w):
    nx, ny, nz = _dims
    _chk(load().mgb_correct_uvw(*_uvw_ptrs(nx, ny, nz, u, v, w)))


def relax(lev, nsweeps): _chk(load().mgb_relax(lev, nsweeps))
def fine2coarse(lev): _chk(load().mgb_fine2coarse(lev))
def coarse2fine(lev): _chk(load().mgb_coarse2fine(lev))
def Fcycle(): _chk(load().mgb_fcycle())
def Vcycle(lev1): _chk(load().mgb_vcycle(lev1))
def fill_halo(lev, name): _chk(load().mgb_fill_halo(lev, name.encode()))


def compute_residual(lev):
    r = C.c_double()
    _chk(load().mgb_residual(lev, C.byref(r)))
    return r.value


def history():
    res = (C.c_double * 512)(); conv = (C.c_double * 512)()
    n = load().mgb_get_history(res, conv, 512)
    return np.array(res[:n]), np.array(conv[:n])


def write_fort100(path="fort.100"): _chk(load().mgb_write_fort100(path.encode()))
def last_solve_ms(): return load().mgb_last_solve_ms()
def last_solve_iters(): return load().mgb_last_solve_iters()
def relax_path(lev=1) -> str:
    """Which kernels relax(lev, .) uses (diagnostic): small | mid | fused | a+bc | a+b+c | a, + "+halo" / "+exchange"."""
    buf = C.create_string_buffer(64)
    _chk(load().mgb_relax_path(lev, buf, 64))
    return buf.value.decode()


def halo_transport() -> str:
    return {0: "none", 1: "p2p", 2: "nccl"}[load().mgb_halo_transport()]


def launch_count(): return load().mgb_launch_count()
def reset_launch_count(): load().mgb_reset_launch_count()


def profile_level1(reps=10):
    out = (C.c_double * 4)()
    _chk(load().mgb_profile_level1(reps, out))
    return dict(relax_colour_ms=out[0], residual_ms=out[1], restrict_ms=out[2], prolong_ms=out[3])


def tictoc_enable(on=True): _chk(load().mgb_tictoc_enable(1 if on else 0))
def write_tictoc(path=None): _chk(load().mgb_write_tictoc((path or "").encode()))


def dump_field(name, lev=1, path=None):
    """Field dump replacing the reference's write_netcdf debug output (src/mg_netcdf_out.f90): grid(lev)%<name> as .npy."""
    a = get_field(name, lev)
    np.save(path or f"{name}_lev{lev}.npy", a)
    return a


def profile_halo(lev, reps=100):
    out = C.c_double()
    _chk(load().mgb_profile_halo(lev, reps, C.byref(out)))
    return out.value


# ---- topology (host only; works without a GPU) --------------------------------------------------------------------
def plan_levels(nx, ny, nz, npx=1, npy=1, rank=0):
    L = load()
    _chk(L.mgb_plan_levels(nx, ny, nz, npx, npy, rank))
    out = []
    keys = ["nx", "ny", "nz", "npx", "npy", "incx", "incy", "gather", "ngx", "ngy", "key", "color"]
    for lev in range(1, L.mgb_nlevs() + 1):
        info = (C.c_int * 20)()
        _chk(L.mgb_level_info(lev, info))
        d = {k: info[q] for q, k in enumerate(keys)}
        d["neighb"] = [info[12 + q] for q in range(8)]
        m = (C.c_int * 4)()
        n = L.mgb_gather_members(lev, m)
        d["members"] = [m[q] for q in range(max(n, 0))]
        out.append(d)
    return out


def nlevs(): return load().mgb_nlevs()


# ---- state access (grid(lev)%...) ------------------------------------------------------------------------------------
def _shape(lev, name):
    info = (C.c_int * 20)()
    _chk(load().mgb_level_info(lev, info))
    nx, ny, nz, ngx, ngy = info[0], info[1], info[2], info[8], info[9]
    if name == "cA": return (nx + 2, ny + 2, nz, 8)
    if name in ("p", "b", "r", "zxdy", "zydx", "bet", "g"): return (nx + 2, ny + 2, nz)
    if name in ("cw", "dzw"): return (nx + 2, ny + 2, nz + 1)
    if name == "zr": return (nx + 4, ny + 4, nz)
    if name == "zw": return (nx + 4, ny + 4, nz + 1)
    if name in ("dx", "dy", "zeta", "h", "s", "x1", "q1"): return (nx + 2, ny + 2)
    if name == "dummy3": return (nx // ngx + 2, ny // ngy + 2, nz)
    raise MgbError(f"unknown field {name}")


def get_field(name, lev=1):
    out = np.empty(_shape(lev, name))
    n = load().mgb_field_size(lev, name.encode())
    if n != out.size:
        raise MgbError(f"field {name}: size {n} != {out.size}: {load().mgb_last_error().decode()}")
    _chk(load().mgb_get_field(lev, name.encode(), _p(out)))
    return out


def set_field(name, arr, lev=1):
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    if arr.shape != _shape(lev, name):
        raise MgbError(f"field {name}: shape {arr.shape} != {_shape(lev, name)}")
    _chk(load().mgb_set_field(lev, name.encode(), _p(arr)))


# ---- multi-GPU bootstrap (one process per GPU; replaces MPI_Init + MPI_COMM_WORLD) --------------------------------------
def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _chk(load().mgb_comm_unique_id(buf))
    return buf.raw


def comm_init(uid: bytes):
    buf = C.create_string_buffer(uid, 128)
    _chk(load().mgb_comm_init(buf))
