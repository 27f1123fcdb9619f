"""Shared helpers of the parity tests: build the same case in the CPU oracle and in the CUDA product."""
import numpy as np

from mgroms_b200 import testcases as tc
from oracle.pyoracle import Oracle


def geometry(kind, nx, ny, nz, npx=1, npy=1, rank=0, bench_dx=False):
    if bench_dx:
        return tc.bench_geometry(kind, nx, ny, nz, npx, npy, rank)
    if kind == "seamount":
        return tc.seamount(nx, ny, nz, npx, npy, rank), tc.SEAMOUNT_VERT
    if kind == "rndtopo":
        return tc.rndtopo(nx, ny, nz, npx, npy, rank), tc.SEAMOUNT_VERT
    if kind == "cuc":
        return tc.cuc(nx, ny, nz, npx, npy, rank), tc.CUC_VERT
    raise ValueError(kind)


def velocities(mode, nx, ny, nz, npx=1, npy=1, rank=0):
    return tc.uvw_wminus1(nx, ny, nz) if mode == "wminus1" else tc.uvw_random(nx, ny, nz, npx, npy, rank)


def make_oracle(kind, nx, ny, nz, uvw="wminus1", bench_dx=False, **params):
    o = Oracle(nx, ny, nz, 1, 1, **params)
    (dx, dy, zeta, h), vert = geometry(kind, nx, ny, nz, bench_dx=bench_dx)
    o.set_geometry(0, dx, dy, zeta, h, tc.bmask_rmask(nx, ny) if params.get("bmask") else None)
    o.set_uvw(0, *velocities(uvw, nx, ny, nz))
    o.matrices(vert["hc"], vert["theta_b"], vert["theta_s"])
    return o


def make_product(nh, kind, nx, ny, nz, bench_dx=False, **params):
    params.pop("relax_order", None)
    nh.reset_params()
    nh.set_params(**params)
    nh.nhydro_init(nx, ny, nz, 1, 1)
    (dx, dy, zeta, h), vert = geometry(kind, nx, ny, nz, bench_dx=bench_dx)
    nh.nhydro_matrices(dx, dy, zeta, h, tc.bmask_rmask(nx, ny) if params.get("bmask") else None, vert["hc"], vert["theta_b"], vert["theta_s"])
    return nh


def random_interior(shape3, seed):
    """Field (nx+2, ny+2, nz) with random interior and zero halo."""
    rng = np.random.default_rng(seed)
    a = np.zeros(shape3)
    a[1:-1, 1:-1, :] = rng.standard_normal((shape3[0] - 2, shape3[1] - 2, shape3[2]))
    return a


def neumann_halo(a):
    """Homogeneous-Neumann mirror of a (nx+2, ny+2, nz) field: what fill_halo does on one rank."""
    a = a.copy()
    a[0, 1:-1] = a[1, 1:-1]; a[-1, 1:-1] = a[-2, 1:-1]
    a[1:-1, 0] = a[1:-1, 1]; a[1:-1, -1] = a[1:-1, -2]
    a[0, 0] = a[1, 1]; a[0, -1] = a[1, -2]; a[-1, 0] = a[-2, 1]; a[-1, -1] = a[-2, -2]
    return a


def relerr(a, b):
    d = np.abs(np.asarray(a) - np.asarray(b)).max()
    s = np.abs(np.asarray(b)).max()
    return d / s if s > 0 else d


# |d(res/bnorm)| <= 1e-10*res + floor : the north-star 1e-10 relative bar, with the FP64 round-off floor of
# any reordering of the sweep (deviation grows like 1.4e-17/res)
def history_close(h_gpu, h_ref, rtol=1e-10, floor=2e-15):
    h_gpu = np.asarray(h_gpu); h_ref = np.asarray(h_ref)
    assert len(h_gpu) == len(h_ref), (len(h_gpu), len(h_ref), h_gpu, h_ref)
    bad = np.abs(h_gpu - h_ref) > rtol * h_ref + floor
    assert not bad.any(), (h_gpu, h_ref, np.abs(h_gpu - h_ref) / h_ref)
