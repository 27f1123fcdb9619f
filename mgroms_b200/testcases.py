"""Synthetic inputs of the reference's three test drivers, restated per emulated rank.

Follows (does not copy) /root/reference/src/mg_setup_tests.f90:108-157 (setup_seamount),
:160-211 (setup_rndtopo), src/mg_testseamount.f90:76-82,119-123 (constants, u=v=0, w=-1) and
src/mg_testcuc.f90:66-72 (hc/theta of the cuc case, whose NetCDF grid file is not shipped, so a
synthetic shelf/slope bathymetry stands in -- SURVEY.md section 8c/8d).

All 2-D arrays are returned in the layout nhydro_matrices expects: shape (nx+2, ny+2) C-order, i.e.
Fortran (0:ny+1, 0:nx+1) with j fastest.  u, v, w are returned as C-order arrays of shape
(nz, ny+2, nx+1), (nz, ny+1, nx+2), (nz+1, ny+2, nx+2), i.e. Fortran (i,j,k) with i fastest
(nhydro.f90:55-59).
"""
from __future__ import annotations

import numpy as np

DX_BENCH = 156.25  # metres: the 64^3 seamount spacing, kept for the big configs


def rank_coords(rank: int, npx: int) -> tuple[int, int]:
    """(pi, pj) of a rank -- mg_grids.f90:593-594."""
    return rank % npx, rank // npx


def _global_index(n_local: int, p: int) -> np.ndarray:
    return np.arange(0, n_local + 2, dtype=np.float64) + p * n_local


def seamount(nx, ny, nz, npx, npy, rank=0, Lx=1.0e4, Ly=1.0e4, Htot=4.0e3):
    """dx, dy, zeta, h of one rank for the Gaussian seamount (mg_setup_tests.f90:130-148)."""
    nxg, nyg = npx * nx, npy * ny
    pi, pj = rank_coords(rank, npx)
    dxv, dyv = Lx / float(nxg), Ly / float(nyg)
    x = (_global_index(nx, pi) - 0.5) * dxv          # i = 0..nx+1
    y = (_global_index(ny, pj) - 0.5) * dyv          # j = 0..ny+1
    x0, y0 = Lx * 0.5, Ly * 0.5
    X = x[:, None]
    Y = y[None, :]
    h = Htot * (1.0 - 0.5 * np.exp(-(X - x0) ** 2 / (Lx / 5.0) ** 2 - (Y - y0) ** 2 / (Ly / 5.0) ** 2))
    dx = np.full((nx + 2, ny + 2), dxv)
    dy = np.full((nx + 2, ny + 2), dyv)
    zeta = np.zeros((nx + 2, ny + 2))
    return dx, dy, zeta, np.ascontiguousarray(h)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def rndtopo(nx, ny, nz, npx, npy, rank=0, Lx=1.0e4, Ly=1.0e4, Htot=4.0e3, seed=20151123, hmin=1.0):
    """Random bathymetry h = Htot*0.2*U[0,1) (mg_setup_tests.f90:190-199).

    The reference draws from Fortran's unseeded random_number, which is not reproducible; here U is a
    counter-based generator keyed by the GLOBAL cell index so every decomposition sees the same field,
    and h is clipped to >= hmin metres to avoid zero-depth cells.
    """
    nxg, nyg = npx * nx, npy * ny
    pi, pj = rank_coords(rank, npx)
    ig = (np.arange(0, nx + 2, dtype=np.int64) + pi * nx)[:, None]
    jg = (np.arange(0, ny + 2, dtype=np.int64) + pj * ny)[None, :]
    key = (ig.astype(np.uint64) * np.uint64(nyg + 2) + jg.astype(np.uint64)) ^ (np.uint64(seed) << np.uint64(32))
    with np.errstate(over="ignore"):
        bits = _splitmix64(key)
    U = (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    h = np.maximum(Htot * (20.0 / 100.0) * U, hmin)
    dx = np.full((nx + 2, ny + 2), Lx / float(nxg))
    dy = np.full((nx + 2, ny + 2), Ly / float(nyg))
    zeta = np.zeros((nx + 2, ny + 2))
    return dx, dy, zeta, np.ascontiguousarray(h)


def cuc(nx, ny, nz, npx, npy, rank=0, dxm=195.0, Hdeep=4000.0, Hshelf=50.0):
    """Synthetic stand-in for the California-Undercurrent grid of mg_testcuc (file not shipped):
    a tanh shelf/slope in x with an along-shore meander and a canyon, uniform dx."""
    nxg, nyg = npx * nx, npy * ny
    pi, pj = rank_coords(rank, npx)
    Lx, Ly = nxg * dxm, nyg * dxm
    x = ((_global_index(nx, pi) - 0.5) * dxm)[:, None]
    y = ((_global_index(ny, pj) - 0.5) * dxm)[None, :]
    xs = 0.65 * Lx + 0.05 * Lx * np.sin(2.0 * np.pi * y / Ly)
    w = 0.08 * Lx
    h = Hshelf + 0.5 * (Hdeep - Hshelf) * (1.0 - np.tanh((x - xs) / w))
    h = h + 0.25 * (Hdeep - Hshelf) * np.exp(-((y - 0.4 * Ly) / (0.04 * Ly)) ** 2) * 0.5 * (1.0 + np.tanh((x - 0.5 * Lx) / w))
    h = np.minimum(h, Hdeep)
    dx = np.full((nx + 2, ny + 2), dxm)
    dy = np.full((nx + 2, ny + 2), dxm)
    zeta = np.zeros((nx + 2, ny + 2))
    return dx, dy, zeta, np.ascontiguousarray(h)


CUC_VERT = dict(hc=250.0, theta_b=6.0, theta_s=6.0)       # mg_testcuc.f90:66-72
SEAMOUNT_VERT = dict(hc=4.0e3, theta_b=0.0, theta_s=0.0)  # mg_testseamount.f90:80-82


def uvw_wminus1(nx, ny, nz):
    """u = v = 0, w(:,:,0) = 0, w(:,:,1:nz) = -1  (mg_testseamount.f90:119-123)."""
    u = np.zeros((nz, ny + 2, nx + 1))
    v = np.zeros((nz, ny + 1, nx + 2))
    w = np.full((nz + 1, ny + 2, nx + 2), -1.0)
    w[0] = 0.0
    return u, v, w


def uvw_random(nx, ny, nz, npx, npy, rank=0, seed=7):
    """Random velocities keyed by global face index (consistent across ranks' shared faces)."""
    pi, pj = rank_coords(rank, npx)
    nxg, nyg = npx * nx, npy * ny

    def field(shape_k, j_lo, j_n, i_lo, i_n, tag):
        k = np.arange(shape_k, dtype=np.uint64)[:, None, None]
        j = (np.arange(j_lo, j_lo + j_n, dtype=np.int64) + pj * ny + 2).astype(np.uint64)[None, :, None]
        i = (np.arange(i_lo, i_lo + i_n, dtype=np.int64) + pi * nx + 2).astype(np.uint64)[None, None, :]
        key = ((k * np.uint64(nyg + 8) + j) * np.uint64(nxg + 8) + i) ^ (np.uint64(seed * 16 + tag) << np.uint64(40))
        with np.errstate(over="ignore"):
            bits = _splitmix64(key)
        return 2.0 * ((bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)) - 1.0

    u = field(nz, 0, ny + 2, 1, nx + 1, 1)
    v = field(nz, 1, ny + 1, 0, nx + 2, 2)
    w = field(nz + 1, 0, ny + 2, 0, nx + 2, 3)
    return np.ascontiguousarray(u), np.ascontiguousarray(v), np.ascontiguousarray(w)


def bench_geometry(kind: str, nx, ny, nz, npx, npy, rank=0):
    """Geometry + vertical-coordinate parameters of a BASELINE.json config.

    C1 (64^3 seamount) uses the literal Lx = Ly = 10 km; every larger grid keeps dx = dy = 156.25 m
    (Lx = nxg * 156.25 m) because the literal domain does not converge when scaled.
    """
    nxg, nyg = npx * nx, npy * ny
    if kind == "seamount":
        Lx, Ly = nxg * DX_BENCH, nyg * DX_BENCH
        return seamount(nx, ny, nz, npx, npy, rank, Lx=Lx, Ly=Ly), SEAMOUNT_VERT
    if kind == "rndtopo":
        Lx, Ly = nxg * DX_BENCH, nyg * DX_BENCH
        return rndtopo(nx, ny, nz, npx, npy, rank, Lx=Lx, Ly=Ly), SEAMOUNT_VERT
    if kind == "cuc":
        return cuc(nx, ny, nz, npx, npy, rank), CUC_VERT
    raise ValueError(kind)


def bmask_rmask(nx, ny, npx=1, npy=1, rank=0):
    """rmask of the reference drivers when bmask=.true. (src/mg_testseamount.f90:97-101): ones, with the halo row / column
    zeroed on every side that has no neighbour rank (fill_halo_2D_bmask, src/mg_mpi_exchange.f90:357-391).  Shape (nx+2, ny+2)."""
    pi, pj = rank % npx, rank // npx
    m = np.ones((nx + 2, ny + 2))
    if pj == 0: m[:, 0] = 0.0          # south
    if pi == npx - 1: m[-1, :] = 0.0   # east
    if pj == npy - 1: m[:, -1] = 0.0   # north
    if pi == 0: m[0, :] = 0.0          # west
    return m
