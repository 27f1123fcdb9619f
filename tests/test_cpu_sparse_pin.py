"""An independent pin for the SWEEP ORDER of the smoother and for the operator application.

The oracle's loops (oracle/mgoracle.cpp: relax, residual) are checked here against linear algebra that shares no code with
them: the stencil coefficients cA are assembled into a SciPy sparse matrix straight from the operator definition of
SURVEY.md Appendix A (= src/mg_relax.f90:263-301, 465-507), over all cells INCLUDING the halo ring, and

  * the residual is compared with  r = b - A p;
  * one red-black sweep in the reference's loop order (`i` outer ascending, src/mg_relax.f90:173-176) is compared with the
    block Gauss-Seidel step it is by definition: for each colour c, with every other cell (other colour, halo) held fixed,
        (D_c + L_c) p_c(new) = b_c - U_c p_c(old) - A_{c,rest} p_rest ,
    where D_c holds the couplings inside a column (the tridiagonal that `tridiag` solves), L_c the couplings to
    same-colour columns visited EARLIER in the loop (the (i-1) diagonal neighbours of the k=1 plane, src/mg_relax.f90:271-276)
    and U_c those visited later; the halo is refreshed (Neumann mirror) after each colour as fill_halo does.
The colour-simultaneous (Jacobi-in-colour) variant is shown to differ, so the test really discriminates the order.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from common import make_oracle, neumann_halo, random_interior


def assemble(cA):
    """Rows: every interior cell; columns: every cell of the (nx+2, ny+2, nz) array.  cA[i, j, k, c-1] = cA(c, k, j, i)."""
    nx2, ny2, nz, _ = cA.shape
    nx, ny = nx2 - 2, ny2 - 2
    idx = np.arange(nx2 * ny2 * nz).reshape(nx2, ny2, nz)
    rows, cols, vals = [], [], []

    def add(i, j, k, di, dj, dk, v):
        rows.append(idx[i, j, k]); cols.append(idx[i + di, j + dj, k + dk]); vals.append(v)
    c = lambda n, k, j, i: cA[i, j, k, n - 1]   # noqa: E731  (Fortran cA(n, k, j, i), k 0-based)
    for i in range(1, nx + 1):
        for j in range(1, ny + 1):
            for k in range(nz):
                lo, hi = k > 0, k < nz - 1
                add(i, j, k, 0, 0, 0, c(1, k, j, i))
                if lo: add(i, j, k, 0, 0, -1, c(2, k, j, i))
                if hi: add(i, j, k, 0, 0, +1, c(2, k + 1, j, i))
                if hi: add(i, j, k, 0, -1, +1, c(3, k, j, i))
                if lo: add(i, j, k, 0, +1, -1, c(3, k - 1, j + 1, i))
                add(i, j, k, 0, -1, 0, c(4, k, j, i))
                add(i, j, k, 0, +1, 0, c(4, k, j + 1, i))
                if lo: add(i, j, k, 0, -1, -1, c(5, k, j, i))
                if hi: add(i, j, k, 0, +1, +1, c(5, k + 1, j + 1, i))
                if hi: add(i, j, k, -1, 0, +1, c(6, k, j, i))
                if lo: add(i, j, k, +1, 0, -1, c(6, k - 1, j, i + 1))
                add(i, j, k, -1, 0, 0, c(7, k, j, i))
                add(i, j, k, +1, 0, 0, c(7, k, j, i + 1))
                if lo: add(i, j, k, -1, 0, -1, c(8, k, j, i))
                if hi: add(i, j, k, +1, 0, +1, c(8, k + 1, j, i + 1))
                if k == 0:      # bottom level: the four diagonal neighbours of the k=1 plane
                    add(i, j, k, -1, +1, 0, c(5, 0, j, i))
                    add(i, j, k, +1, -1, 0, c(5, 0, j - 1, i + 1))
                    add(i, j, k, -1, -1, 0, c(8, 0, j, i))
                    add(i, j, k, +1, +1, 0, c(8, 0, j + 1, i + 1))
    n = nx2 * ny2 * nz
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, n)), idx


CASES = [("seamount", 12, 10, 8), ("rndtopo", 8, 12, 4), ("cuc", 10, 8, 8)]


@pytest.mark.parametrize("kind,nx,ny,nz", CASES)
def test_residual_is_b_minus_A_p(kind, nx, ny, nz):
    o = make_oracle(kind, nx, ny, nz)
    A, idx = assemble(o.get("cA"))
    shp = o.get("p").shape
    p = neumann_halo(random_interior(shp, 5)); b = random_interior(shp, 6)
    o.set("p", p); o.set("b", b); o.residual(1)
    r = (b.ravel() - A @ p.ravel()).reshape(shp)
    got = o.get("r")
    scale = np.abs(A).max() * np.abs(p).max()
    assert np.abs(got[1:-1, 1:-1] - r[1:-1, 1:-1]).max() < 1e-13 * scale


def colour_step(A, idx, p, b, rb, ordered=True):
    """One colour of the sweep as the triangular solve described in the module docstring (ordered) or with every
    same-colour coupling taken at its OLD value (ordered=False: what a naive parallel sweep would compute)."""
    nx2, ny2, nz = idx.shape
    nx, ny = nx2 - 2, ny2 - 2
    cols_c, order = [], {}
    for i in range(1, nx + 1):                       # the reference's loop: i outer, j = 1 + mod(i + rb, 2), ny, 2
        for j in range(1 + (i + rb) % 2, ny + 1, 2):
            order[(i, j)] = len(order)
            cols_c.append((i, j))
    unk = np.array([idx[i, j, k] for (i, j) in cols_c for k in range(nz)])
    visit = np.repeat(np.arange(len(cols_c)), nz)    # position of each unknown's column in the loop
    M = A[unk][:, unk].tocoo()
    lower = visit[M.col] <= visit[M.row] if ordered else visit[M.col] == visit[M.row]
    DL = sp.csr_matrix((M.data[lower], (M.row[lower], M.col[lower])), shape=M.shape)
    U = sp.csr_matrix((M.data[~lower], (M.row[~lower], M.col[~lower])), shape=M.shape)
    x = p.ravel().copy()
    rest = x.copy(); rest[unk] = 0.0
    rhs = b.ravel()[unk] - U @ x[unk] - (A[unk] @ rest)
    x[unk] = spla.spsolve(DL.tocsc(), rhs)
    return neumann_halo(x.reshape(p.shape))          # fill_halo after the colour


@pytest.mark.parametrize("kind,nx,ny,nz", CASES)
def test_red_black_sweep_is_the_ordered_block_gauss_seidel_of_A(kind, nx, ny, nz):
    o = make_oracle(kind, nx, ny, nz, relax_order=0)              # the reference's sequential order
    o1 = make_oracle(kind, nx, ny, nz, relax_order=1)             # the rank-1 form the CUDA path uses
    A, idx = assemble(o.get("cA"))
    shp = o.get("p").shape
    p0 = neumann_halo(random_interior(shp, 7)); b = random_interior(shp, 8)
    for orc in (o, o1):
        orc.set("p", p0); orc.set("b", b); orc.relax(1, 1)
    want = colour_step(A, idx, colour_step(A, idx, p0, b, 1), b, 2)
    naive = colour_step(A, idx, colour_step(A, idx, p0, b, 1, ordered=False), b, 2, ordered=False)
    scale = np.abs(want).max()
    assert np.abs(o.get("p") - want).max() < 1e-12 * scale
    assert np.abs(o1.get("p") - want).max() < 1e-12 * scale
    # the order matters at a level far above the tolerance used here: the test discriminates
    assert np.abs(naive - want).max() > 1e-7 * scale
