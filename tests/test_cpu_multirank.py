"""CPU tests of the emulated-rank machinery of the oracle: halo rules, gather/split, decomposition invariance
(what the reference's bit-rotted src/old_tests/mg_testhalo.f90 / mg_testgather.f90 checked by eye)."""
import numpy as np
import pytest

from mgroms_b200 import testcases as tc
from oracle.pyoracle import Oracle, make_case


def assemble(o, name, lev, npx, npy):
    """Global interior array (nxg, nyg, nz) from the ranks' interiors."""
    g = o.level_info(lev)
    nx, ny = g["nx"], g["ny"]
    rows = []
    for pi in range(npx):
        rows.append(np.concatenate([o.get(name, lev, pj * npx + pi)[1:-1, 1:-1] for pj in range(npy)], axis=1))
    return np.concatenate(rows, axis=0)


@pytest.mark.parametrize("npx,npy", [(2, 1), (2, 2), (4, 2)])
def test_halo_is_the_global_neumann_extension(npx, npy):
    nx, ny, nz = 8, 8, 4
    o = make_case("seamount", nx, ny, nz, npx, npy)
    rng = np.random.default_rng(0)
    glob = rng.standard_normal((npx * nx, npy * ny, nz))
    for r in range(npx * npy):
        pi, pj = r % npx, r // npx
        a = np.full((nx + 2, ny + 2, nz), np.nan)
        a[1:-1, 1:-1] = glob[pi * nx:(pi + 1) * nx, pj * ny:(pj + 1) * ny]
        o.set("p", np.nan_to_num(a, nan=-7.0), 1, r)
    o.fill_halo(1, "p")
    ext = np.pad(glob, ((1, 1), (1, 1), (0, 0)), mode="edge")        # mirror = nearest interior value, corners included
    for r in range(npx * npy):
        pi, pj = r % npx, r // npx
        want = ext[pi * nx:(pi + 1) * nx + 2, pj * ny:(pj + 1) * ny + 2]
        assert np.array_equal(o.get("p", 1, r), want), r


def test_nh2_halo_extrapolates_at_physical_edges():
    o = make_case("seamount", 8, 8, 4, 2, 1)
    for r in range(2):
        zr = o.get("zr", 1, r)                       # (nx+4, ny+4, nz), index 0,1 = rings -1,0
        if r == 0:                                   # west physical edge: ring2 = 2A - B, ring1 = A
            assert np.allclose(zr[0, 2:-2], 2 * zr[2, 2:-2] - zr[3, 2:-2], rtol=0, atol=0)
            assert np.array_equal(zr[1, 2:-2], zr[2, 2:-2])
        else:                                        # west neighbour: both rings come from rank 0's last two columns
            zl = o.get("zr", 1, 0)
            assert np.array_equal(zr[0:2, 2:-2], zl[-4:-2, 2:-2])


def test_gather_split_roundtrip_and_family_layout():
    # C4-like: 4x2 ranks, nsmall raised so that levels gather (SURVEY.md appendix B worked example, scaled down)
    o = make_case("seamount", 16, 32, 8, 4, 2, nsmall=16)
    gath = [o.level_info(l)["gather"] for l in range(1, o.nlevs + 1)]
    assert 1 in gath
    lev = gath.index(1) + 1
    g = o.level_info(lev)
    assert g["ngx"] == 2 and g["ngy"] == 2 and g["npx"] == 2 and g["npy"] == 1
    assert o.gather_members(lev, 0) == [0, 1, 4, 5] and o.gather_members(lev, 2) == [2, 3, 6, 7]
    # every member of a family holds the same gathered matrix, and split gives back the own quarter incl. halo
    for fam in ([0, 1, 4, 5], [2, 3, 6, 7]):
        ref = o.get("cA", lev, fam[0])
        for r in fam[1:]:
            assert np.array_equal(o.get("cA", lev, r), ref)
    rng = np.random.default_rng(1)
    shp = o.get("p", lev, 0).shape
    for r in range(8):
        o.set("p", rng.standard_normal(shp) if r in (0, 2) else o.get("p", lev, 0 if r in (1, 4, 5) else 2), lev, r)
    o.split(lev)
    nxs, nys = g["nx"] // 2, g["ny"] // 2
    for r, (l, m) in zip([0, 1, 4, 5], [(0, 0), (1, 0), (0, 1), (1, 1)]):
        big = o.get("p", lev, r)
        assert np.array_equal(o.get("dummy3", lev, r), big[l * nxs:l * nxs + nxs + 2, m * nys:m * nys + nys + 2])


@pytest.mark.parametrize("bmask", [False, True])
@pytest.mark.parametrize("npx,npy", [(2, 2), (4, 2)])
def test_operators_are_decomposition_invariant(npx, npy, bmask):
    """Everything except the red-black bottom recurrence (which restarts at sub-domain edges, as in the reference) must
    not depend on the decomposition: matrices, rhs, residual, restriction, prolongation."""
    nxg, nyg, nz = 32, 32, 8
    o1 = make_case("rndtopo", nxg, nyg, nz, 1, 1, uvw="random", nsmall=4, bmask=bmask)
    on = make_case("rndtopo", nxg // npx, nyg // npy, nz, npx, npy, uvw="random", nsmall=4, bmask=bmask)      # nsmall=4: no gathering at level 2
    assert np.array_equal(assemble(on, "cA", 1, npx, npy), o1.get("cA")[1:-1, 1:-1])
    assert np.array_equal(assemble(on, "cA", 2, npx, npy), o1.get("cA", 2)[1:-1, 1:-1])
    o1.compute_rhs(); on.compute_rhs()
    assert np.array_equal(assemble(on, "b", 1, npx, npy), o1.get("b")[1:-1, 1:-1])
    r1, rn = o1.residual(1), on.residual(1)
    assert np.array_equal(assemble(on, "r", 1, npx, npy), o1.get("r")[1:-1, 1:-1])
    assert abs(r1 - rn) <= 1e-13 * r1
    o1.fine2coarse(1); on.fine2coarse(1)
    assert np.array_equal(assemble(on, "b", 2, npx, npy), o1.get("b", 2)[1:-1, 1:-1])
    rng = np.random.default_rng(5)
    pc = rng.standard_normal((nxg // 2, nyg // 2, nz // 2))
    for o, px, py in ((o1, 1, 1), (on, npx, npy)):
        nxc, nyc = nxg // 2 // px, nyg // 2 // py
        for r in range(px * py):
            pi, pj = r % px, r // px
            a = np.zeros((nxc + 2, nyc + 2, nz // 2)); a[1:-1, 1:-1] = pc[pi * nxc:(pi + 1) * nxc, pj * nyc:(pj + 1) * nyc]
            o.set("p", a, 2, r)
        o.fill_halo(2, "p"); o.coarse2fine(1)
    assert np.array_equal(assemble(on, "p", 1, npx, npy), o1.get("p")[1:-1, 1:-1])


def test_multirank_solve_converges_like_single_rank():
    o1 = make_case("seamount", 32, 32, 16, 1, 1, solver_prec=1e-10); o1.compute_rhs(); o1.solve_p()
    o4 = make_case("seamount", 16, 16, 16, 2, 2, solver_prec=1e-10); o4.compute_rhs(); o4.solve_p()
    og = make_case("seamount", 16, 16, 16, 2, 2, solver_prec=1e-10, nsmall=16); og.compute_rhs(); og.solve_p()
    h1, h4, hg = o1.history()[0], o4.history()[0], og.history()[0]
    assert len(h1) == len(h4) == len(hg)
    assert np.all(np.abs(h4 - h1) <= 5e-3 * h1) and np.all(np.abs(hg - h1) <= 5e-3 * h1)   # legit 1e-4-level differences
    assert np.any(h4 != h1)
