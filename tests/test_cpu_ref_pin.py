"""The oracle pinned against the REFERENCE'S OWN SOURCE.

tests/golden/ref_golden.* hold output of the reference's Fortran routines, translated statement by statement to C
(oracle/f90toc.py; no Fortran compiler in the image) and run on one rank by oracle/refdriver.py.  The CPU oracle
(oracle/mgoracle.cpp, relax_order=0 = the reference's sweep order) must reproduce them BIT FOR BIT: stencil coefficients on
every level, right-hand side, the whole residual history, the final pressure and the corrected velocities, for the
reference's own 64^3 seamount case (BASELINE configs[0]) and for every geometry / namelist switch of the other drivers.
The rank-1 form the CUDA path uses (relax_order=1) must stay within the north-star tolerances of the same vectors.
When the translated library is present (this container, or prebuilt on the GPU box) a live sweep over more sizes runs too,
and the translator itself is checked on hand-written Fortran snippets whose results are known in closed form."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from common import history_close
from refgolden import digest, load, load_ops, seeded_field, sub

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = load()


def oracle_case(c, relax_order):
    from oracle.pyoracle import make_case
    nx, ny, nz = c["dims"]
    return make_case(c["kind"], nx, ny, nz, 1, 1, uvw=c["uvw"], bench_dx=bool(c.get("bench_dx")), relax_order=relax_order,
                     solver_prec=c["tol"], solver_maxiter=c["maxite"], **c["params"])


@pytest.mark.parametrize("c", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_oracle_reproduces_the_reference_bit_for_bit(c):
    o = oracle_case(c, 0)
    assert o.nlevs == c["nlevs"]
    for lev in range(1, o.nlevs + 1):
        li = o.level_info(lev)
        assert [li["nx"], li["ny"], li["nz"]] == c["level_dims"][lev - 1]
        assert digest(o.get("cA", lev)) == c["sha_cA"][lev - 1], ("cA", lev)
    o.compute_rhs()
    assert digest(o.get("b")[1:-1, 1:-1]) == c["sha_b_interior"]
    o.solve_p()
    res, conv = o.history()
    assert [float(x) for x in res] == c["hist_res"], (list(res), c["hist_res"])          # exact equality of every norm
    assert [float(x) for x in conv[1:len(res)]] == c["hist_conv"]
    assert digest(o.get("p")) == c["sha_p"]
    o.correct_uvw()
    u, v, w = o.get_uvw(0)
    assert (digest(u), digest(v), digest(w)) == (c["sha_u"], c["sha_v"], c["sha_w"])
    o.close()


@pytest.mark.parametrize("c", [c for c in G["cases"] if c["params"].get("relax_method", "RB") == "RB"], ids=lambda c: c["name"])
def test_rank1_form_of_the_smoother_stays_within_tolerance_of_the_reference(c):
    """relax_order=1 (parallel column solves + rank-1 correction, the form of the CUDA kernels) against REFERENCE output:
    per-cycle norms to 1e-10 relative (+ round-off floor), final p to 1e-9 relative -- the north-star tolerances."""
    o = oracle_case(c, 1)
    o.compute_rhs(); o.solve_p()
    history_close(o.history()[0], np.array(c["hist_res"]))
    d = np.abs(sub(o.get("p"), c["sub_stride"]) - load_ops()[c["arrays"] + "/p_sub"]).max()
    assert d <= 1e-9 * c["p_max"], d / c["p_max"]
    o.close()


def oracle_multi(c, relax_order):
    from oracle.pyoracle import make_case
    nx, ny, nz = c["dims"]; npx, npy = c["grid"]
    return make_case(c["kind"], nx, ny, nz, npx, npy, uvw=c["uvw"], relax_order=relax_order, **c["params"])


@pytest.mark.parametrize("c", G["multi_cases"], ids=[c["key"] for c in G["multi_cases"]])
def test_oracle_multirank_reproduces_the_reference_bit_for_bit(c):
    """The decomposed runs: the reference's own halo exchange (8 neighbours, corner precedence), gather / split families,
    global_sum and level planning -- its Fortran text over an emulated MPI -- against the oracle's rank emulation."""
    o = oracle_multi(c, 0)
    N = c["grid"][0] * c["grid"][1]
    assert o.nlevs == c["nlevs"]
    for lev in range(1, o.nlevs + 1):
        li = o.level_info(lev, 0)
        assert {k: li[k] for k in c["levels"][lev - 1]} == c["levels"][lev - 1], lev
        assert [digest(o.get("cA", lev, r)) for r in range(N)] == c["sha_cA"][lev - 1], ("cA", lev)
    o.compute_rhs()
    assert [digest(o.get("b", 1, r)[1:-1, 1:-1]) for r in range(N)] == c["sha_b_interior"]
    o.solve_p()
    res, conv = o.history()
    assert [float(x) for x in res] == c["hist_res"]
    assert [float(x) for x in conv[1:len(res)]] == c["hist_conv"]
    assert [digest(o.get("p", 1, r)) for r in range(N)] == c["sha_p"]
    o.correct_uvw()
    assert [digest(o.get_uvw(r)[0]) for r in range(N)] == c["sha_u"]
    o.close()


@pytest.mark.parametrize("c", [c for c in G["multi_cases"] if c["hist_res"][-1] < 1.0], ids=lambda c: c["key"])
def test_rank1_form_multirank_stays_within_tolerance_of_the_reference(c):
    o = oracle_multi(c, 1)
    N = c["grid"][0] * c["grid"][1]
    o.compute_rhs(); o.solve_p()
    history_close(o.history()[0], np.array(c["hist_res"]))
    for r in range(N):
        d = np.abs(sub(o.get("p", 1, r), c["sub_stride"]) - load_ops()[f"{c['arrays']}/p_sub_r{r}"]).max()
        assert d <= 1e-9 * c["p_max"], (r, d / c["p_max"])
    o.close()


@pytest.mark.parametrize("oc", G["ops_cases"], ids=[c["name"] for c in G["ops_cases"]])
def test_oracle_operators_one_by_one(oc):
    from oracle.pyoracle import make_case
    from oracle.refdriver import neumann_nh1
    ops = load_ops(); key = oc["key"]
    nx, ny, nz = oc["dims"]
    o = make_case(oc["kind"], nx, ny, nz, 1, 1, uvw="random", relax_order=0, **oc["params"])
    for lev in range(1, o.nlevs + 1):
        assert np.array_equal(o.get("cA", lev), ops[f"{key}/cA{lev}"]), lev
    o.compute_rhs()
    assert np.array_equal(o.get("b")[1:-1, 1:-1], ops[f"{key}/b"][1:-1, 1:-1])
    p0 = seeded_field((nx + 2, ny + 2, nz), oc["p0_seed"]); neumann_nh1(p0)
    o.set("b", ops[f"{key}/b"]); o.set("p", p0)
    o.relax(1, 1)
    assert np.array_equal(o.get("p"), ops[f"{key}/p_after_relax1"])
    assert o.residual(1) == float(ops[f"{key}/res_after_relax1"][0])
    assert np.array_equal(o.get("r"), ops[f"{key}/r"])
    o.fine2coarse(1)
    assert np.array_equal(o.get("b", 2), ops[f"{key}/b2"])
    p2 = seeded_field(o.get("p", 2).shape, oc["p2_seed"]); neumann_nh1(p2)
    o.set("p", p2, 2)
    o.coarse2fine(1)
    assert np.array_equal(o.get("p"), ops[f"{key}/p_after_prolong"])
    assert np.array_equal(o.get("r"), ops[f"{key}/r_interp"])
    o.set("p", p0); o.set("b", ops[f"{key}/b"])
    o.vcycle(1)
    assert np.array_equal(o.get("p"), ops[f"{key}/p_after_vcycle"])
    o.close()


def _ref_available():
    from oracle import refdriver
    return refdriver.available()


@pytest.mark.skipif(not _ref_available(), reason="needs oracle/_ref/libmgref.so (built from /root/reference by oracle/build_ref.py)")
@pytest.mark.parametrize("kind,nx,ny,nz,prm", [
    ("seamount", 24, 40, 12, {}), ("rndtopo", 48, 16, 32, {}), ("cuc", 16, 16, 64, {}), ("cuc", 40, 24, 16, dict(bmask=True)),
    ("rndtopo", 32, 32, 8, dict(relax_method="FC", interp_type="nearest")), ("seamount", 16, 48, 16, dict(relax_method="GS", bmask=True)),
])
def test_live_translated_reference_against_oracle(kind, nx, ny, nz, prm):
    """Sizes / switches beyond the committed vectors, reference and oracle run side by side (non-cubic, non-power-of-two)."""
    from mgroms_b200 import testcases as tc
    from oracle.pyoracle import make_case
    from oracle.refdriver import RefSolver
    geo = {"seamount": tc.seamount, "rndtopo": tc.rndtopo, "cuc": tc.cuc}[kind](nx, ny, nz, 1, 1, 0)
    vert = tc.CUC_VERT if kind == "cuc" else tc.SEAMOUNT_VERT
    R = RefSolver(nx, ny, nz, **prm)
    rm = tc.bmask_rmask(nx, ny) if prm.get("bmask") else None
    R.matrices(*geo, rm, vert["hc"], vert["theta_b"], vert["theta_s"])
    o = make_case(kind, nx, ny, nz, 1, 1, uvw="random", relax_order=0, solver_prec=1e-11, solver_maxiter=12, **prm)
    assert R.nlevs == o.nlevs
    for lev in range(1, R.nlevs + 1):
        for name in ("dx", "dy", "zeta", "h", "zr", "zw", "cw", "cA"):
            assert np.array_equal(R.get(name, lev), o.get(name, lev)), (name, lev)
    u, v, w = tc.uvw_random(nx, ny, nz, 1, 1, 0)
    R.compute_rhs(rm, u, v, w); o.compute_rhs()
    assert np.array_equal(R.get("b")[1:-1, 1:-1], o.get("b")[1:-1, 1:-1])
    hr, cr = R.solve_p(1e-11, 12); o.solve_p()
    ho, co = o.history()
    assert np.array_equal(hr, ho) and np.array_equal(cr[1:], co[1:len(ho)])
    assert np.array_equal(R.get("p"), o.get("p")) and np.array_equal(R.get("r"), o.get("r"))
    R.correct_uvw(u, v, w); o.correct_uvw()
    for a, b in zip((u, v, w), o.get_uvw(0)):
        assert np.array_equal(a, b)
    o.close()


@pytest.mark.skipif(not _ref_available(), reason="needs oracle/_ref/libmgref.so")
@pytest.mark.parametrize("kind,nx,ny,nz,npx,npy,prm", [
    ("cuc", 16, 16, 8, 4, 2, dict(nsmall=32)), ("seamount", 24, 16, 12, 2, 2, {}), ("rndtopo", 16, 16, 16, 4, 2, dict(bmask=True, nsmall=16)),
    ("seamount", 16, 16, 8, 2, 2, dict(relax_method="FC", interp_type="nearest")),
])
def test_live_translated_reference_multirank_against_oracle(kind, nx, ny, nz, npx, npy, prm):
    """Decompositions beyond the committed vectors, translated reference (threads + emulated MPI) and oracle side by side:
    every field of every level on every rank, including the halo rings the exchange fills."""
    from mgroms_b200 import testcases as tc
    from oracle.pyoracle import make_case
    from oracle.refdriver import RefWorld
    PR = RefWorld.PerRank
    N = npx * npy
    gen = {"seamount": tc.seamount, "rndtopo": tc.rndtopo, "cuc": tc.cuc}[kind]
    vert = tc.CUC_VERT if kind == "cuc" else tc.SEAMOUNT_VERT
    W = RefWorld(nx, ny, nz, npx, npy, **prm)
    try:
        geos = [gen(nx, ny, nz, npx, npy, r) for r in range(N)]
        rms = PR([tc.bmask_rmask(nx, ny, npx, npy, r) if prm.get("bmask") else None for r in range(N)])
        W.call("matrices", *[PR([g[q] for g in geos]) for q in range(4)], rms, vert["hc"], vert["theta_b"], vert["theta_s"])
        o = make_case(kind, nx, ny, nz, npx, npy, uvw="random", relax_order=0, solver_prec=1e-11, solver_maxiter=8, **prm)
        assert W.nlevs == o.nlevs
        for lev in range(1, W.nlevs + 1):
            infos = W.call("level_info", lev)
            for r in range(N):
                oi = o.level_info(lev, r)
                assert {k: infos[r][k] for k in ("nx", "ny", "nz", "npx", "npy", "incx", "incy", "gather", "ngx", "ngy")} == \
                       {k: oi[k] for k in ("nx", "ny", "nz", "npx", "npy", "incx", "incy", "gather", "ngx", "ngy")}
                assert [x if x >= 0 else -1 for x in infos[r]["neighb"]] == [x if x >= 0 else -1 for x in oi["neighb"]], (lev, r)
            for name in ("dx", "dy", "zeta", "h", "zr", "zw", "cw", "cA"):
                got = W.get(name, lev)
                for r in range(N):
                    assert np.array_equal(got[r], o.get(name, lev, r)), (name, lev, r)
        uvw = [tc.uvw_random(nx, ny, nz, npx, npy, r) for r in range(N)]
        W.call("compute_rhs", rms, *[PR([x[q] for x in uvw]) for q in range(3)]); o.compute_rhs()
        hist = W.call("solve_p", 1e-11, 8)[0]; o.solve_p()
        assert np.array_equal(hist[0], o.history()[0])
        for name in ("p", "r"):
            got = W.get(name)
            for r in range(N):
                assert np.array_equal(got[r], o.get(name, 1, r)), (name, r)
        # one exchange of a field that differs on every rank and every cell: edges, corners, physical mirrors
        for lev in (1, W.nlevs):
            shp = o.get("p", lev, 0).shape
            fields = [seeded_field(shp, 100 * lev + r) for r in range(N)]
            for r in range(N):
                o.set("p", fields[r], lev, r)
            views = W.call("get", "p", lev)
            for r in range(N):
                views[r][...] = fields[r]
            W.call("fill_halo_relax", lev, "p"); o.fill_halo(lev, "p")
            got = W.get("p", lev)
            for r in range(N):
                assert np.array_equal(got[r], o.get("p", lev, r)), ("halo", lev, r)
        o.close()
    finally:
        W.close()


@pytest.mark.skipif(not _ref_available(), reason="needs oracle/_ref/libmgref.so")
def test_seamount_inputs_equal_the_reference_drivers_own_setup():
    """mgroms_b200/testcases.py::seamount against the translated setup_seamount of the reference's driver
    (src/mg_setup_tests.f90:108-157), one rank and 2x2 ranks: dx, dy, zeta bit for bit; h to 1 ulp (NumPy's vectorised exp and
    libm's exp differ in the last bit in a few cells -- it is an INPUT: every implementation under test receives the same array)."""
    from mgroms_b200 import testcases as tc
    from oracle.refdriver import RefSolver, RefWorld

    def same(got, want, where):
        for n, a, b in zip(("dx", "dy", "zeta", "h"), got, want):
            if n == "h":
                assert np.abs(a - b).max() <= 2.3e-16 * np.abs(b).max(), (where, n)
            else:
                assert np.array_equal(a, b), (where, n)
    R = RefSolver(16, 24, 8)
    same(R.setup_seamount(1, 1), tc.seamount(16, 24, 8, 1, 1, 0), "1x1")
    W = RefWorld(16, 8, 8, 2, 2)
    try:
        got = W.call("setup_seamount", 2, 2)
        for r in range(4):
            same(got[r], tc.seamount(16, 8, 8, 2, 2, r), f"2x2 rank {r}")
    finally:
        W.close()


@pytest.mark.skipif(not _ref_available(), reason="needs oracle/_ref/libmgref.so")
@pytest.mark.parametrize("cfg", [(16, 16, 8, 2, 1, 8), (16, 32, 16, 4, 2, 16), (8, 8, 16, 4, 2, 8), (32, 16, 8, 2, 2, 32), (16, 16, 16, 8, 4, 16),
                                 (16, 16, 16, 1, 2, 16), (8, 16, 32, 4, 4, 64), (64, 32, 16, 2, 4, 8), (16, 16, 64, 2, 2, 8)])
def test_product_level_planner_against_the_translated_reference(nh, cfg):
    """PRODUCT host logic (mgroms_b200/csrc/host_plan.cpp through mgb_plan_levels, no device needed) against the reference's
    own define_grids / define_grid_dims / define_neighbours / define_gather_informations (src/mg_grids.f90:468-738), translated:
    number of levels, sizes, process grid, increments, gather flags and factors, the 8 neighbours and the gather key of
    every rank on every level."""
    from oracle.refdriver import RefWorld
    nx, ny, nz, npx, npy, nsmall = cfg
    nh.set_params(nsmall=nsmall)
    W = RefWorld(nx, ny, nz, npx, npy, nsmall=nsmall)
    try:
        keys = ["nx", "ny", "nz", "npx", "npy", "incx", "incy", "gather", "ngx", "ngy"]
        plans = [nh.plan_levels(nx, ny, nz, npx, npy, r) for r in range(npx * npy)]
        assert all(len(p) == W.nlevs for p in plans)
        for lev in range(1, W.nlevs + 1):
            infos = W.call("level_info", lev)
            for r in range(npx * npy):
                d = plans[r][lev - 1]
                assert {k: d[k] for k in keys} == {k: infos[r][k] for k in keys}, (r, lev)
                assert [x if x >= 0 else -1 for x in d["neighb"]] == [x if x >= 0 else -1 for x in infos[r]["neighb"]], (r, lev)
                if d["gather"]:
                    assert d["key"] == infos[r]["key"], (r, lev)
    finally:
        W.close()


SNIPPET = """
module snip
  implicit none
  real(kind=8), parameter :: hlf = 0.5_8
contains
  subroutine kernel(n, m, a, b, s, t)
    integer(kind=4), intent(in) :: n, m
    real(kind=8), dimension(0:n+1,-1:m+2), intent(inout) :: a
    real(kind=8), dimension(n), intent(out) :: b
    real(kind=8), intent(out) :: s
    integer(kind=4), intent(out) :: t
    integer(kind=4) :: i, j, k
    real(kind=8), dimension(:,:), pointer :: w
    allocate(w(0:n+1,2))
    w(:,:) = 2._8
    ! section copies with negative strides and the 2A - B extrapolation
    a(1:n,m+1:m+2) = a(1:n,m:m-1:-1)
    a(1:n,-1) = 2._8 * a(1:n,1) - a(1:n,2)
    a(0,:) = -1.5_8
    s = 0._8
    t = 0
    do i = n, 1, -2            ! negative step
       t = t + i / 2           ! integer division
    enddo
    do j = 1, m
       do i = 1, n
          k = 1 + mod(i + j, 2)
          if (k == 1 .and. .not. (i > 3)) then
             s = s - a(i,j)**2 * hlf + w(i,1)
          elseif (k == 2) then
             s = s + (a(i,j) - a(i-1,j)) / real(i + j, kind=8)
          else
             s = s + 1.e-1          ! default-real literal: single precision 0.1
          endif
       enddo
    enddo
    b(:) = 0._8
    do i = 1, n
       b(i) = sum(a(i,1:m)) + max(a(i,1), a(i,m), 0.25_8)
    enddo
    select case (trim(mode))
    case ('x', 'double')
       s = s * 2._8
    case default
       s = -s
    end select
    deallocate(w)
  end subroutine kernel
end module snip
"""


def test_translator_semantics_on_a_known_snippet(tmp_path):
    """f90toc on hand-written Fortran (not reference code) whose result is recomputed in NumPy: array sections with negative
    strides, bounds with negative lower limits, integer division, mod, **2, mixed-precision literals, do with a negative
    step, if / elseif / else, select case on strings, sum(), max(), allocate / deallocate."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import f90toc
    src = tmp_path / "snip.f90"; src.write_text(SNIPPET)
    f90toc.GLOBALS["mode"] = "char"
    tr = f90toc.Translator({}); tr.add(str(src), ["kernel"])
    c = tmp_path / "snip.c"
    c.write_text('const char* g_mode = "double";\n' + tr.emit())
    (tmp_path / "ref_types.h").write_text(tr.emit_types() + "\n")
    so = tmp_path / "snip.so"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-std=gnu11", "-I", os.path.join(ROOT, "oracle", "ref"), "-I", str(tmp_path),
                           str(c), "-lm", "-o", str(so)])
    L = ctypes.CDLL(str(so))
    n, m = 7, 5
    rng = np.random.default_rng(3)
    a = rng.standard_normal((m + 4, n + 2))            # C-order reversed: a[j + 1, i]
    a0 = a.copy()
    b = np.zeros(n); s = ctypes.c_double(); t = ctypes.c_int()
    PD = ctypes.POINTER(ctypes.c_double)
    L.ref_kernel(ctypes.c_int(n), ctypes.c_int(m), a.ctypes.data_as(PD), b.ctypes.data_as(PD), ctypes.byref(s), ctypes.byref(t))
    # the same statements in NumPy (Fortran a(i,j) == e[j + 1, i])
    e = a0.copy()
    e[m + 2, 1:n + 1] = e[m + 1, 1:n + 1]; e[m + 3, 1:n + 1] = e[m, 1:n + 1]
    e[0, 1:n + 1] = 2.0 * e[2, 1:n + 1] - e[3, 1:n + 1]
    e[:, 0] = -1.5
    assert np.array_equal(a, e)
    assert t.value == sum(i // 2 for i in range(n, 0, -2))
    acc = 0.0
    for j in range(1, m + 1):
        for i in range(1, n + 1):
            k = 1 + (i + j) % 2
            if k == 1 and not i > 3:
                acc = acc - e[j + 1, i] * e[j + 1, i] * 0.5 + 2.0
            elif k == 2:
                acc = acc + (e[j + 1, i] - e[j + 1, i - 1]) / float(i + j)
            else:
                acc = acc + float(np.float32(0.1))
    assert s.value == acc * 2.0
    for i in range(1, n + 1):
        row = 0.0
        for j in range(1, m + 1):
            row = row + e[j + 1, i]
        assert b[i - 1] == row + max(e[2, i], e[m + 1, i], 0.25)


SNIPPET2 = """
module snip2
  implicit none
  type box_type
     integer(kind=4) :: n
     integer(kind=4),dimension(3) :: tag
     real(kind=8),dimension(:,:),pointer :: a => null()
  end type box_type
  type(box_type), dimension(:), pointer :: box
  interface twice
     module procedure twice_1d, twice_2d
  end interface twice
contains
  subroutine twice_1d(x, n)
    integer(kind=4), intent(in) :: n
    real(kind=8), dimension(n), intent(inout) :: x
    x(:) = 2._8 * x(:)
  end subroutine twice_1d
  subroutine twice_2d(x, scale)
    real(kind=8), dimension(:,:), pointer, intent(inout) :: x
    real(kind=8), optional, intent(in) :: scale
    if (present(scale)) then
       x(:,:) = scale * x(:,:)
    else
       x(:,:) = 2._8 * x(:,:)
    endif
  end subroutine twice_2d
  subroutine driver(lev, m, total, count)
    integer(kind=4), intent(in) :: lev, m
    real(kind=8), intent(out) :: total
    integer(kind=4) :: count
    integer(kind=4), dimension(4) :: hits
    real(kind=8), dimension(m) :: v
    integer(kind=4) :: i, j
    box(lev)%n = m
    box(lev)%tag(2) = 7
    allocate(box(lev)%a(0:m+1,-1:1))
    do j = -1, 1
       do i = 0, m+1
          box(lev)%a(i,j) = real(i,kind=8) + 10._8 * real(j,kind=8)
       enddo
    enddo
    call twice(box(lev)%a)                  ! generic -> twice_2d (rank 2), optional argument absent
    call twice(box(lev)%a, scale=0.5_8)     ! keyword argument
    v(:) = box(lev)%a(1:m,1)
    call twice(v, m)                        ! generic -> twice_1d
    hits(:) = 0
    i = 0
    do while (i < 2*m)
       i = i + 3
       hits(1 + mod(i,4)) = hits(1 + mod(i,4)) + 1
    enddo
    count = hits(1) + 10*hits(2) + 100*hits(3) + 1000*hits(4) + box(lev)%tag(2)
    total = sum(v(1:m)) + box(lev)%a(0,-1) + size(box(lev)%a,dim=1)
    deallocate(box(lev)%a)
  end subroutine driver
end module snip2
"""


def test_translator_derived_types_generics_and_integer_arrays(tmp_path):
    """Second self-test of f90toc on hand-written Fortran: derived type with scalar, fixed-size and pointer members,
    allocate / deallocate on a member, generic interface resolved by rank, optional and keyword arguments, integer work
    arrays, do while, by-reference scalar results, size()."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import f90toc
    src = tmp_path / "snip2.f90"; src.write_text(SNIPPET2)
    tr = f90toc.Translator({}); tr.add(str(src), ["twice_1d", "twice_2d", "driver"])
    (tmp_path / "ref_types.h").write_text(tr.emit_types() + "\n")
    c = tmp_path / "snip2.c"
    c.write_text(tr.emit() + "\n__thread box_type BOX_[REF_MAXLEV];\n")
    so = tmp_path / "snip2.so"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-std=gnu11", "-I", os.path.join(ROOT, "oracle", "ref"), "-I", str(tmp_path),
                           str(c), "-lm", "-o", str(so)])
    L = ctypes.CDLL(str(so))
    m = 6
    total = ctypes.c_double(); count = ctypes.c_int()
    L.ref_driver(ctypes.c_int(3), ctypes.c_int(m), ctypes.byref(total), ctypes.byref(count))
    a = np.array([[i + 10.0 * j for j in (-1, 0, 1)] for i in range(m + 2)])        # a[i, j+1]
    a = 0.5 * (2.0 * a)
    v = 2.0 * a[1:m + 1, 2]
    hits = [0, 0, 0, 0]
    i = 0
    while i < 2 * m:
        i += 3
        hits[i % 4] += 1
    assert count.value == hits[0] + 10 * hits[1] + 100 * hits[2] + 1000 * hits[3] + 7
    assert total.value == float(np.sum(v)) + a[0, 0] + (m + 2)
