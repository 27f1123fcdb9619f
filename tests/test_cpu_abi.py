"""CPU tests of the product's host side: the C-ABI library loads without a GPU and exports every symbol include/mgb.h
declares; namelist reader; level/neighbour/gather planning against the oracle; compute calls fail loudly without a GPU."""
import ctypes
import os
import re

import pytest

from oracle.pyoracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(nh):
    hdr = open(os.path.join(ROOT, "include", "mgb.h")).read()
    names = sorted(set(re.findall(r"\b(mgb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 35
    lib = ctypes.CDLL(os.path.join(ROOT, "mgroms_b200", "libmgb.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in nh.load().mgb_version()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mgroms_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "mgoracle" not in txt and "import oracle" not in txt, f


NAMELIST = """!- Default value are available in mg_namelist.f90 file -!
&nhparam
       solver_prec    = 1.d-12,         !- Solver precision
       solver_maxiter = 50,            !- maximum of solver iterations
       nsmall         = 8,
       ns_coarsest    = 40,
       ns_pre         = 3,
       ns_post        = 2,
       cmatrix        = 'real',        !- Take the 'simple' or 'real' matric coefs
       relax_method   = 'RB',          !- 'Gauss-Seidel','GS','Red-Black','RB','Four-Color','FC'
       aggressive     = .false.,
       interp_type    = 'linear',
       restrict_type  = 'avg',
       netcdf_output  = .true.,
       bmask          = .false.,       !- true = Nico Jr version using a mask on boundary
/
"""


def test_namelist_reader(nh, tmp_path):
    d = nh.namelist_probe(str(tmp_path / "absent"))           # no file: defaults of mg_namelist.f90:11-35
    assert d["file"] == "0" and float(d["solver_prec"]) == 1e-6 and d["ns_coarsest"] == "40" and d["relax_method"] == "RB"
    f = tmp_path / "nh_namelist"; f.write_text(NAMELIST)
    d = nh.namelist_probe(str(f))
    assert d["file"] == "1" and float(d["solver_prec"]) == 1e-12 and d["netcdf_output"] == "1" and d["interp_type"] == "linear"
    nh.set_params(solver_prec=1e-10, nsmall=32)               # programmatic overrides win over the file
    d = nh.namelist_probe(str(f))
    assert float(d["solver_prec"]) == 1e-10 and d["nsmall"] == "32"
    nh.reset_params()
    bad = tmp_path / "stale"; bad.write_text("&nhparam\n nhalo = 1,\n red_black = .true.\n/\n")   # examples/namelist/nh_namelist is stale like this
    with pytest.raises(nh.MgbError, match="unknown key"):
        nh.namelist_probe(str(bad))
    lin = tmp_path / "lin"; lin.write_text("&nhparam\n interp_type='linear', restrict_type='linear'\n/\n")
    with pytest.raises(nh.MgbError, match="not permitted"):
        nh.namelist_probe(str(lin))


CONFIGS = [  # (nx, ny, nz local, npx, npy, nsmall)
    (64, 64, 64, 1, 1, 8), (512, 512, 64, 1, 1, 8), (1024, 1024, 128, 1, 1, 8), (512, 1024, 128, 4, 2, 8), (512, 1024, 128, 4, 2, 32),
    (1024, 2048, 128, 4, 2, 8), (16, 16, 32, 2, 2, 16), (16, 32, 8, 4, 2, 16), (256, 512, 64, 2, 1, 64), (512, 512, 64, 4, 2, 64),
]


@pytest.mark.parametrize("cfg", CONFIGS)
def test_level_plan_matches_oracle(nh, cfg):
    nx, ny, nz, npx, npy, nsmall = cfg
    big = nx * ny * nz > 10 ** 7
    keys = ["nx", "ny", "nz", "npx", "npy", "incx", "incy", "gather", "ngx", "ngy"]
    nh.set_params(nsmall=nsmall)
    if big:
        # the oracle allocates its fields; for big grids compare against the closed forms of SURVEY.md section 8 instead
        plan = nh.plan_levels(nx, ny, nz, npx, npy, 0)
        import math
        nl = min(1 + int(math.log2(min(nx * npx, ny * npy) // 4)), 1 + int(math.log2(nz // 2)))
        assert len(plan) == nl and plan[-1]["nz"] == nz >> (nl - 1)
        return
    o = Oracle(nx, ny, nz, npx, npy, nsmall=nsmall)
    for rank in range(npx * npy):
        plan = nh.plan_levels(nx, ny, nz, npx, npy, rank)
        assert len(plan) == o.nlevs
        for lev, d in enumerate(plan, 1):
            ref = o.level_info(lev, rank)
            assert {k: d[k] for k in keys} == {k: ref[k] for k in keys}, (rank, lev)
            assert d["neighb"] == ref["neighb"], (rank, lev)
            if d["gather"]:
                assert d["key"] == ref["key"] and d["members"] == o.gather_members(lev, rank), (rank, lev)


def test_c4_gather_worked_example(nh):
    """8 ranks 4x2, local 512x1024x128, nsmall=32 -> lev 6 = 32x64x4 on 2x1 (gather 2x2),
    lev 7 = 32x32x2 on 1x1 (gather 2x1); lev-6 families {0,1,4,5},{2,3,6,7}; lev-7 pairs {0,2},{1,3},{4,6},{5,7}."""
    nh.set_params(nsmall=32)
    p0 = nh.plan_levels(512, 1024, 128, 4, 2, 0)
    assert len(p0) == 7
    assert (p0[5]["nx"], p0[5]["ny"], p0[5]["nz"], p0[5]["npx"], p0[5]["npy"], p0[5]["gather"]) == (32, 64, 4, 2, 1, 1)
    assert (p0[6]["nx"], p0[6]["ny"], p0[6]["nz"], p0[6]["npx"], p0[6]["npy"], p0[6]["gather"]) == (32, 32, 2, 1, 1, 1)
    assert p0[5]["members"] == [0, 1, 4, 5] and p0[6]["members"] == [0, 2]
    p3 = nh.plan_levels(512, 1024, 128, 4, 2, 3)
    assert p3[5]["members"] == [2, 3, 6, 7] and p3[6]["members"] == [1, 3]
    p6 = nh.plan_levels(512, 1024, 128, 4, 2, 6)
    assert p6[6]["members"] == [4, 6] and p6[6]["key"] == 1


def test_bad_arguments_are_rejected(nh):
    for args in [(48, 64, 64, 1, 1, 0), (64, 64, 64, 3, 1, 0), (64, 64, 64, 2, 2, 4), (2, 2, 64, 1, 1, 0), (64, 64, 1, 1, 1, 0)]:
        with pytest.raises(nh.MgbError):
            nh.plan_levels(*args)
    nh.set_params(aggressive=True)      # unfinished in the reference (src/mg_intergrids.f90:243): refused, not emulated
    with pytest.raises(nh.MgbError, match="aggressive"):
        nh.plan_levels(64, 64, 64, 1, 1, 0)
    nh.reset_params(); nh.set_params(bmask=True)
    assert len(nh.plan_levels(64, 64, 64, 1, 1, 0)) == 5


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu(nh):
    with pytest.raises(nh.MgbError, match="no usable CUDA device"):
        nh.nhydro_init(16, 16, 16)
    with pytest.raises(nh.MgbError):
        nh.solve_p()


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_plain_c_driver_builds_and_fails_loudly_without_gpu(tmp_path):
    """examples/mg_testseamount.c compiles against include/mgb.h + libmgb.so with gcc alone and, without a device, stops with the
    library's message (the reference `stop`s) instead of computing anything on the CPU."""
    import subprocess
    import __graft_entry__ as ge
    ge.build()
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "mg_testseamount")
    out = subprocess.run([exe, "16", "16", "16"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 and "no usable CUDA device" in out.stderr


@pytest.mark.skipif(not os.path.exists("/root/reference/src/nh_namelist"), reason="needs the reference tree (this container)")
def test_namelist_reader_on_the_references_own_file(nh):
    """The namelist file the reference ships (src/nh_namelist), read in place by the product's parser."""
    d = nh.namelist_probe("/root/reference/src/nh_namelist")
    assert d["file"] == "1"
    assert float(d["solver_prec"]) == 1e-12 and d["solver_maxiter"] == "50" and d["nsmall"] == "8"
    assert (d["ns_coarsest"], d["ns_pre"], d["ns_post"]) == ("40", "3", "2")
    assert (d["cmatrix"], d["relax_method"], d["interp_type"], d["restrict_type"]) == ("real", "RB", "linear", "avg")
    assert (d["aggressive"], d["netcdf_output"], d["bmask"]) == ("0", "1", "0")
