"""Kernel-selection variants against the oracle, each in its own process (the library reads its knobs once per process):
the fused one-launch-per-colour relaxation (kernels_relax.cu) forced onto small grids, with one and with several tiles per
CTA and in every warp layout; and the unfused fall-back paths (separate halo launches, A/B/C launches) that the default
configuration never takes on one GPU (MGB_NO_FUSED_HALO left the k=1 plane's halo stale)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

FUSED = dict(MGB_FUSED_RELAX="1", MGB_NO_SMALL_KERNEL="1", MGB_NO_MID_KERNEL="1", MGB_FUSED_MIN_TILES="32", MGB_EXPECT_PATH="fused")
# (levels with nz = 16 or fewer layers never take the fused kernel; coarse vertical grids need 12-14 warm-up rows of the bottom
# recurrence, more than its one-batch cone fetch holds -- these grids need 5-10)
GRIDS = [("cuc", 64, 128, 64), ("cuc", 128, 64, 64), ("seamount", 64, 64, 128), ("seamount", 32, 128, 128, "bench"), ("rndtopo", 64, 64, 128)]


def run(env, *case):
    e = dict(os.environ); e.update(env); e.setdefault("MGB_SPIN_TIMEOUT_MS", "3000")
    r = subprocess.run([sys.executable, os.path.join(HERE, "variant_worker.py"), *map(str, case)], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, (env, case, r.stdout[-2000:], r.stderr[-2000:])
    return r.stdout


@pytest.mark.parametrize("case", GRIDS)
@pytest.mark.parametrize("cap", ["0", "32", "40"])
def test_fused_relax_forced_on_small_grids(case, cap):
    run(dict(FUSED, MGB_F_GRID_CAP=cap), *case)


@pytest.mark.parametrize("variant", ["1", "2", "3"])
def test_fused_relax_warp_layouts(variant):
    run(dict(FUSED, MGB_F_VARIANT=variant, MGB_F_GRID_CAP="40", MGB_THOMAS="1"), "cuc", 128, 64, 64, "solve")
    run(dict(FUSED, MGB_F_VARIANT=variant), "seamount", 64, 64, 128)


def test_fused_relax_full_solve():
    run(dict(FUSED, MGB_F_GRID_CAP="40"), "cuc", 64, 128, 64, "solve")
    run(dict(FUSED, MGB_NO_FUSED_HALO="1"), "seamount", 64, 64, 128, "solve")      # mirror = 0: halo and k=1 ring by separate launches


@pytest.mark.parametrize("extra", [dict(MGB_FUSED_RELAX="1"), dict(), dict(MGB_NO_FUSED_BC="1"), dict(MGB_THOMAS="1")])
def test_unfused_fallback_paths(extra):
    env = dict(MGB_NO_FUSED_HALO="1", MGB_NO_MID_KERNEL="1", MGB_NO_SMALL_KERNEL="1", MGB_FUSED_MIN_TILES="32"); env.update(extra)
    env["MGB_EXPECT_PATH"] = "fused+halo" if "MGB_FUSED_RELAX" in extra else "a+b+c+halo"
    run(env, "cuc", 64, 128, 64, "solve")
    run(dict(MGB_NO_GRAPH="1"), "seamount", 32, 32, 32, "solve")


@pytest.mark.parametrize("case", [("seamount", 64, 64, 16), ("rndtopo", 32, 64, 32), ("cuc", 64, 32, 64), ("seamount", 32, 32, 128), ("rndtopo", 16, 40, 32)])
def test_bulk_copy_prolongation_forced_on_small_grids(case):
    """coarse2fine with the fine field staged by cp.async.bulk (kernels_bulk.cu), normally used on big levels only: the
    residual history and the final pressure of a full solve stay those of the oracle (1e-12: only the norm's summation differs)."""
    if case[2] == 40:
        pytest.skip("the reference requires power-of-two horizontal sizes")
    run(dict(MGB_BULK_PROLONG_MIN="0"), *case, "solve")
    run(dict(MGB_BULK_PROLONG_MIN="0", MGB_BULK_PROLONG_CTAS="1"), *case, "solve")
