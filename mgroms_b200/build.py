"""Builds libmgb.so (the CUDA product library) in-tree with nvcc for sm_100a.

    python -m mgroms_b200.build [--force] [--verbose]

Every translation unit is compiled to an object file in parallel (only the ones whose sources changed), then linked.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "obj")
SO = os.path.join(HERE, "libmgb.so")
SOURCES = ["host_plan.cpp", "kernels_solve.cu", "kernels_relax.cu", "kernels_bulk.cu", "kernels_setup.cu", "solver.cu"]
HEADERS = ["mgb_internal.h", "kernels.cuh", "device_utils.cuh", os.path.join("..", "..", "include", "mgb.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newest_header() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS + [os.path.join("..", "build.py")])


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if os.environ.get("MGB_BUILD_DEFS"):
        force = True                      # experiment builds (-D...) always recompile
    if not force and not needs_build():
        return SO
    os.makedirs(OBJ, exist_ok=True)
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    base = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
            "-fmad=false", "-Xcompiler", "-fPIC", "-ccbin", ccbin]
    if os.environ.get("MGB_BUILD_DEFS"):
        base += ["-D" + d for d in os.environ["MGB_BUILD_DEFS"].split(",")]
        force = True
    if verbose:
        base += ["-Xptxas", "-v"]
    hdr_t = _newest_header()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        path = os.path.join(CSRC, src)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(path), hdr_t):
            cmd = base + ["-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-ccbin", ccbin, "-o", SO] + objs + ["-ldl"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
