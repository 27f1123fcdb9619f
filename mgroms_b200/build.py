"""Builds libmgb.so (the CUDA product library) in-tree with nvcc for sm_100a.

    python -m mgroms_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libmgb.so")
SOURCES = ["host_plan.cpp", "kernels_solve.cu", "kernels_setup.cu", "solver.cu"]
HEADERS = ["mgb_internal.h", "kernels.cuh", os.path.join("..", "..", "include", "mgb.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-ccbin", ccbin,
           "-o", SO] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    if os.environ.get("MGB_BUILD_DEFS"):
        cmd[1:1] = ["-D" + d for d in os.environ["MGB_BUILD_DEFS"].split(",")]
    if verbose:
        cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
