"""fortran/nhydro_mgb.f90 cannot be compiled here (no Fortran compiler), so its ISO_C_BINDING interface blocks are checked
against include/mgb.h instead: every bind(C) function the shim declares must exist in the header with the same number of
arguments, the same argument kinds in the same order (by value int / double, pointer to double, pointer to char / bytes) and
the same result kind, the library must export it, and the shim must define the five procedures of the reference's
`module nhydro` (src/nhydro.f90:18,36,53,105,137) with the reference's argument lists."""
import ctypes
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import f90toc  # noqa: E402  (only its free-form source reader: comments, continuations, case)

SHIM = os.path.join(ROOT, "fortran", "nhydro_mgb.f90")


def shim_interfaces():
    lines = f90toc.logical_lines(SHIM)
    out, cur = {}, None
    for ln in lines:
        m = re.match(r"(integer\(c_int\)\s+)?function\s+(\w+)\s*\((.*?)\)\s*bind\(c,\s*name='(\w+)'\)(\s*result\((\w+)\))?", ln)
        if m:
            cur = dict(args=[a.strip() for a in m.group(3).split(",") if a.strip()], kinds={}, ret="int" if m.group(1) else None, res=m.group(6))
            out[m.group(4)] = cur
            continue
        if cur is None:
            continue
        if re.match(r"end\s+function", ln):
            cur = None
            continue
        m = re.match(r"(integer\(c_int\)|real\(c_double\)|character\(kind=c_char\)|type\(c_ptr\))\s*(,\s*value)?\s*(,\s*dimension\([^)]*\))?\s*::\s*(.*)$", ln)
        if m:
            base = {"integer(c_int)": "int", "real(c_double)": "double", "character(kind=c_char)": "char", "type(c_ptr)": "ptr"}[m.group(1)]
            for name in (x.strip() for x in m.group(4).split(",")):
                if name == cur.get("res"):
                    cur["ret"] = "ptr"
                else:
                    assert m.group(2) or m.group(3), f"{name}: a scalar without VALUE would be passed by reference"
                    cur["kinds"][name] = base if m.group(2) else base + "*"
    return out


def header_prototypes():
    text = open(os.path.join(ROOT, "include", "mgb.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(mgb_\w+)\s*\(([^)]*)\)\s*;", text):
        ret = "ptr" if "*" in m.group(1) else ("int" if re.search(r"\bint\b", m.group(1)) else m.group(1).strip())
        args = []
        for a in (x.strip() for x in m.group(3).split(",")):
            if a in ("", "void"):
                continue
            if "*" in a:
                args.append("double*" if "double" in a else "char*")         # const char* / void* / const void*: bytes
            else:
                args.append("double" if "double" in a else "int")
        protos[m.group(2)] = (ret, args)
    return protos


def test_bind_c_interfaces_match_the_header_and_the_library():
    from mgroms_b200 import nhydro
    lib = ctypes.CDLL(nhydro.load()._name) if hasattr(nhydro.load(), "_name") else nhydro.load()
    shim, protos = shim_interfaces(), header_prototypes()
    assert {"mgb_init", "mgb_comm_unique_id", "mgb_comm_init", "mgb_matrices", "mgb_solve", "mgb_check_nondivergence", "mgb_clean",
            "mgb_write_fort100", "mgb_last_error", "mgb_get_field"} <= set(shim)
    for name, f in shim.items():
        assert name in protos, f"{name} is not declared in include/mgb.h"
        ret, args = protos[name]
        assert f["ret"] == ret, (name, f["ret"], ret)
        got = [f["kinds"][a] for a in f["args"]]
        assert got == args, (name, got, args)
        assert hasattr(lib, name), f"libmgb.so does not export {name}"


def test_shim_defines_the_reference_api():
    lines = f90toc.logical_lines(SHIM)
    subs = {m.group(1): [a.strip() for a in m.group(2).split(",") if a.strip()] for m in (re.match(r"subroutine\s+(\w+)\s*\((.*)\)$", ln) for ln in lines) if m}
    subs.update({m.group(1): [] for m in (re.match(r"subroutine\s+(\w+)\s*(\(\s*\))?$", ln) for ln in lines) if m})
    want = {"nhydro_init": ["nx", "ny", "nz", "npxg", "npyg"],
            "nhydro_matrices": ["dx", "dy", "zeta", "h", "rmask", "hc", "theta_b", "theta_s"],
            "nhydro_solve": ["nx", "ny", "nz", "rmaska", "ua", "va", "wa"],
            "nhydro_check_nondivergence": ["nx", "ny", "nz", "rmaska", "ua", "va", "wa"],
            "nhydro_clean": []}
    for name, args in want.items():
        assert subs.get(name) == args, (name, subs.get(name))
    if os.path.isdir("/root/reference/src"):          # the same argument lists as the reference's module nhydro
        ref = f90toc.logical_lines("/root/reference/src/nhydro.f90")
        rsubs = {m.group(1): [a.strip() for a in (m.group(3) or "").split(",") if a.strip()] for m in (re.match(r"subroutine\s+(\w+)\s*(\((.*)\))?$", ln) for ln in ref) if m}
        for name, args in want.items():
            assert rsubs[name] == args, (name, rsubs[name])
