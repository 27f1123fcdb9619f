#!/usr/bin/env python
"""f90toc -- a statement-by-statement translator from the Fortran 90 subset used by the reference's arithmetic routines to C.

TEST INFRASTRUCTURE (oracle/): there is no Fortran compiler in this image (gfortran, flang, ifort, nvfortran: all absent),
so the reference cannot be compiled natively.  This tool reads the reference's OWN source files where they lie under
/root/reference/src, translates the named subroutines mechanically -- one Fortran statement becomes one C statement, every
expression keeps its operand order and its parentheses, `x**2` becomes `x*x` (what every Fortran compiler emits), integer
and real literals keep their types -- and writes the C text to oracle/_ref/ (git-ignored: no reference source, translated
or not, enters the repository).  `gcc -O2 -ffp-contract=off` then gives the IEEE semantics of the Fortran text, and
tests/golden/make_ref_golden.py runs the result on small inputs to produce the golden vectors that pin oracle/mgoracle.cpp.

Supported: subroutine headers; integer / real / logical / character declarations with dimension(...), pointer, intent,
parameter, save, optional; explicit-shape, automatic and deferred-shape (pointer) arrays with arbitrary lower bounds;
allocate / deallocate; pointer assignment (=>), also to members of the `grid(lev)%...` derived type; do / do while / if /
else if / else; scalar, element and array-SECTION assignments (elementwise, with strides); calls to other translated
subroutines, to call-backs supplied by the harness, and to a drop list (timers, NetCDF output, MPI); the intrinsics mod,
max, min, abs, sqrt, exp, log, cosh, sinh, tanh, real, int, size, lbound, trim, present, associated.
select case on strings; sum(array expression); `write(100,*) a, b` (the residual history) becomes a call-back.
Not supported (not needed): where, functions, array constructors, implied do, other I/O (write / print are dropped).
"""
from __future__ import annotations

import re
import sys

# --------------------------------------------------------------------------------------------------------------------
# source reading
# --------------------------------------------------------------------------------------------------------------------


def _strip_comment(line: str) -> str:
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch; out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def _lower_outside_strings(s: str) -> str:
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch; out.append(ch)
        else:
            out.append(ch.lower())
    return "".join(out)


def logical_lines(path: str) -> list[str]:
    """Comment-free, continuation-joined, lower-cased (outside strings) statements of a free-form source file."""
    res, cur = [], ""
    for raw in open(path, encoding="latin-1").read().splitlines():
        ln = _strip_comment(raw).strip()
        if not ln:
            continue
        if ln.startswith("&"):
            ln = ln[1:].lstrip()
        if ln.endswith("&"):
            cur += ln[:-1] + " "
            continue
        cur += ln
        for part in split_top(_lower_outside_strings(cur), ";"):      # `a = 1; b = 2` and stray trailing semicolons
            if part.strip():
                res.append(part.strip())
        cur = ""
    return res


def extract_unit(lines: list[str], name: str) -> list[str]:
    start = None
    for n, ln in enumerate(lines):
        if re.match(rf"subroutine\s+{name}\s*(\(|$)", ln):
            start = n
        elif start is not None and re.match(rf"end\s*subroutine(\s+{name})?\s*$", ln):
            return lines[start:n + 1]
    raise KeyError(f"subroutine {name} not found")


def module_decls(lines: list[str]) -> list[str]:
    """Declarations between `module x` and `contains` (module-level parameters such as hlf, qrt)."""
    out = []
    for ln in lines:
        if ln == "contains":
            break
        if re.match(r"(integer|real|logical|character)\b", ln):
            out.append(ln)
    return out


# --------------------------------------------------------------------------------------------------------------------
# tokens and expressions
# --------------------------------------------------------------------------------------------------------------------
_DOT = r"and|or|not|eq|ne|lt|le|gt|ge|eqv|neqv|true|false"
TOKEN_RE = re.compile(
    r"\s*(?:(?P<num>(?:\d+\.(?!(?:" + _DOT + r")\.)\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_\w+)?)"
    r"|(?P<name>[a-z_]\w*)"
    r"|(?P<str>'(?:[^']|'')*'|\"[^\"]*\")"
    r"|(?P<dot>\.(?:" + _DOT + r")\.)"
    r"|(?P<op>\*\*|==|/=|<=|>=|=>|::|[-+*/(),:=<>%]))")


def tokenize(s: str) -> list[tuple[str, str]]:
    toks, pos = [], 0
    s = s.rstrip()
    while pos < len(s):
        m = TOKEN_RE.match(s, pos)
        if not m or m.end() == pos:
            raise SyntaxError(f"cannot tokenize at {s[pos:pos + 30]!r} in {s!r}")
        kind = m.lastgroup
        toks.append((kind, m.group(kind)))
        pos = m.end()
    return toks


class P:
    """Recursive-descent parser for Fortran expressions (precedence of the standard)."""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek(); self.i += 1; return tok

    def accept(self, text):
        if self.peek()[1] == text and self.peek()[0] in ("op", "dot"):
            self.i += 1; return True
        return False

    def expect(self, text):
        if not self.accept(text):
            raise SyntaxError(f"expected {text!r}, got {self.peek()} in {self.t}")

    def expr(self):
        e = self.and_()
        while self.peek()[1] == ".or.":
            self.next(); e = ("bin", "||", e, self.and_())
        return e

    def and_(self):
        e = self.not_()
        while self.peek()[1] == ".and.":
            self.next(); e = ("bin", "&&", e, self.not_())
        return e

    def not_(self):
        if self.peek()[1] == ".not.":
            self.next(); return ("un", "!", self.not_())
        return self.rel()

    REL = {"==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">=", ".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}

    def rel(self):
        e = self.add()
        if self.peek()[1] in self.REL and self.peek()[0] in ("op", "dot"):
            op = self.REL[self.next()[1]]
            e = ("bin", op, e, self.add())
        return e

    def add(self):
        if self.peek() == ("op", "-"):
            self.next(); e = ("un", "-", self.mul())
        elif self.peek() == ("op", "+"):
            self.next(); e = ("un", "+", self.mul())
        else:
            e = self.mul()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.next()[1]; e = ("bin", op, e, self.mul())
        return e

    def mul(self):
        e = self.pow()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.next()[1]; e = ("bin", op, e, self.pow())
        return e

    def pow(self):
        e = self.primary()
        if self.peek() == ("op", "**"):
            self.next()
            # right associative; the exponent may carry a sign
            if self.peek() == ("op", "-"):
                self.next(); r = ("un", "-", self.pow())
            else:
                r = self.pow()
            e = ("pow", e, r)
        return e

    def args(self):
        out = []
        if self.accept(")"):
            return out
        while True:
            out.append(self.arg())
            if self.accept(")"):
                return out
            self.expect(",")

    def arg(self):
        if self.peek()[0] == "name" and self.peek(1) == ("op", "="):
            key = self.next()[1]; self.next()
            return ("kw", key, self.expr())
        lo = None
        if self.peek() != ("op", ":"):
            lo = self.expr()
            if self.peek() != ("op", ":"):
                return lo
        self.expect(":")
        hi = st = None
        if self.peek()[1] not in (",", ")", ":") or self.peek()[0] != "op":
            hi = self.expr()
        if self.accept(":"):
            st = self.expr()
        return ("sec", lo, hi, st)

    def primary(self):
        k, v = self.next()
        if k == "num":
            return ("num", v)
        if k == "str":
            return ("str", v[1:-1])
        if k == "dot" and v in (".true.", ".false."):
            return ("num", "1" if v == ".true." else "0")
        if (k, v) == ("op", "("):
            e = self.expr(); self.expect(")"); return ("par", e)
        if k == "name":
            node = ("name", v)
            if self.accept("("):
                node = ("call", node, self.args())
            while self.accept("%"):
                m = self.next()
                assert m[0] == "name", m
                node = ("mem", node, m[1])
                if self.accept("("):
                    node = ("call", node, self.args())
            return node
        raise SyntaxError(f"unexpected token {(k, v)} in {self.t}")


def parse_expr(s: str):
    p = P(tokenize(s))
    e = p.expr()
    if p.peek()[0] != "eof":
        raise SyntaxError(f"trailing tokens {p.t[p.i:]} in {s!r}")
    return e


def split_top(s: str, sep: str = ",") -> list[str]:
    out, depth, cur, q = [], 0, "", None
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch; cur += ch; continue
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def match_paren(s: str, start: int) -> int:
    depth, q = 0, None
    for n in range(start, len(s)):
        ch = s[n]
        if q:
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return n
    raise SyntaxError(f"unbalanced parentheses in {s!r}")


# --------------------------------------------------------------------------------------------------------------------
# symbols
# --------------------------------------------------------------------------------------------------------------------
class Sym:
    def __init__(self, name, base, rank=0, bounds=None, **attrs):
        self.name, self.base, self.rank, self.bounds = name, base, rank, bounds      # base: int | real | char
        self.pointer = attrs.get("pointer", False)
        self.intent = attrs.get("intent")
        self.parameter = attrs.get("parameter")       # initialiser text or None
        self.save = attrs.get("save")
        self.optional = attrs.get("optional", False)
        self.isarg = False
        self.byref = False

    @property
    def deferred(self):
        return self.rank > 0 and (self.bounds is None or all(b == (None, None) for b in self.bounds))


def parse_types(lines: list[str]):
    """Derived-type definitions (`type name ... end type`) and module variables of those types
    (`type(name), dimension(:), pointer :: var`): ({type: {member: Sym}}, {var: type})."""
    types, dvars, cur = {}, {}, None
    for ln in lines:
        if ln == "contains":
            break
        m = re.match(r"type\s+([a-z_]\w*)$", ln)
        if m:
            cur = types.setdefault(m.group(1), {}); continue
        if re.match(r"end\s*type", ln):
            cur = None; continue
        if cur is not None:
            parse_decl(ln, cur); continue
        m = re.match(r"type\s*\(\s*([a-z_]\w*)\s*\).*::\s*([a-z_]\w*)$", ln)
        if m:
            dvars[m.group(2)] = m.group(1)
    return types, dvars


def parse_interfaces(lines: list[str]):
    """Generic interfaces: {generic name: [specific procedures]}."""
    out, cur = {}, None
    for ln in lines:
        m = re.match(r"interface\s+([a-z_]\w*)$", ln)
        if m:
            cur = out.setdefault(m.group(1), []); continue
        if re.match(r"end\s*interface", ln):
            cur = None; continue
        if cur is not None:
            m = re.match(r"module\s+procedure\s+(.*)$", ln)
            if m:
                cur.extend(x.strip() for x in m.group(1).split(","))
    return out


# module-level variables the translated routines read (src/mg_namelist.f90, mg_mpi.f90, mg_grids.f90): name -> base type
GLOBALS = {"myrank": "int", "nprocs": "int", "bmask": "int", "netcdf_output": "int", "aggressive": "int", "nlevs": "int", "ns_pre": "int",
           "ns_post": "int", "ns_coarsest": "int", "mpi_proc_null": "int", "nsmall": "int", "hlim": "real", "nhtheta_b": "real", "nhtheta_s": "real",
           "cmatrix": "char", "relax_method": "char", "interp_type": "char", "rp": "kind", "ip": "kind", "lg": "kind",
           "solver_prec": "real", "solver_maxiter": "int",
           # MPI constants (oracle/ref/ref_prelude.h defines them for the thread-emulated MPI of oracle/ref/ref_mpi.c)
           "mpi_comm_world": "int", "mpi_request_null": "int", "mpi_double_precision": "int", "mpi_sum": "int", "mpi_status_size": "int"}

# MPI procedures -> run-time functions of oracle/ref/ref_mpi.c; per Fortran argument: buf (address of an array / section),
# val (integer expression), iref / rref (address of an integer / real variable or element), skip
MPI_MAP = {
    "mpi_irecv": ("rmpi_irecv", ["buf", "val", "skip", "val", "val", "val", "iref", "skip"]),
    "mpi_isend": ("rmpi_isend", ["buf", "val", "skip", "val", "val", "val", "iref", "skip"]),
    "mpi_waitany": ("rmpi_waitany", ["val", "buf", "iref", "skip", "skip"]),
    "mpi_allreduce": ("rmpi_allreduce_sum", ["rref", "rref", "val", "skip", "skip", "val", "skip"]),
    "mpi_allgather": ("rmpi_allgather", ["buf", "val", "skip", "buf", "skip", "skip", "val", "skip"]),
    "mpi_comm_split": ("rmpi_comm_split", ["val", "val", "val", "iref", "skip"]),
    "mpi_barrier": ("rmpi_barrier", ["val", "skip"]),
}

DROP_CALLS = {"tic", "toc", "write_netcdf", "cpu_time", "print_grids", "read_nhnamelist", "mg_mpi_init",
              "mpi_comm_rank", "mpi_comm_size"}      # myrank / nprocs are set when the rank thread attaches
INTRINSIC_REAL = {"sqrt": "sqrt", "exp": "exp", "log": "log", "cosh": "cosh", "sinh": "sinh", "tanh": "tanh"}


def parse_decl(line: str, table: dict):
    """`type[, attrs] :: a, b(…), c = init` -> Sym entries."""
    m = re.match(r"(integer|real|logical|character)\s*(\([^)]*\))?\s*(.*)$", line)
    if not m:
        return False
    base = {"integer": "int", "real": "real", "logical": "int", "character": "char"}[m.group(1)]
    rest = m.group(3).strip()
    if "::" in rest:
        attr_s, names_s = rest.split("::", 1)
    else:
        attr_s, names_s = "", rest
    attrs, dim = {}, None
    for a in split_top(attr_s.lstrip(",")):
        a = a.strip()
        if not a:
            continue
        if a.startswith("dimension"):
            dim = a[a.index("(") + 1:match_paren(a, a.index("("))]
        elif a.startswith("intent"):
            attrs["intent"] = re.sub(r"\s", "", a[a.index("(") + 1:-1])
        elif a in ("pointer", "optional", "target", "save", "allocatable"):
            attrs[a] = True
        elif a == "parameter":
            attrs["parameter"] = True
        else:
            raise SyntaxError(f"attribute {a!r} in {line!r}")
    for item in split_top(names_s):
        init = None
        if "=>" in item:
            item, init = [x.strip() for x in item.split("=>", 1)]; init = None
        elif "=" in item:
            item, init = [x.strip() for x in item.split("=", 1)]
        mm = re.match(r"([a-z_]\w*)\s*(\((.*)\))?$", item)
        if not mm:
            raise SyntaxError(f"declarator {item!r} in {line!r}")
        name, d = mm.group(1), (mm.group(3) if mm.group(2) else dim)
        rank, bounds = 0, None
        if d is not None:
            bounds = []
            for b in split_top(d):
                if b.strip() == ":":
                    bounds.append((None, None))
                elif ":" in split_top(b, ":")[0] or len(split_top(b, ":")) == 2:
                    lo, hi = split_top(b, ":"); bounds.append((lo, hi))
                else:
                    bounds.append(("1", b))
            rank = len(bounds)
        s = Sym(name, base, rank, bounds, **attrs)
        if attrs.get("parameter"):
            s.parameter = init
        elif init is not None:
            s.save = init
        table[name] = s
    return True


# --------------------------------------------------------------------------------------------------------------------
# the translator proper
# --------------------------------------------------------------------------------------------------------------------
class Unit:
    def __init__(self, name, lines, modsyms):
        self.name, self.lines = name, lines
        self.syms: dict[str, Sym] = dict(modsyms)
        hdr = lines[0]
        m = re.match(rf"subroutine\s+{name}\s*(\((.*)\))?$", hdr)
        self.args = [a.strip() for a in m.group(2).split(",")] if m.group(2) and m.group(2).strip() else []
        self.body = []
        for ln in lines[1:-1]:
            if re.match(r"(use\b|implicit\b)", ln):
                continue
            if re.match(r"(integer|real|logical|character)\b", ln) and ("::" in ln or re.match(r"(integer|real|logical)\s+[a-z]", ln)) \
                    and not re.match(r"real\s*\(.*\)\s*=", ln):
                if parse_decl(ln, self.syms):
                    continue
            self.body.append(ln)
        for a in self.args:
            s = self.syms[a]
            s.isarg = True
            if s.rank == 0 and s.base in ("int", "real") and s.intent != "in":
                s.byref = True

    def c_params(self):
        out = []
        for a in self.args:
            s = self.syms[a]
            if s.base == "char":
                out.append(f"const char* {a}")
            elif s.rank == 0:
                t = "int" if s.base == "int" else "double"
                out.append(f"{t}* {a}" if s.byref else f"{t} {a}")
            elif s.deferred:
                out.append(f"{'iarr' if s.base == 'int' else 'farr'} {a}")
            else:
                out.append(f"double* {a}_a" if s.base == "real" else f"int* {a}_a")
        return ", ".join(out) if out else "void"


class Translator:
    def __init__(self, callbacks=None, unreachable=()):
        self.units: dict[str, Unit] = {}
        self.unreachable = set(unreachable)   # callees of branches the supported configurations never take (nz == 1, aggressive)
        self.callbacks = callbacks or {}      # fortran name -> (c name, [kinds], {keyword: position})
        self.tmp = 0
        self.types: dict[str, dict[str, Sym]] = {}     # derived types
        self.dvars: dict[str, str] = {}                # module variables of derived type: name -> type
        self.generics: dict[str, list[str]] = {}

    def member(self, node) -> Sym:
        """Sym of `var(idx)%member`."""
        base = node[1]
        if not (base[0] == "call" and base[1][0] == "name" and base[1][1] in self.dvars):
            raise KeyError(f"member access on {base}")
        return self.types[self.dvars[base[1][1]]][node[2]]

    def add(self, path, names):
        lines = logical_lines(path)
        t, d = parse_types(lines)
        self.types.update(t); self.dvars.update(d)
        for g, specs in parse_interfaces(lines).items():
            self.generics.setdefault(g, []).extend(specs)
        modsyms = {}
        for d in module_decls(lines):
            parse_decl(d, modsyms)
        modsyms = {k: v for k, v in modsyms.items() if v.parameter is not None}     # module parameters only (hlf, qrt, ...)
        for n in names:
            self.units[n] = Unit(n, extract_unit(lines, n), modsyms)

    # ---- types -------------------------------------------------------------------------------------------------
    def typeof(self, u: Unit, e) -> str:
        k = e[0]
        if k == "num":
            return "real" if re.search(r"[.ed]", e[1].split("_")[0]) else "int"
        if k == "str":
            return "char"
        if k == "par":
            return self.typeof(u, e[1])
        if k == "un":
            return "int" if e[1] == "!" else self.typeof(u, e[2])
        if k == "pow":
            return self.typeof(u, e[1])
        if k == "bin":
            if e[1] in ("&&", "||", "==", "!=", "<", "<=", ">", ">="):
                return "int"
            a, b = self.typeof(u, e[2]), self.typeof(u, e[3])
            return "real" if "real" in (a, b) else "int"
        if k == "name":
            if e[1] in u.syms:
                return u.syms[e[1]].base
            if e[1] in GLOBALS:
                return GLOBALS[e[1]]
            raise KeyError(f"{u.name}: unknown name {e[1]}")
        if k == "mem":
            return self.member(e).base
        if k == "call":
            f = e[1]
            if f[0] == "mem":
                return self.member(f).base
            n = f[1]
            if n in u.syms:
                return u.syms[n].base
            if n in ("real",) or n in INTRINSIC_REAL:
                return "real"
            if n in ("int", "size", "lbound", "ubound", "present", "associated", "floor"):
                return "int"
            if n == "sum":
                return "real"
            if n == "trim":
                return "char"
            if n in ("mod", "max", "min", "abs"):
                ts = [self.typeof(u, a) for a in e[2] if a[0] != "kw"]
                return "real" if "real" in ts else "int"
            raise KeyError(f"{u.name}: unknown function {n}")
        raise KeyError(e)

    # ---- expressions -------------------------------------------------------------------------------------------
    def num(self, text: str) -> str:
        body, _, kind = text.partition("_")
        if not re.search(r"[.ed]", body):
            return body
        isdouble = bool(kind) or "d" in body
        body = body.replace("d", "e")
        if re.match(r"^\d+$", body.split("e")[0]):
            body = body.replace("e", ".e", 1) if "e" in body else body + "."
        return body if isdouble else f"((double){body}f)"

    def array_desc(self, u: Unit, node):
        """C expression of the descriptor of an array-valued node and its rank; None if the node is not an array."""
        if node[0] == "name" and node[1] in u.syms and u.syms[node[1]].rank > 0:
            return node[1], u.syms[node[1]].rank, u.syms[node[1]].base
        if node[0] == "mem" and self.member(node).rank > 0:
            m = self.member(node)
            return f"{self.ex(u, node[1])}.{node[2]}", m.rank, m.base
        return None

    def ex(self, u: Unit, e, secmap=None) -> str:
        """secmap: list of loop-index names for section dimensions (array-section assignment in progress)."""
        k = e[0]
        if k == "num":
            return self.num(e[1])
        if k == "str":
            return '"' + e[1] + '"'
        if k == "par":
            return "(" + self.ex(u, e[1], secmap) + ")"
        if k == "un":
            return f"!({self.ex(u, e[2], secmap)})" if e[1] == "!" else e[1] + self.ex(u, e[2], secmap)
        if k == "pow":
            base = self.ex(u, e[1], secmap)
            if e[2] == ("num", "2"):
                t = "SQI" if self.typeof(u, e[1]) == "int" else "SQD"
                return f"{t}({base})"
            return f"pow({base}, {self.ex(u, e[2], secmap)})"
        if k == "bin":
            a, b = e[2], e[3]
            if e[1] in ("==", "!=") and "char" in (self.typeof(u, a), self.typeof(u, b)):
                c = f"fstr_eq({self.ex(u, a, secmap)}, {self.ex(u, b, secmap)})"
                return c if e[1] == "==" else "!" + c
            return f"{self.ex(u, a, secmap)} {e[1]} {self.ex(u, b, secmap)}"
        if k == "name":
            n = e[1]
            if n in u.syms:
                s = u.syms[n]
                if s.rank > 0:
                    if secmap is not None:      # whole array inside an array expression == all-colon section
                        return self.element(u, e, [("sec", None, None, None)] * s.rank, secmap)
                    return n
                return f"(*{n})" if s.byref else n
            if n in GLOBALS:
                return "g_" + n
            raise KeyError(f"{u.name}: unknown name {n}")
        if k == "mem":
            base = self.ex(u, e[1])
            if self.member(e).rank > 0 and secmap is not None:
                return self.element(u, e, [("sec", None, None, None)] * self.member(e).rank, secmap)
            return f"{base}.{e[2]}"
        if k == "call":
            f, args = e[1], e[2]
            if f[0] == "name" and f[1] in self.dvars and f[1] not in u.syms:
                assert len(args) == 1
                return f"{f[1].upper()}_[{self.ex(u, args[0])}]"
            d = self.array_desc(u, f)
            if d:
                if f[0] == "mem" and not self.member(f).pointer:        # fixed-size member array (neighb(8)): a plain C array
                    return f"{self.ex(u, f[1])}.{f[2]}[{self.ex(u, args[0])}]"
                return self.element(u, f, args, secmap)
            n = f[1]
            pos = [a for a in args if a[0] != "kw"]
            kw = {a[1]: a[2] for a in args if a[0] == "kw"}
            if n in INTRINSIC_REAL:
                return f"{INTRINSIC_REAL[n]}({self.ex(u, pos[0], secmap)})"
            if n == "real":
                return f"((double)({self.ex(u, pos[0], secmap)}))"
            if n == "int":
                return f"((int)({self.ex(u, pos[0], secmap)}))"
            if n == "floor":
                return f"((int)floor({self.ex(u, pos[0], secmap)}))"
            if n == "sum":      # sequential accumulation in array-element order (what an inlined SUM does without fast-math)
                secs = self.find_sections(u, pos[0])
                assert secs, (u.name, e)
                self.tmp += 1
                t = self.tmp
                names = [f"s{t}_{q}" for q in range(len(secs))]
                code = "({ double acc_ = 0.; "
                for q, (desc, dim, a) in enumerate(secs):
                    lo = self.ex(u, a[1]) if a[1] is not None else f"{desc}.lb[{dim}]"
                    hi = self.ex(u, a[2]) if a[2] is not None else f"({desc}.lb[{dim}] + {desc}.n[{dim}] - 1)"
                    st = self.ex(u, a[3]) if a[3] is not None else "1"
                    code += f"const int n{t}_{q} = SECCOUNT({lo}, {hi}, {st}); "
                for q in reversed(range(len(secs))):
                    code += f"for (int {names[q]} = 0; {names[q]} < n{t}_{q}; ++{names[q]}) "
                code += f"acc_ = acc_ + {self.ex(u, pos[0], names)}; acc_; }})"
                return code
            if n == "mod":
                if self.typeof(u, e) == "int":
                    return f"(({self.ex(u, pos[0], secmap)}) % ({self.ex(u, pos[1], secmap)}))"
                return f"fmod({self.ex(u, pos[0], secmap)}, {self.ex(u, pos[1], secmap)})"
            if n in ("max", "min"):
                t = "I" if self.typeof(u, e) == "int" else "D"
                out = self.ex(u, pos[0], secmap)
                for a in pos[1:]:
                    out = f"{n.upper()}{t}({out}, {self.ex(u, a, secmap)})"
                return out
            if n == "abs":
                return ("abs" if self.typeof(u, e) == "int" else "fabs") + f"({self.ex(u, pos[0], secmap)})"
            if n in ("size", "lbound", "ubound"):
                node = pos[0]
                if node[0] == "call" and self.array_desc(u, node[1]) and all(a == ("sec", None, None, None) for a in node[2]):
                    node = node[1]             # size(a(:,:,:),dim=1)
                d = self.array_desc(u, node)
                dim = kw.get("dim") or pos[1]
                idx = f"({self.ex(u, dim)}) - 1"
                if n == "size":
                    return f"{d[0]}.n[{idx}]"
                if n == "lbound":
                    return f"{d[0]}.lb[{idx}]"
                return f"({d[0]}.lb[{idx}] + {d[0]}.n[{idx}] - 1)"
            if n == "trim":
                return self.ex(u, pos[0])
            if n == "present":
                sy = u.syms[pos[0][1]]
                return f"({pos[0][1]}_a != 0)" if (sy.rank > 0 and not sy.deferred) else (f"({pos[0][1]}.a != 0)" if sy.rank > 0 else f"({pos[0][1]} != 0)")
            if n == "associated":
                return f"({self.array_desc(u, pos[0])[0]}.a != 0)"
            raise KeyError(f"{u.name}: unknown function {n}")
        raise KeyError(e)

    def element(self, u: Unit, arr, args, secmap) -> str:
        desc, rank, _ = self.array_desc(u, arr)
        assert len(args) == rank, (u.name, arr, args)
        idx, used = [], 0
        for d, a in enumerate(args):
            if a[0] == "sec":
                if secmap is None:
                    raise SyntaxError(f"{u.name}: array section outside an array assignment: {arr}")
                lo = self.ex(u, a[1]) if a[1] is not None else f"{desc}.lb[{d}]"
                st = self.ex(u, a[3]) if a[3] is not None else "1"
                idx.append(f"({lo}) + {secmap[used]} * ({st})")
                used += 1
            else:
                idx.append(self.ex(u, a, secmap))
        return f"FA{rank}({desc}, {', '.join(idx)})"

    # ---- statements --------------------------------------------------------------------------------------------
    def section_dims(self, u: Unit, node):
        """[(desc, dim, sec)] of the section dimensions of the first sectioned array reference, or [] (scalar)."""
        if node[0] == "call":
            d = self.array_desc(u, node[1])
            if d and any(a[0] == "sec" for a in node[2]):
                return [(d[0], n, a) for n, a in enumerate(node[2]) if a[0] == "sec"]
        d = self.array_desc(u, node)
        if d and node[0] in ("name", "mem"):
            return [(d[0], n, ("sec", None, None, None)) for n in range(d[1])]
        return []

    def find_sections(self, u: Unit, node):
        d = self.section_dims(u, node)
        if d:
            return d
        if node[0] in ("par", "un"):
            return self.find_sections(u, node[-1])
        if node[0] == "pow":
            return self.find_sections(u, node[1]) or self.find_sections(u, node[2])
        if node[0] == "bin":
            return self.find_sections(u, node[2]) or self.find_sections(u, node[3])
        if node[0] == "call" and not self.array_desc(u, node[1]):
            for a in node[2]:
                r = self.find_sections(u, a[2] if a[0] == "kw" else a)
                if r:
                    return r
        return []

    def assign(self, u: Unit, lhs, rhs, ind) -> list[str]:
        secs = self.section_dims(u, lhs)
        if not secs:
            return [f"{ind}{self.ex(u, lhs)} = {self.ex(u, rhs)};"]
        self.tmp += 1
        t = self.tmp
        names = [f"s{t}_{n}" for n in range(len(secs))]
        out = [ind + "{"]
        for n, (desc, dim, a) in enumerate(secs):
            lo = self.ex(u, a[1]) if a[1] is not None else f"{desc}.lb[{dim}]"
            hi = self.ex(u, a[2]) if a[2] is not None else f"({desc}.lb[{dim}] + {desc}.n[{dim}] - 1)"
            st = self.ex(u, a[3]) if a[3] is not None else "1"
            out.append(f"{ind}  const int n{t}_{n} = SECCOUNT({lo}, {hi}, {st});")
        for n in reversed(range(len(secs))):
            out.append(f"{ind}  for (int {names[n]} = 0; {names[n]} < n{t}_{n}; ++{names[n]})")
        if lhs[0] == "call":
            l = self.element(u, lhs[1], lhs[2], names)
        else:
            l = self.ex(u, lhs, names)
        out.append(f"{ind}    {l} = {self.ex(u, rhs, names)};")
        out.append(ind + "}")
        return out

    def call_stmt(self, u: Unit, text: str, ind: str) -> list[str]:
        e = parse_expr(text)
        if e[0] == "name":
            e = ("call", e, [])
        name, args = e[1][1], e[2]
        if name in DROP_CALLS:
            return [f"{ind}/* call {name}(...) dropped */"]
        if name in self.unreachable:
            return [f"{ind}abort(); /* {name}: branch outside the supported configurations (2-D / aggressive coarsening) */"]
        pos = [a for a in args if a[0] != "kw"]
        kws = {a[1]: a[2] for a in args if a[0] == "kw"}
        if name in MPI_MAP:
            cname, spec = MPI_MAP[name]
            assert len(pos) == len(spec) and not kws, (u.name, name, len(pos))
            cargs = []
            for kind, a in zip(spec, pos):
                if kind == "skip":
                    continue
                if kind == "val":
                    cargs.append(self.ex(u, a))
                elif kind == "buf":
                    cargs.append(self.array_pointer(u, a))
                else:
                    cargs.append(self.address(u, a))
            return [f"{ind}{cname}({', '.join(cargs)});"]
        if name.startswith("mpi_"):
            raise KeyError(f"{u.name}: MPI procedure {name} has no run-time mapping")
        if name not in self.units and name in self.generics:
            name = self.resolve_generic(u, name, pos, kws)
        if name in self.unreachable:
            return [f"{ind}abort(); /* {name}: branch outside the supported configurations (2-D / aggressive coarsening) */"]
        if name in self.units:
            cal = self.units[name]
            actual = list(pos) + [None] * (len(cal.args) - len(pos))
            for k_, v in kws.items():
                actual[cal.args.index(k_)] = v
            cargs = []
            for formal, a in zip(cal.args, actual):
                fs = cal.syms[formal]
                cargs.append(self.actual(u, a, fs))
            return [f"{ind}ref_{name}({', '.join(cargs)});"]
        # call-backs: resolved by name and number of positional arguments
        key = (name, len(pos))
        if key in self.callbacks:
            cname, kinds, kwpos = self.callbacks[key]
            actual = list(pos) + [None] * (len(kinds) - len(pos))
            for k_, v in kws.items():
                actual[kwpos[k_]] = v
            cargs = []
            for kind, a in zip(kinds, actual):
                if a is None:
                    cargs.append("0")
                elif kind == "int" or kind == "real":
                    cargs.append(self.ex(u, a))
                elif kind == "str":
                    cargs.append(self.ex(u, a))
                elif kind == "ptr":
                    cargs.append(self.array_pointer(u, a))
                elif kind == "farr":
                    cargs.append(self.array_desc(u, a)[0])
                elif kind == "realref":
                    s = u.syms[a[1]]
                    cargs.append(a[1] if s.byref else "&" + a[1])
            return [f"{ind}{cname}({', '.join(cargs)});"]
        raise KeyError(f"{u.name}: call to unknown subroutine {name} with {len(pos)} positional arguments")

    def resolve_generic(self, u: Unit, name, pos, kws) -> str:
        """The specific procedure of a generic interface: by number of arguments and by the rank of array arguments."""
        for cand in self.generics[name]:
            if cand in self.unreachable:
                continue
            if cand not in self.units:
                continue
            cal = self.units[cand]
            if len(pos) > len(cal.args) or any(k not in cal.args for k in kws):
                continue
            given = set(cal.args[:len(pos)]) | set(kws)
            if any(not cal.syms[f].optional and f not in given for f in cal.args):
                continue
            ok = True
            for f, a in zip(cal.args, pos):
                d = self.array_desc(u, a)
                fr = cal.syms[f].rank
                if d and d[1] != fr:
                    ok = False
                if not d and fr > 0 and a[0] != "call":
                    ok = False
            if ok:
                return cand
        raise KeyError(f"{u.name}: no specific procedure of generic {name} matches {len(pos)} positional arguments")

    def address(self, u: Unit, a) -> str:
        """Address of a scalar variable or of an array element (actual argument of a by-reference dummy)."""
        if a[0] == "name" and a[1] in u.syms and u.syms[a[1]].rank == 0:
            return a[1] if u.syms[a[1]].byref else "&" + a[1]
        if a[0] == "call" and self.array_desc(u, a[1]) and not any(x[0] == "sec" for x in a[2]):
            return "&" + self.ex(u, a)
        if a[0] == "mem" and self.member(a).rank == 0:
            return "&" + self.ex(u, a)
        raise SyntaxError(f"{u.name}: cannot take the address of {a}")

    def array_pointer(self, u: Unit, a) -> str:
        d = self.array_desc(u, a)
        if d:
            return f"{d[0]}.a"
        if a[0] == "call":      # element or column section: address of its first element
            dd = self.array_desc(u, a[1])
            args = []
            for n, x in enumerate(a[2]):
                if x[0] == "sec":
                    args.append(self.ex(u, x[1]) if x[1] is not None else f"{dd[0]}.lb[{n}]")
                else:
                    args.append(self.ex(u, x))
            return f"&FA{dd[1]}({dd[0]}, {', '.join(args)})"
        raise SyntaxError(f"{u.name}: cannot take the address of {a}")

    def actual(self, u: Unit, a, fs: Sym) -> str:
        if a is None:
            return "0"
        if fs.base == "char":
            return self.ex(u, a)
        if fs.rank == 0:
            if not fs.byref:
                return self.ex(u, a)
            return self.address(u, a)
        if fs.deferred:
            d = self.array_desc(u, a)
            if not d:
                raise SyntaxError(f"{u.name}: pointer dummy needs an array actual, got {a}")
            return d[0]
        return self.array_pointer(u, a)

    def bounds_args(self, u: Unit, bounds) -> str:
        return ", ".join(f"{self.ex(u, parse_expr(lo))}, {self.ex(u, parse_expr(hi))}" for lo, hi in bounds)

    def translate_unit(self, u: Unit) -> str:
        out = [f"void ref_{u.name}({u.c_params()}) {{"]
        frees = []
        # declarations
        for n, s in u.syms.items():
            if s.parameter is not None:
                t = "int" if s.base == "int" else "double"
                if s.base == "kind":
                    continue
                out.append(f"  const {t} {n} = {self.ex(u, parse_expr(s.parameter))};")
        for n, s in u.syms.items():
            if s.parameter is not None:
                continue
            if s.isarg:
                if s.rank > 0 and not s.deferred:
                    out.append(f"  {'iarr' if s.base == 'int' else 'farr'} {n} = {'imk' if s.base == 'int' else 'mk'}{s.rank}({n}_a, {self.bounds_args(u, s.bounds)});")
                continue
            if s.base == "char":
                out.append(f"  const char* {n} = \"\";")
            elif s.rank == 0:
                t = "int" if s.base == "int" else "double"
                if s.save is not None:
                    out.append(f"  static {t} {n} = {self.ex(u, parse_expr(s.save))};")
                else:
                    out.append(f"  {t} {n} = 0;")
            elif s.deferred:
                out.append(f"  {'iarr' if s.base == 'int' else 'farr'} {n} = {{0}};")
            elif s.base == "real":
                out.append(f"  farr {n} = fa_alloc{s.rank}({self.bounds_args(u, s.bounds)});")
                frees.append(n)
            else:
                out.append(f"  iarr {n} = ia_alloc{s.rank}({self.bounds_args(u, s.bounds)});")
                frees.append(n)
        for n, s in u.syms.items():
            if not s.isarg and s.parameter is None and s.rank == 0 and s.base != "char":
                out.append(f"  (void){n};")
        depth = 1
        sel = []
        for ln in u.body:
            ind = "  " * depth
            # labels / format / io
            if re.match(r"\d+\s+format", ln):
                out.append(f"{ind}/* format dropped */"); continue
            m = re.match(r"select\s*case\s*\((.*)\)$", ln)
            if m:
                sel.append([parse_expr(m.group(1)), True]); continue
            m = re.match(r"case\s*\((.*)\)$", ln)
            if m or ln == "case default":
                if m:
                    cond = " || ".join(self.ex(u, ("bin", "==", sel[-1][0], parse_expr(v))) for v in split_top(m.group(1)))
                if sel[-1][1]:
                    out.append(f"{ind}if ({cond}) {{" if m else f"{ind}{{"); depth += 1; sel[-1][1] = False
                else:
                    out.append(f"{'  ' * (depth - 1)}}} else if ({cond}) {{" if m else f"{'  ' * (depth - 1)}}} else {{")
                continue
            if re.match(r"end\s*select", ln):
                if not sel.pop()[1]:
                    depth -= 1; out.append(f"{'  ' * depth}}}")
                continue
            m = re.match(r"if\s*\(", ln)
            if m:
                close = match_paren(ln, ln.index("("))
                cond, rest = ln[ln.index("(") + 1:close], ln[close + 1:].strip()
                c = self.ex(u, parse_expr(cond))
                if rest == "then":
                    out.append(f"{ind}if ({c}) {{"); depth += 1
                else:
                    out.append(f"{ind}if ({c}) {{")
                    out.extend(self.simple(u, rest, ind + "  "))
                    out.append(f"{ind}}}")
                continue
            m = re.match(r"else\s*if\s*\(", ln)
            if m:
                close = match_paren(ln, ln.index("("))
                c = self.ex(u, parse_expr(ln[ln.index("(") + 1:close]))
                out.append(f"{'  ' * (depth - 1)}}} else if ({c}) {{"); continue
            if ln == "else":
                out.append(f"{'  ' * (depth - 1)}}} else {{"); continue
            if re.match(r"end\s*(if|do)\b", ln):
                depth -= 1; out.append(f"{'  ' * depth}}}"); continue
            m = re.match(r"do\s+while\s*\(", ln)
            if m:
                close = match_paren(ln, ln.index("("))
                out.append(f"{ind}while ({self.ex(u, parse_expr(ln[ln.index('(') + 1:close]))}) {{"); depth += 1; continue
            m = re.match(r"do\s+([a-z_]\w*)\s*=\s*(.*)$", ln)
            if m:
                v, parts = m.group(1), split_top(m.group(2))
                lo, hi = (self.ex(u, parse_expr(x)) for x in parts[:2])
                self.tmp += 1
                e_ = f"e{self.tmp}"
                if len(parts) == 2 or re.match(r"^\d+$", parts[2]):
                    st = parts[2] if len(parts) > 2 else "1"
                    out.append(f"{ind}{{ const int {e_} = {hi}; for ({v} = {lo}; {v} <= {e_}; {v} += {st}) {{")
                elif re.match(r"^-\s*\d+$", parts[2]):
                    out.append(f"{ind}{{ const int {e_} = {hi}; for ({v} = {lo}; {v} >= {e_}; {v} -= {parts[2].lstrip('-').strip()}) {{")
                else:
                    st = self.ex(u, parse_expr(parts[2]))
                    out.append(f"{ind}{{ const int {e_} = {hi}, t{self.tmp} = {st}; for ({v} = {lo}; t{self.tmp} > 0 ? {v} <= {e_} : {v} >= {e_}; {v} += t{self.tmp}) {{")
                out.append(f"{ind}/*do*/"); depth += 1; continue
            out.extend(self.simple(u, ln, ind))
        # close the double brace of do loops: every `do` opened "{ for (...) {" -- patch the matching closers
        text = self.fix_do_braces(out)
        for n in frees:
            text.append(f"  {'ia' if u.syms[n].base == 'int' else 'fa'}_free({n});")
        text.append("}")
        return "\n".join(text)

    @staticmethod
    def fix_do_braces(lines: list[str]) -> list[str]:
        """`do` emits two opening braces (scope of the end value + loop body); its closing `}` must be doubled."""
        stack, out = [], []
        for ln in lines:
            s = ln.strip()
            if s == "/*do*/":
                stack[-1] = "do"; continue
            if s.endswith("{") and not s.startswith("}"):
                stack.append("x")
                out.append(ln); continue
            if s.startswith("}") and s.endswith("{"):
                out.append(ln); continue
            if s == "}":
                kind = stack.pop()
                out.append(ln + (" }" if kind == "do" else "")); continue
            out.append(ln)
        return out

    def simple(self, u: Unit, ln: str, ind: str) -> list[str]:
        if ln in ("return",):
            raise SyntaxError(f"{u.name}: `return` is not supported")
        m = re.match(r"write\s*\(\s*100\s*,\s*\*\s*\)(.*)$", ln)      # the residual history (unit 100, src/mg_solvers.f90:59,72)
        if m:
            a, b = (self.ex(u, parse_expr(x)) for x in split_top(m.group(1)))
            return [f"{ind}cb_history((double)({a}), (double)({b}));"]
        if re.match(r"(write|print)\b", ln):
            return [f"{ind}/* i/o dropped */"]
        if ln == "stop" or ln.startswith("stop "):
            return [f"{ind}abort();"]
        if ln.startswith("call "):
            return self.call_stmt(u, ln[5:].strip(), ind)
        m = re.match(r"(de)?allocate\s*\((.*)\)$", ln)
        if m:
            out = []
            for item in split_top(m.group(2)):
                e = parse_expr(item)
                if m.group(1):
                    if e[0] == "name" and e[1] in self.dvars:
                        out.append(f"{ind}/* deallocate({item}): static storage */"); continue
                    d = self.array_desc(u, e)
                    out.append(f"{ind}fa_free({d[0]}); {d[0]}.a = 0;")
                    continue
                target, dims = e[1], e[2]
                if target[0] == "name" and target[1] in self.dvars:
                    out.append(f"{ind}/* allocate({item}): {target[1].upper()}_ is static storage */"); continue
                d = self.array_desc(u, target)
                b = ", ".join((f"{self.ex(u, x[1])}, {self.ex(u, x[2])}" if x[0] == "sec" else f"1, {self.ex(u, x)}") for x in dims)
                out.append(f"{ind}{d[0]} = {'ia' if d[2] == 'int' else 'fa'}_alloc{len(dims)}({b});")
            return out
        toks = tokenize(ln)
        depth = 0
        for n, (k, v) in enumerate(toks):
            if (k, v) == ("op", "("):
                depth += 1
            elif (k, v) == ("op", ")"):
                depth -= 1
            elif depth == 0 and k == "op" and v in ("=", "=>"):
                lhs = P(toks[:n]).expr()
                if v == "=>":
                    r = toks[n + 1:]
                    if [t[1] for t in r] == ["null", "(", ")"]:
                        return [f"{ind}{self.array_desc(u, lhs)[0]} = (farr){{0}};"]
                    rhs = P(r).expr()
                    return [f"{ind}{self.array_desc(u, lhs)[0]} = {self.array_desc(u, rhs)[0]};"]
                rhs = P(toks[n + 1:]).expr()
                if self.typeof(u, lhs) == "char":
                    return [f"{ind}{self.ex(u, lhs)} = {self.ex(u, rhs)};"]
                return self.assign(u, lhs, rhs, ind)
        raise SyntaxError(f"{u.name}: cannot translate statement {ln!r}")

    def emit_types(self) -> str:
        """C struct of every derived type + the (thread-local) module arrays of those types; generated, not committed."""
        out = ["/* generated from the reference's derived types by oracle/f90toc.py */"]
        for tname, members in self.types.items():
            out.append("typedef struct {")
            for n, m in members.items():
                if m.rank == 0:
                    out.append(f"  {'int' if m.base == 'int' else 'double'} {n};")
                elif m.pointer:
                    out.append(f"  {'iarr' if m.base == 'int' else 'farr'} {n};")
                else:                     # fixed-size array member, 1-based
                    out.append(f"  {'int' if m.base == 'int' else 'double'} {n}[{m.bounds[0][1]} + 1];")
            out.append(f"}} {tname};")
        for v, t in self.dvars.items():
            out.append(f"extern __thread {t} {v.upper()}_[REF_MAXLEV];")
        return "\n".join(out)

    def layout(self) -> dict:
        """Member lists for the ctypes mirror in oracle/refdriver.py."""
        lay = {}
        for tname, members in self.types.items():
            lay[tname] = [[n, ("int" if m.base == "int" else "double") if m.rank == 0 else (("iarr" if m.base == "int" else "farr") if m.pointer else f"int[{int(m.bounds[0][1]) + 1}]")]
                          for n, m in members.items()]
        return {"types": lay, "vars": dict(self.dvars)}

    def emit(self) -> str:
        out = ['#include "ref_prelude.h"', '#include "ref_types.h"', ""]
        for u in self.units.values():
            out.append(f"void ref_{u.name}({u.c_params()});")
        out.append("")
        for u in self.units.values():
            out.append(self.translate_unit(u)); out.append("")
        return "\n".join(out)


if __name__ == "__main__":
    print("use oracle/build_ref.py", file=sys.stderr)
