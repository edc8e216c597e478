#!/usr/bin/env python3
"""f90_to_c.py -- mechanical Fortran-90-subset -> C translator.  TEST INFRASTRUCTURE (oracle/).

Purpose: tie the parity oracle to the reference's SOURCE TEXT.  No Fortran compiler exists in the
build image, so the reference (`/root/reference/subs.f90`, `/root/reference/set3d.f90`) cannot be
compiled as it is.  This script reads those two files as text and emits C that performs the same
operations in the same order (`oracle/_ref/ref_subs.c`, `oracle/_ref/ref_set3d.c`, never committed:
`oracle/_ref/` is git-ignored); `oracle/Makefile` compiles the result with
`gcc -O2 -ffp-contract=off -fno-fast-math -fwrapv` into `oracle/_ref/libref.so`.

What "the same operations" means (the arithmetic contract of `gfortran -O3 -fdefault-real-8` on
x86-64, reference Makefile:4,8):
  * default REAL and every real literal are IEEE binary64; REAL*4 is binary32; INTEGER is int32
    (wrapping on overflow, -fwrapv); INTEGER*2 is int16;
  * expressions are parsed with Fortran's precedence/associativity and emitted FULLY PARENTHESISED,
    so the C compiler evaluates them in exactly the Fortran order; no FMA contraction, no
    re-association, IEEE division and sqrt;
  * `x**2` / `x**3` with a literal exponent become repeated multiplication (what gfortran emits);
  * MAX/MIN follow gfortran's expansion (ref_runtime.h: f_max/f_min);
  * mixed INTEGER/REAL arithmetic and assignment conversions are the same in C and Fortran
    (integer division truncates toward zero, real->integer assignment truncates);
  * DO loops evaluate their bounds once; EXIT leaves the innermost DO; arguments are passed by
    reference (expressions through a temporary); INTENT(IN) scalars are read once at entry, which
    the Fortran aliasing rules permit;
  * arrays keep Fortran layout (first index fastest) and their declared lower bounds;
  * local variables start at zero (undefined in Fortran); unreferenced automatic arrays are not
    allocated (minMax's unused gradPhi, subs.f90:422).

The main program is emitted as ONE function whose top-level statements are each guarded by the
source line they start on, so a harness can execute any line range of set3d.f90 against program
variables it has set up (`ref_program_exec(first_line, last_line)`, `ref_var`), e.g. the inline
sign search (:196-268) or the inline min/max loop (:394-462) -- the code executed is still the
translated text of those lines, nothing restated by hand.

Supported subset = what the two files use; anything else raises TranslateError (the build fails
rather than silently dropping a statement).  Usage:
    python oracle/f90_to_c.py /root/reference oracle/_ref
"""
from __future__ import annotations

import os
import re
import sys


class TranslateError(Exception):
    pass


# ======================================================================================= source
def _strip_comment(line: str) -> str:
    q = None
    for i, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return line[:i]
    return line


def _lower_outside_strings(s: str) -> str:
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        else:
            if ch in "'\"":
                q = ch
                out.append(ch)
            else:
                out.append(ch.lower())
    return "".join(out)


def _split_semicolons(s: str):
    parts, cur, q = [], [], None
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == ";":
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur))
    return [p.strip() for p in parts if p.strip()]


def read_statements(path: str):
    """-> list of (line_number_of_first_line, lower-cased statement text)"""
    stmts = []
    with open(path, "r", errors="replace") as f:
        lines = f.read().split("\n")
    i = 0
    while i < len(lines):
        start = i + 1
        text = _strip_comment(lines[i]).rstrip()
        i += 1
        while text.endswith("&"):
            text = text[:-1]
            while i < len(lines) and not _strip_comment(lines[i]).strip():
                i += 1
            nxt = _strip_comment(lines[i]).strip()
            i += 1
            if nxt.startswith("&"):
                nxt = nxt[1:]
            text = text + " " + nxt
            text = text.rstrip()
        text = text.strip()
        if not text:
            continue
        for part in _split_semicolons(text):
            stmts.append((start, _lower_outside_strings(part)))
    return stmts


# ======================================================================================= lexer
_DOTOPS = ("and", "or", "not", "eq", "ne", "lt", "le", "gt", "ge", "true", "false")
_TOKEN_RE = re.compile(
    r"""\s*(?:
      (?P<dotop>\.(?:and|or|not|eq|ne|lt|le|gt|ge|true|false)\.)
    | (?P<real>(?:\d+\.(?![a-z]+\.)\d*(?:[ed][+-]?\d+)?)|(?:\.\d+(?:[ed][+-]?\d+)?)|(?:\d+[ed][+-]?\d+))
    | (?P<int>\d+)
    | (?P<name>[a-z_][a-z0-9_]*)
    | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
    | (?P<op>\(/|/\)|\*\*|//|==|/=|<=|>=|::|[-+*/(),:=<>%])
    )""",
    re.X,
)


def tokenize(s: str):
    toks, pos = [], 0
    s = s.rstrip()
    while pos < len(s):
        m = _TOKEN_RE.match(s, pos)
        if not m or m.end() == pos:
            raise TranslateError(f"cannot tokenize at: {s[pos:pos+30]!r} in {s!r}")
        pos = m.end()
        kind = m.lastgroup
        val = m.group(kind)
        if kind == "str":
            q = val[0]
            val = val[1:-1].replace(q + q, q)
        toks.append((kind, val))
    return toks


# ======================================================================================= parser
class Parser:
    def __init__(self, toks):
        self.t = toks
        self.p = 0

    def peek(self, k=0):
        return self.t[self.p + k] if self.p + k < len(self.t) else ("eof", None)

    def next(self):
        tok = self.peek()
        self.p += 1
        return tok

    def at_op(self, v):
        k, x = self.peek()
        return k == "op" and x == v

    def expect_op(self, v):
        k, x = self.next()
        if k != "op" or x != v:
            raise TranslateError(f"expected {v!r}, got {x!r} (tokens {self.t})")

    def done(self):
        return self.p >= len(self.t)

    # precedence climbing, Fortran 90 rules
    def expr(self):
        a = self.and_expr()
        while self.peek() == ("dotop", ".or."):
            self.next()
            a = ("bin", "||", a, self.and_expr())
        return a

    def and_expr(self):
        a = self.not_expr()
        while self.peek() == ("dotop", ".and."):
            self.next()
            a = ("bin", "&&", a, self.not_expr())
        return a

    def not_expr(self):
        if self.peek() == ("dotop", ".not."):
            self.next()
            return ("un", "!", self.rel_expr())
        return self.rel_expr()

    _REL = {"==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">=",
            ".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}

    def rel_expr(self):
        a = self.concat_expr()
        k, v = self.peek()
        if (k in ("op", "dotop")) and v in self._REL:
            self.next()
            return ("rel", self._REL[v], a, self.concat_expr())
        return a

    def concat_expr(self):
        a = self.add_expr()
        while self.at_op("//"):
            self.next()
            a = ("concat", a, self.add_expr())
        return a

    def add_expr(self):
        if self.at_op("-"):
            self.next()
            a = ("un", "-", self.mul_expr())
        elif self.at_op("+"):
            self.next()
            a = self.mul_expr()
        else:
            a = self.mul_expr()
        while self.at_op("+") or self.at_op("-"):
            op = self.next()[1]
            a = ("bin", op, a, self.mul_expr())
        return a

    def mul_expr(self):
        a = self.pow_expr()
        while self.at_op("*") or self.at_op("/"):
            op = self.next()[1]
            a = ("bin", op, a, self.pow_expr())
        return a

    def pow_expr(self):
        a = self.primary()
        if self.at_op("**"):
            self.next()
            b = self.pow_expr()          # right associative
            return ("pow", a, b)
        return a

    def primary(self):
        k, v = self.next()
        if k == "int":
            return ("num_i", v)
        if k == "real":
            return ("num_r", v)
        if k == "str":
            return ("str", v)
        if k == "dotop" and v in (".true.", ".false."):
            return ("num_i", "1" if v == ".true." else "0")
        if k == "name":
            if self.at_op("("):
                self.next()
                args = self.arg_list(")")
                return ("call", v, args)
            return ("name", v)
        if k == "op" and v == "(":
            a = self.expr()
            self.expect_op(")")
            return ("paren", a)
        if k == "op" and v == "(/":
            items = self.arg_list("/)")
            return ("arrcons", items)
        raise TranslateError(f"unexpected token {v!r} in expression (tokens {self.t})")

    def arg_list(self, closer):
        args = []
        if self.at_op(closer):
            self.next()
            return args
        while True:
            args.append(self.arg())
            if self.at_op(","):
                self.next()
                continue
            self.expect_op(closer)
            return args

    def arg(self):
        # keyword=expr | [expr]:[expr] | expr
        if self.peek()[0] == "name" and self.peek(1) == ("op", "="):
            kw = self.next()[1]
            self.next()
            return ("kw", kw, self.expr())
        lo = None
        if not self.at_op(":"):
            lo = self.expr()
            if not self.at_op(":"):
                return lo
        self.next()  # ':'
        hi = None
        if not (self.at_op(",") or self.at_op(")")):
            hi = self.expr()
        return ("range", lo, hi)


# ======================================================================================= symbols
class Sym:
    def __init__(self, name, base, kind):
        self.name, self.base, self.kind = name, base, kind   # base: int|real|char ; kind: bytes / char length
        self.dims = None        # list of (lb_ast or None, ub_ast) for explicit shape
        self.rank = 0
        self.alloc = False
        self.dummy = False
        self.intent = None
        self.init = None
        self.used = False

    @property
    def ctype(self):
        if self.base == "int":
            return {2: "int16_t", 4: "int32_t"}[self.kind]
        if self.base == "real":
            return {4: "float", 8: "double"}[self.kind]
        if self.base == "char":
            return "char"
        raise TranslateError(self.base)

    @property
    def typ(self):
        return (self.base, self.kind)


def _split_top(s: str, sep=","):
    parts, depth, cur, q = [], 0, [], None
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return [p for p in parts if p != ""]


_DECL_RE = re.compile(r"^(real|integer|character|logical)\b(.*)$")


def parse_decl(text, default_real=8):
    """-> list of Sym or None if `text` is not a type declaration"""
    m = _DECL_RE.match(text)
    if not m:
        return None
    base_kw, rest = m.group(1), m.group(2).strip()
    base = {"real": "real", "integer": "int", "character": "char", "logical": "int"}[base_kw]
    kind = {"real": default_real, "int": 4, "char": 1}[base]
    # type parameter directly after the keyword
    m2 = re.match(r"^\*\s*(\d+)(.*)$", rest)
    if m2:
        kind, rest = int(m2.group(1)), m2.group(2).strip()
    else:
        m2 = re.match(r"^\(\s*(?:len\s*=\s*)?(\d+)\s*\)(.*)$", rest)
        if m2:
            kind, rest = int(m2.group(1)), m2.group(2).strip()
    attrs = []
    if "::" in rest:
        left, right = rest.split("::", 1)
        attrs = _split_top(left.strip().lstrip(","))
        ents = right.strip()
    else:
        ents = rest.lstrip(",").strip()
    dim_attr, alloc, intent = None, False, None
    for a in attrs:
        if a == "allocatable":
            alloc = True
        elif a.startswith("dimension"):
            dim_attr = a[a.index("(") + 1: a.rindex(")")]
        elif a.startswith("intent"):
            intent = a[a.index("(") + 1: a.rindex(")")].strip()
        else:
            raise TranslateError(f"attribute {a!r} not supported: {text}")
    syms = []
    for ent in _split_top(ents):
        init = None
        if "=" in ent:
            ent, init_s = ent.split("=", 1)
            init = Parser(tokenize(init_s)).expr()
            ent = ent.strip()
        ek = kind
        m3 = re.match(r"^(.*)\*\s*(\d+)$", ent)
        if m3:
            ent, ek = m3.group(1).strip(), int(m3.group(2))
        dims_s = dim_attr
        m4 = re.match(r"^([a-z_][a-z0-9_]*)\s*\((.*)\)$", ent)
        if m4:
            ent, dims_s = m4.group(1), m4.group(2)
        s = Sym(ent, base, ek)
        s.alloc, s.intent, s.init = alloc, intent, init
        if dims_s is not None:
            dims = []
            for d in _split_top(dims_s):
                if d == ":":
                    dims.append(None)
                elif ":" in d:
                    lo, hi = d.split(":", 1)
                    dims.append((Parser(tokenize(lo)).expr(), Parser(tokenize(hi)).expr()))
                else:
                    dims.append((("num_i", "1"), Parser(tokenize(d)).expr()))
            s.rank = len(dims)
            s.dims = None if alloc else dims
        syms.append(s)
    return syms


# ======================================================================================= units
class Unit:
    def __init__(self, kind, name, dummies, line, srcfile):
        self.kind, self.name, self.dummies, self.line, self.srcfile = kind, name, dummies, line, srcfile
        self.syms = {}
        self.decl_order = []
        self.body = []          # (line, text)

    def add(self, s: Sym):
        if s.name in self.syms:
            raise TranslateError(f"{self.name}: {s.name} declared twice")
        self.syms[s.name] = s
        self.decl_order.append(s.name)


def split_units(stmts, srcfile):
    units, cur = [], None
    for line, text in stmts:
        if re.match(r"^(module\b(?!\s*procedure)|end\s*module|contains$|implicit\s+none$|use\b)", text):
            continue
        m = re.match(r"^subroutine\s+([a-z_][a-z0-9_]*)\s*\((.*)\)$", text)
        if m:
            cur = Unit("sub", m.group(1), [a.strip() for a in m.group(2).split(",") if a.strip()], line, srcfile)
            units.append(cur)
            continue
        m = re.match(r"^program\s+([a-z_][a-z0-9_]*)$", text)
        if m:
            cur = Unit("prog", m.group(1), [], line, srcfile)
            units.append(cur)
            continue
        if re.match(r"^end\s*(subroutine|program)\b", text):
            cur = None
            continue
        if cur is None:
            raise TranslateError(f"{srcfile}:{line}: statement outside a program unit: {text}")
        decl = parse_decl(text) if not cur.body else None
        if decl is not None:
            for s in decl:
                cur.add(s)
        else:
            cur.body.append((line, text))
    for u in units:
        for d in u.dummies:
            if d not in u.syms:
                raise TranslateError(f"{u.name}: dummy {d} undeclared")
            u.syms[d].dummy = True
    return units


# ======================================================================================= codegen
def c_string(s: str) -> str:
    out = []
    for ch in s:
        o = ord(ch)
        if ch == "\\":
            out.append("\\\\")
        elif ch == '"':
            out.append('\\"')
        elif 32 <= o < 127:
            out.append(ch)
        else:
            out.append("\\%03o" % o)
    return '"' + "".join(out) + '"'


def c_real_literal(txt: str) -> str:
    t = txt.replace("d", "e")
    if "." not in t and "e" not in t:
        t += "."
    if t.startswith("."):
        t = "0" + t
    return t.upper() if False else t


INT4, REAL8, LOGICAL = ("int", 4), ("real", 8), ("log", 0)


class Gen:
    def __init__(self, unit: Unit, all_subs: dict, srcname: str):
        self.u, self.subs, self.src = unit, all_subs, srcname
        self.out = []
        self.ind = 1
        self.sect = None        # (extent_code) while translating a statement with array sections
        self.tmp_id = 0
        self.cleanup = []

    # ---------------------------------------------------------------- naming
    def is_prog(self):
        return self.u.kind == "prog"

    def sym(self, name) -> Sym:
        s = self.u.syms.get(name)
        if s is None:
            raise TranslateError(f"{self.src}: {self.u.name}: unknown symbol {name!r}")
        s.used = True
        return s

    def scalar_ref(self, s: Sym) -> str:
        if self.is_prog():
            return f"G.v_{s.name}"
        if s.dummy and s.rank == 0 and s.base != "char":
            return f"v_{s.name}" if s.intent == "in" else f"(*p_{s.name})"
        return f"v_{s.name}"

    def desc(self, s: Sym) -> str:
        if self.is_prog():
            return f"G.v_{s.name}"
        if s.dummy:
            return f"(*d_{s.name})"
        return f"v_{s.name}"

    def array_base(self, s: Sym) -> str:
        if s.alloc:
            return f"(({s.ctype}*){self.desc(s)}.base)"
        if self.is_prog():
            return f"G.v_{s.name}"
        return f"v_{s.name}"

    def bound(self, s: Sym, r: int, what: str) -> str:
        if s.alloc:
            return f"{self.desc(s)}.{'lb' if what == 'l' else 'ext'}[{r}]"
        return f"{s.name}_{what}{r + 1}"

    def total_size(self, s: Sym) -> str:
        return "(" + "*".join(self.bound(s, r, "e") for r in range(s.rank)) + ")"

    def elem(self, s: Sym, idx_codes) -> str:
        if len(idx_codes) != s.rank:
            raise TranslateError(f"{s.name}: rank mismatch")
        off = None
        for r in reversed(range(s.rank)):
            term = f"((long)({idx_codes[r]}) - {self.bound(s, r, 'l')})"
            off = term if off is None else f"({term} + {self.bound(s, r, 'e')}*{off})"
        return f"{self.array_base(s)}[{off}]"

    # ---------------------------------------------------------------- expressions
    def strval(self, node):
        code, t = self.expr(node)
        if t[0] != "char":
            raise TranslateError("character expression expected")
        return code

    def expr(self, n):
        k = n[0]
        if k == "num_i":
            return n[1], INT4
        if k == "num_r":
            return c_real_literal(n[1]), REAL8
        if k == "str":
            return f"f_lit({c_string(n[1])}, {len(n[1])})", ("char", len(n[1]))
        if k == "paren":
            c, t = self.expr(n[1])
            return f"({c})", t
        if k == "name":
            s = self.sym(n[1])
            if s.rank:
                raise TranslateError(f"whole-array reference to {s.name} in a scalar expression")
            if s.base == "char":
                return f"f_var({self.scalar_ref(s)}, {s.kind})", ("char", s.kind)
            return self.scalar_ref(s), s.typ
        if k == "un":
            c, t = self.expr(n[2])
            if n[1] == "!":
                return f"(!({c}))", LOGICAL
            return f"(-({c}))", t
        if k == "bin":
            a, ta = self.expr(n[2])
            b, tb = self.expr(n[3])
            if n[1] in ("&&", "||"):
                return f"(({a}) {n[1]} ({b}))", LOGICAL
            return f"(({a}) {n[1]} ({b}))", self.arith_type(ta, tb)
        if k == "rel":
            a, ta = self.expr(n[2])
            b, tb = self.expr(n[3])
            if ta[0] == "char" or tb[0] == "char":
                if n[1] not in ("==", "!="):
                    raise TranslateError("character ordering comparison not supported")
                return (f"f_str_eq({a}, {b})" if n[1] == "==" else f"(!f_str_eq({a}, {b}))"), LOGICAL
            return f"(({a}) {n[1]} ({b}))", LOGICAL
        if k == "concat":
            return f"f_concat({self.strval(n[1])}, {self.strval(n[2])})", ("char", None)
        if k == "pow":
            a, ta = self.expr(n[1])
            if n[2][0] != "num_i" or n[2][1] not in ("2", "3"):
                raise TranslateError("only **2 and **3 are supported")
            e = int(n[2][1])
            code = f"(({a})*({a}))" if e == 2 else f"((({a})*({a}))*({a}))"
            return code, ta
        if k == "call":
            return self.call_expr(n)
        raise TranslateError(f"expression node {k} not supported here")

    @staticmethod
    def arith_type(ta, tb):
        if ta[0] == "real" or tb[0] == "real":
            ka = ta[1] if ta[0] == "real" else 0
            kb = tb[1] if tb[0] == "real" else 0
            return ("real", max(ka, kb))
        return ("int", max(ta[1] if ta[0] == "int" else 4, tb[1] if tb[0] == "int" else 4))

    def call_expr(self, n):
        name, args = n[1], n[2]
        s = self.u.syms.get(name)
        if s is not None:
            s.used = True
            if s.rank:
                idx = []
                for r, a in enumerate(args):
                    if a[0] == "range":
                        if a[1] is not None or a[2] is not None or self.sect is None:
                            raise TranslateError(f"array section of {name} not supported here")
                        ext = self.bound(s, r, "e")
                        if self.sect["extent"] is None:
                            self.sect["extent"] = ext
                        idx.append(f"({self.bound(s, r, 'l')} + _q)")
                    else:
                        c, t = self.expr(a)
                        if t[0] != "int":
                            raise TranslateError(f"non-integer subscript of {name}")
                        idx.append(c)
                return self.elem(s, idx), s.typ
            if s.base == "char" and len(args) == 1 and args[0][0] == "range":
                lo = self.expr(args[0][1])[0] if args[0][1] is not None else "1"
                hi = self.expr(args[0][2])[0] if args[0][2] is not None else str(s.kind)
                return f"f_substr(f_var({self.scalar_ref(s)}, {s.kind}), {lo}, {hi})", ("char", None)
            raise TranslateError(f"{name} is not an array")
        # intrinsics
        av = [self.expr(a) for a in args]
        def real_kind():
            ks = [t[1] for _, t in av if t[0] == "real"]
            return max(ks) if ks else 0
        if name == "sqrt":
            (c, t), = av
            return (f"sqrt({c})" if t[1] == 8 else f"sqrtf({c})"), t
        if name == "abs":
            (c, t), = av
            if t[0] == "int":
                return f"abs({c})", t
            return (f"fabs({c})" if t[1] == 8 else f"fabsf({c})"), t
        if name in ("max", "min"):
            if len(av) < 2:
                raise TranslateError("max/min need two arguments")
            rk = real_kind()
            fn = {0: f"f_i{name}", 4: f"f_{name}f", 8: f"f_{name}"}[rk]
            code = av[0][0]
            for c, _ in av[1:]:
                code = f"{fn}({code}, {c})"
            return code, (("real", rk) if rk else INT4)
        if name == "floor":
            (c, t), = av
            return f"((int32_t)floor({c}))", INT4
        if name == "ceiling":
            (c, t), = av
            return f"((int32_t)ceil({c}))", INT4
        if name == "isnan":
            (c, t), = av
            return f"(isnan({c}) != 0)", LOGICAL
        if name == "len_trim":
            return f"f_len_trim({av[0][0]})", INT4
        if name == "trim":
            return f"f_trim({av[0][0]})", ("char", None)
        if name == "char":
            return f"f_char({av[0][0]})", ("char", 1)
        raise TranslateError(f"{self.src}: function/array {name!r} not known")

    # ---------------------------------------------------------------- output helpers
    def emit(self, s):
        self.out.append("    " * self.ind + s)

    def tmp(self):
        self.tmp_id += 1
        return f"_t{self.tmp_id}"

    # ---------------------------------------------------------------- statements
    def lvalue(self, node):
        if node[0] == "name":
            s = self.sym(node[1])
            return self.scalar_ref(s), s
        if node[0] == "call":
            code, _ = self.call_expr(node)
            return code, self.sym(node[1])
        raise TranslateError("bad assignment target")

    def assignment(self, lhs, rhs):
        if lhs[0] == "name" and self.sym(lhs[1]).rank:
            return self.whole_array_assign(self.sym(lhs[1]), rhs)
        if lhs[0] == "name" and self.sym(lhs[1]).base == "char":
            s = self.sym(lhs[1])
            self.emit(f"f_str_assign({self.scalar_ref(s)}, {s.kind}, {self.strval(rhs)});")
            return
        self.sect = {"extent": None}
        lcode, _ = self.lvalue(lhs)
        rcode, _ = self.expr(rhs)
        ext = self.sect["extent"]
        self.sect = None
        if ext is not None:
            self.emit(f"for (long _q = 0; _q < {ext}; ++_q) {lcode} = {rcode};")
        else:
            self.emit(f"{lcode} = {rcode};")

    def whole_array_assign(self, s: Sym, rhs):
        if rhs[0] == "name" and self.u.syms.get(rhs[1]) is not None and self.sym(rhs[1]).rank:
            r = self.sym(rhs[1])
            if s.alloc and r.alloc:
                self.emit(f"f_assign_alloc(&{self.desc(s)}, &{self.desc(r)});")
                return
            if s.alloc or r.rank != s.rank:
                raise TranslateError(f"array assignment {s.name} = {r.name} not supported")
            self.emit(f"for (long _q = 0; _q < {self.total_size(s)}; ++_q) "
                      f"{self.array_base(s)}[_q] = {self.array_base(r)}[_q];")
            return
        if rhs[0] == "arrcons":
            if s.rank != 1:
                raise TranslateError("array constructor needs a rank-1 target")
            self.emit("{")
            self.ind += 1
            names = []
            for it in rhs[1]:                      # the whole right-hand side is evaluated first
                t = self.tmp()
                self.emit(f"const {s.ctype} {t} = {self.expr(it)[0]};")
                names.append(t)
            for q, t in enumerate(names):
                self.emit(f"{self.array_base(s)}[{q}] = {t};")
            self.ind -= 1
            self.emit("}")
            return
        code, _ = self.expr(rhs)                   # scalar broadcast
        t = self.tmp()
        self.emit(f"{{ const {s.ctype} {t} = {code}; for (long _q = 0; _q < {self.total_size(s)}; ++_q) "
                  f"{self.array_base(s)}[_q] = {t}; }}")

    def call_stmt(self, name, args):
        if name == "cpu_time":
            s = self.sym(args[0][1])
            self.emit(f"f_cpu_time(&{self.scalar_ref(s)});")
            return
        if name == "getarg":
            s = self.sym(args[1][1])
            self.emit(f"f_getarg({self.expr(args[0])[0]}, {self.scalar_ref(s)}, {s.kind});")
            return
        callee = self.subs.get(name)
        if callee is None:
            raise TranslateError(f"CALL of unknown subroutine {name}")
        if len(args) != len(callee.dummies):
            raise TranslateError(f"CALL {name}: argument count")
        actuals = []
        for a, dname in zip(args, callee.dummies):
            d = callee.syms[dname]
            if d.rank:
                if a[0] != "name" or not self.sym(a[1]).rank:
                    raise TranslateError(f"CALL {name}: array actual expected for {dname}")
                s = self.sym(a[1])
                if d.alloc:
                    if not s.alloc:
                        raise TranslateError("allocatable dummy needs allocatable actual")
                    actuals.append(f"&{self.desc(s)}")
                else:
                    if s.typ != d.typ:
                        raise TranslateError(f"CALL {name}: type mismatch on {dname}")
                    actuals.append(self.array_base(s))
            elif d.base == "char":
                s = self.sym(a[1])
                if s.kind != d.kind:
                    raise TranslateError("character length mismatch")
                actuals.append(self.scalar_ref(s))
            else:
                code, t = self.expr(a)
                is_var = a[0] == "name" or (a[0] == "call" and a[1] in self.u.syms)
                if is_var:
                    if t != d.typ:
                        raise TranslateError(f"CALL {name}: type mismatch on {dname} ({t} vs {d.typ})")
                    actuals.append(f"({d.ctype}*)&{code}")
                else:
                    actuals.append(f"&({d.ctype}){{{code}}}")
        self.emit(f"ref_{name}({', '.join(actuals)});")

    # ---- I/O
    def io_items(self, p: Parser):
        items = []
        while not p.done():
            if p.at_op(","):
                p.next()
                continue
            items.append(self.io_item(p))
        return items

    def io_item(self, p: Parser):
        if p.at_op("("):
            # implied DO if an '=' sits at depth 1
            depth, q, is_do = 0, p.p, False
            while q < len(p.t):
                k, v = p.t[q]
                if k == "op" and v in ("(", "(/"):
                    depth += 1
                elif k == "op" and v in (")", "/)"):
                    depth -= 1
                    if depth == 0:
                        break
                elif k == "op" and v == "=" and depth == 1:
                    is_do = True
                q += 1
            if is_do:
                p.next()
                inner = []
                while True:
                    if p.peek()[0] == "name" and p.peek(1) == ("op", "="):
                        var = p.next()[1]
                        p.next()
                        lo = p.expr()
                        p.expect_op(",")
                        hi = p.expr()
                        step = None
                        if p.at_op(","):
                            p.next()
                            step = p.expr()
                        p.expect_op(")")
                        return ("impdo", inner, var, lo, hi, step)
                    inner.append(self.io_item(p))
                    p.expect_op(",")
        return p.expr()

    def emit_io_item(self, it, emit_one):
        """emit_one(code, type, sym_or_None)"""
        if it[0] == "impdo":
            _, inner, var, lo, hi, step = it
            v = self.scalar_ref(self.sym(var))
            st = self.expr(step)[0] if step else "1"
            self.emit(f"for ({v} = {self.expr(lo)[0]}; ({st}) > 0 ? {v} <= ({self.expr(hi)[0]}) : {v} >= ({self.expr(hi)[0]}); {v} += {st}) {{")
            self.ind += 1
            for x in inner:
                self.emit_io_item(x, emit_one)
            self.ind -= 1
            self.emit("}")
            return
        self.sect = {"extent": None}
        code, t = self.expr(it)
        ext = self.sect["extent"]
        self.sect = None
        if ext is not None:
            self.emit(f"for (long _q = 0; _q < {ext}; ++_q) {{")
            self.ind += 1
            emit_one(code, t)
            self.ind -= 1
            self.emit("}")
        else:
            emit_one(code, t)

    def io_stmt(self, kw, text, line):
        toks = tokenize(text)
        p = Parser(toks)
        p.next()                                   # keyword
        if kw == "print":
            p.expect_op("*")
            items = self.io_items(p)
            self.emit(f"f_pr_begin({line});")
            def one(code, t):
                if t[0] == "char":
                    self.emit(f"f_pr_s({code});")
                elif t[0] == "int":
                    self.emit(f"f_pr_i({code});")
                else:
                    self.emit(f"f_pr_r({code});")
            for it in items:
                self.emit_io_item(it, one)
            self.emit("f_pr_end();")
            return
        p.expect_op("(")
        ctl, pos = {}, 0
        while True:
            if p.at_op("*"):
                p.next()
                ctl[pos] = "*"
            else:
                a = p.arg()
                if a[0] == "kw":
                    ctl[a[1]] = a[2]
                else:
                    ctl[pos] = a
            pos += 1
            if p.at_op(","):
                p.next()
                continue
            p.expect_op(")")
            break
        unit = ctl.get(0, ctl.get("unit"))
        if kw == "open":
            def s(key, default):
                return self.strval(ctl[key]) if key in ctl else f'f_lit("{default}", {len(default)})'
            self.emit(f"f_open({self.expr(unit)[0]}, {self.strval(ctl['file'])}, {s('status', 'unknown')}, "
                      f"{s('access', 'sequential')}, {s('form', 'formatted')});")
            return
        if kw == "close":
            self.emit(f"f_close({self.expr(unit)[0]});")
            return
        items = self.io_items(p)
        fmt = ctl.get(1, ctl.get("fmt"))
        if kw == "read":
            if fmt is not None:
                raise TranslateError("formatted READ not supported")
            u = self.expr(unit)[0]
            for it in items:
                if it[0] == "name" and self.sym(it[1]).base == "char":
                    s_ = self.sym(it[1])
                    self.emit(f"f_read({u}, {self.scalar_ref(s_)}, {s_.kind});")
                else:
                    code, s_ = self.lvalue(it)
                    self.emit(f"f_read({u}, &{code}, sizeof({s_.ctype}));")
            return
        # WRITE
        internal = unit != "*" and unit[0] == "name" and self.sym(unit[1]).base == "char"
        if fmt is None:                            # unformatted stream
            u = self.expr(unit)[0]
            def one(code, t):
                if t[0] == "char":
                    self.emit(f"f_write_s({u}, {code});")
                else:
                    ct = {("int", 2): "int16_t", ("int", 4): "int32_t", ("real", 4): "float", ("real", 8): "double"}[t]
                    self.emit(f"{{ const {ct} _w = {code}; f_write({u}, &_w, sizeof _w); }}")
            for it in items:
                self.emit_io_item(it, one)
            return
        if fmt == "*":                             # list-directed to a file
            self.emit(f"f_ld_begin({self.expr(unit)[0]}, {line});")
            def one(code, t):
                self.emit({"char": f"f_ld_s({code});", "int": f"f_ld_i({code});", "real": f"f_ld_r({code});"}[t[0]])
            for it in items:
                self.emit_io_item(it, one)
            self.emit("f_ld_end();")
            return
        self.emit(f"f_fmt_begin({self.strval(fmt)});")
        def one(code, t):
            self.emit({"char": f"f_fmt_s({code});", "int": f"f_fmt_i({code});", "real": f"f_fmt_r({code});"}[t[0]])
        for it in items:
            self.emit_io_item(it, one)
        if internal:
            s_ = self.sym(unit[1])
            self.emit(f"f_fmt_end_internal({self.scalar_ref(s_)}, {s_.kind});")
        elif unit == "*":
            self.emit("f_fmt_end_unit(-2);")
        else:
            self.emit(f"f_fmt_end_unit({self.expr(unit)[0]});")

    def allocate_stmt(self, text):
        p = Parser(tokenize(text))
        p.next()
        p.expect_op("(")
        while True:
            name = p.next()[1]
            s = self.sym(name)
            if not s.alloc:
                raise TranslateError(f"ALLOCATE of non-allocatable {name}")
            p.expect_op("(")
            dims = p.arg_list(")")
            lbs, ubs = [], []
            for d in dims:
                if d[0] == "range":
                    lbs.append(self.expr(d[1])[0])
                    ubs.append(self.expr(d[2])[0])
                else:
                    lbs.append("1")
                    ubs.append(self.expr(d)[0])
            self.emit(f"{{ const long _lb[] = {{{', '.join(lbs)}}}, _ub[] = {{{', '.join(ubs)}}}; "
                      f"f_allocate(&{self.desc(s)}, {len(dims)}, sizeof({s.ctype}), _lb, _ub); }}")
            if p.at_op(","):
                p.next()
                continue
            p.expect_op(")")
            break

    def simple_stmt(self, line, text):
        """one non-block statement"""
        if text == "exit":
            self.emit("break;")
        elif text == "stop":
            self.emit("f_stop();")
        elif text.startswith("call "):
            p = Parser(tokenize(text[5:]))
            name = p.next()[1]
            args = []
            if p.at_op("("):
                p.next()
                args = p.arg_list(")")
            self.call_stmt(name, args)
        elif re.match(r"^print\s*\*", text):
            self.io_stmt("print", text, line)
        elif re.match(r"^(open|close|read|write)\s*\(", text):
            self.io_stmt(text[: re.match(r"^[a-z]+", text).end()], text, line)
        elif re.match(r"^allocate\s*\(", text):
            self.allocate_stmt(text)
        elif re.match(r"^deallocate\s*\(", text):
            inner = text[text.index("(") + 1: text.rindex(")")]
            for nm in _split_top(inner):
                self.emit(f"f_deallocate(&{self.desc(self.sym(nm))});")
        else:
            p = Parser(tokenize(text))
            lhs = p.primary()
            p.expect_op("=")
            rhs = p.expr()
            if not p.done():
                raise TranslateError(f"trailing tokens in assignment: {text}")
            self.assignment(lhs, rhs)

    def body(self):
        stack = []
        gate_open = False
        for line, text in self.u.body:
            top = self.is_prog() and not stack
            closing = re.match(r"^(end\s*do|end\s*if|else\b|else\s*if\b|elseif\b)", text) is not None
            if top and not closing:
                self.emit(f"if (REF_GATE({line})) {{")
                self.ind += 1
                gate_open = True
            self.emit(f"/* {self.src}:{line} */")
            m = re.match(r"^do\s+([a-z_][a-z0-9_]*)\s*=(.*)$", text)
            if m:
                v = self.scalar_ref(self.sym(m.group(1)))
                p = Parser(tokenize(m.group(2)))
                lo = self.expr(p.expr())[0]
                p.expect_op(",")
                hi = self.expr(p.expr())[0]
                st = "1"
                if p.at_op(","):
                    p.next()
                    st = self.expr(p.expr())[0]
                e, s_ = self.tmp(), self.tmp()
                self.emit(f"{{ const int32_t {e} = {hi}, {s_} = {st};")
                self.emit(f"for ({v} = {lo}; {s_} > 0 ? {v} <= {e} : {v} >= {e}; {v} += {s_}) {{")
                self.ind += 1
                stack.append("do")
                continue
            if re.match(r"^end\s*do$", text):
                if stack.pop() != "do":
                    raise TranslateError(f"{self.src}:{line}: END DO mismatch")
                self.ind -= 1
                self.emit("} }")
            elif re.match(r"^if\s*\(", text):
                # find the matching paren of the condition
                depth, q = 0, text.index("(")
                start = q
                while True:
                    if text[q] == "(":
                        depth += 1
                    elif text[q] == ")":
                        depth -= 1
                        if depth == 0:
                            break
                    q += 1
                cond = self.expr(Parser(tokenize(text[start + 1: q])).expr())[0]
                rest = text[q + 1:].strip()
                if rest == "then":
                    self.emit(f"if ({cond}) {{")
                    self.ind += 1
                    stack.append("if")
                    continue
                self.emit(f"if ({cond}) {{")
                self.ind += 1
                self.simple_stmt(line, rest)
                self.ind -= 1
                self.emit("}")
            elif re.match(r"^(else\s*if|elseif)\s*\(", text):
                cond_s = text[text.index("(") + 1: text.rindex(")")]
                cond = self.expr(Parser(tokenize(cond_s)).expr())[0]
                self.ind -= 1
                self.emit(f"}} else if ({cond}) {{")
                self.ind += 1
                continue
            elif text == "else":
                self.ind -= 1
                self.emit("} else {")
                self.ind += 1
                continue
            elif re.match(r"^end\s*if$", text):
                if stack.pop() != "if":
                    raise TranslateError(f"{self.src}:{line}: END IF mismatch")
                self.ind -= 1
                self.emit("}")
            else:
                self.simple_stmt(line, text)
            if self.is_prog() and not stack and gate_open:
                self.ind -= 1
                self.emit("}")
                gate_open = False
        if stack:
            raise TranslateError(f"{self.u.name}: unterminated block")

    # ---------------------------------------------------------------- bounds
    def bounds_consts(self, s: Sym):
        lines = []
        for r, (lb, ub) in enumerate(s.dims):
            l = self.expr(lb)[0]
            u_ = self.expr(ub)[0]
            lines.append(f"const long {s.name}_l{r + 1} = {l}, {s.name}_e{r + 1} = (long)({u_}) - (long)({l}) + 1;")
        return lines


def mark_used(u: Unit):
    """which symbols does the body reference (used to skip dead automatic arrays)"""
    names = set()
    for _, text in u.body:
        for k, v in tokenize(text):
            if k == "name":
                names.add(v)
    return names


def gen_sub(u: Unit, subs, srcname):
    g = Gen(u, subs, srcname)
    used = mark_used(u)
    params = []
    for dn in u.dummies:
        d = u.syms[dn]
        if d.rank:
            params.append(f"f_desc *d_{dn}" if d.alloc else f"{d.ctype} *restrict v_{dn}")
        elif d.base == "char":
            params.append(f"char *v_{dn}")
        else:
            params.append(f"{d.ctype} *p_{dn}")
    # Fortran's no-alias rule for dummies is what `restrict` states; INOUT arrays that the caller passes
    # twice do not occur in the two files.
    head = f"void ref_{u.name}({', '.join(params)})"
    pre = []
    for dn in u.dummies:
        d = u.syms[dn]
        if not d.rank and d.base != "char" and d.intent == "in":
            pre.append(f"const {d.ctype} v_{dn} = *p_{dn};")
    autos = []
    for name in u.decl_order:
        s = u.syms[name]
        if s.dims is not None and (s.dummy or name in used):
            pre.extend(g.bounds_consts(s))
    for name in u.decl_order:
        s = u.syms[name]
        if s.dummy:
            continue
        if s.alloc:
            pre.append(f"f_desc v_{name} = {{0}};")
        elif s.dims is not None:
            if name not in used:
                pre.append(f"/* automatic array {name} is never referenced: not allocated */")
                continue
            size = "*".join(f"{name}_e{r + 1}" for r in range(s.rank))
            guard = f"(4*{name}_e1*{name}_e2 + 64)" if s.rank >= 3 else "64"
            pre.append(f"{s.ctype} *v_{name} = ({s.ctype}*)f_auto({size}, sizeof({s.ctype}), {guard});")
            autos.append((name, s.ctype, guard))
        elif s.base == "char":
            pre.append(f"char v_{name}[{s.kind}]; memset(v_{name}, ' ', {s.kind});")
        elif s.init is not None:
            pre.append(f"static {s.ctype} v_{name} = {g.expr(s.init)[0]};   /* initialised => SAVE */")
        else:
            pre.append(f"{s.ctype} v_{name} = 0; (void)v_{name};")
    g.body()
    lines = [f"/* {srcname}:{u.line}  SUBROUTINE {u.name} */", head, "{"]
    lines += ["    " + x for x in pre]
    lines += g.out
    for name, ct, guard in autos:
        lines.append(f"    f_auto_free(v_{name}, sizeof({ct}), {guard});")
    lines.append("}")
    # guarded entry for the test harness: STOP comes back as return value 1
    args = []
    for dn in u.dummies:
        d = u.syms[dn]
        args.append(f"d_{dn}" if (d.rank and d.alloc) else (f"v_{dn}" if (d.rank or d.base == 'char') else f"p_{dn}"))
    lines += ["", f"int ref_call_{u.name}({', '.join(p.replace('restrict ', '') for p in params)})", "{",
              "    int rc = 0;", "    ref_stop_armed = 1;",
              f"    if (setjmp(ref_stop_jmp) == 0) ref_{u.name}({', '.join(args)}); else rc = 1;",
              "    ref_stop_armed = 0;", "    return rc;", "}", ""]
    return head + ";", "\n".join(lines)


def gen_prog(u: Unit, subs, srcname):
    g = Gen(u, subs, srcname)
    fields, consts, inits, table = [], [], [], []
    for name in u.decl_order:
        s = u.syms[name]
        if s.alloc:
            fields.append(f"f_desc v_{name};")
            table.append(f'{{"{name}", 2, &G.v_{name}, {s.rank}, {REF_ELT[s.typ]}}},')
        elif s.dims is not None:
            n = 1
            for r, (lb, ub) in enumerate(s.dims):
                if lb[0] != "num_i" or ub[0] != "num_i":
                    raise TranslateError("program arrays need constant bounds")
                consts.append(f"static const long {name}_l{r + 1} = {lb[1]}, {name}_e{r + 1} = {int(ub[1]) - int(lb[1]) + 1};")
                n *= int(ub[1]) - int(lb[1]) + 1
            fields.append(f"{s.ctype} v_{name}[{n}];")
            table.append(f'{{"{name}", 4, G.v_{name}, {n}, {REF_ELT[s.typ]}}},')
        elif s.base == "char":
            fields.append(f"char v_{name}[{s.kind}];")
            inits.append(f"memset(G.v_{name}, ' ', {s.kind});")
            if s.init is not None:
                inits.append(f"f_str_assign(G.v_{name}, {s.kind}, {g.strval(s.init)});")
            table.append(f'{{"{name}", 3, G.v_{name}, {s.kind}, 0}},')
        else:
            fields.append(f"{s.ctype} v_{name};")
            if s.init is not None:
                inits.append(f"G.v_{name} = {g.expr(s.init)[0]};")
            table.append(f'{{"{name}", {0 if s.base == "int" else 1}, &G.v_{name}, 1, {REF_ELT[s.typ]}}},')
    g.body()
    L = [f"/* {srcname}:{u.line}  PROGRAM {u.name}: all program variables live in G */",
         "static struct {"] + ["    " + f for f in fields] + ["} G;", ""] + consts + [
        "", "static int ref_first_line, ref_last_line, ref_inited;",
        "#define REF_GATE(L) ((L) >= ref_first_line && (L) <= ref_last_line)", "",
        "static void ref_program_body(void)", "{",
        "    if (!ref_inited) {", "        ref_inited = 1;"] + ["        " + i for i in inits] + ["    }"]
    L += g.out + ["}", "",
                  "/* executes the top-level statements of the program that START on lines first..last */",
                  "int ref_program_exec(int first_line, int last_line)", "{",
                  "    int rc = 0;", "    ref_first_line = first_line; ref_last_line = last_line;",
                  "    ref_stop_armed = 1;",
                  "    if (setjmp(ref_stop_jmp) == 0) ref_program_body(); else rc = 1;",
                  "    ref_stop_armed = 0;", "    return rc;", "}", "",
                  "typedef struct { const char *name; int kind; void *ptr; long n; int elt; } ref_var_entry;",
                  "/* kind: 0 INTEGER scalar, 1 REAL scalar, 2 allocatable (f_desc), 3 CHARACTER(n), 4 fixed array(n);",
                  "   elt: 0 int32, 1 double, 2 float, 3 int16 */",
                  "static const ref_var_entry ref_vars[] = {"] + ["    " + t for t in table] + ["    {0, 0, 0, 0, 0}", "};", "",
                  "const ref_var_entry *ref_var(const char *name)", "{",
                  "    for (const ref_var_entry *e = ref_vars; e->name; ++e) if (!strcmp(e->name, name)) return e;",
                  "    return 0;", "}", "",
                  "void ref_program_reset(void)", "{",
                  "    for (const ref_var_entry *e = ref_vars; e->name; ++e) if (e->kind == 2) f_deallocate((f_desc*)e->ptr);",
                  "    memset(&G, 0, sizeof G); ref_inited = 0;", "}", ""]
    return "\n".join(L)


REF_ELT = {("int", 4): 0, ("real", 8): 1, ("real", 4): 2, ("int", 2): 3, ("char", 1): 0}


HEADER = """/* GENERATED by oracle/f90_to_c.py from {src} -- do not edit, do not commit.
 * A statement-by-statement translation of the reference's Fortran text (see the script's docstring
 * for the arithmetic contract).  TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's CPU legs
 * may load the library built from this file. */
#include "ref_runtime.h"
"""


def translate(ref_dir: str, out_dir: str):
    os.makedirs(out_dir, exist_ok=True)
    subs_src, prog_src = os.path.join(ref_dir, "subs.f90"), os.path.join(ref_dir, "set3d.f90")
    mod_units = split_units(read_statements(subs_src), "subs.f90")
    subs = {u.name: u for u in mod_units}
    protos, bodies = [], []
    for u in mod_units:
        proto, body = gen_sub(u, subs, "subs.f90")
        protos.append(proto)
        bodies.append(body)
    with open(os.path.join(out_dir, "ref_subs.h"), "w") as f:
        f.write(HEADER.format(src="subs.f90") + "\n".join(protos) + "\n")
    with open(os.path.join(out_dir, "ref_subs.c"), "w") as f:
        f.write(HEADER.format(src="subs.f90") + '#include "ref_subs.h"\n\n' + "\n".join(bodies))
    prog_units = split_units(read_statements(prog_src), "set3d.f90")
    if len(prog_units) != 1 or prog_units[0].kind != "prog":
        raise TranslateError("set3d.f90: expected exactly one PROGRAM")
    with open(os.path.join(out_dir, "ref_set3d.c"), "w") as f:
        f.write(HEADER.format(src="set3d.f90") + '#include "ref_subs.h"\n\n' + gen_prog(prog_units[0], subs, "set3d.f90"))
    n_stmt = sum(len(u.body) for u in mod_units) + len(prog_units[0].body)
    print(f"f90_to_c: translated {len(mod_units)} subroutines + 1 program, {n_stmt} executable statements -> {out_dir}")


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    translate(sys.argv[1], sys.argv[2])
