#!/usr/bin/env python
"""go2cpp.py — a syntax-directed translator from the Go subset the reference is written in to C++17.

TEST INFRASTRUCTURE ONLY (part of oracle/).  There is no Go toolchain in this image, so the reference
(tbogdala/cubez: rigidbody.go, colliders.go, contact.go, math/*.go) cannot be compiled as Go.  This tool
reads those files WHERE THEY LIE under /root/reference plus the headless harness mains of go/harness/,
translates them statement by statement into one C++ translation unit, and oracle/Makefile compiles the
result into oracle/_ref/ (git-ignored; never committed).  The translation is mechanical — the translator
knows Go syntax, not physics — so the resulting binary executes the reference's own source text:
expression trees are kept exactly (every binary expression is parenthesised as Go parses it, no
re-association, compiled with -ffp-contract=off like Go on amd64), value/pointer semantics of Go arrays,
structs, slices and interfaces are mapped onto C++ value types, pointers, a shared-backing slice and
abstract classes.  It is how the hand-written restatement (oracle/cubez_oracle.hpp) is pinned against
the reference itself: tests/test_oracle_vs_transpiled_reference.py compares their dumps bit for bit.

What is NOT the reference here: the Go compiler and runtime (replaced by this translator + g++), and
Go's math.Pow (C pow(); the library takes the three Pow factors as host inputs for exactly that reason).

Supported subset (everything the reference and the harnesses use): package/import clauses, const/var/type
declarations (named basic, array, struct and interface types), functions and methods (pointer and value
receivers, multiple and named results), if/else, the three for forms and range loops, expression and type
switches, short variable declarations, tuple assignment, inc/dec, composite literals, conversions, type
assertions, append/len/new, untyped-constant folding in exact rational arithmetic.
Usage: go2cpp.py -o out.cpp pkgpath=dir[,file...] ... (packages in dependency order; the last one is main)
"""
from __future__ import annotations

import os
import re
import sys
from fractions import Fraction

KEYWORDS = {"break", "case", "chan", "const", "continue", "default", "defer", "else", "fallthrough", "for", "func", "go", "goto",
            "if", "import", "interface", "map", "package", "range", "return", "select", "struct", "switch", "type", "var"}
CPP_RESERVED = {"this", "new", "delete", "class", "template", "namespace", "register", "union", "auto", "operator", "private", "public",
                "protected", "friend", "virtual", "typename", "using", "static", "extern", "inline", "double", "float", "int", "long",
                "short", "char", "signed", "unsigned", "void", "bool", "do", "while", "try", "catch", "throw", "enum", "typedef", "volatile",
                "const_cast", "and", "or", "not", "xor", "near", "far", "index", "max", "min", "abs", "time", "main"}
OPS = ["<<=", ">>=", "&^=", "...", "&&", "||", "<-", "++", "--", "==", "!=", "<=", ">=", ":=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=",
       "<<", ">>", "&^", "+", "-", "*", "/", "%", "&", "|", "^", "<", ">", "=", "!", "(", ")", "[", "]", "{", "}", ",", ";", ".", ":"]
BASIC = {"bool": "bool", "int": "long long", "int8": "int8_t", "int16": "int16_t", "int32": "int32_t", "int64": "long long", "uint": "unsigned long long",
         "uint8": "uint8_t", "uint16": "uint16_t", "uint32": "uint32_t", "uint64": "unsigned long long", "byte": "uint8_t", "float32": "float",
         "float64": "double", "string": "std::string", "uintptr": "uintptr_t", "rune": "int32_t"}
STD_NS = {"math": "gomath", "fmt": "gofmt", "os": "goos", "strconv": "gostrconv", "time": "gotime", "testing": "gotesting"}


class Tok:
    __slots__ = ("kind", "val", "line")

    def __init__(self, kind, val, line):
        self.kind, self.val, self.line = kind, val, line

    def __repr__(self):
        return f"{self.kind}:{self.val}@{self.line}"


def lex(src: str, fname: str):
    toks, i, line, n = [], 0, 1, len(src)

    def auto_semi():
        if not toks:
            return False
        t = toks[-1]
        if t.kind in ("ident", "int", "float", "string", "char"):
            return True
        if t.kind == "kw" and t.val in ("break", "continue", "fallthrough", "return"):
            return True
        return t.kind == "op" and t.val in ("++", "--", ")", "]", "}")

    while i < n:
        c = src[i]
        if c == "\n":
            if auto_semi():
                toks.append(Tok("op", ";", line))
            line += 1
            i += 1
        elif c in " \t\r":
            i += 1
        elif src.startswith("//", i):
            while i < n and src[i] != "\n":
                i += 1
        elif src.startswith("/*", i):
            j = src.index("*/", i + 2)
            nl = src.count("\n", i, j)
            if nl and auto_semi():
                toks.append(Tok("op", ";", line))
            line += nl
            i = j + 2
        elif c.isalpha() or c == "_":
            j = i
            while j < n and (src[j].isalnum() or src[j] == "_"):
                j += 1
            w = src[i:j]
            toks.append(Tok("kw" if w in KEYWORDS else "ident", w, line))
            i = j
        elif c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit()):
            m = re.match(r"0[xX][0-9a-fA-F_]+|(\d[\d_]*)?\.\d*([eE][+-]?\d+)?|\d[\d_]*[eE][+-]?\d+|\d[\d_]*", src[i:])
            s = m.group(0)
            isf = not s.lower().startswith("0x") and any(ch in s for ch in ".eE")
            toks.append(Tok("float" if isf else "int", s.replace("_", ""), line))
            i += len(s)
        elif c == '"':
            j = i + 1
            while src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            toks.append(Tok("string", src[i:j + 1], line))
            i = j + 1
        elif c == "`":
            j = src.index("`", i + 1)
            toks.append(Tok("string", '"' + src[i + 1:j].replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n") + '"', line))
            line += src.count("\n", i, j)
            i = j + 1
        elif c == "'":
            j = i + 1
            while src[j] != "'":
                j += 2 if src[j] == "\\" else 1
            toks.append(Tok("char", src[i:j + 1], line))
            i = j + 1
        else:
            for op in OPS:
                if src.startswith(op, i):
                    toks.append(Tok("op", op, line))
                    i += len(op)
                    break
            else:
                raise SyntaxError(f"{fname}:{line}: unexpected character {c!r}")
    if auto_semi():
        toks.append(Tok("op", ";", line))
    toks.append(Tok("eof", "", line))
    return toks


class Expr:
    """A translated expression: C++ text, optional exact constant value, and what kind of thing it names."""
    __slots__ = ("cpp", "const", "kind", "pkg", "tname")

    def __init__(self, cpp, const=None, kind="value", pkg=None, tname=None):
        self.cpp, self.const, self.kind, self.pkg, self.tname = cpp, const, kind, pkg, tname


def const_cpp(c):
    """C++ literal of an exact untyped constant (Fraction, 'int' | 'float')."""
    v, k = c
    if k == "int":
        iv = int(v)
        return f"{iv}LL" if abs(iv) < 2 ** 63 else f"{iv}ULL"
    f = float(v)    # Fraction -> float is correctly rounded (round-half-even), like Go's conversion of the exact constant
    if f != f or f in (float("inf"), float("-inf")):
        raise ValueError("constant overflows float64")
    # An untyped constant has no type until it meets a typed operand: `0.3 < x` is a float32 comparison when x is float32
    # and the constant is then the exact value rounded ONCE to float32.  C++ would promote to double instead, so the
    # literal carries both roundings and picks by the type it meets (struct gouf in the prelude).
    return f"gouf({f.hex()}, {frac_to_f32(v).hex()}f)"


def frac_to_f32(v):
    """The float32 nearest to the exact rational v (round-half-even), returned as a Python float holding exactly that value."""
    v = Fraction(v)
    if v == 0:
        return 0.0
    sign, a = (-1.0 if v < 0 else 1.0), abs(v)
    e = a.numerator.bit_length() - a.denominator.bit_length()      # 2^(e-1) < a < 2^(e+1)
    if Fraction(2) ** e > a:
        e -= 1                                                      # now 2^e <= a < 2^(e+1)
    q_exp = max(e - 23, -149)                                       # spacing of float32 at this magnitude (subnormals: 2^-149)
    scaled = a / (Fraction(2) ** q_exp)
    q = scaled.numerator // scaled.denominator
    rem = scaled - q
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (q & 1)):
        q += 1
    r = Fraction(q) * (Fraction(2) ** q_exp)
    if r >= Fraction(2) ** 128:
        raise ValueError("constant overflows float32")
    return sign * float(r)      # exact: q has at most 25 bits


class Package:
    def __init__(self, path, ns):
        self.path, self.ns = path, ns
        self.types = {}        # name -> ("basic", cpp) | ("array", elem, n) | ("struct", fields) | ("interface", methods) | ("named", cpp)
        self.methods = {}      # type name -> list of (name, proto, ptr_receiver)
        self.consts = {}       # untyped constants: name -> (Fraction, kind)
        self.globals = set()
        self.funcs = set()
        self.out_types, self.out_protos, self.out_vars, self.out_funcs = [], [], [], []


class Translator:
    def __init__(self):
        self.packages = {}      # import path -> Package
        self.pkg = None
        self.toks, self.p = [], 0
        self.imports = {}       # alias -> namespace (per file)
        self.scopes = []
        self.fname = ""
        self.no_lit = 0
        self.tmp = 0
        self.results = None     # (types, names) of the function being translated

    # ---- token helpers ---------------------------------------------------------------------
    @property
    def t(self):
        return self.toks[self.p]

    def peek(self, k=1):
        return self.toks[min(self.p + k, len(self.toks) - 1)]

    def err(self, msg):
        raise SyntaxError(f"{self.fname}:{self.t.line}: {msg} (at {self.t!r})")

    def is_op(self, v):
        return self.t.kind == "op" and self.t.val == v

    def is_kw(self, v):
        return self.t.kind == "kw" and self.t.val == v

    def accept(self, v):
        if self.t.kind in ("op", "kw") and self.t.val == v:
            self.p += 1
            return True
        return False

    def expect(self, v):
        if not self.accept(v):
            self.err(f"expected {v!r}")

    def ident(self):
        if self.t.kind != "ident":
            self.err("expected identifier")
        v = self.t.val
        self.p += 1
        return v

    def skip_semis(self):
        while self.is_op(";"):
            self.p += 1

    def fresh(self, base="_t"):
        self.tmp += 1
        return f"{base}{self.tmp}"

    @staticmethod
    def cname(name):
        return name + "_" if name in CPP_RESERVED else name

    # ---- scopes ----------------------------------------------------------------------------
    def push(self):
        self.scopes.append(set())

    def pop(self):
        self.scopes.pop()

    def declare(self, name):
        if name != "_":
            self.scopes[-1].add(name)

    def is_local(self, name):
        return any(name in s for s in self.scopes)

    # ---- types -----------------------------------------------------------------------------
    def named_type(self, pkg: Package, name: str) -> str:
        q = "" if pkg is self.pkg else pkg.ns + "::"
        if name in pkg.types and pkg.types[name][0] == "interface":
            return f"{q}{name}*"
        return f"{q}{name}"

    def parse_type(self) -> str:
        if self.accept("*"):
            return self.parse_type() + "*"
        if self.accept("("):
            ty = self.parse_type()
            self.expect(")")
            return ty
        if self.accept("["):
            if self.accept("]"):
                return f"GoSlice<{self.parse_type()}>"
            n = self.parse_expr()
            self.expect("]")
            if n.const is None:
                self.err("array length must be a constant")
            return f"GoArray<{self.parse_type()}, {int(n.const[0])}>"
        if self.accept("map"):
            self.expect("[")
            k = self.parse_type()
            self.expect("]")
            return f"GoMap<{k}, {self.parse_type()}>"
        if self.is_kw("chan") or self.is_kw("func") or self.is_kw("struct") or self.is_kw("interface"):
            self.err("type literal not supported here")
        name = self.ident()
        if self.is_op(".") and name in self.imports and not self.is_local(name):
            self.p += 1
            tn = self.ident()
            ns = self.imports[name]
            pkg = next((p for p in self.packages.values() if p.ns == ns), None)
            if pkg is not None:
                return self.named_type(pkg, tn)
            return f"{ns}::{tn}"
        if name in BASIC and name not in self.pkg.types:
            return BASIC[name]
        return self.named_type(self.pkg, name)

    def at_type_start(self):
        """Does the current token begin a type literal that cannot be an expression ([]T, [N]T)?"""
        return self.is_op("[") or self.is_kw("map")

    # ---- expressions -----------------------------------------------------------------------
    PREC = {"||": 1, "&&": 2, "==": 3, "!=": 3, "<": 3, "<=": 3, ">": 3, ">=": 3, "+": 4, "-": 4, "|": 4, "^": 4,
            "*": 5, "/": 5, "%": 5, "<<": 5, ">>": 5, "&": 5, "&^": 5}

    def parse_expr(self, prec=1) -> Expr:
        lhs = self.parse_unary()
        while self.t.kind == "op" and self.t.val in self.PREC and self.PREC[self.t.val] >= prec:
            op = self.t.val
            self.p += 1
            rhs = self.parse_expr(self.PREC[op] + 1)
            lhs = self.binary(op, lhs, rhs)
        return lhs

    def binary(self, op, a: Expr, b: Expr) -> Expr:
        if a.const is not None and b.const is not None and op in ("+", "-", "*", "/", "%", "<<", ">>"):
            (x, kx), (y, ky) = a.const, b.const
            k = "float" if "float" in (kx, ky) else "int"
            if op == "+":
                v = x + y
            elif op == "-":
                v = x - y
            elif op == "*":
                v = x * y
            elif op == "/":
                if y == 0:
                    self.err("constant division by zero")
                v = x / y if k == "float" else Fraction(int(abs(x) // abs(y)) * (1 if (x >= 0) == (y >= 0) else -1))   # Go: integer constants truncate
            elif op == "%":
                v = Fraction(int(x) - int(y) * int(Fraction(int(abs(x) // abs(y)) * (1 if (x >= 0) == (y >= 0) else -1))))
            elif op == "<<":
                v = x * (2 ** int(y))
            else:
                v = Fraction(int(x) >> int(y))
            c = (Fraction(v), k)
            return Expr(const_cpp(c), c)
        if op == "&^":
            return Expr(f"({a.cpp} & ~({b.cpp}))")
        # every binary expression keeps Go's parse tree through explicit parentheses
        return Expr(f"({a.cpp} {op} {b.cpp})")

    def parse_unary(self) -> Expr:
        if self.t.kind == "op" and self.t.val in ("+", "-", "!", "^", "*", "&"):
            op = self.t.val
            self.p += 1
            if op == "&" and self.looks_like_composite():
                e = self.parse_unary()
                return Expr(f"(new {e.cpp})")
            e = self.parse_unary()
            if e.const is not None and op in ("+", "-"):
                c = (e.const[0] if op == "+" else -e.const[0], e.const[1])
                return Expr(const_cpp(c), c)
            if op == "^":
                return Expr(f"(~{e.cpp})")
            if op == "*":
                return Expr(f"(*{e.cpp})")
            if op == "&":
                return Expr(f"(&{e.cpp})")
            return Expr(f"({op}{e.cpp})")
        return self.parse_primary()

    def looks_like_composite(self):
        """&T{...}: scan ahead for an identifier path followed by '{'."""
        k = 0
        if self.peek(k).kind != "ident" and not (self.peek(k).kind == "op" and self.peek(k).val == "["):
            return False
        while self.peek(k).kind == "ident" or (self.peek(k).kind == "op" and self.peek(k).val in (".", "[", "]", "*")) or self.peek(k).kind == "int":
            k += 1
        return self.peek(k).kind == "op" and self.peek(k).val == "{" and self.no_lit == 0

    def type_named(self, name):
        """(pkg, name) if `name` is a type of the current package (and not shadowed)."""
        if self.is_local(name):
            return None
        if name in self.pkg.types:
            return self.pkg
        return None

    def parse_operand(self) -> Expr:
        t = self.t
        if t.kind == "int":
            self.p += 1
            v = int(t.val, 16) if t.val.lower().startswith("0x") else (int(t.val, 8) if len(t.val) > 1 and t.val[0] == "0" and t.val.isdigit() else int(t.val))
            c = (Fraction(v), "int")
            return Expr(const_cpp(c), c)
        if t.kind == "float":
            self.p += 1
            c = (Fraction(t.val), "float")
            return Expr(const_cpp(c), c)
        if t.kind == "string":
            self.p += 1
            return Expr(f"std::string({t.val})")
        if t.kind == "char":
            self.p += 1
            return Expr(t.val)
        if self.accept("("):
            self.no_lit, saved = 0, self.no_lit
            e = self.parse_expr()
            self.no_lit = saved
            self.expect(")")
            if e.kind == "type":
                return e
            return Expr(e.cpp if e.const is not None else f"({e.cpp})", e.const)
        if self.at_type_start():
            ty = self.parse_type()
            return Expr(ty, kind="type")
        if self.is_kw("func"):
            self.err("function literals are not supported")
        name = self.ident()
        if self.is_local(name):
            return Expr(self.cname(name))
        if name in self.imports:
            return Expr(self.imports[name], kind="package", pkg=self.imports[name])
        if name in ("true", "false"):
            return Expr(name)
        if name == "nil":
            return Expr("nullptr")
        if name in ("len", "cap", "append", "new", "panic", "make", "copy"):
            return Expr(name, kind="builtin")
        if name in self.pkg.consts:
            c = self.pkg.consts[name]
            return Expr(const_cpp(c), c)
        if name in self.pkg.types:
            return Expr(self.named_type(self.pkg, name), kind="type", tname=name)
        if name in BASIC:
            return Expr(BASIC[name], kind="type")
        return Expr(self.cname(name))

    def parse_call_args(self):
        args = []
        self.no_lit, saved = 0, self.no_lit
        while not self.is_op(")"):
            if self.at_type_start() or (self.t.kind == "ident" and self.t.val in BASIC and not self.is_local(self.t.val) and self.peek().val != "("):
                args.append(Expr(self.parse_type(), kind="type"))
            else:
                args.append(self.parse_expr())
            if not self.accept(","):
                break
            self.skip_semis()
        self.no_lit = saved
        self.expect(")")
        return args

    def parse_composite(self, ty: str) -> Expr:
        self.expect("{")
        self.no_lit, saved = 0, self.no_lit
        elems, keyed = [], False
        self.skip_semis()
        while not self.is_op("}"):
            if self.is_op("{"):          # element of a composite element type: the type is elided — not needed by the sources
                self.err("nested elided composite literals are not supported")
            if self.t.kind == "ident" and self.peek().kind == "op" and self.peek().val == ":":
                fld = self.ident()
                self.p += 1
                elems.append((fld, self.parse_expr().cpp))
                keyed = True
            else:
                elems.append((None, self.parse_expr().cpp))
            if not self.accept(","):
                self.skip_semis()
                break
            self.skip_semis()
        self.no_lit = saved
        self.expect("}")
        if keyed:
            tmp = self.fresh("_lit")
            body = "".join(f" {tmp}.{self.cname(f)} = {v};" for f, v in elems)
            return Expr(f"([&]{{ {ty} {tmp}{{}};{body} return {tmp}; }}())")
        if ty.startswith("GoSlice<"):
            return Expr(f"{ty}::of({{{', '.join(v for _, v in elems)}}})")
        return Expr(f"{ty}{{{', '.join(v for _, v in elems)}}}")

    def parse_primary(self) -> Expr:
        e = self.parse_operand()
        while True:
            if self.is_op("."):
                self.p += 1
                if self.accept("("):             # type assertion x.(T) / x.(type)
                    if self.accept("type"):
                        self.expect(")")
                        e = Expr(e.cpp, kind="typeswitch")
                    else:
                        ty = self.parse_type()
                        self.expect(")")
                        e = Expr(f"dynamic_cast<{ty}>({e.cpp})", kind="assert")
                    continue
                name = self.ident()
                if e.kind == "package":
                    pkg = next((p for p in self.packages.values() if p.ns == e.pkg), None)
                    if pkg is not None and name in pkg.consts:
                        c = pkg.consts[name]
                        e = Expr(const_cpp(c), c)
                    elif pkg is not None and name in pkg.types:
                        saved = self.pkg
                        e = Expr(self.named_type(pkg, name), kind="type", tname=name)
                        self.pkg = saved
                    else:
                        e = Expr(f"{e.pkg}::{self.cname(name)}")
                else:
                    e = Expr(f"D({e.cpp}).{self.cname(name)}")
            elif self.is_op("["):
                self.p += 1
                self.no_lit, saved = 0, self.no_lit
                lo = None if self.is_op(":") else self.parse_expr()
                if self.accept(":"):
                    hi = None if self.is_op("]") else self.parse_expr()
                    self.no_lit = saved
                    self.expect("]")
                    e = Expr(f"goslice({e.cpp}, {lo.cpp if lo else '0'}, {hi.cpp if hi else '-1'})")
                else:
                    self.no_lit = saved
                    self.expect("]")
                    e = Expr(f"D({e.cpp})[{lo.cpp}]")
            elif self.is_op("("):
                self.p += 1
                args = self.parse_call_args()
                if e.kind == "type":
                    if len(args) != 1:
                        self.err("conversion takes one argument")
                    e = Expr(f"goconv<{e.cpp}>({args[0].cpp})")
                elif e.kind == "builtin":
                    a = [x.cpp for x in args]
                    if e.cpp == "len":
                        e = Expr(f"golen({a[0]})")
                    elif e.cpp == "cap":
                        e = Expr(f"gocap({a[0]})")
                    elif e.cpp == "append":
                        e = Expr(f"goappend({', '.join(a)})")
                    elif e.cpp == "new":
                        e = Expr(f"(new {a[0]}())")
                    elif e.cpp == "panic":
                        e = Expr(f"gopanic({a[0]})")
                    elif e.cpp == "make":
                        e = Expr(f"{a[0]}::make({', '.join(a[1:])})")
                    else:
                        self.err(f"builtin {e.cpp} not supported")
                else:
                    e = Expr(f"{e.cpp}({', '.join(x.cpp for x in args)})")
            elif self.is_op("{") and e.kind == "type" and self.no_lit == 0:
                e = self.parse_composite(e.cpp)
            else:
                return e

    # ---- statements ------------------------------------------------------------------------
    def parse_block(self, out, ind, new_scope=True):
        self.expect("{")
        if new_scope:
            self.push()
        self.skip_semis()
        while not self.is_op("}"):
            self.parse_stmt(out, ind)
            self.skip_semis()
        self.expect("}")
        if new_scope:
            self.pop()

    def parse_expr_list(self):
        es = [self.parse_expr()]
        while self.accept(","):
            es.append(self.parse_expr())
        return es

    def parse_simple(self, out, ind, as_text=False):
        """Simple statement (expression, send, inc/dec, assignment, short var decl).  Emits lines, or returns
        C++ text without the trailing ';' when as_text (for-loop headers)."""
        lhs = self.parse_expr_list()
        pad = " " * ind
        res = []

        def emit(s):
            res.append(s)

        if self.is_op(":="):
            self.p += 1
            names = [x.cpp for x in lhs]
            if self.is_kw("range"):
                return ("range", names)
            rhs = self.parse_expr_list()
            # names introduced by := keep Go names; already-declared ones (same scope) are assigned
            raw = [n[:-1] if n.endswith("_") and n[:-1] in CPP_RESERVED else n for n in names]
            if len(rhs) == 1 and len(names) > 1:
                r = rhs[0]
                if r.kind == "assert":       # v, ok := x.(T)
                    emit(self.bind(raw[0], names[0], r.cpp))
                    emit(self.bind(raw[1], names[1], f"({names[0]} != nullptr)"))
                else:
                    tmp = self.fresh()
                    emit(f"auto {tmp} = {r.cpp}")
                    for k, (rn, n) in enumerate(zip(raw, names)):
                        if rn != "_":
                            emit(self.bind(rn, n, f"std::get<{k}>({tmp})"))
            elif len(rhs) == len(names):
                if len(names) == 1:
                    emit(self.bind(raw[0], names[0], self.init_value(rhs[0])))
                else:
                    tmps = []
                    for r in rhs:
                        tmp = self.fresh()
                        emit(f"auto {tmp} = {self.init_value(r)}")
                        tmps.append(tmp)
                    for rn, n, tmp in zip(raw, names, tmps):
                        if rn != "_":
                            emit(self.bind(rn, n, tmp))
            else:
                self.err("assignment count mismatch")
        elif self.is_op("="):
            self.p += 1
            rhs = self.parse_expr_list()
            if len(lhs) == 1 and len(rhs) == 1:
                emit(f"{lhs[0].cpp} = {rhs[0].cpp}")
            elif len(rhs) == 1:
                r = rhs[0]
                if r.kind == "assert":
                    emit(f"{lhs[0].cpp} = {r.cpp}")
                    emit(f"{lhs[1].cpp} = ({lhs[0].cpp} != nullptr)")
                else:
                    tmp = self.fresh()
                    emit(f"auto {tmp} = {r.cpp}")
                    for k, l in enumerate(lhs):
                        if l.cpp != "_":
                            emit(f"{l.cpp} = std::get<{k}>({tmp})")
            else:
                tmps = []
                for r in rhs:      # Go evaluates every right-hand side before assigning
                    tmp = self.fresh()
                    emit(f"auto {tmp} = {r.cpp}")
                    tmps.append(tmp)
                for l, tmp in zip(lhs, tmps):
                    if l.cpp != "_":
                        emit(f"{l.cpp} = {tmp}")
        elif self.t.kind == "op" and self.t.val in ("+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>="):
            op = self.t.val
            self.p += 1
            r = self.parse_expr()
            # x op= y is x = x op y with one rounding: identical for IEEE arithmetic; keep the compound form
            emit(f"{lhs[0].cpp} {op} {r.cpp}")
        elif self.is_op("&^="):
            self.p += 1
            r = self.parse_expr()
            emit(f"{lhs[0].cpp} &= ~({r.cpp})")
        elif self.is_op("++") or self.is_op("--"):
            op = self.t.val
            self.p += 1
            emit(f"{lhs[0].cpp}{op}")
        else:
            emit(lhs[0].cpp)
        if as_text:
            return ", ".join(res) if len(res) > 1 else (res[0] if res else "")
        for s in res:
            out.append(f"{pad}{s};")
        return None

    def init_value(self, e: Expr) -> str:
        """Initialiser of `x := e`: an untyped constant takes Go's default type (int / float64)."""
        if e.const is not None:
            return f"(long long){e.cpp}" if e.const[1] == "int" else f"(double){e.cpp}"
        return e.cpp

    def bind(self, raw, name, value):
        if raw == "_":
            return f"(void)({value})"
        if raw in self.scopes[-1]:
            return f"{name} = {value}"
        self.declare(raw)
        return f"auto {self.cname(raw)} = {value}"

    def parse_var_decl(self, out, ind, top=False):
        """var name[, name] [Type] [= expr[, expr]]  (one spec)."""
        pad = " " * ind
        names = [self.ident()]
        while self.accept(","):
            names.append(self.ident())
        ty = None
        if not self.is_op("=") and not self.is_op(";"):
            ty = self.parse_type()
        vals = None
        if self.accept("="):
            vals = self.parse_expr_list()
        pre = "static " if top else ""
        for k, n in enumerate(names):
            if top:
                self.pkg.globals.add(n)
            else:
                self.declare(n)
            cn = self.cname(n)
            if vals is None:
                out.append(f"{pad}{pre}{ty} {cn}{{}};")
            elif len(vals) == len(names):
                out.append(f"{pad}{pre}{ty or 'auto'} {cn} = {self.init_value(vals[k]) if ty is None else vals[k].cpp};")
            else:
                self.err("var with a multi-value initialiser is not supported")

    def parse_const_decl(self, out, ind, top=False):
        """const name [Type] = expr (one spec).  Untyped constants are folded exactly; typed ones become C++ consts
        (their arithmetic then happens in the type, as Go's does after the conversion)."""
        pad = " " * ind
        name = self.ident()
        ty = None
        if not self.is_op("="):
            ty = self.parse_type()
        self.expect("=")
        e = self.parse_expr()
        if ty is None and e.const is not None:
            if top:
                self.pkg.consts[name] = e.const
            else:
                # a local untyped constant: shadow through a scoped table
                self.local_consts[-1][name] = e.const
                self.pkg.consts[name] = e.const        # (function-local names do not collide in the sources)
            return
        if top:
            self.pkg.globals.add(name)
        else:
            self.declare(name)
        out.append(f"{pad}{'static ' if top else ''}const {ty or 'auto'} {self.cname(name)} = {e.cpp};")

    def parse_stmt(self, out, ind):
        pad = " " * ind
        if self.is_op("{"):
            out.append(pad + "{")
            self.parse_block(out, ind + 4)
            out.append(pad + "}")
        elif self.accept("var"):
            if self.accept("("):
                self.skip_semis()
                while not self.is_op(")"):
                    self.parse_var_decl(out, ind)
                    self.skip_semis()
                self.expect(")")
            else:
                self.parse_var_decl(out, ind)
        elif self.accept("const"):
            if self.accept("("):
                self.skip_semis()
                while not self.is_op(")"):
                    self.parse_const_decl(out, ind)
                    self.skip_semis()
                self.expect(")")
            else:
                self.parse_const_decl(out, ind)
        elif self.accept("return"):
            if self.is_op(";") or self.is_op("}"):
                if self.results and self.results[1]:
                    names = [self.cname(n) for n in self.results[1]]
                    out.append(f"{pad}return {names[0] if len(names) == 1 else 'std::make_tuple(' + ', '.join(names) + ')'};")
                else:
                    out.append(pad + "return;")
            else:
                es = self.parse_expr_list()
                if len(es) == 1:
                    out.append(f"{pad}return {es[0].cpp};")
                else:
                    tys = self.results[0]
                    out.append(f"{pad}return std::tuple<{', '.join(tys)}>({', '.join(e.cpp for e in es)});")
        elif self.accept("break"):
            out.append(pad + "break;")
        elif self.accept("continue"):
            out.append(pad + "continue;")
        elif self.accept("if"):
            self.parse_if(out, ind)
        elif self.accept("for"):
            self.parse_for(out, ind)
        elif self.accept("switch"):
            self.parse_switch(out, ind)
        elif self.t.kind == "kw":
            self.err(f"statement {self.t.val!r} is not supported")
        else:
            self.parse_simple(out, ind)

    def parse_if(self, out, ind):
        pad = " " * ind
        self.push()
        self.no_lit += 1
        opened = False
        start = self.p
        # optional init statement: look for ';' before '{' at depth 0
        depth, k, has_init = 0, self.p, False
        while True:
            tk = self.toks[k]
            if tk.kind == "op" and tk.val in ("(", "["):
                depth += 1
            elif tk.kind == "op" and tk.val in (")", "]"):
                depth -= 1
            elif tk.kind == "op" and tk.val == "{" and depth == 0:
                break
            elif tk.kind == "op" and tk.val == ";" and depth == 0:
                has_init = True
                break
            k += 1
        self.p = start
        if has_init:
            out.append(pad + "{")
            opened = True
            self.parse_simple(out, ind + 4)
            self.expect(";")
            pad2, ind2 = pad + "    ", ind + 4
        else:
            pad2, ind2 = pad, ind
        cond = self.parse_expr()
        self.no_lit -= 1
        out.append(f"{pad2}if ({cond.cpp}) {{")
        self.parse_block(out, ind2 + 4)
        if self.accept("else"):
            if self.accept("if"):
                out.append(f"{pad2}}} else {{")
                self.parse_if(out, ind2 + 4)
                out.append(f"{pad2}}}")
            else:
                out.append(f"{pad2}}} else {{")
                self.parse_block(out, ind2 + 4)
                out.append(f"{pad2}}}")
        else:
            out.append(f"{pad2}}}")
        if opened:
            out.append(pad + "}")
        self.pop()

    def parse_for(self, out, ind):
        pad = " " * ind
        self.push()
        self.no_lit += 1
        if self.is_op("{"):                       # for { }
            self.no_lit -= 1
            out.append(f"{pad}for (;;) {{")
            self.parse_block(out, ind + 4)
            out.append(pad + "}")
            self.pop()
            return
        if self.is_kw("range"):                   # for range x
            self.p += 1
            x = self.parse_expr()
            self.no_lit -= 1
            r, i = self.fresh("_r"), self.fresh("_i")
            out.append(f"{pad}{{ auto {r} = {x.cpp}; for (long long {i} = 0; {i} < golen({r}); {i}++) {{")
            self.parse_block(out, ind + 4)
            out.append(pad + "} }")
            self.pop()
            return
        init = None
        if not self.is_op(";"):
            init = self.parse_simple(out, ind, as_text=True)
        if isinstance(init, tuple):               # for k, v := range x
            names = init[1]
            self.p += 1                           # 'range'
            x = self.parse_expr()
            self.no_lit -= 1
            r, i = self.fresh("_r"), self.fresh("_i")
            out.append(f"{pad}{{ auto {r} = {x.cpp}; for (long long {i} = 0; {i} < golen({r}); {i}++) {{")
            if names[0] != "_":
                self.declare(names[0])
                out.append(f"{pad}    long long {names[0]} = {i};")
            if len(names) > 1 and names[1] != "_":
                raw = names[1][:-1] if names[1].endswith("_") and names[1][:-1] in CPP_RESERVED else names[1]
                self.declare(raw)
                out.append(f"{pad}    auto {names[1]} = {r}[{i}];")
            self.parse_block(out, ind + 4)
            out.append(pad + "} }")
            self.pop()
            return
        if self.is_op("{"):                       # for cond { }
            self.no_lit -= 1
            out.append(f"{pad}while ({init}) {{")
            self.parse_block(out, ind + 4)
            out.append(pad + "}")
            self.pop()
            return
        self.expect(";")
        cond = "" if self.is_op(";") else self.parse_expr().cpp
        self.expect(";")
        post = "" if self.is_op("{") else self.parse_simple(out, ind, as_text=True)
        self.no_lit -= 1
        out.append(f"{pad}for ({init or ''}; {cond}; {post}) {{")
        self.parse_block(out, ind + 4)
        out.append(pad + "}")
        self.pop()

    def parse_switch(self, out, ind):
        pad = " " * ind
        self.push()
        self.no_lit += 1
        tag = None
        if not self.is_op("{"):
            tag = self.parse_expr()
        self.no_lit -= 1
        self.expect("{")
        self.skip_semis()
        tv = self.fresh("_sw")
        out.append(pad + "do {")       # `break` inside a Go switch leaves the switch: a do { } while (0) gives it that meaning
        if tag is not None:
            out.append(f"{pad}    auto {tv} = {tag.cpp};")
        first = True
        default_body = None
        while not self.is_op("}"):
            body = []
            if self.accept("default"):
                self.expect(":")
                self.push()
                while not (self.is_kw("case") or self.is_kw("default") or self.is_op("}")):
                    self.parse_stmt(body, ind + 8)
                    self.skip_semis()
                self.pop()
                default_body = body
                continue
            self.expect("case")
            conds = []
            while True:
                if tag is not None and tag.kind == "typeswitch":
                    ty = self.parse_type()
                    conds.append(f"dynamic_cast<{ty}>({tv}) != nullptr")
                else:
                    e = self.parse_expr()
                    conds.append(e.cpp if tag is None else f"({tv} == {e.cpp})")
                if not self.accept(","):
                    break
            self.expect(":")
            self.push()
            while not (self.is_kw("case") or self.is_kw("default") or self.is_op("}")):
                self.parse_stmt(body, ind + 8)
                self.skip_semis()
            self.pop()
            out.append(f"{pad}    {'if' if first else 'else if'} ({' || '.join(conds)}) {{")
            out.extend(body)
            out.append(f"{pad}    }}")
            first = False
        self.expect("}")
        if default_body is not None:
            out.append(f"{pad}    {'{' if first else 'else {'}")
            out.extend(default_body)
            out.append(f"{pad}    }}")
        out.append(pad + "} while (0);")
        self.pop()

    # ---- declarations ----------------------------------------------------------------------
    def parse_params(self):
        """(a, b T, c *U) or (T, U) -> list of (name | None, type)."""
        self.expect("(")
        groups, pending = [], []
        while not self.is_op(")"):
            # try "name Type" / "name, name Type"; fall back to a bare type
            if self.t.kind == "ident" and self.peek().kind == "op" and self.peek().val == ",":
                pending.append(self.ident())
                self.p += 1
                continue
            if self.t.kind == "ident" and not (self.peek().kind == "op" and self.peek().val in (")", ".")):
                name = self.ident()
                ty = self.parse_type()
                for n in pending + [name]:
                    groups.append((n, ty))
                pending = []
            else:
                for n in pending:          # they were types, not names
                    groups.append((None, self.type_from_name(n)))
                pending = []
                groups.append((None, self.parse_type()))
            if not self.accept(","):
                break
        for n in pending:
            groups.append((None, self.type_from_name(n)))
        self.expect(")")
        return groups

    def type_from_name(self, name):
        if name in BASIC and name not in self.pkg.types:
            return BASIC[name]
        return self.named_type(self.pkg, name)

    def parse_signature(self):
        params = self.parse_params()
        rtypes, rnames = [], []
        if self.is_op("("):
            res = self.parse_params()
            rtypes = [t for _, t in res]
            rnames = [n for n, _ in res if n]
        elif not self.is_op("{") and not self.is_op(";") and not self.is_op("}"):
            rtypes = [self.parse_type()]
        return params, rtypes, rnames

    @staticmethod
    def ret_type(rtypes):
        if not rtypes:
            return "void"
        return rtypes[0] if len(rtypes) == 1 else f"std::tuple<{', '.join(rtypes)}>"

    def parse_func(self):
        recv = None
        if self.is_op("("):
            r = self.parse_params()
            recv = r[0]
        name = self.ident()
        params, rtypes, rnames = self.parse_signature()
        plist = ", ".join(f"{t} {self.cname(n) if n and n != '_' else self.fresh('_p')}" for n, t in params)
        rt = self.ret_type(rtypes)
        body = []
        self.push()
        self.local_consts = [{}]
        for n, _ in params:
            if n:
                self.declare(n)
        self.results = (rtypes, rnames)
        saved_consts = dict(self.pkg.consts)
        if recv is not None:
            rname, rtype = recv
            ptr = rtype.endswith("*")
            tname = rtype.rstrip("*")
            if rname and rname != "_":
                self.declare(rname)
                body.append(f"    auto {self.cname(rname)} = {'this' if ptr else '*this'};")
        for n, t in zip(rnames, rtypes):
            self.declare(n)
            body.append(f"    {t} {self.cname(n)}{{}};")
        if self.is_op("{"):
            self.parse_block(body, 4, new_scope=False)
        self.pop()
        self.pkg.consts = saved_consts
        self.results = None
        cn = self.cname(name)
        if recv is not None:
            q = "" if ptr else " const"
            self.pkg.methods.setdefault(tname, []).append((name, f"{rt} {cn}({plist}){q}", ptr, params, rtypes))
            self.pkg.out_funcs.append(f"{rt} {tname}::{cn}({plist}){q} {{\n" + "\n".join(body) + "\n}\n")
        else:
            self.pkg.funcs.add(name)
            self.pkg.out_protos.append(f"{rt} {cn}({plist});")
            self.pkg.out_funcs.append(f"{rt} {cn}({plist}) {{\n" + "\n".join(body) + "\n}\n")

    def parse_type_decl(self):
        name = self.ident()
        if self.accept("struct"):
            self.expect("{")
            self.skip_semis()
            fields = []
            while not self.is_op("}"):
                names = [self.ident()]
                while self.accept(","):
                    names.append(self.ident())
                ty = self.parse_type()
                if self.t.kind == "string":
                    self.p += 1     # field tag
                for n in names:
                    fields.append((n, ty))
                self.skip_semis()
            self.expect("}")
            self.pkg.types[name] = ("struct", fields)
        elif self.accept("interface"):
            self.expect("{")
            self.skip_semis()
            methods = []
            while not self.is_op("}"):
                mn = self.ident()
                params, rtypes, _ = self.parse_signature()
                methods.append((mn, params, rtypes))
                self.skip_semis()
            self.expect("}")
            self.pkg.types[name] = ("interface", methods)
        elif self.is_op("["):
            self.p += 1
            n = self.parse_expr()
            self.expect("]")
            self.pkg.types[name] = ("array", self.parse_type(), int(n.const[0]))
        else:
            under = self.parse_type()
            # a named type whose underlying type is a named array type shares its layout (type Quat Vector4)
            base = under.split("::")[-1]
            if base in self.pkg.types and self.pkg.types[base][0] == "array":
                self.pkg.types[name] = self.pkg.types[base]
            else:
                self.pkg.types[name] = ("named", under)

    def prescan_types(self, toks):
        """Register the type names of a file before translating bodies (Go has no declaration order)."""
        for k, t in enumerate(toks):
            if t.kind == "kw" and t.val == "type" and toks[k + 1].kind == "ident" and (k == 0 or toks[k - 1].val in (";", "(")):
                nm, nx = toks[k + 1].val, toks[k + 2]
                kind = "interface" if nx.kind == "kw" and nx.val == "interface" else ("struct" if nx.kind == "kw" and nx.val == "struct" else "named")
                self.pkg.types.setdefault(nm, (kind, [] if kind != "named" else "?"))

    def translate_file(self, path):
        self.fname = path
        self.toks, self.p = lex(read_source(path), path), 0
        self.imports = {}
        self.scopes = [set()]
        self.skip_semis()
        self.expect("package")
        self.ident()
        self.skip_semis()
        while self.accept("import"):
            specs = []
            if self.accept("("):
                self.skip_semis()
                while not self.is_op(")"):
                    alias = self.ident() if self.t.kind == "ident" else None
                    specs.append((alias, self.t.val.strip('"')))
                    self.p += 1
                    self.skip_semis()
                self.expect(")")
            else:
                alias = self.ident() if self.t.kind == "ident" else None
                specs.append((alias, self.t.val.strip('"')))
                self.p += 1
            for alias, path_ in specs:
                if path_ in self.packages:
                    ns = self.packages[path_].ns
                elif path_ in STD_NS:
                    ns = STD_NS[path_]
                else:
                    raise SyntaxError(f"{path}: import of {path_} is not available to the translator")
                self.imports[alias or path_.split("/")[-1]] = ns
            self.skip_semis()
        while self.t.kind != "eof":
            if self.accept("func"):
                self.parse_func()
            elif self.accept("type"):
                if self.accept("("):
                    self.skip_semis()
                    while not self.is_op(")"):
                        self.parse_type_decl()
                        self.skip_semis()
                    self.expect(")")
                else:
                    self.parse_type_decl()
            elif self.accept("var"):
                if self.accept("("):
                    self.skip_semis()
                    while not self.is_op(")"):
                        self.parse_var_decl(self.pkg.out_vars, 0, top=True)
                        self.skip_semis()
                    self.expect(")")
                else:
                    self.parse_var_decl(self.pkg.out_vars, 0, top=True)
            elif self.accept("const"):
                self.local_consts = [{}]
                if self.accept("("):
                    self.skip_semis()
                    while not self.is_op(")"):
                        self.parse_const_decl(self.pkg.out_vars, 0, top=True)
                        self.skip_semis()
                    self.expect(")")
                else:
                    self.parse_const_decl(self.pkg.out_vars, 0, top=True)
            else:
                self.err("unexpected top-level token")
            self.skip_semis()

    def translate_package(self, path, files):
        ns = "pkg_" + re.sub(r"\W", "_", path.split("/")[-1]) if path != "main" else "pkg_main"
        if any(p.ns == ns for p in self.packages.values()):
            ns += str(len(self.packages))
        self.pkg = Package(path, ns)
        self.packages[path] = self.pkg
        for f in files:
            self.fname = f
            self.prescan_types(lex(read_source(f), f))
        for f in files:
            self.translate_file(f)
        return self.emit_package()

    def emit_package(self):
        pkg = self.pkg
        o = [f"// ===== package {pkg.path} =====", f"namespace {pkg.ns} {{"]
        structs = [n for n, t in pkg.types.items() if t[0] in ("struct", "array", "interface")]
        for n in structs:
            o.append(f"struct {n};")
        for n, t in pkg.types.items():
            if t[0] == "named":
                o.append(f"typedef {t[1]} {n};")
        # interfaces first (abstract classes), then value types in declaration order
        impl = {}
        for n, t in pkg.types.items():
            if t[0] != "interface":
                continue
            o.append(f"struct {n} {{")
            o.append(f"    virtual ~{n}() {{}}")
            for mn, params, rtypes in t[1]:
                plist = ", ".join(f"{ty} {self.cname(pn) if pn else 'p' + str(k)}" for k, (pn, ty) in enumerate(params))
                o.append(f"    virtual {self.ret_type(rtypes)} {self.cname(mn)}({plist}) = 0;")
            o.append("};")
            want = {mn for mn, _, _ in t[1]}
            for sn, st in pkg.types.items():
                if st[0] == "struct" and want <= {m[0] for m in pkg.methods.get(sn, []) if m[2]}:
                    impl.setdefault(sn, []).append(n)
        for n, t in pkg.types.items():
            if t[0] == "array":
                o.append(f"struct {n} {{")
                o.append(f"    {t[1]} e[{t[2]}];")
                o.append(f"    {t[1]}& operator[](long long i) {{ return e[i]; }}")
                o.append(f"    const {t[1]}& operator[](long long i) const {{ return e[i]; }}")
                for m in pkg.methods.get(n, []):
                    o.append(f"    {m[1]};")
                o.append("};")
                o.append(f"inline long long golen(const {n}&) {{ return {t[2]}; }}")
            elif t[0] == "struct":
                bases = impl.get(n, [])
                o.append(f"struct {n}{' : ' + ', '.join(bases) if bases else ''} {{")
                for fn, fty in t[1]:
                    o.append(f"    {fty} {self.cname(fn)}{{}};")
                for m in pkg.methods.get(n, []):
                    o.append(f"    {m[1]}{' override' if any(m[0] in {x[0] for x in pkg.types[b][1]} for b in bases) else ''};")
                o.append("};")
        o.extend(pkg.out_protos)
        o.extend(pkg.out_vars)
        o.extend(pkg.out_funcs)
        o.append(f"}}  // namespace {pkg.ns}\n")
        return "\n".join(o)


PRELUDE = r'''// GENERATED by oracle/go2cpp.py from the reference's Go sources — do not edit, do not commit.
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>
#include <type_traits>

// ---- Go value / pointer plumbing ----------------------------------------------------------------
template <class T> inline T &D(T *p) { return *p; }                 // Go's implicit dereference in selectors and indexing
template <class T> inline T &D(T &r) { return r; }
template <class T> inline const T &D(const T &r) { return r; }
template <class T, long long N> struct GoArray {
    T e[N > 0 ? N : 1]{};
    T &operator[](long long i) { return e[i]; }
    const T &operator[](long long i) const { return e[i]; }
};
template <class T, long long N> inline long long golen(const GoArray<T, N> &) { return N; }
// slice: shared backing store + length (copying a slice shares the store; append grows it like Go's runtime when len == cap)
template <class T> struct GoSlice {
    std::shared_ptr<std::vector<T>> store;
    long long off = 0, len = 0;
    GoSlice() {}
    GoSlice(std::nullptr_t) {}
    T &operator[](long long i) const { return (*store)[off + i]; }
    bool isnil() const { return !store; }
    static GoSlice of(std::initializer_list<T> l) { GoSlice s; s.store = std::make_shared<std::vector<T>>(l); s.len = (long long)l.size(); return s; }
    static GoSlice make(long long n, long long cap = 0) { GoSlice s; s.store = std::make_shared<std::vector<T>>(); s.store->reserve(cap > n ? cap : n); s.store->resize(n); s.len = n; return s; }
};
template <class K, class V> struct GoMap {
    std::shared_ptr<std::unordered_map<K, V>> m;
    static GoMap make(long long = 0) { GoMap r; r.m = std::make_shared<std::unordered_map<K, V>>(); return r; }
    V &operator[](const K &k) const { return (*m)[k]; }
};
template <class K, class V> inline long long golen(const GoMap<K, V> &m) { return m.m ? (long long)m.m->size() : 0; }
template <class T> inline bool operator==(const GoSlice<T> &s, std::nullptr_t) { return s.isnil(); }
template <class T> inline bool operator!=(const GoSlice<T> &s, std::nullptr_t) { return !s.isnil(); }
template <class T> inline long long golen(const GoSlice<T> &s) { return s.len; }
template <class T> inline long long gocap(const GoSlice<T> &s) { return s.store ? (long long)s.store->capacity() - s.off : 0; }
inline long long golen(const std::string &s) { return (long long)s.size(); }
template <class T, class U> inline GoSlice<T> goappend(GoSlice<T> s, U v) {
    if (!s.store) s.store = std::make_shared<std::vector<T>>();
    if ((long long)s.store->size() != s.off + s.len) {          // appending in the middle of a shared store: copy (Go would overwrite within cap;
        auto n = std::make_shared<std::vector<T>>(s.store->begin() + s.off, s.store->begin() + s.off + s.len);   // the sources never rely on that)
        s.store = n; s.off = 0;
    }
    s.store->push_back((T)v);
    s.len++;
    return s;
}
template <class T> inline GoSlice<T> goslice(GoSlice<T> s, long long lo, long long hi) { if (hi < 0) hi = s.len; s.off += lo; s.len = hi - lo; return s; }
template <class To, class From> inline To goconv(const From &v) {
    if constexpr (std::is_arithmetic<To>::value || std::is_pointer<To>::value) return (To)v;
    else if constexpr (std::is_same<To, From>::value) return v;
    else { static_assert(sizeof(To) == sizeof(From), "conversion between named types of identical layout only"); To t; std::memcpy((void *)&t, (const void *)&v, sizeof(To)); return t; }
}
[[noreturn]] inline void gopanic(const std::string &m) { std::fprintf(stderr, "panic: %s\n", m.c_str()); std::abort(); }

// ---- the slices of the Go standard library the sources use ------------------------------------------
// An untyped Go float constant: takes the type of the operand it meets (see const_cpp in go2cpp.py).
struct gouf {
    double d; float f;
    constexpr gouf(double d_, float f_) : d(d_), f(f_) {}
    constexpr operator double() const { return d; }
    constexpr operator float() const { return f; }
    template <class T> constexpr T as() const { if constexpr (std::is_same<T, float>::value) return f; else return (T)d; }
};
#define GOUF_BIN(OP) \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> constexpr auto operator OP(gouf c, T x) -> decltype(x OP x) { return c.as<T>() OP x; } \
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> constexpr auto operator OP(T x, gouf c) -> decltype(x OP x) { return x OP c.as<T>(); }
GOUF_BIN(+) GOUF_BIN(-) GOUF_BIN(*) GOUF_BIN(/) GOUF_BIN(<) GOUF_BIN(>) GOUF_BIN(<=) GOUF_BIN(>=) GOUF_BIN(==) GOUF_BIN(!=)
#undef GOUF_BIN
#define GOUF_ASSIGN(OP) template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> inline T &operator OP(T &x, gouf c) { return x OP c.as<T>(); }
GOUF_ASSIGN(+=) GOUF_ASSIGN(-=) GOUF_ASSIGN(*=) GOUF_ASSIGN(/=)
#undef GOUF_ASSIGN

namespace gomath {
const double MaxFloat64 = std::numeric_limits<double>::max();
const double SmallestNonzeroFloat64 = std::numeric_limits<double>::denorm_min();
const double MaxFloat32 = std::numeric_limits<float>::max();
const double Pi = 3.14159265358979323846264338327950288419716939937510582097494459;
inline double Sqrt(double x) { return std::sqrt(x); }
inline double Abs(double x) { return std::fabs(x); }
inline double Pow(double x, double y) { return std::pow(x, y); }   // NOT Go's Pow: <= 1 ulp apart; the library takes Pow factors as inputs
inline double Sin(double x) { return std::sin(x); }
inline double Cos(double x) { return std::cos(x); }
inline double Acos(double x) { return std::acos(x); }
inline double Floor(double x) { return std::floor(x); }
inline double Inf(long long sign) { return sign >= 0 ? HUGE_VAL : -HUGE_VAL; }
inline double NaN() { return std::numeric_limits<double>::quiet_NaN(); }
inline bool IsNaN(double x) { return x != x; }
inline unsigned long long Float64bits(double x) { unsigned long long b; std::memcpy(&b, &x, 8); return b; }
inline double Float64frombits(unsigned long long b) { double x; std::memcpy(&x, &b, 8); return x; }
}  // namespace gomath
namespace goos {
static GoSlice<std::string> Args;
static FILE *Stdout = stdout, *Stderr = stderr;
[[noreturn]] inline void Exit(long long c) { std::exit((int)c); }
}  // namespace goos
namespace gostrconv {
inline std::tuple<long long, const char *> Atoi(const std::string &s) { char *e = nullptr; long long v = std::strtoll(s.c_str(), &e, 10); return {v, (e && *e == 0 && !s.empty()) ? nullptr : "syntax"}; }
}
namespace gotime {
struct Duration { double s; double Seconds() const { return s; } };
struct Time { std::chrono::steady_clock::time_point t; };
inline Time Now() { return Time{std::chrono::steady_clock::now()}; }
inline Duration Since(Time t0) { return Duration{std::chrono::duration<double>(std::chrono::steady_clock::now() - t0.t).count()}; }
}  // namespace gotime
namespace gotesting {   // the reference's own tests only call t.Errorf: count the failures
struct T {
    int failed = 0;
    template <class... A> void Errorf(const std::string &fmt, A...) { failed++; std::fprintf(stderr, "    FAIL: %s\n", fmt.c_str()); }
};
}  // namespace gotesting
namespace gofmt {
struct Arg { int k; long long i; unsigned long long u; double d; std::string s; };
inline Arg mk(bool v) { return Arg{3, v, 0, 0, ""}; }
inline Arg mk(double v) { return Arg{2, 0, 0, v, ""}; }
inline Arg mk(float v) { return Arg{2, 0, 0, v, ""}; }
inline Arg mk(const std::string &v) { return Arg{4, 0, 0, 0, v}; }
inline Arg mk(const char *v) { return Arg{4, 0, 0, 0, v}; }
template <class T> inline typename std::enable_if<std::is_integral<T>::value && std::is_signed<T>::value, Arg>::type mk(T v) { return Arg{0, (long long)v, 0, 0, ""}; }
template <class T> inline typename std::enable_if<std::is_integral<T>::value && !std::is_signed<T>::value && !std::is_same<T, bool>::value, Arg>::type mk(T v) { return Arg{1, 0, (unsigned long long)v, 0, ""}; }
inline void vprint(FILE *f, const std::string &fmt, const std::vector<Arg> &a) {
    size_t k = 0;
    for (size_t i = 0; i < fmt.size(); i++) {
        if (fmt[i] != '%') { std::fputc(fmt[i], f); continue; }
        size_t j = i + 1;
        std::string spec = "%";
        while (j < fmt.size() && std::strchr("0123456789.+- #", fmt[j])) spec += fmt[j++];
        const char c = fmt[j];
        i = j;
        if (c == '%') { std::fputc('%', f); continue; }
        if (k >= a.size()) { std::fputs("%!(MISSING)", f); continue; }
        const Arg &x = a[k++];
        if (c == 'd') { if (x.k == 1) std::fprintf(f, (spec + "llu").c_str(), x.u); else std::fprintf(f, (spec + "lld").c_str(), x.i); }
        else if (c == 'x') { std::fprintf(f, (spec + "llx").c_str(), x.k == 1 ? x.u : (unsigned long long)x.i); }
        else if (c == 'f' || c == 'g' || c == 'e') { std::fprintf(f, (spec + c).c_str(), x.d); }
        else if (c == 's') { std::fprintf(f, (spec + "s").c_str(), x.s.c_str()); }
        else if (c == 't' || c == 'v') {
            if (x.k == 3) std::fputs(x.i ? "true" : "false", f);
            else if (x.k == 0) std::fprintf(f, "%lld", x.i);
            else if (x.k == 1) std::fprintf(f, "%llu", x.u);
            else if (x.k == 2) std::fprintf(f, "%g", x.d);
            else std::fputs(x.s.c_str(), f);
        }
    }
}
template <class... A> inline void Printf(const std::string &fmt, A... a) { vprint(stdout, fmt, std::vector<Arg>{mk(a)...}); }
template <class... A> inline void Fprintf(FILE *f, const std::string &fmt, A... a) { vprint(f, fmt, std::vector<Arg>{mk(a)...}); }
template <class... A> inline void Println(A... a) { std::vector<Arg> v{mk(a)...}; std::string fmt; for (size_t i = 0; i < v.size(); i++) fmt += i ? " %v" : "%v"; vprint(stdout, fmt + "\n", v); }
}  // namespace gofmt
'''


# --real=float32: the reference's own switch for single precision is a one-line source edit ("type Real float64",
# math/math.go:23).  The translator applies that edit to the text it reads — nothing else changes.
REAL_EDIT = None


def read_source(path):
    text = open(path).read()
    if REAL_EDIT and path.endswith(os.path.join("math", "math.go")):
        assert "type Real float64" in text, "math.go no longer declares `type Real float64`"
        text = text.replace("type Real float64", "type Real " + REAL_EDIT, 1)
    return text


def main(argv):
    global REAL_EDIT
    out = None
    specs = []
    run_tests = None
    i = 0
    while i < len(argv):
        if argv[i] == "-o":
            out = argv[i + 1]
            i += 2
        elif argv[i].startswith("--real="):
            REAL_EDIT = argv[i].split("=", 1)[1]
            i += 1
        elif argv[i].startswith("--run-tests="):      # translate <pkg>'s *_test.go too and emit a main that runs every Test* function
            run_tests = argv[i].split("=", 1)[1]
            i += 1
        else:
            specs.append(argv[i])
            i += 1
    tr = Translator()
    parts = [PRELUDE]
    for spec in specs:
        path, _, where = spec.partition("=")
        files = []
        for w in where.split(","):
            if os.path.isdir(w):
                files += sorted(os.path.join(w, f) for f in os.listdir(w) if f.endswith(".go") and (path == run_tests or not f.endswith("_test.go")))
            else:
                files.append(w)
        parts.append(tr.translate_package(path, files))
    if run_tests:
        pkg = tr.packages[run_tests]
        tests = sorted(f for f in pkg.funcs if f.startswith("Test"))
        body = "".join(f'    {{ gotesting::T t; {pkg.ns}::{f}(&t); std::printf("%s %s\\n", t.failed ? "FAIL" : "ok  ", "{f}"); failed += t.failed ? 1 : 0; }}\n' for f in tests)
        parts.append(f'int main() {{\n    int failed = 0;\n{body}    std::printf("%d tests, %d failed\\n", {len(tests)}, failed);\n    return failed ? 1 : 0;\n}}\n')
    else:
        parts.append("int main(int argc, char **argv) {\n    for (int i = 0; i < argc; i++) goos::Args = goappend(goos::Args, std::string(argv[i]));\n"
                     "    pkg_main::main_();\n    return 0;\n}\n")
    text = "\n".join(parts)
    if out:
        with open(out, "w") as f:
            f.write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main(sys.argv[1:])
