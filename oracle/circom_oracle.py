"""ORACLE (test infrastructure, never shipped, never imported by the product).

A from-scratch CPU restatement of what the reference's witness path does:
`wasm_tester(circuit).calculateWitness(input)` followed by
`checkConstraints(w)` (/root/reference/test/automatisationTest.js:37-51).
The reference delegates that work to the un-vendored circom compiler + wasm
runtime (package.json:29-63), so this file restates the
published circom 2.1.x language semantics as a direct big-integer interpreter
of the reference's own `.circom` files: it parses them, instantiates the
template tree and evaluates every signal of one witness with Python ints,
asserting every `===` / `<==` constraint on the way.

PARITY STATUS: "parity unpinned" against the real circom wasm (neither node nor
circom exist in this image and the reference ships no compiled artefacts).
What *is* pinned: Poseidon against the circomlib vector and test/poseidon.js,
SHA-1/2 digests against hashlib, RSA against `cryptography`, BabyJubjub against
an independent affine implementation (tests/test_cpu_host.py, tests/test_cpu_compiler_vs_oracle.py,
tests/golden/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import it.

Semantics restated (SURVEY.md section 8a footnote):
  values canonical in [0,p); + - * mod p; `/` = mul by inverse, x/0 = 0;
  `\\` `%` integer quotient / remainder on canonical representatives; `**`
  modular power; < <= > >= on signed representatives (v > p/2 -> v - p);
  >> << & | ^ ~ on canonical representatives masked to 254 bits then mod p;
  && || ! on "non-zero is true"; unassigned signals read 0; a sub-component
  body runs when its last input signal has been assigned.
"""
from __future__ import annotations

import os
import re
import sys

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
HALF = P >> 1
MASK = (1 << 254) - 1

sys.setrecursionlimit(20000)


class CircomError(Exception):
    pass


class AssertFailed(CircomError):
    """Runtime `assert` / `===` failure: the wasm would throw "Assert Failed."."""


# --------------------------------------------------------------------------
# Lexer
# --------------------------------------------------------------------------
_TOKEN_RE = re.compile(
    r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<hex>0x[0-9a-fA-F]+)
  | (?P<num>\d+)
  | (?P<id>[A-Za-z_$][A-Za-z0-9_$]*)
  | (?P<str>"(?:[^"\\]|\\.)*")
  | (?P<op><==|==>|<--|-->|===|\*\*=|<<=|>>=|\*\*|\+\+|--|&&|\|\||==|!=|<=|>=|<<|>>|\+=|-=|\*=|/=|\\=|%=|&=|\|=|\^=|[-+*/\\%<>=!~&|^?:;,.(){}\[\]])
    """,
    re.X | re.S,
)

KEYWORDS = {
    "pragma", "include", "template", "function", "component", "signal", "var",
    "input", "output", "public", "if", "else", "for", "while", "return",
    "assert", "log", "main", "parallel", "custom",
}


def tokenize(src: str, fname: str):
    toks = []
    pos = 0
    n = len(src)
    m = _TOKEN_RE.match
    while pos < n:
        mo = m(src, pos)
        if mo is None:
            line = src.count("\n", 0, pos) + 1
            raise CircomError(f"{fname}:{line}: bad character {src[pos]!r}")
        kind = mo.lastgroup
        if kind != "ws":
            txt = mo.group()
            if kind == "hex":
                toks.append(("num", int(txt, 16), pos))
            elif kind == "num":
                toks.append(("num", int(txt), pos))
            elif kind == "id":
                toks.append(("id", txt, pos))
            elif kind == "str":
                toks.append(("str", txt[1:-1], pos))
            else:
                toks.append(("op", txt, pos))
        pos = mo.end()
    toks.append(("eof", None, n))
    return toks


# --------------------------------------------------------------------------
# Parser -> tuple AST
#   expressions: ('num',v) ('var',name) ('idx',base,e) ('mem',base,name)
#                ('call',name,[args]) ('arr',[e]) ('un',op,e) ('bin',op,l,r)
#                ('tern',c,a,b)
#   statements:  ('block',[s]) ('sigdecl',kind,[(name,[dims],init_op,init)])
#                ('vardecl',[(name,[dims],init)]) ('compdecl',[(name,[dims],init)])
#                ('assign',op,lhs,rhs) ('constr',l,r) ('if',c,a,b)
#                ('for',init,cond,step,body) ('while',c,body) ('return',e)
#                ('assert',e) ('log',) ('incdec',lhs,delta)
# --------------------------------------------------------------------------
BINPREC = {
    "||": 1, "&&": 2,
    "==": 3, "!=": 3, "<": 3, ">": 3, "<=": 3, ">=": 3,
    "|": 4, "^": 5, "&": 6, "<<": 7, ">>": 7,
    "+": 8, "-": 8, "*": 9, "/": 9, "\\": 9, "%": 9, "**": 10,
}
ASSIGN_OPS = {"=", "<==", "<--", "+=", "-=", "*=", "/=", "\\=", "%=", "<<=", ">>=",
              "&=", "|=", "^=", "**="}


class Parser:
    def __init__(self, src: str, fname: str):
        self.fname = fname
        self.src = src
        self.t = tokenize(src, fname)
        self.i = 0

    def err(self, msg):
        pos = self.t[self.i][2]
        line = self.src.count("\n", 0, pos) + 1
        raise CircomError(f"{self.fname}:{line}: {msg} (at {self.t[self.i][:2]})")

    def peek(self, k=0):
        return self.t[self.i + k]

    def isop(self, v, k=0):
        t = self.t[self.i + k]
        return t[0] == "op" and t[1] == v

    def iskw(self, v, k=0):
        t = self.t[self.i + k]
        return t[0] == "id" and t[1] == v

    def eat(self, v):
        t = self.t[self.i]
        if t[0] == "op" and t[1] == v:
            self.i += 1
            return True
        return False

    def expect(self, v):
        if not self.eat(v):
            self.err(f"expected {v!r}")

    def ident(self):
        t = self.t[self.i]
        if t[0] != "id":
            self.err("expected identifier")
        self.i += 1
        return t[1]

    def line(self):
        return self.src.count("\n", 0, self.t[self.i][2]) + 1

    # ---- top level
    def parse_file(self):
        includes, templates, functions, main = [], {}, {}, None
        while self.peek()[0] != "eof":
            if self.iskw("pragma"):
                while not self.eat(";"):
                    self.i += 1
            elif self.iskw("include"):
                self.i += 1
                t = self.peek()
                if t[0] != "str":
                    self.err("expected include path")
                self.i += 1
                self.expect(";")
                includes.append(t[1])
            elif self.iskw("template"):
                self.i += 1
                while self.iskw("custom") or self.iskw("parallel"):
                    self.i += 1
                ln = self.line()
                name = self.ident()
                params = self.params()
                body = self.block()
                templates[name] = (name, params, body, self.fname, ln)
            elif self.iskw("function"):
                self.i += 1
                ln = self.line()
                name = self.ident()
                params = self.params()
                body = self.block()
                functions[name] = (name, params, body, self.fname, ln)
            elif self.iskw("component"):
                self.i += 1
                if not self.iskw("main"):
                    self.err("only `component main` allowed at top level")
                self.i += 1
                public = []
                if self.eat("{"):
                    if not self.iskw("public"):
                        self.err("expected public")
                    self.i += 1
                    self.expect("[")
                    while not self.eat("]"):
                        public.append(self.ident())
                        self.eat(",")
                    self.expect("}")
                self.expect("=")
                call = self.expr()
                self.expect(";")
                main = (public, call)
            else:
                self.err("unexpected top-level token")
        return includes, templates, functions, main

    def params(self):
        self.expect("(")
        ps = []
        while not self.eat(")"):
            ps.append(self.ident())
            self.eat(",")
        return ps

    def block(self):
        self.expect("{")
        stmts = []
        while not self.eat("}"):
            stmts.append(self.stmt())
        return ("block", stmts)

    # ---- statements
    def stmt(self):
        t = self.peek()
        if t[0] == "op" and t[1] == "{":
            return self.block()
        if t[0] == "id":
            kw = t[1]
            if kw == "signal":
                s = self.decl_signal()
                self.expect(";")
                return s
            if kw == "var":
                s = self.decl_var()
                self.expect(";")
                return s
            if kw == "component":
                s = self.decl_comp()
                self.expect(";")
                return s
            if kw == "if":
                self.i += 1
                self.expect("(")
                c = self.expr()
                self.expect(")")
                a = self.stmt()
                b = None
                if self.iskw("else"):
                    self.i += 1
                    b = self.stmt()
                return ("if", c, a, b)
            if kw == "for":
                self.i += 1
                self.expect("(")
                init = self.simple_stmt()
                self.expect(";")
                cond = self.expr()
                self.expect(";")
                step = self.simple_stmt()
                self.expect(")")
                body = self.stmt()
                return ("for", init, cond, step, body)
            if kw == "while":
                self.i += 1
                self.expect("(")
                c = self.expr()
                self.expect(")")
                return ("while", c, self.stmt())
            if kw == "return":
                self.i += 1
                e = self.expr()
                self.expect(";")
                return ("return", e)
            if kw == "assert":
                self.i += 1
                self.expect("(")
                e = self.expr()
                self.expect(")")
                self.expect(";")
                return ("assert", e, self.fname, self.line())
            if kw == "log":
                self.i += 1
                self.expect("(")
                depth = 1
                while depth:
                    if self.isop("("):
                        depth += 1
                    elif self.isop(")"):
                        depth -= 1
                    self.i += 1
                self.expect(";")
                return ("log",)
        s = self.simple_stmt()
        self.expect(";")
        return s

    def simple_stmt(self):
        if self.iskw("var"):
            return self.decl_var()
        ln = self.line()
        if self.isop("++") or self.isop("--"):
            d = 1 if self.peek()[1] == "++" else -1
            self.i += 1
            lhs = self.expr()
            return ("incdec", lhs, d)
        e = self.expr()
        t = self.peek()
        if t[0] == "op":
            op = t[1]
            if op in ASSIGN_OPS:
                self.i += 1
                rhs = self.expr()
                return ("assign", op, e, rhs, self.fname, ln)
            if op == "==>":
                self.i += 1
                lhs = self.expr()
                return ("assign", "<==", lhs, e, self.fname, ln)
            if op == "-->":
                self.i += 1
                lhs = self.expr()
                return ("assign", "<--", lhs, e, self.fname, ln)
            if op == "===":
                self.i += 1
                rhs = self.expr()
                return ("constr", e, rhs, self.fname, ln)
            if op == "++":
                self.i += 1
                return ("incdec", e, 1)
            if op == "--":
                self.i += 1
                return ("incdec", e, -1)
        self.err("expected statement")

    def dims(self):
        ds = []
        while self.eat("["):
            ds.append(self.expr())
            self.expect("]")
        return ds

    def decl_signal(self):
        self.i += 1
        kind = "mid"
        if self.iskw("input"):
            kind = "in"
            self.i += 1
        elif self.iskw("output"):
            kind = "out"
            self.i += 1
        if self.isop("{"):  # tags
            while not self.eat("}"):
                self.i += 1
        ln = self.line()
        items = []
        while True:
            name = self.ident()
            ds = self.dims()
            op = init = None
            if self.isop("<==") or self.isop("<--"):
                op = self.peek()[1]
                self.i += 1
                init = self.expr()
            items.append((name, ds, op, init))
            if not self.eat(","):
                break
        return ("sigdecl", kind, items, self.fname, ln)

    def decl_var(self):
        self.i += 1
        items = []
        while True:
            name = self.ident()
            ds = self.dims()
            init = None
            if self.eat("="):
                init = self.expr()
            items.append((name, ds, init))
            if not self.eat(","):
                break
        return ("vardecl", items)

    def decl_comp(self):
        self.i += 1
        while self.iskw("parallel"):
            self.i += 1
        ln = self.line()
        items = []
        while True:
            name = self.ident()
            ds = self.dims()
            init = None
            if self.eat("="):
                init = self.expr()
            items.append((name, ds, init))
            if not self.eat(","):
                break
        return ("compdecl", items, self.fname, ln)

    # ---- expressions
    def expr(self):
        c = self.binexpr(1)
        if self.eat("?"):
            a = self.expr()
            self.expect(":")
            b = self.expr()
            return ("tern", c, a, b)
        return c

    def binexpr(self, minprec):
        lhs = self.unary()
        while True:
            t = self.peek()
            if t[0] != "op":
                break
            prec = BINPREC.get(t[1])
            if prec is None or prec < minprec:
                break
            self.i += 1
            rhs = self.binexpr(prec + 1)
            lhs = ("bin", t[1], lhs, rhs)
        return lhs

    def unary(self):
        t = self.peek()
        if t[0] == "op" and t[1] in ("-", "!", "~"):
            self.i += 1
            e = self.unary()
            if t[1] == "-" and e[0] == "num":
                return ("num", (-e[1]) % P)
            return ("un", t[1], e)
        return self.postfix()

    def postfix(self):
        t = self.peek()
        if t[0] == "num":
            self.i += 1
            e = ("num", t[1] % P)
        elif t[0] == "id":
            if t[1] == "parallel":
                self.i += 1
                return self.postfix()
            self.i += 1
            if self.isop("("):
                self.i += 1
                args = []
                while not self.eat(")"):
                    args.append(self.expr())
                    self.eat(",")
                e = ("call", t[1], args)
            else:
                e = ("var", t[1])
        elif t[0] == "op" and t[1] == "(":
            self.i += 1
            e = self.expr()
            self.expect(")")
        elif t[0] == "op" and t[1] == "[":
            self.i += 1
            items = []
            while not self.eat("]"):
                items.append(self.expr())
                self.eat(",")
            e = ("arr", items)
        else:
            self.err("expected expression")
        while True:
            if self.eat("["):
                ix = self.expr()
                self.expect("]")
                e = ("idx", e, ix)
            elif self.isop(".") and self.peek(1)[0] == "id":
                self.i += 1
                e = ("mem", e, self.ident())
            else:
                break
        return e


# --------------------------------------------------------------------------
# Program = all parsed files reachable through `include`
# --------------------------------------------------------------------------
class Program:
    def __init__(self, main_path: str, include_dirs=(), lazy_pattern=r"/powers/"):
        self.templates = {}
        self.functions = {}
        self.main = None
        self.seen = set()
        self.deferred = []
        self.include_dirs = list(include_dirs)
        self.lazy = re.compile(lazy_pattern) if lazy_pattern else None
        self.missing = []
        self._load(os.path.realpath(main_path), root=True)

    def _load(self, path, root=False, force=False):
        if path in self.seen:
            return
        if not os.path.exists(path):
            self.missing.append(path)
            self.seen.add(path)
            return
        if self.lazy is not None and not force and self.lazy.search(path):
            self.deferred.append(path)
            self.seen.add(path)
            return
        self.seen.add(path)
        with open(path) as f:
            src = f.read()
        incs, tmpls, funcs, main = Parser(src, path).parse_file()
        d = os.path.dirname(path)
        for inc in incs:
            cand = os.path.realpath(os.path.join(d, inc))
            if not os.path.exists(cand):
                for idir in self.include_dirs:
                    c2 = os.path.realpath(os.path.join(idir, inc))
                    if os.path.exists(c2):
                        cand = c2
                        break
            self._load(cand)
        self.templates.update(tmpls)
        self.functions.update(funcs)
        if root:
            self.main = main

    def _load_deferred(self):
        todo, self.deferred = self.deferred, []
        for p in todo:
            self.seen.discard(p)
            self._load(p, force=True)

    def get_template(self, name):
        t = self.templates.get(name)
        if t is None and self.deferred:
            self._load_deferred()
            t = self.templates.get(name)
        return t

    def get_function(self, name):
        f = self.functions.get(name)
        if f is None and self.deferred:
            self._load_deferred()
            f = self.functions.get(name)
        return f


# --------------------------------------------------------------------------
# Field helpers
# --------------------------------------------------------------------------
def _signed(v):
    return v - P if v > HALF else v


def _shl(a, b):
    if b > HALF:
        return _shr(a, P - b)
    if b >= 254:
        return 0
    return ((a << b) & MASK) % P


def _shr(a, b):
    if b > HALF:
        return _shl(a, P - b)
    if b >= 254:
        return 0
    return a >> b


def binop(op, a, b):
    if op == "+":
        return (a + b) % P
    if op == "-":
        return (a - b) % P
    if op == "*":
        return (a * b) % P
    if op == "<":
        return 1 if _signed(a) < _signed(b) else 0
    if op == "==":
        return 1 if a == b else 0
    if op == "\\":
        if b == 0:
            raise CircomError("integer division by zero")
        return a // b
    if op == "%":
        if b == 0:
            raise CircomError("modulo by zero")
        return a % b
    if op == "<<":
        return _shl(a, b)
    if op == ">>":
        return _shr(a, b)
    if op == "&":
        return (a & b) % P
    if op == ">":
        return 1 if _signed(a) > _signed(b) else 0
    if op == "<=":
        return 1 if _signed(a) <= _signed(b) else 0
    if op == ">=":
        return 1 if _signed(a) >= _signed(b) else 0
    if op == "!=":
        return 1 if a != b else 0
    if op == "&&":
        return 1 if (a != 0 and b != 0) else 0
    if op == "||":
        return 1 if (a != 0 or b != 0) else 0
    if op == "/":
        if b == 0:
            return 0
        return (a * pow(b, -1, P)) % P
    if op == "**":
        return pow(a, b, P)
    if op == "|":
        return ((a | b) & MASK) % P
    if op == "^":
        return ((a ^ b) & MASK) % P
    raise CircomError(f"unknown operator {op}")


class _Unknown:
    """Phase-A placeholder for anything that depends on a signal value."""
    __slots__ = ()

    def __repr__(self):
        return "UNK"


UNK = _Unknown()


def _has_unk(v):
    if v is UNK:
        return True
    if isinstance(v, list):
        for x in v:
            if _has_unk(x):
                return True
    return False


def _zeros(dims):
    if not dims:
        return 0
    if len(dims) == 1:
        return [0] * dims[0]
    return [_zeros(dims[1:]) for _ in range(dims[0])]


def _copy(v):
    if v.__class__ is list:
        if v and v[0].__class__ is list:
            return [_copy(x) for x in v]
        return v[:]   # circom arrays are rectangular: a list whose first element is a scalar holds scalars only
    return v


def _argkey(v):
    if isinstance(v, list):
        return tuple(_argkey(x) for x in v)
    return v


def _prod(dims):
    n = 1
    for d in dims:
        n *= d
    return n


# --------------------------------------------------------------------------
# Layout (phase A result, cached per (template, args))
# --------------------------------------------------------------------------
class Layout:
    __slots__ = ("tname", "args", "sigs", "order", "own", "total", "children",
                 "n_inputs", "comp_dims", "child_order")

    def __init__(self, tname, args):
        self.tname = tname
        self.args = args
        self.sigs = {}        # name -> [rel_off, dims, kind]
        self.order = []       # declaration order of signal names
        self.children = {}    # (name, flat_idx) -> (tname, argkey, args, rel_base, Layout)
        self.child_order = []
        self.comp_dims = {}   # name -> dims
        self.own = 0
        self.total = 0
        self.n_inputs = 0


class _Return(Exception):
    def __init__(self, v):
        self.v = v


class SigRef:
    """An l-value / r-value handle on a (sub-)array of signals of one component."""
    __slots__ = ("comp", "off", "dims", "kind", "name")

    def __init__(self, comp, off, dims, kind, name):
        self.comp = comp
        self.off = off
        self.dims = dims
        self.kind = kind
        self.name = name


class CompArr:
    __slots__ = ("name", "dims", "path")

    def __init__(self, name, dims, path=()):
        self.name = name
        self.dims = dims
        self.path = path


class Comp:
    __slots__ = ("lay", "base", "pending", "ran", "kids", "name", "parent")

    def __init__(self, lay, base, name, parent):
        self.lay = lay
        self.base = base
        self.pending = lay.n_inputs
        self.ran = False
        self.kids = {}
        self.name = name
        self.parent = parent


# --------------------------------------------------------------------------
# Interpreter
# --------------------------------------------------------------------------
class Circuit:
    """One instantiated circuit. `calculate_witness(inputs)` evaluates all signals."""

    def __init__(self, main_path, include_dirs=(), main_override=None, translate_functions=True):
        self.prog = Program(main_path, include_dirs)
        # functions are run as translated Python (FunctionTranslator) unless translate_functions=False
        self.translator = FunctionTranslator(self) if translate_functions else None
        self.layouts = {}
        main = main_override or self.prog.main
        if main is None:
            raise CircomError("no `component main` in root file")
        self.public_inputs, call = main
        if call[0] != "call":
            raise CircomError("main must be a template call")
        self.phase = "A"
        self.W = None
        args = [self.eval_expr(a, [{}], None) for a in call[2]]
        self.main_layout = self.layout_of(call[1], args)
        self.n_signals = self.main_layout.total
        self.constraint_count = 0
        self.fn_depth = 0

    # ---------------------------------------------------------------- phase A
    def layout_of(self, tname, args):
        key = (tname, _argkey(args))
        lay = self.layouts.get(key)
        if lay is not None:
            return lay
        tmpl = self.prog.get_template(tname)
        if tmpl is None:
            raise CircomError(f"unknown template {tname}")
        _, params, body, fname, ln = tmpl
        if len(params) != len(args):
            raise CircomError(f"{tname}: expected {len(params)} args, got {len(args)}")
        lay = Layout(tname, args)
        saved = self.phase
        self.phase = "A"
        env = [dict(zip(params, [_copy(a) for a in args]))]
        self._cur_lay_stack = getattr(self, "_cur_lay_stack", [])
        self._cur_lay_stack.append(lay)
        try:
            self.exec_stmt(body, env, lay)
        except _Return:
            raise CircomError(f"return inside template {tname}")
        finally:
            self._cur_lay_stack.pop()
            self.phase = saved
        # offsets: outputs, inputs, intermediates (declaration order), then children
        off = 0
        for kind in ("out", "in", "mid"):
            for name in lay.order:
                s = lay.sigs[name]
                if s[2] == kind:
                    s[0] = off
                    off += _prod(s[1])
                    if kind == "in":
                        lay.n_inputs += _prod(s[1])
        lay.own = off
        for ck in lay.child_order:
            c = lay.children[ck]
            c[3] = off
            off += c[4].total
        lay.total = off
        self.layouts[key] = lay
        return lay

    # ---------------------------------------------------------------- naming
    def signal_names(self):
        """Qualified name for every signal index (``main.a.b[3].c[1][2]``)."""
        names = [None] * self.n_signals

        def idx_suffixes(dims):
            if not dims:
                return [""]
            out = [""]
            for d in dims:
                out = [p + f"[{i}]" for p in out for i in range(d)]
            return out

        def walk(lay, base, prefix):
            for name in lay.order:
                off, dims, _ = lay.sigs[name]
                for k, suf in enumerate(idx_suffixes(dims)):
                    names[base + off + k] = f"{prefix}.{name}{suf}"
            for (cname, flat) in lay.child_order:
                c = lay.children[(cname, flat)]
                dims = lay.comp_dims[cname]
                suf = ""
                if dims:
                    rem = flat
                    parts = []
                    for d in reversed(dims):
                        parts.append(rem % d)
                        rem //= d
                    suf = "".join(f"[{i}]" for i in reversed(parts))
                walk(c[4], base + c[3], f"{prefix}.{cname}{suf}")

        walk(self.main_layout, 0, "main")
        return names

    def main_io(self):
        """[(name, dims, kind, offset)] for main's input/output signals."""
        lay = self.main_layout
        return [(n, tuple(lay.sigs[n][1]), lay.sigs[n][2], lay.sigs[n][0])
                for n in lay.order if lay.sigs[n][2] in ("in", "out")]

    # ---------------------------------------------------------------- witness
    def calculate_witness(self, inputs: dict, check=True):
        """Mirror of witness_calculator.js calculateWitness(input, sanityCheck)
        (call site /root/reference/test/automatisationTest.js:40-50).  Returns the
        signal vector (index = our O0 signal numbering, see signal_names())."""
        self.W = [None] * self.n_signals
        self.check = check
        self.phase = "B"
        self.constraint_count = 0
        lay = self.main_layout
        main = Comp(lay, 0, "main", None)
        n_set = 0
        for name, val in inputs.items():
            s = lay.sigs.get(name)
            if s is None or s[2] != "in":
                raise CircomError(f"Signal not found: {name}")
            flat = _flatten_input(val)
            need = _prod(s[1])
            if len(flat) < need:
                raise CircomError(f"Not enough values for input signal {name}")
            if len(flat) > need:
                raise CircomError(f"Too many values for input signal {name}")
            for k, v in enumerate(flat):
                self.W[s[0] + k] = v % P
            n_set += need
        if n_set != lay.n_inputs:
            raise CircomError(f"Not all inputs have been set. Only {n_set} out of {lay.n_inputs}")
        main.pending = 0
        self.run_comp(main)
        W = self.W
        return [0 if v is None else v for v in W]

    def run_comp(self, comp):
        if comp.ran:
            raise CircomError(f"component {comp.name} executed twice")
        comp.ran = True
        lay = comp.lay
        tmpl = self.prog.get_template(lay.tname)
        env = [dict(zip(tmpl[1], [_copy(a) for a in lay.args]))]
        self.exec_stmt(tmpl[2], env, comp)

    # ---------------------------------------------------------------- statements
    def exec_stmt(self, s, env, ctx):
        k = s[0]
        if k == "block":
            env.append({})
            try:
                for st in s[1]:
                    self.exec_stmt(st, env, ctx)
            finally:
                env.pop()
        elif k == "assign":
            self.exec_assign(s, env, ctx)
        elif k == "for":
            env.append({})
            try:
                self.exec_stmt(s[1], env, ctx)
                body = s[4]
                cond = s[2]
                step = s[3]
                n_iter = 0
                while True:
                    c = self.eval_expr(cond, env, ctx)
                    if c is UNK:
                        raise CircomError("loop condition depends on a signal")
                    if c == 0:
                        break
                    self.exec_stmt(body, env, ctx)
                    self.exec_stmt(step, env, ctx)
                    n_iter += 1
            finally:
                env.pop()
        elif k == "if":
            c = self.eval_expr(s[1], env, ctx)
            if c is UNK:
                # phase A: both branches only touch vars; poison what they assign
                for name in _assigned_vars(s[2]) | (_assigned_vars(s[3]) if s[3] else set()):
                    self.set_var(env, name, UNK, missing_ok=True)
                return
            if c != 0:
                self.exec_stmt(s[2], env, ctx)
            elif s[3] is not None:
                self.exec_stmt(s[3], env, ctx)
        elif k == "vardecl":
            for name, dims, init in s[1]:
                if init is not None:
                    v = self.eval_expr(init, env, ctx)
                    v = _copy(v)
                else:
                    dd = [self.eval_expr(d, env, ctx) for d in dims]
                    if any(d is UNK for d in dd):
                        raise CircomError(f"var {name}: dimension depends on a signal")
                    v = _zeros(dd)
                env[-1][name] = v
        elif k == "sigdecl":
            self.exec_sigdecl(s, env, ctx)
        elif k == "compdecl":
            self.exec_compdecl(s, env, ctx)
        elif k == "incdec":
            lhs = s[1]
            cur = self.eval_expr(lhs, env, ctx)
            nv = UNK if cur is UNK else (cur + s[2]) % P
            self.assign_var(lhs, nv, env, ctx)
        elif k == "constr":
            if self.phase == "A":
                return
            a = self.eval_expr(s[1], env, ctx)
            b = self.eval_expr(s[2], env, ctx)
            self._constrain_eq(a, b, s)
        elif k == "while":
            while True:
                c = self.eval_expr(s[1], env, ctx)
                if c is UNK:
                    raise CircomError("while condition depends on a signal")
                if c == 0:
                    break
                self.exec_stmt(s[2], env, ctx)
        elif k == "return":
            raise _Return(self.eval_expr(s[1], env, ctx))
        elif k == "assert":
            c = self.eval_expr(s[1], env, ctx)
            if c is UNK:
                return
            if c == 0:
                raise AssertFailed(f"assert failed at {s[2]}:{s[3]}")
        elif k == "log":
            pass
        else:
            raise CircomError(f"unknown statement {k}")

    def _constrain_eq(self, a, b, s):
        self.constraint_count += 1
        if isinstance(a, list) or isinstance(b, list):
            fa, fb = _flatten_input(a), _flatten_input(b)
            if len(fa) != len(fb):
                raise CircomError(f"{s[3]}:{s[4]}: array constraint size mismatch")
            self.constraint_count += len(fa) - 1
            if self.check and fa != fb:
                raise AssertFailed(f"Assert Failed. constraint at {s[3]}:{s[4]}")
        elif self.check and a != b:
            raise AssertFailed(f"Assert Failed. constraint at {s[3]}:{s[4]}")

    def exec_sigdecl(self, s, env, ctx):
        kind = s[1]
        for name, dims, op, init in s[2]:
            if self.phase == "A":
                dd = [self.eval_expr(d, env, ctx) for d in dims]
                if any(d is UNK for d in dd):
                    raise CircomError(f"signal {name}: dimension depends on a signal")
                if name in ctx.sigs:
                    raise CircomError(f"{s[3]}:{s[4]}: signal {name} declared twice")
                ctx.sigs[name] = [0, dd, kind]
                ctx.order.append(name)
            elif init is not None:
                self.exec_assign(("assign", op, ("var", name), init, s[3], s[4]), env, ctx)

    def exec_compdecl(self, s, env, ctx):
        for name, dims, init in s[1]:
            if self.phase == "A":
                dd = [self.eval_expr(d, env, ctx) for d in dims]
                if any(d is UNK for d in dd):
                    raise CircomError(f"component {name}: dimension depends on a signal")
                if name not in ctx.comp_dims:
                    ctx.comp_dims[name] = dd
            if init is not None:
                self.exec_assign(("assign", "=", ("var", name), init, s[2], s[3]), env, ctx)

    # ---------------------------------------------------------------- l-values
    def find_var(self, env, name):
        for scope in reversed(env):
            if name in scope:
                return scope
        return None

    def set_var(self, env, name, v, missing_ok=False):
        sc = self.find_var(env, name)
        if sc is None:
            if missing_ok:
                return
            raise CircomError(f"unknown variable {name}")
        sc[name] = v

    def assign_var(self, lhs, v, env, ctx):
        """`lhs = v` where lhs is var or var[idx]...; v already evaluated."""
        if lhs[0] == "var":
            sc = self.find_var(env, lhs[1])
            if sc is None:
                raise CircomError(f"assignment to unknown variable {lhs[1]}")
            sc[lhs[1]] = _copy(v)
            return
        if lhs[0] == "idx":
            idxs = []
            base = lhs
            while base[0] == "idx":
                idxs.append(self.eval_expr(base[2], env, ctx))
                base = base[1]
            if base[0] != "var":
                raise CircomError("bad assignment target")
            idxs.reverse()
            sc = self.find_var(env, base[1])
            if sc is None:
                raise CircomError(f"assignment to unknown variable {base[1]}")
            if any(i is UNK for i in idxs):
                sc[base[1]] = UNK
                return
            arr = sc[base[1]]
            if arr is UNK:
                return
            for i in idxs[:-1]:
                arr = arr[i]
                if arr is UNK:
                    return
            arr[idxs[-1]] = _copy(v)
            return
        raise CircomError("bad assignment target")

    def resolve_lhs(self, e, env, ctx):
        """Resolve a signal/component l-value expression to SigRef / CompArr / Comp,
        or None when it names a var."""
        k = e[0]
        if k == "var":
            name = e[1]
            if self.find_var(env, name) is not None:
                return None
            lay = ctx if self.phase == "A" else ctx.lay
            s = lay.sigs.get(name)
            if s is not None:
                return SigRef(ctx, s[0], s[1], s[2], name)
            if name in lay.comp_dims:
                return CompArr(name, lay.comp_dims[name])
            raise CircomError(f"unknown identifier {name}")
        if k == "idx":
            base = self.resolve_lhs(e[1], env, ctx)
            if base is None:
                return None
            if base is UNK:
                return UNK
            i = self.eval_expr(e[2], env, ctx)
            if i is UNK:
                raise CircomError("signal index depends on a signal")
            if isinstance(base, SigRef):
                if not base.dims:
                    raise CircomError(f"too many indices on signal {base.name}")
                d0 = base.dims[0]
                if i >= d0:
                    raise CircomError(f"index {i} out of range for {base.name} (dim {d0})")
                rest = base.dims[1:]
                return SigRef(base.comp, base.off + i * _prod(rest), rest, base.kind, base.name)
            if isinstance(base, CompArr):
                if not base.dims:
                    raise CircomError(f"too many indices on component {base.name}")
                if i >= base.dims[0]:
                    raise CircomError(f"component index {i} out of range for {base.name}")
                return CompArr(base.name, base.dims[1:], base.path + (i,))
            raise CircomError("cannot index this")
        if k == "mem":
            base = self.resolve_lhs(e[1], env, ctx)
            if not isinstance(base, CompArr) or base.dims:
                raise CircomError("member access on a non-component")
            if self.phase == "A":
                return UNK
            child = self.get_child(ctx, base)
            s = child.lay.sigs.get(e[2])
            if s is None:
                raise CircomError(f"component {child.name} has no signal {e[2]}")
            return SigRef(child, s[0], s[1], s[2], e[2])
        raise CircomError("bad l-value")

    def _flat_index(self, lay, ca):
        dims = lay.comp_dims[ca.name]
        flat = 0
        for d, i in zip(dims, ca.path):
            flat = flat * d + i
        return flat

    def get_child(self, comp, ca):
        key = (ca.name, self._flat_index(comp.lay, ca))
        child = comp.kids.get(key)
        if child is None:
            raise CircomError(f"component {comp.name}.{ca.name}{list(ca.path)} used before instantiation")
        return child

    # ---------------------------------------------------------------- assignment
    def exec_assign(self, s, env, ctx):
        op, lhs, rhs = s[1], s[2], s[3]
        if self.phase == "A" and (op == "<==" or op == "<--"):
            return
        target = self.resolve_lhs(lhs, env, ctx) if op in ("=", "<==", "<--") else None
        if target is None:
            # plain var assignment (or compound)
            if op in ("<==", "<--"):
                raise CircomError(f"{s[4]}:{s[5]}: signal assignment to a var")
            v = self.eval_expr(rhs, env, ctx)
            if op != "=":
                cur = self.eval_expr(lhs, env, ctx)
                if cur is UNK or v is UNK:
                    v = UNK
                else:
                    v = binop(op[:-1], cur, v)
            self.assign_var(lhs, v, env, ctx)
            return
        if isinstance(target, CompArr):
            if op != "=":
                raise CircomError(f"{s[4]}:{s[5]}: components are assigned with =")
            self.instantiate(target, rhs, env, ctx)
            return
        if self.phase == "A":
            return
        if op == "=":
            raise CircomError(f"{s[4]}:{s[5]}: signals are assigned with <== or <--")
        v = self.eval_expr(rhs, env, ctx)
        self.store_signal(target, v, s)

    def store_signal(self, ref, v, s):
        W = self.W
        comp = ref.comp
        base = comp.base + ref.off
        if ref.dims:
            flat = _flatten_input(v)
            if len(flat) != _prod(ref.dims):
                raise CircomError(f"{s[4]}:{s[5]}: array assignment size mismatch for {ref.name}")
        else:
            if isinstance(v, list):
                raise CircomError(f"{s[4]}:{s[5]}: array assigned to scalar signal {ref.name}")
            flat = (v,)
        for k, x in enumerate(flat):
            if W[base + k] is not None:
                raise CircomError(f"{s[4]}:{s[5]}: signal {comp.name}.{ref.name} assigned twice")
            W[base + k] = x
        if s[1] == "<==":
            self.constraint_count += len(flat)
        if ref.kind == "in" and comp.parent is not None and comp is not None:
            comp.pending -= len(flat)
            if comp.pending == 0:
                self.run_comp(comp)

    def instantiate(self, ca, rhs, env, ctx):
        if ca.dims:
            raise CircomError(f"component array {ca.name} assigned as a whole")
        if rhs[0] != "call":
            raise CircomError("component initialiser must be a template call")
        args = [self.eval_expr(a, env, ctx) for a in rhs[2]]
        if self.phase == "A":
            if any(_has_unk(a) for a in args):
                raise CircomError(f"template argument of {rhs[1]} depends on a signal")
            lay = ctx
            key = (ca.name, self._flat_index(lay, ca))
            if key in lay.children:
                raise CircomError(f"component {ca.name}{list(ca.path)} instantiated twice")
            sub = self.layout_of(rhs[1], args)
            lay.children[key] = [rhs[1], _argkey(args), args, 0, sub]
            lay.child_order.append(key)
            return
        comp = ctx
        key = (ca.name, self._flat_index(comp.lay, ca))
        c = comp.lay.children.get(key)
        if c is None or c[0] != rhs[1] or c[1] != _argkey(args):
            raise CircomError(f"phase mismatch instantiating {ca.name}{list(ca.path)}")
        suffix = "".join(f"[{i}]" for i in ca.path)
        child = Comp(c[4], comp.base + c[3], f"{comp.name}.{ca.name}{suffix}", comp)
        comp.kids[key] = child
        if child.pending == 0:
            self.run_comp(child)

    # ---------------------------------------------------------------- expressions
    def read_signal(self, ref):
        if self.phase == "A":
            return UNK
        W = self.W
        base = ref.comp.base + ref.off
        if not ref.dims:
            v = W[base]
            return 0 if v is None else v
        return self._read_arr(base, ref.dims)

    def _read_arr(self, base, dims):
        W = self.W
        if len(dims) == 1:
            return [0 if v is None else v for v in W[base:base + dims[0]]]
        stride = _prod(dims[1:])
        return [self._read_arr(base + i * stride, dims[1:]) for i in range(dims[0])]

    def eval_expr(self, e, env, ctx):
        k = e[0]
        if k == "num":
            return e[1]
        if k == "var":
            name = e[1]
            for scope in reversed(env):
                if name in scope:
                    return scope[name]
            if ctx is None:
                raise CircomError(f"unknown identifier {name}")
            r = self.resolve_lhs(e, env, ctx)
            if isinstance(r, SigRef):
                return self.read_signal(r)
            raise CircomError(f"component {name} used as a value")
        if k == "bin":
            op = e[1]
            a = self.eval_expr(e[2], env, ctx)
            if op == "&&":
                if a is not UNK and a == 0:
                    return 0
            elif op == "||":
                if a is not UNK and a != 0:
                    return 1
            b = self.eval_expr(e[3], env, ctx)
            if a is UNK or b is UNK:
                return UNK
            if isinstance(a, list) or isinstance(b, list):
                raise CircomError(f"operator {op} on arrays")
            return binop(op, a, b)
        if k == "idx":
            # fast path for var arrays
            base = e
            idxs = []
            while base[0] == "idx":
                idxs.append(base[2])
                base = base[1]
            if base[0] == "var":
                name = base[1]
                for scope in reversed(env):
                    if name in scope:
                        v = scope[name]
                        for ie in reversed(idxs):
                            i = self.eval_expr(ie, env, ctx)
                            if v is UNK or i is UNK:
                                return UNK
                            if not isinstance(v, list):
                                raise CircomError(f"indexing scalar var {name}")
                            if i >= len(v):
                                raise CircomError(f"index {i} out of range for var {name}")
                            v = v[i]
                        return v
            r = self.resolve_lhs(e, env, ctx)
            if r is UNK:
                return UNK
            if isinstance(r, SigRef):
                return self.read_signal(r)
            raise CircomError("component used as a value")
        if k == "mem":
            r = self.resolve_lhs(e, env, ctx)
            if r is UNK:
                return UNK
            return self.read_signal(r)
        if k == "call":
            return self.call_function(e, env, ctx)
        if k == "tern":
            c = self.eval_expr(e[1], env, ctx)
            if c is UNK:
                return UNK
            return self.eval_expr(e[2] if c != 0 else e[3], env, ctx)
        if k == "un":
            a = self.eval_expr(e[2], env, ctx)
            if a is UNK:
                return UNK
            op = e[1]
            if op == "-":
                return (-a) % P
            if op == "!":
                return 1 if a == 0 else 0
            if op == "~":
                return ((a ^ MASK) & MASK) % P
        if k == "arr":
            return [self.eval_expr(x, env, ctx) for x in e[1]]
        raise CircomError(f"bad expression {k}")

    def call_function(self, e, env, ctx):
        name = e[1]
        if self.prog.get_function(name) is None:
            raise CircomError(f"unknown function {name}")
        args = [self.eval_expr(a, env, ctx) for a in e[2]]
        if any(_has_unk(a) for a in args):
            return UNK
        if self.translator is not None:
            return self.translator.call(name, args)
        return self.interpret_function(name, args)

    def interpret_function(self, name, args):
        """tree-walking execution of a function body (the definition of the semantics; FunctionTranslator is the
        fast path and falls back to this for anything outside its subset)"""
        fn = self.prog.get_function(name)
        key = None
        if all(not isinstance(a, list) for a in args):
            key = (name, tuple(args))
            hit = self._fn_cache.get(key) if hasattr(self, "_fn_cache") else None
            if hit is not None:
                return _copy(hit)
        _, params, body, fname, ln = fn
        if len(params) != len(args):
            raise CircomError(f"{name}: expected {len(params)} args")
        fenv = [dict(zip(params, [_copy(a) for a in args]))]
        try:
            self.exec_stmt(body, fenv, None)
        except _Return as r:
            if key is not None:
                if not hasattr(self, "_fn_cache"):
                    self._fn_cache = {}
                self._fn_cache[key] = _copy(r.v)
            return r.v
        raise CircomError(f"function {name} did not return")


# --------------------------------------------------------------------------
# Function translator: circom `function` bodies -> Python functions
# --------------------------------------------------------------------------
class _Unsupported(Exception):
    pass


def _idiv(a, b):
    if b == 0:
        raise CircomError("integer division by zero")
    return a // b


def _imod(a, b):
    if b == 0:
        raise CircomError("modulo by zero")
    return a % b


def _fdiv(a, b):
    if b == 0:
        return 0
    return (a * pow(b, -1, P)) % P


class FunctionTranslator:
    """The witness-time hint functions of the big-integer library (long_div, mod_inv = a^(p-2) by 380 rounds of
    prod + long_div, ... /root/reference/circuits/lib/circuits/bigInt/bigIntFunc.circom:126-660) are pure functions
    of vars; walking their syntax trees costs ~4 us per node, which puts one ECDSA witness (362 mod_inv calls) at
    hours.  This translates a function's syntax tree ONCE into Python source with exactly the interpreter's
    semantics (same field operators, block scoping resolved statically, copy-on-assign arrays, lazy && || ?:) and
    executes that instead.  It is still the reference's own source text that runs - no function is restated by hand;
    tests/test_cpu_compiler_vs_oracle.py::test_translated_functions_equal_the_tree_walker compares both execution
    modes on random operands.  Functions using anything outside the subset keep running in the tree walker."""
    MAX_STMTS = 4000   # the 70 000-line generator tables are evaluated once by the tree walker

    def __init__(self, circuit):
        self.circuit = circuit
        self.fns = {}
        self.n_calls = 0

    def get(self, name):
        f = self.fns.get(name, False)
        if f is not False:
            return f
        self.fns[name] = None
        fn = self.circuit.prog.get_function(name)
        if fn is None:
            return None
        try:
            src = self._translate(name, fn[1], fn[2])
            glb = {"P": P, "HALF": HALF, "MASK": MASK, "_cp": _copy, "_zeros": _zeros, "_shl": _shl, "_shr": _shr,
                   "_idiv": _idiv, "_imod": _imod, "_fdiv": _fdiv, "_call": self.call, "CircomError": CircomError,
                   "AssertFailed": AssertFailed}
            exec(compile(src, f"<circom function {name}>", "exec"), glb)
            self.fns[name] = glb["_f"]
        except _Unsupported:
            self.fns[name] = None
        return self.fns[name]

    def call(self, name, args):
        f = self.get(name)
        if f is None:
            return self.circuit.interpret_function(name, args)
        self.n_calls += 1
        return f(*[_copy(a) for a in args])

    # ---- translation
    def _translate(self, name, params, body):
        self.n_stmts = 0
        self.uid = 0
        self.lines = []
        self.scopes = [{}]
        pn = []
        for q in params:
            pn.append(self._declare(q))
        self.lines.append("def _f(" + ", ".join(pn) + "):")
        self._stmt(body, 1)
        self.lines.append(f"    raise CircomError('function {name} did not return')")
        return "\n".join(self.lines) + "\n"

    def _declare(self, name):
        self.uid += 1
        py = f"v{self.uid}_{re.sub(r'[^A-Za-z0-9_]', '_', name)}"
        self.scopes[-1][name] = py
        return py

    def _lookup(self, name):
        for sc in reversed(self.scopes):
            if name in sc:
                return sc[name]
        raise _Unsupported(name)

    def _emit(self, ind, text):
        self.lines.append("    " * ind + text)

    def _sg(self, x):
        return f"({x} - P if {x} > HALF else {x})" if re.fullmatch(r"[A-Za-z0-9_]+", x) else f"(lambda t: t - P if t > HALF else t)({x})"

    def _cond(self, e):
        """Python truth value of `e != 0`."""
        if e[0] == "bin":
            op = e[1]
            if op in ("<", ">", "<=", ">="):
                return f"({self._sg(self._expr(e[2]))} {op} {self._sg(self._expr(e[3]))})"
            if op in ("==", "!="):
                return f"({self._expr(e[2])} {op} {self._expr(e[3])})"
            if op == "&&":
                return f"({self._cond(e[2])} and {self._cond(e[3])})"
            if op == "||":
                return f"({self._cond(e[2])} or {self._cond(e[3])})"
        if e[0] == "un" and e[1] == "!":
            return f"(not {self._cond(e[2])})"
        return f"({self._expr(e)} != 0)"

    def _expr(self, e):
        k = e[0]
        if k == "num":
            return str(e[1])
        if k == "var":
            return self._lookup(e[1])
        if k == "bin":
            op = e[1]
            if op in ("<", ">", "<=", ">=", "==", "!=", "&&", "||"):
                return f"(1 if {self._cond(e)} else 0)"
            a, b = self._expr(e[2]), self._expr(e[3])
            if op in ("+", "-", "*"):
                return f"(({a} {op} {b}) % P)"
            if op == "\\":
                return f"_idiv({a}, {b})"
            if op == "%":
                return f"_imod({a}, {b})"
            if op == "<<":
                return f"_shl({a}, {b})"
            if op == ">>":
                return f"_shr({a}, {b})"
            if op == "&":
                return f"(({a} & {b}) % P)"
            if op in ("|", "^"):
                return f"((({a} {op} {b}) & MASK) % P)"
            if op == "/":
                return f"_fdiv({a}, {b})"
            if op == "**":
                return f"pow({a}, {b}, P)"
            raise _Unsupported(op)
        if k == "un":
            a = self._expr(e[2])
            if e[1] == "-":
                return f"((-{a}) % P)"
            if e[1] == "!":
                return f"(1 if {a} == 0 else 0)"
            return f"((({a} ^ MASK) & MASK) % P)"
        if k == "tern":
            return f"({self._expr(e[2])} if {self._cond(e[1])} else {self._expr(e[3])})"
        if k == "idx":
            idxs = []
            base = e
            while base[0] == "idx":
                idxs.append(base[2])
                base = base[1]
            if base[0] != "var":
                raise _Unsupported("indexed expression")
            return self._lookup(base[1]) + "".join(f"[{self._expr(i)}]" for i in reversed(idxs))
        if k == "call":
            return f"_call({e[1]!r}, [{', '.join(self._expr(a) for a in e[2])}])"
        if k == "arr":
            return "[" + ", ".join(self._expr(x) for x in e[1]) + "]"
        raise _Unsupported(k)

    def _rhs(self, e):
        """assignment copies arrays; expressions that can only be scalars skip the copy"""
        x = self._expr(e)
        return x if e[0] in ("num", "bin", "un") else f"_cp({x})"

    def _lvalue(self, lhs):
        if lhs[0] == "var":
            return self._lookup(lhs[1])
        if lhs[0] == "idx":
            return self._expr(lhs)
        raise _Unsupported("assignment target")

    def _stmt(self, s, ind):
        self.n_stmts += 1
        if self.n_stmts > self.MAX_STMTS:
            raise _Unsupported("too large")
        k = s[0]
        if k == "block":
            self.scopes.append({})
            if not s[1]:
                self._emit(ind, "pass")
            for st in s[1]:
                self._stmt(st, ind)
            self.scopes.pop()
        elif k == "vardecl":
            for name, dims, init in s[1]:
                if init is not None:
                    val = self._rhs(init)
                elif dims:
                    val = "_zeros([" + ", ".join(self._expr(d) for d in dims) + "])"
                else:
                    val = "0"
                self._emit(ind, f"{self._declare(name)} = {val}")
        elif k == "assign":
            op, lhs, rhs = s[1], s[2], s[3]
            if op == "=":
                self._emit(ind, f"{self._lvalue(lhs)} = {self._rhs(rhs)}")
            elif op.endswith("=") and op not in ("<==", "<--"):
                fake = ("bin", op[:-1], lhs, rhs)
                self._emit(ind, f"{self._lvalue(lhs)} = {self._expr(fake)}")
            else:
                raise _Unsupported(op)
        elif k == "incdec":
            lv = self._lvalue(s[1])
            self._emit(ind, f"{lv} = ({lv} + {s[2]}) % P")
        elif k == "for":
            self.scopes.append({})
            self._stmt(s[1], ind)
            self._emit(ind, "while True:")
            self._emit(ind + 1, f"if not {self._cond(s[2])}:")
            self._emit(ind + 2, "break")
            self._stmt(s[4], ind + 1)
            self._stmt(s[3], ind + 1)
            self.scopes.pop()
        elif k == "while":
            self._emit(ind, "while True:")
            self._emit(ind + 1, f"if not {self._cond(s[1])}:")
            self._emit(ind + 2, "break")
            self._stmt(s[2], ind + 1)
        elif k == "if":
            self._emit(ind, f"if {self._cond(s[1])}:")
            self.scopes.append({})
            self._stmt(s[2], ind + 1)
            self.scopes.pop()
            if s[3] is not None:
                self._emit(ind, "else:")
                self.scopes.append({})
                self._stmt(s[3], ind + 1)
                self.scopes.pop()
        elif k == "return":
            self._emit(ind, f"return {self._expr(s[1])}")
        elif k == "assert":
            self._emit(ind, f"if not {self._cond(s[1])}:")
            self._emit(ind + 1, f"raise AssertFailed('assert failed at {s[2]}:{s[3]}')")
        elif k == "log":
            self._emit(ind, "pass")
        else:
            raise _Unsupported(k)


def _assigned_vars(s):
    out = set()
    if s is None:
        return out
    k = s[0]
    if k == "block":
        for st in s[1]:
            out |= _assigned_vars(st)
    elif k == "assign":
        b = s[2]
        while b[0] == "idx":
            b = b[1]
        if b[0] == "var":
            out.add(b[1])
    elif k == "incdec":
        b = s[1]
        while b[0] == "idx":
            b = b[1]
        if b[0] == "var":
            out.add(b[1])
    elif k == "if":
        out |= _assigned_vars(s[2]) | _assigned_vars(s[3])
    elif k == "for":
        out |= _assigned_vars(s[1]) | _assigned_vars(s[3]) | _assigned_vars(s[4])
    elif k == "while":
        out |= _assigned_vars(s[2])
    elif k in ("sigdecl", "compdecl"):
        raise CircomError("declaration under a signal-dependent condition")
    return out


def _flatten_input(v):
    if isinstance(v, (list, tuple)):
        out = []
        for x in v:
            out.extend(_flatten_input(x))
        return out
    if isinstance(v, str):
        return [int(v, 16) if v.startswith(("0x", "0X")) else int(v)]
    return [int(v)]
