#!/usr/bin/env python
"""Dev-time generator for the 1-D forward-transform networks (Daala lifting DCT/DST).

rav1e states its 1-D forward transforms as ~1400 lines of straight-line lifting steps over a
`TxOperations` value type (src/transform/forward_shared.rs:399-1796).  Hand-transcribing them
is error-prone, so this tool RESTATES them mechanically: it parses that restricted Rust
subset (let-bindings, tuple patterns, method chains, calls, slices), symbolically executes
every top-level 1-D transform with the primitive kernels implemented natively below
(butterflies :349-396, rotations :220-345, restated by hand from the reference), and emits a
flattened single-assignment op list per transform:

    oracle/txfm_networks.h                 plain C, used by the CPU oracle
    rav1e_b200/csrc/txfm_networks.cuh      CUDA device functions, used by the product kernels

Each emitted op carries the reference line number of the statement it came from.  The tool
needs /root/reference and is run by hand when (if ever) the networks change; its outputs are
committed.  Nothing at build, test or run time depends on it or on the reference tree.
"""
import os
import re
import sys

REF = "/root/reference/src/transform/forward_shared.rs"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# ------------------------------------------------------------------------- tokenizer
TOK = re.compile(r"\s*(?:(//[^\n]*)|(\d+)|([A-Za-z_][A-Za-z_0-9]*!?)|(::<|::|->|\.\.|&mut|[-+*/=<>(){}\[\],;:.&#$]))")


def tokenize(text, line0):
    toks, pos, line = [], 0, line0
    while pos < len(text):
        m = TOK.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                break
            raise SyntaxError(f"bad token at line {line}: {text[pos:pos + 40]!r}")
        line += text[pos:m.end()].count("\n") - (text[m.start(0):m.end()].lstrip() != text[m.start(0):m.end()] and 0)
        comment, num, ident, punct = m.groups()
        # recompute the line of the token start precisely
        tok_line = line0 + text[:m.end()].count("\n")
        if comment is None:
            if num is not None:
                toks.append(("num", int(num), tok_line))
            elif ident is not None:
                toks.append(("id", ident, tok_line))
            else:
                toks.append(("p", punct, tok_line))
        pos = m.end()
    return toks


# ------------------------------------------------------------------------- parser
class P:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", None, -1)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, kind, val=None):
        tk = self.peek()
        if tk[0] == kind and (val is None or tk[1] == val):
            self.i += 1
            return tk
        return None

    def expect(self, kind, val=None):
        tk = self.accept(kind, val)
        if not tk:
            raise SyntaxError(f"expected {kind} {val}, got {self.peek()}")
        return tk

    # ---- items
    def parse_fn(self):
        self.expect("id", "fn")
        name = self.expect("id")[1]
        if self.accept("p", "<"):
            depth = 1
            while depth:
                tk = self.next()
                if tk[1] == "<":
                    depth += 1
                elif tk[1] == ">":
                    depth -= 1
        self.expect("p", "(")
        params = []
        while not self.accept("p", ")"):
            pname = self.expect("id")[1]
            self.expect("p", ":")
            self.skip_type()
            params.append(pname)
            self.accept("p", ",")
        if self.accept("p", "->"):
            self.skip_type()
        body = self.parse_block()
        return name, params, body

    def skip_type(self):
        depth = 0
        while True:
            tk = self.peek()
            if depth == 0 and tk[1] in (",", ")", "{", "=", ";"):
                return
            if tk[1] in ("(", "[", "<"):
                depth += 1
            elif tk[1] in (")", "]", ">"):
                depth -= 1
            self.next()

    def skip_attrs(self):
        while self.peek()[1] == "#":
            self.next()
            self.expect("p", "[")
            depth = 1
            while depth:
                tk = self.next()
                if tk[1] == "[":
                    depth += 1
                elif tk[1] == "]":
                    depth -= 1
        # `$($s)*` macro fragment in front of fn
        while self.peek()[1] == "$":
            self.next()
            self.expect("p", "(")
            depth = 1
            while depth:
                tk = self.next()
                if tk[1] == "(":
                    depth += 1
                elif tk[1] == ")":
                    depth -= 1
            self.expect("p", "*")

    def parse_block(self):
        self.expect("p", "{")
        stmts = []
        while not self.accept("p", "}"):
            self.skip_attrs()
            tk = self.peek()
            if tk == ("eof", None, -1):
                raise SyntaxError("eof in block")
            if tk[1] == "fn":
                stmts.append(("fn", self.parse_fn(), tk[2]))
            elif tk[1] == "{":
                stmts.append(("block", self.parse_block(), tk[2]))
            elif tk[1] == "let":
                self.next()
                pat = self.parse_pat()
                if self.accept("p", ":"):
                    self.skip_type()
                self.expect("p", "=")
                e = self.parse_expr()
                self.expect("p", ";")
                stmts.append(("let", pat, e, tk[2]))
            elif tk[1] == "assert!":
                while self.next()[1] != ";":
                    pass
            elif tk[1] == "store_coeffs!":
                self.next()
                self.expect("p", "(")
                args = []
                while not self.accept("p", ")"):
                    args.append(self.parse_expr())
                    self.accept("p", ",")
                self.accept("p", ";")
                stmts.append(("store", args, tk[2]))
            else:
                e = self.parse_expr()
                if self.accept("p", "="):
                    rhs = self.parse_expr()
                    self.expect("p", ";")
                    stmts.append(("assign", e, rhs, tk[2]))
                else:
                    if not self.accept("p", ";"):
                        # trailing expression = return value
                        stmts.append(("ret", e, tk[2]))
                    else:
                        stmts.append(("expr", e, tk[2]))
        return stmts

    def parse_pat(self):
        if self.accept("p", "("):
            items = []
            while not self.accept("p", ")"):
                items.append(self.parse_pat())
                self.accept("p", ",")
            return ("tuple", items)
        self.accept("id", "mut")
        return ("name", self.expect("id")[1])

    # ---- expressions: additive over multiplicative over postfix
    def parse_expr(self):
        e = self.parse_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            e = ("bin", op, e, self.parse_mul())
        return e

    def parse_mul(self):
        e = self.parse_postfix()
        while self.peek()[1] == "*":
            self.next()
            e = ("bin", "*", e, self.parse_postfix())
        return e

    def parse_args(self):
        args = []
        while not self.accept("p", ")"):
            args.append(self.parse_expr())
            self.accept("p", ",")
        return args

    def parse_generics(self):
        gens = []
        while not self.accept("p", ">"):
            gens.append(self.expect("num")[1])
            self.accept("p", ",")
        return gens

    def parse_postfix(self):
        tk = self.next()
        if tk[0] == "num":
            e = ("num", tk[1])
        elif tk[1] == "&mut":
            return ("mutref", self.parse_postfix())
        elif tk[1] == "(":
            items = []
            while not self.accept("p", ")"):
                items.append(self.parse_expr())
                self.accept("p", ",")
            e = items[0] if len(items) == 1 else ("tuple", items)
        elif tk[1] == "[":  # array repeat literal [x; N]
            depth = 1
            while depth:
                t2 = self.next()
                if t2[1] == "[":
                    depth += 1
                elif t2[1] == "]":
                    depth -= 1
            e = ("arrayinit",)
        elif tk[0] == "id":
            path = [tk[1]]
            gens = []
            while True:
                if self.accept("p", "::"):
                    path.append(self.expect("id")[1])
                elif self.accept("p", "::<"):
                    gens = self.parse_generics()
                else:
                    break
            if self.accept("p", "("):
                e = ("call", path, gens, self.parse_args())
            else:
                e = ("var", path[0]) if len(path) == 1 else ("path", path)
        else:
            raise SyntaxError(f"unexpected token {tk}")
        while True:
            if self.accept("p", "."):
                t2 = self.next()
                if t2[0] == "num":
                    e = ("field", e, t2[1])
                else:
                    gens = []
                    if self.accept("p", "::<"):
                        gens = self.parse_generics()
                    self.expect("p", "(")
                    e = ("method", e, t2[1], gens, self.parse_args())
            elif self.accept("p", "["):
                lo = self.parse_expr()
                if self.accept("p", ".."):
                    hi = self.parse_expr()
                    self.expect("p", "]")
                    e = ("slice", e, lo, hi)
                else:
                    self.expect("p", "]")
                    e = ("index", e, lo)
            else:
                return e


# ------------------------------------------------------------------------- symbolic machine
class Sym:
    __slots__ = ("id",)

    def __init__(self, i):
        self.id = i


class View:
    """&mut [T] into a python list."""

    def __init__(self, base, off=0):
        self.base, self.off = base, off


class Machine:
    def __init__(self, fns):
        self.fns = fns
        self.ops = []   # (dst_id, op, a_id, b_id|None, imm tuple, line)
        self.n = 0
        self.line = 0

    def new_input(self):
        s = Sym(self.n)
        self.n += 1
        return s

    def emit(self, op, a, b=None, imm=()):
        s = Sym(self.n)
        self.n += 1
        self.ops.append((s.id, op, a.id, None if b is None else b.id, imm, self.line))
        return s

    # primitives ---------------------------------------------------------------
    def add(self, a, b): return self.emit("add", a, b)
    def sub(self, a, b): return self.emit("sub", a, b)
    def add_avg(self, a, b): return self.emit("add_avg", a, b)
    def sub_avg(self, a, b): return self.emit("sub_avg", a, b)
    def rshift1(self, a): return self.emit("rshift1", a)
    def tx_mul(self, a, shift, mul): return self.emit("tx_mul", a, None, (mul, shift))

    # forward_shared.rs:349-396
    def butterfly_add(self, p0, p1):
        p0 = self.add(p0, p1)
        p0h = self.rshift1(p0)
        p1h = self.sub(p1, p0h)
        return ((p0h, p0), p1h)

    def butterfly_sub(self, p0, p1):
        p0 = self.sub(p0, p1)
        p0h = self.rshift1(p0)
        p1h = self.add(p1, p0h)
        return ((p0h, p0), p1h)

    def butterfly_neg(self, p0, p1):
        p1 = self.sub(p0, p1)
        p1h = self.rshift1(p1)
        p0h = self.sub(p0, p1h)
        return (p0h, (p1h, p1))

    def butterfly_add_asym(self, p0, p1h):
        p1 = self.add(p1h, p0[0])
        p0n = self.sub(p0[1], p1)
        return (p0n, p1)

    def butterfly_sub_asym(self, p0, p1h):
        p1 = self.sub(p1h, p0[0])
        p0n = self.add(p0[1], p1)
        return (p0n, p1)

    def butterfly_neg_asym(self, p0h, p1):
        p0 = self.add(p0h, p1[0])
        p1n = self.sub(p0, p1[1])
        return (p0, p1n)

    # forward_shared.rs:220-345
    PI4 = {"RotatePi4Add": ("add", "sub"), "RotatePi4AddAvg": ("add_avg", "sub"),
           "RotatePi4Sub": ("sub", "add"), "RotatePi4SubAvg": ("sub_avg", "add")}
    ROT = {"RotateAdd": ("add", "sub", False), "RotateAddAvg": ("add_avg", "sub", False),
           "RotateAddShift": ("add", "sub", True), "RotateSub": ("sub", "add", False),
           "RotateSubAvg": ("sub_avg", "add", False), "RotateSubShift": ("sub", "add", True)}
    NEG = {"RotateNeg": "sub", "RotateNegAvg": "sub_avg"}

    def rotate(self, kind, fn, gens, args):
        if kind in self.PI4:
            ADD, SUB = self.PI4[kind]
            p0, p1, m = args
            t = getattr(self, ADD)(p1, p0)
            a = self.tx_mul(p0, gens[0], m[0])
            out0 = self.tx_mul(t, gens[1], m[1])
            out1 = getattr(self, SUB)(a, out0)
            return (out0, out1)
        if kind in self.ROT:
            ADD, SUB, SHIFT = self.ROT[kind]
            p0, p1, m = args
            if fn == "kernel":
                p0 = (p0, p0)
            t = getattr(self, ADD)(p1, p0[0])
            a = self.tx_mul(p0[1], gens[0], m[0])
            b = self.tx_mul(p1, gens[1], m[1])
            c = self.tx_mul(t, gens[2], m[2])
            out0 = self.add(b, c)
            shifted = self.rshift1(c) if SHIFT else c
            out1 = getattr(self, SUB)(a, shifted)
            return (out0, out1)
        if kind in self.NEG:
            ADD = self.NEG[kind]
            p0, p1, m = args
            t = getattr(self, ADD)(p0, p1)
            a = self.tx_mul(p0, gens[0], m[0])
            b = self.tx_mul(p1, gens[1], m[1])
            c = self.tx_mul(t, gens[2], m[2])
            return (self.sub(b, c), self.sub(c, a))
        raise KeyError(kind)

    # interpreter --------------------------------------------------------------
    def call(self, name, args):
        if hasattr(self, name) and name.startswith("butterfly"):
            return getattr(self, name)(*args)
        params, body = self.fns[name]
        env = dict(zip(params, args))
        return self.run_block(body, env, dict(self.fns))

    def run_block(self, stmts, env, fns):
        saved = self.fns
        self.fns = fns
        ret = None
        for st in stmts:
            kind = st[0]
            self.line = st[-1]
            if kind == "fn":
                name, params, body = st[1]
                fns[name] = (params, body)
            elif kind == "block":
                self.run_block(st[1], env, fns)
            elif kind == "let":
                self.bind(st[1], self.ev(st[2], env), env)
            elif kind == "store":
                tgt = self.ev(st[1][0], env)
                for i, e in enumerate(st[1][1:]):
                    self.store(tgt, i, self.ev(e, env))
            elif kind == "assign":
                lhs = st[1]
                assert lhs[0] == "index"
                self.store(self.ev(lhs[1], env), self.ev(lhs[2], env), self.ev(st[2], env))
            elif kind == "expr":
                self.ev(st[1], env)
            elif kind == "ret":
                ret = self.ev(st[1], env)
        self.fns = saved
        return ret

    def store(self, tgt, i, val):
        if isinstance(tgt, View):
            tgt.base[tgt.off + i] = val
        else:
            tgt[i] = val

    def bind(self, pat, val, env):
        if pat[0] == "name":
            env[pat[1]] = val
        else:
            assert len(pat[1]) == len(val), (pat, val)
            for p, v in zip(pat[1], val):
                self.bind(p, v, env)

    def ev(self, e, env):
        k = e[0]
        if k == "num":
            return e[1]
        if k == "var":
            return env[e[1]]
        if k == "tuple":
            return tuple(self.ev(x, env) for x in e[1])
        if k == "arrayinit":
            return [None] * 64
        if k == "mutref":
            v = self.ev(e[1], env)
            return v if isinstance(v, View) else View(v, 0)
        if k == "field":
            return self.ev(e[1], env)[e[2]]
        if k == "index":
            arr, i = self.ev(e[1], env), self.ev(e[2], env)
            return arr.base[arr.off + i] if isinstance(arr, View) else arr[i]
        if k == "slice":
            arr, lo = self.ev(e[1], env), self.ev(e[2], env)
            hi = self.ev(e[3], env)
            v = View(arr.base, arr.off + lo) if isinstance(arr, View) else View(arr, lo)
            v.len = hi - lo
            return v
        if k == "bin":
            a, b = self.ev(e[2], env), self.ev(e[3], env)
            return {"+": a + b, "-": a - b, "*": a * b}[e[1]]
        if k == "method":
            recv, name, gens, args = self.ev(e[1], env), e[2], e[3], [self.ev(x, env) for x in e[4]]
            if name == "reverse":
                n = recv.len
                seg = recv.base[recv.off:recv.off + n]
                recv.base[recv.off:recv.off + n] = seg[::-1]
                return None
            if name == "tx_mul":
                return self.tx_mul(recv, gens[0], args[0])
            if name == "rshift1":
                return self.rshift1(recv)
            if name in ("add", "sub", "add_avg", "sub_avg"):
                return getattr(self, name)(recv, args[0])
            raise KeyError(name)
        if k == "call":
            path, gens, args = e[1], e[2], [self.ev(x, env) for x in e[3]]
            if len(path) == 2 and path[1] in ("kernel", "half_kernel"):
                return self.rotate(path[0], path[1], gens, args)
            assert len(path) == 1, path
            return self.call(path[0], args)
        raise KeyError(e)


# ------------------------------------------------------------------------- driver
TOP = [("fdct4", "daala_fdct4", 4), ("fdct8", "daala_fdct8", 8), ("fdct16", "daala_fdct16", 16),
       ("fdct32", "daala_fdct32", 32), ("fdct64", "daala_fdct64", 64),
       ("fdst_vii_4", "daala_fdst_vii_4", 4), ("fdst8", "daala_fdst8", 8),
       ("fdst16", "daala_fdst16", 16), ("fwht4", "fwht4", 4)]


def load_fns():
    src = open(REF).read()
    start = src.index("fn daala_fdct_ii_2_asym")
    line0 = src[:start].count("\n") + 1
    # walk back to the beginning of that line's attribute block is unnecessary: fns parse from `fn`
    text = src[start:]
    fns = {}
    # split on top-level `fn ` occurrences at macro depth: parse sequentially
    toks = tokenize(text, line0)
    p = P(toks)
    while True:
        p.skip_attrs()
        tk = p.peek()
        if tk[1] == "fn":
            name, params, body = p.parse_fn()
            fns[name] = (params, body)
        elif tk[1] == "}" or tk[0] == "eof":
            break
        else:
            raise SyntaxError(f"unexpected top-level token {tk}")
    return fns


def trace(fns, fn_name, n):
    m = Machine(fns)
    ins = [m.new_input() for _ in range(n)]
    coeffs = list(ins)
    m.call(fn_name, [View(coeffs, 0)])
    return m, [s.id for s in coeffs]


def dce(ops, outs):
    live = set(outs)
    keep = []
    for op in reversed(ops):
        if op[0] in live:
            keep.append(op)
            live.add(op[2])
            if op[3] is not None:
                live.add(op[3])
    return keep[::-1]


def emit_body(name, n, ops, outs, indent="  "):
    def v(i):
        return f"c[{i}]" if False else (f"i{i}" if i < n else f"v{i}")
    lines = []
    lines.append(f"{indent}const TXV " + ", ".join(f"i{k} = c[{k}]" for k in range(n)) + ";")
    last_line = None
    for dst, op, a, b, imm, line in ops:
        cite = f"  /* :{line} */" if line != last_line else ""
        last_line = line
        if op == "tx_mul":
            lines.append(f"{indent}const TXV v{dst} = TX_MUL({v(a)}, {imm[0]}, {imm[1]});{cite}")
        elif op == "rshift1":
            lines.append(f"{indent}const TXV v{dst} = TX_RSHIFT1({v(a)});{cite}")
        else:
            lines.append(f"{indent}const TXV v{dst} = TX_{op.upper()}({v(a)}, {v(b)});{cite}")
    for k, o in enumerate(outs):
        lines.append(f"{indent}c[{k}] = {v(o)};")
    return lines


HEADER = """/* GENERATED by tools/gen_txfm_networks.py -- do not edit by hand.
 *
 * Flattened single-assignment restatement of rav1e's 1-D forward transform networks
 * (Daala lifting DCT-II / DST-IV / DST-VII / WHT), src/transform/forward_shared.rs:399-1796
 * @ 564ae3b.  The trailing `:NNNN` comments give the reference line of the statement each op
 * was derived from; primitives follow :220-396 and the i32 TxOperations of
 * src/transform/forward.rs:37-65 (TX_MUL = (x*m + (1<<s>>1)) >> s with wrapping multiply,
 * TX_RSHIFT1 rounds toward zero, TX_ADD_AVG / TX_SUB_AVG floor).
 * Output is in natural frequency order (the reference's final bit-reversal is folded in).
 */
"""


def main():
    fns = load_fns()
    bodies = []
    stats = []
    for short, full, n in TOP:
        m, outs = trace(fns, full, n)
        ops = dce(m.ops, outs)
        stats.append((short, n, len(ops)))
        bodies.append((short, n, emit_body(short, n, ops, outs)))

    c_out = [HEADER, "#ifndef ORC_TXFM_NETWORKS_H", "#define ORC_TXFM_NETWORKS_H", "#include <stdint.h>",
             "typedef int32_t TXV;",
             "#define TX_ADD(a, b) ((TXV)((uint32_t)(a) + (uint32_t)(b)))",
             "#define TX_SUB(a, b) ((TXV)((uint32_t)(a) - (uint32_t)(b)))",
             "#define TX_ADD_AVG(a, b) (TX_ADD(a, b) >> 1)",
             "#define TX_SUB_AVG(a, b) (TX_SUB(a, b) >> 1)",
             "#define TX_RSHIFT1(a) (((a) + ((a) < 0)) >> 1)",
             "#define TX_MUL(a, m, s) ((TXV)((uint32_t)(a) * (uint32_t)(m) + (1u << (s) >> 1)) >> (s))",
             ""]
    for short, n, lines in bodies:
        c_out.append(f"static inline void orc_{short}(TXV *c) {{")
        c_out += lines
        c_out.append("}")
        c_out.append("")
    c_out.append("#endif")
    with open(os.path.join(ROOT, "oracle", "txfm_networks.h"), "w") as f:
        f.write("\n".join(c_out) + "\n")

    cu = [HEADER, "#pragma once",
          "typedef int TXV;",
          "#define TX_ADD(a, b) ((a) + (b))",
          "#define TX_SUB(a, b) ((a) - (b))",
          "#define TX_ADD_AVG(a, b) (((a) + (b)) >> 1)",
          "#define TX_SUB_AVG(a, b) (((a) - (b)) >> 1)",
          "#define TX_RSHIFT1(a) (((a) + (int)((unsigned)(a) >> 31)) >> 1)",
          "#define TX_MUL(a, m, s) (((a) * (m) + (1 << (s) >> 1)) >> (s))",
          ""]
    for short, n, lines in bodies:
        cu.append(f"__device__ __forceinline__ void tx_{short}(TXV (&c)[{n}]) {{")
        cu += lines
        cu.append("}")
        cu.append("")
    with open(os.path.join(ROOT, "rav1e_b200", "csrc", "txfm_networks.cuh"), "w") as f:
        f.write("\n".join(cu) + "\n")
    for s in stats:
        print("%-12s n=%2d ops=%d" % s)


if __name__ == "__main__":
    main()
