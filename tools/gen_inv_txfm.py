#!/usr/bin/env python
"""Dev-time generator for the 1-D INVERSE transform networks (AV1-normative butterflies).

rav1e states av1_idct4..64 / av1_iadst4..16 as ~1500 lines of staged array literals
(src/transform/inverse.rs:71-1591): `let stgK = [half_btf(..), clamp_value(a + b, range), ..]`,
nested calls into the half-size DCT, and `output[i] = ..`.  Like tools/gen_txfm_networks.py for
the forward direction, this tool RESTATES them mechanically: it parses that restricted Rust subset,
symbolically executes every function (arrays become lists of value ids, nested calls are inlined)
and emits one flattened single-assignment program per transform into

    oracle/inv_txfm_networks.h              plain C, used by the CPU oracle (oracle/inv_txfm.c)
    rav1e_b200/csrc/inv_txfm_networks.cuh   CUDA device functions (rav1e_b200/csrc/inv_txfm.cu)

with every op tagged by the reference line it came from.  The primitives (half_btf with its
wrapping arithmetic, clamp_value, round_shift: transform/mod.rs:296-315) are written by hand in
oracle/inv_txfm.c; the cosine tables are regenerated from their closed forms and compared with the
reference's at generation time.  Needs /root/reference; run by hand; outputs are committed.
Nothing at build, test or run time depends on this tool or on the reference tree.
"""
import math
import os
import re

REF = "/root/reference/src/transform/inverse.rs"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "oracle", "inv_txfm_networks.h")
OUT_CU = os.path.join(ROOT, "rav1e_b200", "csrc", "inv_txfm_networks.cuh")

TOK = re.compile(r"\s*(?:(//[^\n]*)|(\d+)|([A-Za-z_][A-Za-z_0-9]*!?)|(\.\.|&mut|[-+*/=<>(){}\[\],;:.&|]))")


def tokenize(text, line0):
    toks, pos = [], 0
    while pos < len(text):
        m = TOK.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                break
            raise SyntaxError(f"bad token near line {line0 + text[:pos].count(chr(10))}: {text[pos:pos + 40]!r}")
        comment, num, ident, punct = m.groups()
        line = line0 + text[:m.end()].count("\n")
        if comment is None:
            toks.append(("num", int(num), line) if num is not None else
                        ("id", ident, line) if ident is not None else ("p", punct, line))
        pos = m.end()
    return toks


class Parser:
    """Statements of one function body -> a list of (kind, ...) tuples with expression trees."""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", None, -1)

    def next(self):
        self.i += 1
        return self.t[self.i - 1]

    def accept(self, kind, val=None):
        tk = self.peek()
        if tk[0] == kind and (val is None or tk[1] == val):
            self.i += 1
            return tk
        return None

    def expect(self, kind, val=None):
        tk = self.accept(kind, val)
        if not tk:
            raise SyntaxError(f"expected {kind} {val!r}, got {self.peek()}")
        return tk

    def statements(self):
        out = []
        while self.peek()[0] != "eof":
            tk = self.peek()
            if tk == ("id", "assert!", tk[2]):
                self.next()
                self.skip_parens()
                self.expect("p", ";")
            elif tk[0] == "id" and tk[1] == "let":
                self.next()
                self.accept("id", "mut")
                name = self.expect("id")[1]
                if self.accept("p", ":"):                     # `: [i32; K]`
                    self.expect("p", "[")
                    self.expect("id", "i32")
                    self.expect("p", ";")
                    self.expect("num")
                    self.expect("p", "]")
                self.expect("p", "=")
                line = self.peek()[2]
                if self.peek() == ("p", "[", self.peek()[2]):
                    self.next()
                    first = self.expr()
                    if self.accept("p", ";"):                 # `[0; K]`
                        k = self.expect("num")[1]
                        self.expect("p", "]")
                        out.append(("zeros", name, k, line))
                    else:
                        items = [first]
                        while self.accept("p", ","):
                            if self.peek()[1] == "]":
                                break
                            items.append(self.expr())
                        self.expect("p", "]")
                        out.append(("array", name, items, line))
                else:
                    out.append(("scalar", name, self.expr(), line))
                self.expect("p", ";")
            elif tk[0] == "id" and tk[1] == "output" and self.peek(1)[1] == "[" and self.peek(2)[0] == "num":
                self.next()
                self.next()
                idx = self.expect("num")[1]
                self.expect("p", "]")
                self.expect("p", "=")
                line = self.peek()[2]
                out.append(("out", idx, self.expr(), line))
                self.expect("p", ";")
            elif tk[0] == "id" and tk[1].startswith("av1_"):   # nested call f(&a, &mut b, range)
                fn = self.next()[1]
                self.expect("p", "(")
                self.expect("p", "&")
                src = self.expect("id")[1]
                self.expect("p", ",")
                self.expect("p", "&mut")
                dst = self.expect("id")[1]
                self.expect("p", ",")
                self.expect("id", "range")
                self.expect("p", ")")
                self.expect("p", ";")
                out.append(("call", fn, src, dst, tk[2]))
            else:
                raise SyntaxError(f"unexpected statement at {tk}")
        return out

    def skip_parens(self):
        self.expect("p", "(")
        depth = 1
        while depth:
            tk = self.next()
            depth += (tk[1] == "(") - (tk[1] == ")")

    # expr := term (('+'|'-') term)* ; term := unary ('*' unary)* ; unary := '-' unary | atom
    def expr(self):
        node = self.term()
        while self.peek()[0] == "p" and self.peek()[1] in "+-":
            op = self.next()[1]
            node = ("bin", op, node, self.term())
        return node

    def term(self):
        node = self.unary()
        while self.peek() == ("p", "*", self.peek()[2]):
            self.next()
            node = ("bin", "*", node, self.unary())
        return node

    def unary(self):
        if self.accept("p", "-"):
            return ("neg", self.unary())
        return self.atom()

    def atom(self):
        tk = self.next()
        if tk[0] == "num":
            return ("const", tk[1])
        if tk == ("p", "(", tk[2]):
            node = self.expr()
            self.expect("p", ")")
            return node
        if tk[0] == "id":
            name = tk[1]
            if self.accept("p", "["):
                idx = self.expect("num")[1]
                self.expect("p", "]")
                return ("index", name, idx)
            if self.accept("p", "("):
                args = []
                if not self.accept("p", ")"):
                    args.append(self.expr())
                    while self.accept("p", ","):
                        args.append(self.expr())
                    self.expect("p", ")")
                return ("call", name, args)
            return ("var", name)
        raise SyntaxError(f"bad atom {tk}")


def cospi_table():
    return [int(round(4096 * math.cos(k * math.pi / 128))) for k in range(64)]


def sinpi_table():
    return [0] + [int(round(4096 * 2 * math.sqrt(2) / 3 * math.sin(k * math.pi / 9))) for k in range(1, 5)]


class Emitter:
    """Symbolic execution: every computed value gets an id and one emitted C line."""

    def __init__(self, fns, cospi, sinpi):
        self.fns, self.cospi, self.sinpi = fns, cospi, sinpi

    def run(self, fn, n):
        self.lines, self.nid = [], 0
        inputs = [f"in[{i}]" for i in range(n)]
        outs = self.exec_fn(fn, inputs, n)
        return self.lines, outs

    def fresh(self, rhs, line):
        name = f"v{self.nid}"
        self.nid += 1
        self.lines.append(f"  const int32_t {name} = {rhs}; /* L{line} */")
        return name

    def exec_fn(self, fn, inputs, n):
        env = {"input": list(inputs)}
        outputs = [None] * n
        for st in self.fns[fn]:
            kind = st[0]
            if kind == "zeros":
                env[st[1]] = ["0"] * st[2]
            elif kind == "array":
                env[st[1]] = [self.ev(e, env, st[3]) for e in st[2]]
            elif kind == "scalar":
                env[st[1]] = self.ev(st[2], env, st[3])
            elif kind == "out":
                outputs[st[1]] = self.ev(st[2], env, st[3])
            elif kind == "call":
                _, callee, src, dst, line = st
                res = self.exec_fn(callee, env[src], len(env[src]))
                env[dst] = res
        assert all(o is not None for o in outputs), fn
        return outputs

    def ev(self, e, env, line):
        k = e[0]
        if k == "const":
            return str(e[1])
        if k == "var":
            if e[1] in ("range", "INV_COS_BIT"):
                return e[1]
            return env[e[1]]
        if k == "index":
            if e[1] == "COSPI_INV":
                return str(self.cospi[e[2]])
            if e[1] == "SINPI_INV":
                return str(self.sinpi[e[2]])
            return env[e[1]][e[2]]
        if k == "neg":
            v = self.ev(e[1], env, line)
            if re.fullmatch(r"-?\d+", v):
                return str(-int(v))
            return self.fresh(f"WNEG({v})", line)
        if k == "bin":
            a, b = self.ev(e[2], env, line), self.ev(e[3], env, line)
            return self.fresh({"+": "WADD", "-": "WSUB", "*": "WMUL"}[e[1]] + f"({a}, {b})", line)
        if k == "call":
            args = [self.ev(a, env, line) for a in e[2]]
            if e[1] == "half_btf":
                assert args[4] == "INV_COS_BIT"
                return self.fresh(f"HALF_BTF({args[0]}, {args[1]}, {args[2]}, {args[3]})", line)
            if e[1] == "clamp_value":
                assert args[1] == "range"
                return self.fresh(f"CLAMP_VALUE({args[0]}, range)", line)
            if e[1] == "round_shift":
                return self.fresh(f"ROUND_SHIFT({args[0]}, {args[1]})", line)
        raise ValueError(e)


def main():
    src = open(REF).read()
    # the closed forms reproduce the reference's tables (compared here, at generation time)
    ref_cos = [int(x) for x in re.findall(r"\d+", src[src.index("static COSPI_INV"):].split("];")[0].split("=")[1])]
    ref_sin = [int(x) for x in re.findall(r"\d+", src[src.index("static SINPI_INV"):].split("];")[0].split("=")[1])]
    assert ref_cos == cospi_table(), "COSPI_INV closed form mismatch"
    assert ref_sin == sinpi_table(), "SINPI_INV closed form mismatch"
    names = {"av1_idct4": 4, "av1_idct8": 8, "av1_idct16": 16, "av1_idct32": 32, "av1_idct64": 64,
             "av1_iadst4": 4, "av1_iadst8": 8, "av1_iadst16": 16}
    fns = {}
    for name in names:
        m = re.search(r"fn %s\(input: &\[i32\], output: &mut \[i32\], _?range: usize\) \{\n" % name, src)
        start = m.end()
        end = src.index("\n}\n", start)
        line0 = src[:start].count("\n") + 1
        fns[name] = Parser(tokenize(src[start:end], line0)).statements()
    em = Emitter(fns, cospi_table(), sinpi_table())
    out = ["/* GENERATED by tools/gen_inv_txfm.py from the staged butterfly listings of rav1e",
           " * src/transform/inverse.rs:71-1591 (xiph/rav1e @ 564ae3b) - do not edit.  One flattened",
           " * single-assignment program per 1-D inverse transform; `L<n>` = reference line of the",
           " * statement an op came from.  Primitives (WADD/WSUB/WMUL/WNEG wrap like Rust release builds;",
           " * HALF_BTF, CLAMP_VALUE, ROUND_SHIFT: transform/mod.rs:296-315) are defined by the includer.",
           " * TEST INFRASTRUCTURE ONLY (see oracle.h). */", ""]
    cu = ["// GENERATED by tools/gen_inv_txfm.py from the staged butterfly listings of rav1e",
          "// src/transform/inverse.rs:71-1591 (xiph/rav1e @ 564ae3b) - do not edit.  One flattened",
          "// single-assignment program per 1-D inverse transform over a register array; `L<n>` = reference",
          "// line of the statement an op came from.  B200_HDI (function qualifiers), WADD/WSUB/WMUL/WNEG, HALF_BTF,",
          "// CLAMP_VALUE, ROUND_SHIFT are defined by the includer (inv_txfm.cu).", ""]
    for name, n in names.items():
        lines, outs = em.run(name, n)
        cu.append(f"B200_HDI void d_{name}(const int (&in)[{n}], int (&out)[{n}], int range) {{")
        if not any("range" in l for l in lines):
            cu.append("  (void)range;")
        cu += [l.replace("const int32_t", "const int") for l in lines]
        cu += [f"  out[{i}] = {o};" for i, o in enumerate(outs)]
        cu += ["}", ""]
        uses_range = any("range" in l for l in lines)
        out.append(f"static void {name}(const int32_t *in, int32_t *out, int range) {{")
        if not uses_range:
            out.append("  (void)range;")
        if name == "av1_iadst4":
            lines = [l.replace("bit", "12") for l in lines]
        out += lines
        for i, o in enumerate(outs):
            out.append(f"  out[{i}] = {o};")
        out.append("}")
        out.append("")
        print(f"{name}: {len(lines)} ops")
    with open(OUT, "w") as f:
        f.write("\n".join(out))
    with open(OUT_CU, "w") as f:
        f.write("\n".join(cu))
    print("wrote", OUT, "and", OUT_CU)


if __name__ == "__main__":
    main()
