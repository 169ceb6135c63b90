#!/usr/bin/env python3
"""Extract the piecewise fit data of the reference's Rys::root1..root5 into neutral tables.

TEST-INFRASTRUCTURE / BUILD TOOL.  Runs only where /root/reference exists (this container).

The reference evaluates Rys roots/weights for 1..5 roots with the classic piecewise fits in
X (Rys.cpp:314-2197).  Bit-for-tolerance parity (1e-12 per integral) needs the same fit
*coefficients* in the CHECKER; oracle/rys_roots_oracle.c is emitted from here.  The product's evaluator
(unomol_b200/csrc/rys_roots.cuh + tools/gen_rys_tables.py) is independent of this script and of the
reference's fits.  This script:

  1. parses the five functions with a small C-expression parser,
  2. symbolically executes each X band (probe x between consecutive breakpoints),
  3. rewrites every assignment as an expression over *Laurent polynomials* in one variable
     (x, or the shifted y = x - x0); the polynomial coefficients are pulled out exactly
     (Horner steps only shift/insert coefficients, no arithmetic on them; the script asserts that
     no power ever receives two contributions),
  4. emits (a) a human-readable IR listing (--ir) used to write the evaluators by hand and
     (b) the oracle's C source (--emit-c PATH).

Usage:  python oracle/tools/rys_extract.py --ref /root/reference --ir
"""
import argparse
import re
import sys
from fractions import Fraction

TOK = re.compile(r"\s*(?:(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+(?:[eE][-+]?\d+)?)|([A-Za-z_]\w*)|(<=|[-+*/()=;{}\[\],<>]))")


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    s = re.sub(r"//[^\n]*", " ", s)
    return s


def tokenize(s):
    out, pos = [], 0
    s = s.strip()
    while pos < len(s):
        m = TOK.match(s, pos)
        if not m:
            if s[pos:].strip() == "":
                break
            raise SyntaxError("bad token at %r" % s[pos:pos + 40])
        if m.group(1) is not None:
            out.append(("num", m.group(1)))
        elif m.group(2) is not None:
            out.append(("id", m.group(2)))
        else:
            out.append(("op", m.group(3)))
        pos = m.end()
    return out


# ---------------------------------------------------------------- AST
class Num:
    def __init__(self, text): self.text = text; self.val = float(text)
    def __repr__(self): return self.text
class Var:
    def __init__(self, name): self.name = name
    def __repr__(self): return self.name
class Bin:
    def __init__(self, op, a, b): self.op, self.a, self.b = op, a, b
    def __repr__(self): return "(%r %s %r)" % (self.a, self.op, self.b)
class Neg:
    def __init__(self, a): self.a = a
    def __repr__(self): return "(-%r)" % (self.a,)
class Call:
    def __init__(self, fn, a): self.fn, self.a = fn, a
    def __repr__(self): return "%s(%r)" % (self.fn, self.a)


class Parser:
    def __init__(self, toks): self.t, self.i = toks, 0
    def peek(self): return self.t[self.i] if self.i < len(self.t) else (None, None)
    def next(self): tok = self.t[self.i]; self.i += 1; return tok
    def expect(self, v):
        tok = self.next()
        if tok[1] != v: raise SyntaxError("expected %s got %s at %d" % (v, tok, self.i))
    # expression grammar
    def expr(self):
        a = self.term()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]; b = self.term(); a = Bin(op, a, b)
        return a
    def term(self):
        a = self.unary()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]; b = self.unary(); a = Bin(op, a, b)
        return a
    def unary(self):
        if self.peek()[1] == "-":
            self.next(); return Neg(self.unary())
        if self.peek()[1] == "+":
            self.next(); return self.unary()
        return self.atom()
    def atom(self):
        k, v = self.next()
        if k == "num": return Num(v)
        if k == "id":
            if self.peek()[1] == "(":
                self.next(); a = self.expr(); self.expect(")"); return Call(v, a)
            if self.peek()[1] == "[":
                self.next(); idx = self.next(); self.expect("]"); return Var("%s[%s]" % (v, idx[1]))
            return Var(v)
        if v == "(":
            a = self.expr(); self.expect(")"); return a
        raise SyntaxError("atom %s" % (v,))
    # statements
    def block(self):
        stmts = []
        self.expect("{")
        while self.peek()[1] != "}":
            stmts.append(self.stmt())
        self.expect("}")
        return stmts
    def stmt(self):
        k, v = self.peek()
        if v == "{": return ("block", self.block())
        if v == "if":
            self.next(); self.expect("("); lhs = self.next(); self.expect("<="); c = self.next(); self.expect(")")
            assert lhs[1] == "x"
            th = self.stmt()
            el = None
            if self.peek()[1] == "else":
                self.next(); el = self.stmt()
            return ("if", float(c[1]), th, el)
        if v == "return":
            self.next(); self.expect(";"); return ("return",)
        if v in ("const", "double"):
            # declarations: 'const double a = 1.0;' or 'double a, b, c;'
            consts = []
            if v == "const": self.next()
            self.expect("double")
            while True:
                name = self.next()[1]
                if self.peek()[1] == "=":
                    self.next(); e = self.expr(); consts.append((name, e))
                if self.peek()[1] == ",": self.next(); continue
                break
            self.expect(";")
            return ("decl", consts)
        # assignment (also 'a -= b')
        k, name = self.next()
        if self.peek()[1] == "[":
            self.next(); idx = self.next(); self.expect("]"); name = "%s[%s]" % (name, idx[1])
        op = self.next()[1]
        if op in ("-", "+"):   # -= / +=
            self.expect("="); e = self.expr(); self.expect(";")
            return ("assign", name, Bin(op, Var(name), e))
        assert op == "=", (name, op)
        e = self.expr(); self.expect(";")
        return ("assign", name, e)


def function_body(src, n):
    m = re.search(r"Rys::root%d\s*\(double x\)\s*noexcept\s*\{" % n, src)
    i = m.end() - 1
    depth, j = 0, i
    while True:
        if src[j] == "{": depth += 1
        elif src[j] == "}":
            depth -= 1
            if depth == 0: break
        j += 1
    return src[i:j + 1]


def breakpoints(stmts, acc):
    for s in stmts:
        if s[0] == "if":
            acc.add(s[1])
            breakpoints([s[2]], acc)
            if s[3]: breakpoints([s[3]], acc)
        elif s[0] == "block":
            breakpoints(s[1], acc)
    return acc


class Return(Exception):
    pass


def trace(stmts, x, out, consts):
    """Collect the assignments executed for probe value x."""
    for s in stmts:
        if s[0] == "block": trace(s[1], x, out, consts)
        elif s[0] == "decl":
            for name, e in s[1]: consts[name] = e
        elif s[0] == "assign": out.append((s[1], s[2]))
        elif s[0] == "return": raise Return()
        elif s[0] == "if":
            if x <= s[1]: trace([s[2]], x, out, consts)
            elif s[3]: trace([s[3]], x, out, consts)


# ---------------------------------------------------------------- Laurent-polynomial extraction
class Poly:
    """Laurent polynomial sum c_k v^k, coefficients kept as the ORIGINAL decimal strings (with sign)."""
    def __init__(self, var, coef): self.var, self.coef = var, dict(coef)
    def __repr__(self):
        ks = sorted(self.coef)
        return "P[%s;%d..%d]" % (self.var, ks[0], ks[-1])


def negstr(s): return s[1:] if s.startswith("-") else "-" + s


def to_poly(e, env):
    """Return Poly if e is a Laurent polynomial in a single variable with literal coefficients, else None.
    env maps variable names to ('var', base) for x / y / xinv."""
    if isinstance(e, Num): return Poly(None, {0: e.text})
    if isinstance(e, Var):
        if e.name == "x": return Poly("x", {1: "1"})
        if e.name == "xinv": return Poly("x", {-1: "1"})
        if e.name == "y": return Poly("y", {1: "1"})
        if e.name in env: return Poly(None, {0: env[e.name]})
        return None
    if isinstance(e, Neg):
        p = to_poly(e.a, env)
        if p is None: return None
        return Poly(p.var, {k: negstr(v) for k, v in p.coef.items()})
    if isinstance(e, Bin):
        a, b = to_poly(e.a, env), to_poly(e.b, env)
        if a is None or b is None: return None
        var = a.var or b.var
        if a.var and b.var and a.var != b.var: return None
        if e.op in "+-":
            co = dict(a.coef)
            for k, v in b.coef.items():
                if k in co:
                    # x + x  -> 2x is the only legal merge (exact)
                    if co[k] == "1" and v == "1" and e.op == "+": co[k] = "2"; continue
                    return None
                co[k] = v if e.op == "+" else negstr(v)
            return Poly(var, co)
        if e.op == "*":
            # monomial * poly only (Horner step) or const * poly with const == +-1/monomial
            for m, p in ((a, b), (b, a)):
                if len(m.coef) == 1:
                    (km, cm), = m.coef.items()
                    if cm in ("1",):
                        return Poly(var, {k + km: v for k, v in p.coef.items()})
            # literal * monomial(1)  e.g. 'x * -8.36e-8' handled above (monomial x, poly const). else:
            return None
        if e.op == "/":
            if len(b.coef) == 1:
                (kb, cb), = b.coef.items()
                if cb == "1":
                    return Poly(var, {k - kb: v for k, v in a.coef.items()})
            return None
    return None


def abstract(e, env, polys):
    """Replace maximal polynomial subtrees (degree span >= 2 terms, in x or y) by table references."""
    p = to_poly(e, env)
    if p is not None and p.var is not None and len(p.coef) >= 2:
        polys.append(p)
        return Var("@P%d" % (len(polys) - 1))
    if isinstance(e, Bin): return Bin(e.op, abstract(e.a, env, polys), abstract(e.b, env, polys))
    if isinstance(e, Neg): return Neg(abstract(e.a, env, polys))
    if isinstance(e, Call): return Call(e.fn, abstract(e.a, env, polys))
    return e


def analyse(ref_dir):
    src = strip_comments(open(ref_dir + "/Rys.cpp").read())
    result = {}
    for n in range(1, 6):
        body = function_body(src, n)
        stmts = Parser(tokenize(body)).block()
        bps = sorted(breakpoints(stmts, set()))
        bands = []
        edges = [0.0] + bps + [float("inf")]
        for lo, hi in zip(edges[:-1], edges[1:]):
            probe = (lo + hi) / 2 if hi != float("inf") else lo * 2 + 1
            if lo == 0.0: probe = hi / 2
            out, consts = [], {}
            try: trace(stmts, probe, out, consts)
            except Return: pass
            env = {k: repr(v) if not isinstance(v, Num) else v.text for k, v in consts.items()}
            y0 = None
            rows = []
            for name, e in out:
                if name == "y":
                    # y = x - c
                    assert isinstance(e, Bin) and e.op == "-" and isinstance(e.a, Var) and e.a.name == "x"
                    y0 = e.b.text; continue
                polys = []
                ae = abstract(e, env, polys)
                rows.append((name, ae, polys))
            bands.append(dict(lo=lo, hi=hi, y0=y0, rows=rows, consts=env))
        result[n] = bands
    return result


# ---------------------------------------------------------------- emitters
CONST_TABLE = None   # when a list, lit() appends the literal and returns a table reference (CUDA constant bank)


def lit(s):
    """decimal literal text -> C double literal (or a reference into the constant table for device code)"""
    if CONST_TABLE is not None:
        v = _lit(s)
        import struct
        if struct.unpack("<Q", struct.pack("<d", float(v)))[0] & 0xffffffff == 0:
            return v          # fits the 32-bit FP64 immediate form (1.0, 0.5, 7.5, ...)
        if v not in CONST_INDEX:
            CONST_INDEX[v] = len(CONST_TABLE); CONST_TABLE.append(v)
        return "RC(%d, %s)" % (CONST_INDEX[v], v)
    return _lit(s)


CONST_INDEX = {}


def _lit(s):
    neg = s.startswith("-")
    if neg: s = s[1:]
    if s.startswith("."): s = "0" + s
    if s.endswith("."): s = s + "0"
    if not re.search(r"[.eE]", s): s = s + ".0"
    return ("-" if neg else "") + s


def horner(coefs_desc, v, fma):
    """coefs_desc: literal strings, highest power first."""
    acc = lit(coefs_desc[0])
    for c in coefs_desc[1:]:
        acc = ("fma(%s, %s, %s)" % (acc, v, lit(c))) if fma else ("(%s * %s + %s)" % (acc, v, lit(c)))
    return acc


def emit_poly(p, fma):
    """Laurent polynomial -> C expression.  Non-negative powers: Horner in v; negative: Horner in xinv."""
    v = p.var
    ks = sorted(p.coef)
    lo, hi = ks[0], ks[-1]
    parts = []
    if hi >= 0:
        pos = [p.coef.get(k, "0") for k in range(hi, -1, -1)]
        parts.append(horner(pos, v, fma))
    if lo < 0:
        assert v == "x"
        neg = [p.coef.get(k, "0") for k in range(lo, 0)]       # c_lo .. c_-1  (highest |power| first)
        parts.append("(xinv * %s)" % horner(neg, "xinv", fma))
    return "(" + " + ".join(parts) + ")"


def emit_expr(e, polys, env, fma, names):
    if isinstance(e, Num): return lit(e.text)
    if isinstance(e, Var):
        if e.name.startswith("@P"): return emit_poly(polys[int(e.name[2:])], fma)
        if e.name in env: return lit(env[e.name])
        m = re.match(r"(roots|weights)\[(\d)\]", e.name)
        if m: return "%s[%s]" % (names[m.group(1)], m.group(2))
        return e.name
    if isinstance(e, Neg): return "(-%s)" % emit_expr(e.a, polys, env, fma, names)
    if isinstance(e, Call): return "%s(%s)" % (e.fn, emit_expr(e.a, polys, env, fma, names))
    return "(%s %s %s)" % (emit_expr(e.a, polys, env, fma, names), e.op, emit_expr(e.b, polys, env, fma, names))


HDR_NOTE = """ * GENERATED by oracle/tools/rys_extract.py -- do not edit by hand.
 * Piecewise fits for Rys roots/weights, 1..5 roots.  The fit COEFFICIENTS and band limits are those the
 * reference uses (reference Rys.cpp:314-2197; the classic fits for the Rys quadrature in X = rho*|PQ|^2),
 * required for 1e-12 per-integral parity; the evaluator is restated table-driven: each band is a short
 * list of Laurent polynomials in x or in y = x - x0 evaluated by Horner's rule, plus the band's closing
 * algebra.  Includes, on purpose, the reference's two-root behaviour for 15 < X <= 33 (Rys.cpp:614-624:
 * no dedicated fit, the (33,40] asymptotic form is used) -- see SURVEY.md section 7.
"""


def emit_function_bodies(res, fma, names, indent="    "):
    """returns {n: [lines]} of the band-dispatch body for each nroots"""
    out = {}
    for n, bands in res.items():
        L = []
        his = [b["hi"] for b in bands[:-1]]
        first = True
        for b in bands:
            cond = "x <= %s" % lit(repr(b["hi"])) if b["hi"] != float("inf") else None
            if cond is None: L.append(indent + "{")
            else: L.append(indent + ("if" if first else "else if") + " (%s) {" % cond) if first else L.append(indent + "else if (%s) {" % cond)
            if b is bands[-1] and cond is None and not first:
                L[-1] = indent + "else {"
            first = False
            declared = set()
            uses_y = b["y0"] is not None
            if uses_y: L.append(indent * 2 + "const double y = x - %s;" % lit(b["y0"]))
            for name, ae, polys in b["rows"]:
                ex = emit_expr(ae, polys, b["consts"], fma, names)
                m = re.match(r"(roots|weights)\[(\d)\]", name)
                if m: L.append(indent * 2 + "%s[%s] = %s;" % (names[m.group(1)], m.group(2), ex))
                elif name in declared: L.append(indent * 2 + "%s = %s;" % (name, ex))
                else:
                    declared.add(name)
                    L.append(indent * 2 + "double %s = %s;" % (name, ex))
            L.append(indent + "}")
        out[n] = L
    return out


def drop_unused_locals(lines):
    """remove 'double xinv = ...;' style locals never read later in the same band (keeps -Wall quiet)"""
    res, i = [], 0
    blocks, cur = [], []
    for ln in lines:
        cur.append(ln)
        if ln.strip() == "}": blocks.append(cur); cur = []
    for blk in blocks:
        keep = []
        for j, ln in enumerate(blk):
            m = re.match(r"\s*double (\w+) = ", ln)
            if m and not any(re.search(r"\b%s\b" % m.group(1), l2.split("=", 1)[1] if "=" in l2 else l2) for l2 in blk[j + 1:]):
                continue
            keep.append(ln)
        res.extend(keep)
    return res


def emit_c(res, path):
    bodies = emit_function_bodies(res, fma=False, names={"roots": "r", "weights": "w"})
    with open(path, "w") as f:
        f.write("/* oracle/rys_roots_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.\n *\n" + HDR_NOTE +
                " * Each function follows the reference routine of the same root count:\n"
                " *   n=1 Rys.cpp:314-411, n=2 :413-631, n=3 :633-970, n=4 :975-1486, n=5 :1488-2197.\n */\n")
        f.write("#include <math.h>\n#include \"unomol_oracle.h\"\n\n")
        for n in sorted(bodies):
            f.write("static void rys_fit_%d(double x, double *r, double *w) {\n" % n)
            f.write("\n".join(drop_unused_locals(bodies[n])) + "\n}\n\n")
        f.write("int oracle_rys_roots(int nroots, double x, double *r, double *w) {\n"
                "    switch (nroots) {\n" +
                "".join("    case %d: rys_fit_%d(x, r, w); return 0;\n" % (n, n) for n in sorted(bodies)) +
                "    default: return -1; /* rootN (6..9 roots, Rys.cpp:231-312) is not restated: it hangs for 2<~X<~15 */\n"
                "    }\n}\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--ir", action="store_true")
    ap.add_argument("--emit-c")
    args = ap.parse_args()
    res = analyse(args.ref)
    if args.emit_c: emit_c(res, args.emit_c)
    if args.ir:
        for n, bands in res.items():
            print("=" * 30, "nroots", n)
            for b in bands:
                print("--- band (%g, %g]  y0=%s" % (b["lo"], b["hi"], b["y0"]))
                for name, ae, polys in b["rows"]:
                    ps = "; ".join("P%d=%s{%d..%d}" % (i, p.var, min(p.coef), max(p.coef)) for i, p in enumerate(polys))
                    print("   %-12s = %r    [%s]" % (name, ae, ps))


if __name__ == "__main__":
    main()
