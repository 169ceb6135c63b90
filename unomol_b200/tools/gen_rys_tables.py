#!/usr/bin/env python3
"""unomol_b200/tools/gen_rys_tables.py -- generates unomol_b200/csrc/rys_tables.inc from first principles.

Nothing here reads the reference: every number is computed with mpmath from the definition of the Rys
quadrature (weight exp(-X t^2) on [0,1]; nodes x_i = t_i^2, the kernels use r_i = x_i/(1-x_i)).

  * Boys grid        {F_MTOP(X_i), exp(-X_i)} on X_i = i/16, i = 0 .. 33*16+1.  The evaluator gets F_m(X_i) for every
                     m < MTOP by the (positive-term, hence stable) downward recursion, F_top(X) by an 8-term Taylor
                     series in X - X_i (dF_m/dX = -F_{m+1}) and the lower orders by downward recursion at X itself.
                     One and two roots are built from these moments in closed form (rys_roots.cuh).
  * 3, 4, 5 roots    piecewise polynomials (monomials in s in [-1,1], converted from Chebyshev interpolants
                     computed at 60 digits) of y_i(X) = t_i^2 and w_i(X) on uniform intervals of [0, XA_n); above XA_n the
                     exp(-X)-free Gauss-Hermite limit is exact to double precision.
  * Hermite limits   R_i = squared positive roots of H_2n, W_i = Gauss-Hermite weights / sqrt(pi) (n = 1..5).

Run:  python unomol_b200/tools/gen_rys_tables.py [--check]      (about a minute)
"""
import os
import sys
import mpmath as mp

mp.mp.dps = 60
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "csrc", "rys_tables.inc")
OUT_CONSTS = os.path.join(HERE, "..", "csrc", "rys_consts.inc")

BOYS_H_INV = 16            # grid points per unit X
BOYS_XMAX = 46             # grid end: two-root moments are used up to here in exact mode; one root switches to
                           # the asymptotic F_0 = sqrt(pi/4X) at X = 35 already (exp(-35)/70 = 9e-18)
BOYS_MTOP = 10             # highest order tabulated (two roots need F_0..F_3, Taylor of F_3 reaches F_10)
BOYS1_XMAX = 35            # second grid for the one-root classes: F_8 only (Taylor of F_1 reaches F_8), up to the asymptotic switch
BOYS1_MTOP = 8
NTERMS = 8                 # Taylor terms: (1/32)^8/8! = 2e-17

# n: (interval width, polynomial degree, XA_n)
PIECE = {3: (1.0, 12, 52.0), 4: (1.0, 12, 58.0), 5: (1.0, 12, 64.0),   # measured: degree 11 leaves 1.1e-16, XA: limit good to 4e-16
         # 6..9 roots (--hi: rys_tables_hi.inc; the reference's Rys::rootN range, Rys.cpp:231-312): Gauss-Hermite limit measured
         # good to 4e-16 (roots) / 1e-15 (weights) at 70 / 76 / 82 / 88
         6: (1.0, 12, 72.0), 7: (1.0, 12, 78.0), 8: (1.0, 12, 84.0), 9: (1.0, 12, 90.0)}


def boys(m, x):
    x = mp.mpf(x)
    if x == 0:
        return mp.mpf(1) / (2 * m + 1)
    return mp.gammainc(m + mp.mpf(1) / 2, 0, x) / (2 * x ** (m + mp.mpf(1) / 2))


def rys_exact(n, x):
    """nodes x_i = t_i^2 (ascending) and weights of the n-point Rys quadrature, from the moments F_k(X), k < 2n"""
    m = [boys(k, x) for k in range(2 * n + 1)]
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i in range(n):
        for j in range(n):
            A[i, j] = m[i + j]
        b[i] = -m[i + n]
    c = mp.lu_solve(A, b)                       # monic orthogonal polynomial x^n + sum c_j x^j
    roots = mp.polyroots([mp.mpf(1)] + [c[j] for j in reversed(range(n))], maxsteps=200, extraprec=400)
    roots = sorted([mp.re(r) for r in roots])
    V = mp.matrix(n, n)
    rhs = mp.matrix(n, 1)
    for k in range(n):
        for i in range(n):
            V[k, i] = roots[i] ** k
        rhs[k] = m[k]
    w = mp.lu_solve(V, rhs)
    return roots, [w[i] for i in range(n)]


def hermite_limits(n):
    """large-X limit: x_i -> R_i / X, w_i -> W_i sqrt(pi/(4X)); R_i = squared positive roots of H_2n"""
    # Gauss-Hermite with 2n points via Golub-Welsch at high precision
    N = 2 * n
    J = mp.matrix(N, N)
    for i in range(N - 1):
        J[i, i + 1] = J[i + 1, i] = mp.sqrt(mp.mpf(i + 1) / 2)
    E, Q = mp.eigsy(J)
    pts = sorted([(E[i], Q[0, i] ** 2) for i in range(N)], key=lambda t: t[0])
    pos = [(x * x, 2 * w) for (x, w) in pts if x > 0]     # weights normalised to sum 1 over the positive half
    pos.sort(key=lambda t: t[0])
    return [p[0] for p in pos], [p[1] for p in pos]


def cheb_fit(fvals, deg):
    """Chebyshev coefficients of the degree-deg interpolant through the deg+1 Chebyshev nodes (values given there)"""
    N = deg + 1
    c = []
    for k in range(N):
        s = mp.mpf(0)
        for j in range(N):
            s += fvals[j] * mp.cos(mp.pi * k * (j + mp.mpf(1) / 2) / N)
        c.append(s * (2 if k else 1) / N)
    return c


def cheb_to_mono(c):
    """Chebyshev series -> monomial coefficients in s (exact rational arithmetic on mp numbers)"""
    n = len(c)
    T0 = [mp.mpf(1)]
    T1 = [mp.mpf(0), mp.mpf(1)]
    out = [mp.mpf(0)] * n
    polys = [T0, T1]
    for k in range(2, n):
        Tk = [mp.mpf(0)] + [2 * a for a in polys[-1]]
        for i, a in enumerate(polys[-2]):
            Tk[i] -= a
        polys.append(Tk)
    for k in range(n):
        for i, a in enumerate(polys[k]):
            out[i] += c[k] * a
    return out


def fmt(v):
    return "%.17e" % float(v)


def gen_piece(n, check):
    width, deg, xa = PIECE[n]
    nint = int(round(xa / width))
    N = deg + 1
    nodes = [mp.cos(mp.pi * (j + mp.mpf(1) / 2) / N) for j in range(N)]
    rows = []          # [interval][k][func], func = y_0..y_{n-1} (y = t^2), w_0..w_{n-1}
    worst = 0.0
    for iv in range(nint):
        x0 = mp.mpf(iv) * width
        xc = x0 + mp.mpf(width) / 2
        vals = []
        for s in nodes:
            xs, ws = rys_exact(n, xc + s * width / 2)
            vals.append(list(xs) + ws)          # nodes as y = t^2 in (0,1): what the recurrences use
        mono = []
        for f in range(2 * n):
            mono.append(cheb_to_mono(cheb_fit([v[f] for v in vals], deg)))
        rows.append(mono)
        if check:
            for s in (mp.mpf(-1), mp.mpf("-0.77"), mp.mpf("-0.31"), mp.mpf("0.123"), mp.mpf("0.5"), mp.mpf("0.93"), mp.mpf(1)):
                xs, ws = rys_exact(n, xc + s * width / 2)
                ex = list(xs) + ws
                sd = float(s)
                for f in range(2 * n):
                    acc = 0.0
                    for k in range(deg, -1, -1):
                        acc = acc * sd + float(mono[f][k])
                    worst = max(worst, abs(acc - float(ex[f])) / abs(float(ex[f])))
    return width, deg, xa, nint, rows, worst


def main():
    check = "--check" in sys.argv
    L = []
    C = []
    C.append("// unomol_b200/csrc/rys_consts.inc -- GENERATED by unomol_b200/tools/gen_rys_tables.py (mpmath, 60 digits): grid")
    C.append("// parameters of rys_tables.inc and the Gauss-Hermite limits.  Do not edit by hand.")
    L.append("// unomol_b200/csrc/rys_tables.inc -- GENERATED by unomol_b200/tools/gen_rys_tables.py (mpmath, 60 digits) from the")
    L.append("// definition of the Rys quadrature; nothing is taken from the reference.  Do not edit by hand.")
    C.append("#define RYS_BOYS_HINV %d" % BOYS_H_INV)
    C.append("#define RYS_BOYS_XMAX %d" % BOYS_XMAX)
    C.append("#define RYS_BOYS_MTOP %d" % BOYS_MTOP)
    nb = BOYS_XMAX * BOYS_H_INV + 2
    C.append("#define RYS_BOYS_NPTS %d" % nb)
    nb1 = BOYS1_XMAX * BOYS_H_INV + 2
    C.append("#define RYS_BOYS1_XMAX %d" % BOYS1_XMAX)
    C.append("#define RYS_BOYS1_MTOP %d" % BOYS1_MTOP)
    C.append("#define RYS_BOYS1_NPTS %d" % nb1)
    C.append("#define RYS_BOYS0_MTOP %d" % (BOYS1_MTOP - 1))
    L.append("RYS_TABLE(rys_boys_tab, 2 * RYS_BOYS_NPTS) = {")
    for i in range(nb):
        x = mp.mpf(i) / BOYS_H_INV
        L.append("    %s, %s," % (fmt(boys(BOYS_MTOP, x)), fmt(mp.exp(-x))))
    L.append("};")
    L.append("RYS_TABLE(rys_boys1_tab, 2 * RYS_BOYS1_NPTS) = {")
    for i in range(nb1):
        x = mp.mpf(i) / BOYS_H_INV
        L.append("    %s, %s," % (fmt(boys(BOYS1_MTOP, x)), fmt(mp.exp(-x))))
    L.append("};")
    L.append("RYS_TABLE(rys_boys0_tab, 2 * RYS_BOYS1_NPTS) = {")
    for i in range(nb1):
        x = mp.mpf(i) / BOYS_H_INV
        L.append("    %s, %s," % (fmt(boys(BOYS1_MTOP - 1, x)), fmt(mp.exp(-x))))
    L.append("};")
    # Hermite limits
    C.append("// large-X limit: r_i = R_i/(X - R_i), w_i = W_i sqrt(pi/(4X)); row n-1 holds n entries")
    hr, hw = [], []
    for n in range(1, 6):
        R, W = hermite_limits(n)
        hr.append(R + [mp.mpf(0)] * (5 - n))
        hw.append(W + [mp.mpf(0)] * (5 - n))
    C.append("RYS_CONST(rys_herm_r, 25) = {")
    for row in hr:
        C.append("    " + ", ".join(fmt(v) for v in row) + ",")
    C.append("};")
    C.append("RYS_CONST(rys_herm_w, 25) = {")
    for row in hw:
        C.append("    " + ", ".join(fmt(v) for v in row) + ",")
    C.append("};")
    for n in (3, 4, 5):
        width, deg, xa, nint, rows, worst = gen_piece(n, check)
        sys.stderr.write("n=%d: %d intervals of width %g, degree %d, XA %g, worst rel err on check points %.2e\n"
                         % (n, nint, width, deg, xa, worst))
        C.append("#define RYS_P%d_WIDTH_INV %.17g" % (n, 1.0 / width))
        C.append("#define RYS_P%d_DEG %d" % (n, deg))
        C.append("#define RYS_P%d_XA %.17g" % (n, xa))
        C.append("#define RYS_P%d_NINT %d" % (n, nint))
        L.append("// layout [interval][k = 0..deg][y_0..y_%d, w_0..w_%d], y = t^2" % (n - 1, n - 1))
        L.append("RYS_TABLE(rys_piece%d_tab, %d) = {" % (n, nint * (deg + 1) * 2 * n))
        for iv in range(nint):
            for k in range(deg + 1):
                L.append("    " + ", ".join(fmt(rows[iv][f][k]) for f in range(2 * n)) + ",")
        L.append("};")
    with open(OUT, "w") as f:
        f.write("\n".join(L) + "\n")
    with open(OUT_CONSTS, "w") as f:
        f.write("\n".join(C) + "\n")
    sys.stderr.write("wrote %s\n" % os.path.normpath(OUT))


def main_hi():
    """rys_tables_hi.inc / rys_consts_hi.inc: 6..9 roots (all-Rys mode for quartets with l_tot > 8).  Separate files so that the
    tables of the default path stay byte-identical."""
    mp.mp.dps = 80                 # the Hankel systems of 9 roots lose ~25 digits
    check = "--check" in sys.argv
    L = ["// unomol_b200/csrc/rys_tables_hi.inc -- GENERATED by unomol_b200/tools/gen_rys_tables.py --hi (mpmath, 80 digits) from the",
         "// definition of the Rys quadrature: 6..9 roots.  Do not edit by hand."]
    C = ["// unomol_b200/csrc/rys_consts_hi.inc -- GENERATED by unomol_b200/tools/gen_rys_tables.py --hi.  Do not edit by hand.",
         "// large-X limit of 6..9 roots: t_i^2 = R_i / X, w_i = W_i sqrt(pi/(4X)); row n-6 holds n entries"]
    hr, hw = [], []
    for n in range(6, 10):
        R, W = hermite_limits(n)
        hr.append(R + [mp.mpf(0)] * (9 - n))
        hw.append(W + [mp.mpf(0)] * (9 - n))
    for name, rows in (("rys_herm_hi_r", hr), ("rys_herm_hi_w", hw)):
        C.append("RYS_CONST(%s, 36) = {" % name)
        for row in rows:
            C.append("    " + ", ".join(fmt(v) for v in row) + ",")
        C.append("};")
    for n in (6, 7, 8, 9):
        width, deg, xa, nint, rows, worst = gen_piece(n, check)
        sys.stderr.write("n=%d: %d intervals of width %g, degree %d, XA %g, worst rel err on check points %.2e\n"
                         % (n, nint, width, deg, xa, worst))
        C.append("#define RYS_P%d_DEG %d" % (n, deg))
        C.append("#define RYS_P%d_XA %.17g" % (n, xa))
        C.append("#define RYS_P%d_NINT %d" % (n, nint))
        L.append("// layout [interval][k = 0..deg][y_0..y_%d, w_0..w_%d], y = t^2" % (n - 1, n - 1))
        L.append("RYS_TABLE(rys_piece%d_tab, %d) = {" % (n, nint * (deg + 1) * 2 * n))
        for iv in range(nint):
            for k in range(deg + 1):
                L.append("    " + ", ".join(fmt(rows[iv][f][k]) for f in range(2 * n)) + ",")
        L.append("};")
    with open(os.path.join(HERE, "..", "csrc", "rys_tables_hi.inc"), "w") as f:
        f.write("\n".join(L) + "\n")
    with open(os.path.join(HERE, "..", "csrc", "rys_consts_hi.inc"), "w") as f:
        f.write("\n".join(C) + "\n")
    sys.stderr.write("wrote rys_tables_hi.inc, rys_consts_hi.inc\n")


def main_poly():
    """rys_tables_poly.inc / rys_consts_poly.inc: Taylor coefficient tables of the Boys function about the points of a coarse grid,
    a_k = F_{m+k}(X_i) / k! (dF_m/dX = -F_{m+1}, so F_m(x) = sum_k a_k (X_i - x)^k), X_i = i / HINV.  One row of coefficients
    replaces the downward recursion + Taylor chain of boys_grid for the order the kernels need most: the near branch of the root
    evaluation runs with few active lanes, so its LENGTH is what costs."""
    mp.mp.dps = 50
    hinv, deg, stride = 4, 10, 12             # |X_i - x| <= 1/8: (1/8)^11 / 11! = 3e-18
    L = ["// unomol_b200/csrc/rys_tables_poly.inc -- GENERATED by unomol_b200/tools/gen_rys_tables.py --poly (mpmath, 50 digits).",
         "// Do not edit by hand."]
    C = ["// unomol_b200/csrc/rys_consts_poly.inc -- GENERATED by unomol_b200/tools/gen_rys_tables.py --poly.  Do not edit by hand.",
         "#define RYS_FP_HINV %d" % hinv, "#define RYS_FP_DEG %d" % deg, "#define RYS_FP_STRIDE %d" % stride]
    for name, m, xmax in (("rys_f0poly_tab", 0, BOYS1_XMAX),):
        npts = xmax * hinv + 2
        C.append("#define %s_NPTS %d" % (name.upper(), npts))
        L.append("// [point][k = 0..%d (+ padding to %d)]: F_%d+k(X_i) / k!" % (deg, stride, m))
        L.append("RYS_TABLE(%s, %d) = {" % (name, npts * stride))
        for i in range(npts):
            x = mp.mpf(i) / hinv
            row = [boys(m + k, x) / mp.factorial(k) for k in range(deg + 1)] + [mp.mpf(0)] * (stride - deg - 1)
            L.append("    " + ", ".join(fmt(v) for v in row) + ",")
        L.append("};")
    # two roots: Taylor rows of F_3 (F_{3+j}(X_i) / j!, j = 0..10) with exp(-X_i) in the twelfth slot, up to the end of the moment
    # range of the exact mode; F_2, F_1, F_0 follow by the downward recursion at x with exp(-x) = exp(-X_i) exp(X_i - x)
    npts = BOYS_XMAX * hinv + 2
    C.append("#define RYS_F3POLY_TAB_NPTS %d" % npts)
    L.append("// [point][j = 0..%d: F_3+j(X_i) / j!][exp(-X_i)]" % deg)
    L.append("RYS_TABLE(rys_f3poly_tab, %d) = {" % (npts * stride))
    for i in range(npts):
        x = mp.mpf(i) / hinv
        row = [boys(3 + k, x) / mp.factorial(k) for k in range(deg + 1)] + [mp.exp(-x)]
        L.append("    " + ", ".join(fmt(v) for v in row) + ",")
    L.append("};")
    with open(os.path.join(HERE, "..", "csrc", "rys_tables_poly.inc"), "w") as f:
        f.write("\n".join(L) + "\n")
    with open(os.path.join(HERE, "..", "csrc", "rys_consts_poly.inc"), "w") as f:
        f.write("\n".join(C) + "\n")
    sys.stderr.write("wrote rys_tables_poly.inc, rys_consts_poly.inc\n")


if __name__ == "__main__":
    if "--poly" in sys.argv:
        main_poly()
    elif "--hi" in sys.argv:
        main_hi()
    else:
        main()
