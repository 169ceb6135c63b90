"""CPU: the PRODUCT's Rys evaluator (unomol_b200/csrc/rys_roots.cuh, generated tables from
unomol_b200/tools/gen_rys_tables.py) compiled for the host by tests/host_emul/rys_host.cpp.

It is an independent evaluator (Boys grid + closed-form 1/2 roots, degree-12 piecewise polynomials for 3..5 roots), not
the reference's piecewise fits, so it is compared
  * with the reference's roots/weights on the committed 1219-point grid (tests/golden/rys_grid.npz) at the tolerance the
    reference's own fits allow: they reproduce the exact Boys moments to 6.4e-14, individual roots
    and weights to a few 1e-13 -- the stated tolerance is 5e-13 relative, which keeps every quartet fixture within 1e-12
    (tests/test_gpu_parity.py, tests/test_highl_emulation.py);
  * with its defining property, sum_i w_i t_i^(2k) = F_k(X) for k < 2n, against closed-form / scipy Boys values, at
    1e-13: this is what "exact" means for rys2_exact = 1;
  * in parity mode with the reference's two-root band 15 < X <= 40 (reference Rys.cpp:614-624) to rounding.
"""
import ctypes
import os
import subprocess
import numpy as np
import pytest
from conftest import GOLDEN, ROOT

_D = ctypes.c_double
_dp = lambda a: a.ctypes.data_as(ctypes.POINTER(_D))


@pytest.fixture(scope="module")
def rys():
    d = os.path.join(ROOT, "tests", "host_emul")
    so = os.path.join(d, "librys_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(d, "rys_host.cpp")])
    L = ctypes.CDLL(so)
    L.unomol_rys_host.argtypes = [ctypes.c_int, _D, ctypes.c_int, ctypes.POINTER(_D), ctypes.POINTER(_D)]
    L.unomol_boys_host.argtypes = [_D, ctypes.POINTER(_D)]
    L.unomol_boys_grid_host.argtypes = [_D, ctypes.POINTER(_D)]
    L.unomol_f0_host.argtypes = [_D]; L.unomol_f0_host.restype = _D

    def f(n, x, exact=0):
        r = np.zeros(9); w = np.zeros(9)
        assert L.unomol_rys_host(n, float(x), exact, _dp(r), _dp(w)) == 0
        return r[:n].copy(), w[:n].copy()
    f.lib = L
    return f


def boys_exact(m, x):
    from scipy.special import gammainc, gamma
    if x < 1e-12:
        return 1.0 / (2 * m + 1)
    return gammainc(m + 0.5, x) * gamma(m + 0.5) / (2.0 * x ** (m + 0.5))


def test_matches_reference_grid_at_the_stated_tolerance(rys):
    g = np.load(os.path.join(GOLDEN, "rys_grid.npz"))
    worst = 0.0
    for n in range(1, 6):
        for i, x in enumerate(g["x"]):
            r, w = rys(n, x)
            worst = max(worst, np.max(np.abs(r / g["r%d" % n][i] - 1.0)), np.max(np.abs(w / g["w%d" % n][i] - 1.0)))
    assert worst < 5e-13, worst


def test_two_root_parity_band_is_the_reference_formula(rys):
    g = np.load(os.path.join(GOLDEN, "rys_grid.npz"))
    band = [(i, x) for i, x in enumerate(g["x"]) if 15.0 < x <= 40.0]
    assert len(band) > 50
    for i, x in band:
        r, w = rys(2, x)
        # same formula evaluated for t^2 = r/(1+r) with one reciprocal (rys2_compat_band_t2) and converted back: rounding only
        np.testing.assert_allclose(r, g["r2"][i], rtol=1e-14)
        np.testing.assert_allclose(w, g["w2"][i], rtol=2e-14)
    # and the defect is real: in exact mode the quadrature differs from the reference there by far more than rounding
    r, w = rys(2, 16.0, exact=1)
    i16 = int(np.argmin(np.abs(g["x"] - 16.0)))
    assert abs(g["x"][i16] - 16.0) < 1e-12
    assert 1e-9 < np.max(np.abs(r / g["r2"][i16] - 1.0)) < 1e-4


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_exact_mode_reproduces_the_boys_moments(rys, n):
    rng = np.random.default_rng(n)
    xs = np.concatenate([rng.uniform(0.0, 70.0, 400), [0.0, 1e-9, 0.03124, 0.03126, 14.999, 15.0, 15.001, 33.0, 34.999, 35.0,
                                                       40.0, 45.999, 46.0, 51.999, 52.0, 57.999, 58.0, 63.999, 64.0, 200.0]])
    worst = 0.0
    for x in xs:
        r, w = rys(n, x, exact=1)
        assert np.all(np.diff(r) > 0) and np.all(w > 0)
        t2 = r / (1.0 + r)
        for k in range(2 * n):
            ex = boys_exact(k, x)
            worst = max(worst, abs(np.sum(w * t2 ** k) - ex) / ex)
    # a root error eps shows up 2k-fold in the k-th moment (k up to 9): 1e-13 on the moments = ~1e-15 on roots/weights,
    # which tools/gen_rys_tables.py --check measures directly against 60-digit values (2.2e-16)
    assert worst < 1e-13, worst


def test_boys_grid_path(rys):
    F = np.zeros(4)
    worst = 0.0
    for x in np.linspace(0.0, 45.99, 2311):
        rys.lib.unomol_boys_host(float(x), _dp(F))
        for m in range(4):
            ex = boys_exact(m, x)
            worst = max(worst, abs(F[m] - ex) / ex)
    assert worst < 2e-14, worst       # scipy's gammainc is the limit here; against mpmath the grid path is good to 7e-16
    # the Taylor-row path the kernels use against the recursion-based grid: two independent evaluations of the same moments
    G = np.zeros(4)
    worst = 0.0
    for x in np.linspace(0.0, 45.99, 4603):
        rys.lib.unomol_boys_host(float(x), _dp(F)); rys.lib.unomol_boys_grid_host(float(x), _dp(G))
        worst = max(worst, float(np.max(np.abs(F / G - 1.0))))
    assert worst < 2e-15, worst


def test_f0_only_path_matches_the_one_root_weight(rys):
    """(ss|ss) uses F_0 from its own {F_7, exp(-X_i)} grid; it must agree with the one-root weight to rounding"""
    worst = 0.0
    for x in np.concatenate([np.linspace(0.0, 60.0, 4801), [34.999999, 35.0, 35.000001]]):
        r, w = rys(1, x)
        worst = max(worst, abs(rys.lib.unomol_f0_host(float(x)) / w[0] - 1.0))
    assert worst < 1e-15, worst


@pytest.mark.parametrize("n", [6, 7, 8, 9])
def test_six_to_nine_roots_reproduce_the_boys_moments(rys, n):
    """the range of the reference's Rys::rootN (Rys.cpp:231-312), here degree-12 tables generated at 80
    digits (tools/gen_rys_tables.py --hi).  Defining property: the n-point rule integrates t^(2k), k < 2n, exactly."""
    import mpmath as mp
    mp.mp.dps = 40

    def boys_mp(m, x):
        if x == 0:
            return 1.0 / (2 * m + 1)
        return float(mp.gammainc(m + mp.mpf(1) / 2, 0, x) / (2 * mp.mpf(x) ** (m + mp.mpf(1) / 2)))
    rng = np.random.default_rng(n)
    xa = {6: 72.0, 7: 78.0, 8: 84.0, 9: 90.0}[n]
    xs = np.concatenate([rng.uniform(0.0, 100.0, 150), [0.0, 1e-9, 0.999999, 1.0, 1.000001, 2.0, 8.5, 15.0, xa - 1e-9, xa, xa + 1e-9, 250.0]])
    worst = 0.0
    for x in xs:
        r, w = rys(n, x)
        assert np.all(np.diff(r) > 0) and np.all(w > 0)
        t2 = r / (1.0 + r)
        for k in range(2 * n):
            ex = boys_mp(k, float(x))
            worst = max(worst, abs(np.sum(w * t2 ** k) - ex) / ex)
    assert worst < 2e-13, worst


def test_six_to_nine_roots_vs_the_reference_rootN_where_it_returns(rys):
    """The reference's general routine smashes its stack or hangs for 2 <~ X <~ 15; the fixture
    (tests/golden/generate_golden.py: rootn_fixture, one child process per point) holds the points where it came back: X <= 2.5
    and X >= 18.  There it is an orthogonal-polynomial construction from Boys moments in double precision, good to ~1e-9 on the
    small weights; the table evaluator is compared at the tolerance that construction allows."""
    g = np.load(os.path.join(GOLDEN, "rys_rootn_grid.npz"))
    # measured differences (the moment tests above put the table evaluator at 2e-13 of the exact quadrature, so these are the
    # reference's own errors): small X 3e-10 / 6e-9 / 2e-7 / 1.2e-5 for 6 / 7 / 8 / 9 roots, X >= 18 7e-13 / 1e-11 / 1e-10 / 7e-10
    tol_small = {6: 1e-9, 7: 2e-8, 8: 1e-6, 9: 5e-5}
    tol_large = {6: 5e-12, 7: 5e-11, 8: 5e-10, 9: 5e-9}
    npts = 0
    for n in range(6, 10):
        for i, x in enumerate(g["x%d" % n]):
            r, w = rys(n, x)
            tol = tol_small[n] if x < 10.0 else tol_large[n]
            assert np.max(np.abs(r / g["r%d" % n][i] - 1.0)) < tol, (n, x)
            assert np.max(np.abs(w - g["w%d" % n][i])) / np.max(w) < tol, (n, x)
            npts += 1
    assert npts > 250
