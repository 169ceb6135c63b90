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
    L.unomol_f0_host.argtypes = [_D]; L.unomol_f0_host.restype = _D

    def f(n, x, exact=0):
        r = np.zeros(5); w = np.zeros(5)
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


def test_f0_only_path_matches_the_one_root_weight(rys):
    """(ss|ss) uses F_0 from its own {F_7, exp(-X_i)} grid; it must agree with the one-root weight to rounding"""
    worst = 0.0
    for x in np.concatenate([np.linspace(0.0, 60.0, 4801), [34.999999, 35.0, 35.000001]]):
        r, w = rys(1, x)
        worst = max(worst, abs(rys.lib.unomol_f0_host(float(x)) / w[0] - 1.0))
    assert worst < 1e-15, worst
