"""CPU: the oracle (oracle/unomol_oracle.c) against the committed fixtures generated from the unmodified
reference (tests/golden/generate_golden.py).  This is what pins the oracle on machines without /root/reference."""
import json
import os
import numpy as np
import pytest
from conftest import GOLDEN, golden_input


def test_rys_roots_match_reference_grid(oracle):
    g = np.load(os.path.join(GOLDEN, "rys_grid.npz"))
    for n in range(1, 6):
        for i, x in enumerate(g["x"]):
            r, w = oracle.rys_roots(n, x)
            # same fit coefficients, different operation order: rounding-level agreement
            np.testing.assert_allclose(r, g["r%d" % n][i], rtol=5e-14, atol=0)
            np.testing.assert_allclose(w, g["w%d" % n][i], rtol=5e-13, atol=1e-300)


def test_rys_two_root_defect_band_is_reproduced(oracle):
    """reference Rys.cpp:614-624: 15 < X <= 33 falls through to the (33,40] asymptotic form.  The exact
    zeroth moment sum(w) = F0(X) = sqrt(pi/4X) erf(sqrt X) holds there by construction, the higher ones do not."""
    from math import erf, sqrt, pi, exp
    x = 16.0
    r, w = oracle.rys_roots(2, x)
    f0 = sqrt(pi / (4 * x)) * erf(sqrt(x))
    assert abs(w.sum() - f0) < 1e-7
    t2 = r / (1 + r)
    f1 = (f0 - exp(-x)) / (2 * x)
    assert 1e-9 < abs((w * t2).sum() - f1) < 1e-5      # the documented defect, not rounding


@pytest.mark.parametrize("name", ["3g.h2o", "631.nh3", "631.co"])
def test_unique_integral_list_matches_reference_cache(oracle, name):
    g = np.load(os.path.join(GOLDEN, "eri_%s.npz" % name.replace(".", "_")))
    b = oracle.basis(golden_input(name))
    vals, ijkl, ncalc = oracle.unique_eris(b)
    assert len(vals) == len(g["vals"])
    assert np.array_equal(ijkl, g["ijkl"])          # same records, same order
    assert np.max(np.abs(vals - g["vals"])) < 1e-13
    assert ncalc == {"3g.h2o": 406, "631.nh3": 108345, "631.co": 108345}[name]


@pytest.mark.parametrize("name", ["3g.h2o", "631.nh3", "631.co", "b.dhdz", "dh95.co2", "fg.h2o"])
def test_g_matrices_match_reference(oracle, name):
    g = np.load(os.path.join(GOLDEN, "g_%s.npz" % name.replace(".", "_")))
    b = oracle.basis(golden_input(name))
    vals, ijkl, _ = oracle.unique_eris(b)
    G = oracle.form_g_rhf(vals, ijkl, g["P"])
    GA, GB = oracle.form_g_uhf(vals, ijkl, g["P"], g["PB"])
    scale = max(1.0, np.max(np.abs(g["G"])))
    assert np.max(np.abs(G - g["G"])) < 1e-12 * scale
    assert np.max(np.abs(GA - g["GA"])) < 1e-12 * scale
    assert np.max(np.abs(GB - g["GB"])) < 1e-12 * scale


@pytest.mark.parametrize("name", ["3g.h2o", "631.nh3", "b.dhdz", "dh95.co2", "dh95.c2h2", "fg.h2o", "fg2.hf"])
def test_quartet_blocks_match_reference(oracle, name):
    g = np.load(os.path.join(GOLDEN, "quartets_%s.npz" % name.replace(".", "_")))
    b = oracle.basis(golden_input(name))
    for q, (i, j, k, l) in enumerate(g["quartets"]):
        ref = g["values"][g["offsets"][q]:g["offsets"][q + 1]]
        blk = oracle.quartet_block(b, int(i), int(j), int(k), int(l)).ravel()
        assert np.max(np.abs(blk - ref)) < 1e-13


def test_high_l_fixture_covers_both_reference_algorithms(oracle):
    """fg.h2o (ours; f and g shells): the fixture holds blocks of both the Rys path (l_tot <= 8) and the
    McMurchie-Davidson path (l_tot > 8, reference TwoElectronInts.cpp:661-665), one-, two- and multi-centre."""
    g = np.load(os.path.join(GOLDEN, "quartets_fg_h2o.npz"))
    b = oracle.basis(golden_input("fg.h2o"))
    assert b.maxl == 4
    ltot = [sum(int(b.lv[s]) for s in q) for q in g["quartets"]]
    ncen = [len({int(b.cen[s]) for s in q}) for q in g["quartets"]]
    assert min(ltot) <= 8 < max(ltot)
    assert {1, 2, 3} <= {n for n, lt in zip(ncen, ltot) if lt > 8}


def test_direct_g_equals_stored_g(oracle):
    b = oracle.basis(golden_input("631.nh3"))
    vals, ijkl, _ = oracle.unique_eris(b)
    rng = np.random.default_rng(3)
    P = rng.standard_normal(b.no2)
    G1 = oracle.form_g_rhf(vals, ijkl, P)
    G2, nq, npq = oracle.direct_g_rhf(b, P)
    assert np.max(np.abs(G1 - G2)) < 1e-12
    assert nq > 0 and npq > 0


def test_digestion_equals_dense_contraction(oracle):
    """G_ij = sum_kl P_kl [2 (ij|kl) - (ik|jl)] for any symmetric P."""
    b = oracle.basis(golden_input("3g.h2o"))
    n = b.nbf
    eri = np.zeros((n, n, n, n))
    sh_of = np.zeros(n, int); comp_of = np.zeros(n, int)
    for s in range(b.nshell):
        nc = (b.lv[s] + 1) * (b.lv[s] + 2) // 2
        for c in range(nc):
            sh_of[b.off[s] + c] = s; comp_of[b.off[s] + c] = c
    for i in range(b.nshell):
        for j in range(b.nshell):
            for k in range(b.nshell):
                for l in range(b.nshell):
                    blk = oracle.quartet_block(b, i, j, k, l)
                    sl = lambda s: slice(b.off[s], b.off[s] + blk.shape[[i, j, k, l].index(s)])
                    eri[b.off[i]:b.off[i] + blk.shape[0], b.off[j]:b.off[j] + blk.shape[1],
                        b.off[k]:b.off[k] + blk.shape[2], b.off[l]:b.off[l] + blk.shape[3]] = blk
    rng = np.random.default_rng(5)
    A = rng.standard_normal((n, n)); Pf = A + A.T
    Gf = 2 * np.einsum("ijkl,kl->ij", eri, Pf) - np.einsum("ikjl,kl->ij", eri, Pf)
    tri = np.tril_indices(n)
    vals, ijkl, _ = oracle.unique_eris(b)
    G = oracle.form_g_rhf(vals, ijkl, Pf[tri])
    assert np.max(np.abs(G - Gf[tri])) < 1e-12


def test_short_dat_fixture_is_consistent_with_fresh_reference_runs():
    short = json.load(open(os.path.join(GOLDEN, "short_dat.json")))
    runs = json.load(open(os.path.join(GOLDEN, "ref_runs.json")))
    for name, r in runs.items():
        if name in short:
            assert abs(r["e_final"] - short[name][1]) < 5e-12, name   # HEAD reproduces goldens to 1.3e-12


def test_reference_mpi_build_over_the_shim():
    """The reference's MPI driver (unmodified sources, oracle/mpi_shim/mpi.h instead of an MPI library) on 3 forked ranks
    reproduces the serial reference's H2O/STO-3G energy: the round-robin split of TwoElectronIntsMPI.cpp:350-354 plus the
    MPI_Reduce of RHF_MPI.hpp:108 lose nothing."""
    import os
    from oracle.oracle import run_reference_mpi, HERE
    if not os.path.exists(os.path.join(HERE, "_ref", "UnomolMPI")):
        pytest.skip("oracle/_ref/UnomolMPI not built (no /root/reference at build time)")
    from unomol_b200.basis import test_input
    r1 = run_reference_mpi(test_input("3g.h2o"), 1)
    r3 = run_reference_mpi(test_input("3g.h2o"), 3)
    assert r3["ranks"] == 3 and r3["iterations"] > 3
    assert abs(r3["energy"] - (-74.962940006948)) < 1e-9      # BASELINE.json config 1
    assert abs(r3["energy"] - r1["energy"]) < 1e-11
