"""GPU (-m gpu): the CUDA path, called through the C ABI (unomol_b200.capi -> libunomol_b200.so), against the
CPU oracle on the same inputs and against the committed reference fixtures.

Tolerances are BASELINE.json's: per-quartet ERIs within 1e-12 absolute; G within 1e-12 relative to max|G|
(G sums up to ~nbf^2 integrals); SCF energies within 1e-9 Eh (tests/test_gpu_scf.py)."""
import os
import numpy as np
import pytest
from conftest import GOLDEN, golden_input

pytestmark = pytest.mark.gpu

ERI_TOL = 1e-12


@pytest.fixture(scope="module")
def capi():
    from unomol_b200 import capi as c
    return c


def _handle(capi, name, **kw):
    from unomol_b200.basis import Basis
    b = Basis.from_patin(golden_input(name))
    return b, capi.Handle(b, **kw)


@pytest.mark.parametrize("name", ["3g.h2o", "631.nh3", "b.dhdz", "dh95.co2", "dh95.c2h2", "tz2p.sf6", "fg.h2o", "fg2.hf"])
def test_quartet_blocks_vs_reference_fixture(capi, name):
    path = os.path.join(GOLDEN, "quartets_%s.npz" % name.replace(".", "_"))
    g = np.load(path)
    b, h = _handle(capi, name)
    worst = 0.0
    for q, (i, j, k, l) in enumerate(g["quartets"]):
        ref = g["values"][g["offsets"][q]:g["offsets"][q + 1]]
        blk = h.eri_quartet(int(i), int(j), int(k), int(l)).ravel()
        worst = max(worst, np.max(np.abs(blk - ref)))
    assert worst < ERI_TOL, worst


@pytest.mark.parametrize("name", ["3g.h2o", "631.nh3", "631.co", "b.dhdz", "dh95.co2", "dh95.c2h2", "tz2p.sf6"])
def test_quartet_blocks_from_the_fock_kernels_vs_reference_fixture(capi, name):
    """option dump_kernel = 1: the block comes from the kernel a Fock build runs for the quartet's class (bra-tile kernel
    for the s/p classes, one-bra register kernel for the small d classes) in its dump mode, not from the generic kernel"""
    g = np.load(os.path.join(GOLDEN, "quartets_%s.npz" % name.replace(".", "_")))
    b, h = _handle(capi, name)
    h.set_option("reg_kernels", 2); h.set_option("tile_kernels", 2)     # small molecule: force the kernels of the long lists
    h.set_option("dump_kernel", 1)
    worst, hot = 0.0, 0
    for q, (i, j, k, l) in enumerate(g["quartets"]):
        ref = g["values"][g["offsets"][q]:g["offsets"][q + 1]]
        blk = h.eri_quartet(int(i), int(j), int(k), int(l)).ravel()
        hot += h.stats()["last_dump_kernel"]
        worst = max(worst, np.max(np.abs(blk - ref)))
    assert worst < ERI_TOL, worst
    assert hot >= len(g["quartets"]) // 2, hot       # most fixture quartets belong to the register / tile classes


@pytest.mark.parametrize("name", ["631.nh3", "631.co"])
def test_unique_integral_list_from_the_fock_kernels(capi, name):
    """every stored integral of the reference, with the tile / register kernels producing the blocks of their classes"""
    g = np.load(os.path.join(GOLDEN, "eri_%s.npz" % name.replace(".", "_")))
    b, h = _handle(capi, name)
    h.set_option("reg_kernels", 2); h.set_option("tile_kernels", 2)
    h.set_option("dump_kernel", 1)
    vals, ijkl = h.dump_eris(0.0)
    got = {tuple(r): v for r, v in zip(ijkl.tolist(), vals)}
    worst = 0.0
    for r, v in zip(g["ijkl"].tolist(), g["vals"]):
        worst = max(worst, abs(got.get(tuple(r), 0.0) - v))
    assert worst < ERI_TOL, worst
    h.set_option("dump_kernel", 0)
    vals0, ijkl0 = h.dump_eris(0.0)
    assert np.array_equal(ijkl, ijkl0) and np.max(np.abs(vals - vals0)) < 1e-13


def test_every_shell_quartet_h2o_sto3g_vs_oracle(capi, oracle):
    b, h = _handle(capi, "3g.h2o")
    ob = oracle.basis(golden_input("3g.h2o"))
    ns = b.nshell
    worst = 0.0
    for i in range(ns):
        for j in range(ns):
            for k in range(ns):
                for l in range(ns):
                    worst = max(worst, np.max(np.abs(h.eri_quartet(i, j, k, l) - oracle.quartet_block(ob, i, j, k, l))))
    assert worst < ERI_TOL, worst


@pytest.mark.parametrize("name", ["dh95.co2", "tz2p.sf6"])
def test_random_d_shell_quartets_vs_oracle(capi, oracle, name):
    """covers every class up to (dd|dd) with all orderings of the four shells"""
    b, h = _handle(capi, name)
    ob = oracle.basis(golden_input(name))
    rng = np.random.default_rng(7)
    by_l = {l: [s for s in range(b.nshell) if b.lv[s] == l] for l in (0, 1, 2)}
    worst = 0.0
    for la in (0, 1, 2):
        for lb in (0, 1, 2):
            for lc in (0, 1, 2):
                for ld in (0, 1, 2):
                    for _ in range(2):
                        i, j, k, l = (int(rng.choice(by_l[x])) for x in (la, lb, lc, ld))
                        d = np.max(np.abs(h.eri_quartet(i, j, k, l) - oracle.quartet_block(ob, i, j, k, l)))
                        worst = max(worst, d)
    assert worst < ERI_TOL, worst


@pytest.mark.parametrize("name", ["3g.h2o", "631.nh3", "631.co"])
def test_unique_integral_list_vs_reference_cache(capi, name):
    """all stored integrals of the reference (|val| > 1e-14), same records in the same order"""
    g = np.load(os.path.join(GOLDEN, "eri_%s.npz" % name.replace(".", "_")))
    b, h = _handle(capi, name)
    # thresholding at 1e-14 can flip for values within rounding of the threshold: compare as dense maps
    vals, ijkl = h.dump_eris(0.0)
    got = {tuple(r): v for r, v in zip(ijkl.tolist(), vals)}
    worst = 0.0
    for r, v in zip(g["ijkl"].tolist(), g["vals"]):
        worst = max(worst, abs(got.get(tuple(r), 0.0) - v))
    assert worst < ERI_TOL, worst
    vals14, ijkl14 = h.dump_eris(1e-14)
    assert abs(len(vals14) - len(g["vals"])) <= 2      # borderline |val| ~ 1e-14 records only
    # everything the reference dropped is below its threshold here too
    ref_keys = set(map(tuple, g["ijkl"].tolist()))
    extra = [abs(v) for r, v in got.items() if r not in ref_keys]
    assert not extra or max(extra) < 1e-14 + ERI_TOL


@pytest.mark.parametrize("name", ["3g.h2o", "631.nh3", "631.co", "b.dhdz", "dh95.co2", "dh95.c2h2", "tz2p.sf6", "fg.h2o", "fg2.hf"])
def test_g_matrices_vs_reference_fixture(capi, name):
    g = np.load(os.path.join(GOLDEN, "g_%s.npz" % name.replace(".", "_")))
    b, h = _handle(capi, name)
    h.set_option("schwarz_tau", 0.0)        # no screening: the reference has none
    G = h.fock_rhf(g["P"])
    GA, GB = h.fock_uhf(g["P"], g["PB"])
    scale = max(1.0, np.max(np.abs(g["G"])))
    assert np.max(np.abs(G - g["G"])) < 1e-12 * scale
    assert np.max(np.abs(GA - g["GA"])) < 1e-12 * scale
    assert np.max(np.abs(GB - g["GB"])) < 1e-12 * scale
    # default Schwarz screening (tau = 1e-12) must stay inside the same tolerance
    h.set_option("schwarz_tau", 1e-12)
    G2 = h.fock_rhf(g["P"])
    assert np.max(np.abs(G2 - g["G"])) < 2e-12 * scale * max(1.0, np.max(np.abs(g["P"])))


def test_g_accumulates_like_the_reference(capi):
    """formGmatrix does G += ... (reference RHF.hpp:89-93: the caller zeroes G)"""
    g = np.load(os.path.join(GOLDEN, "g_631_nh3.npz"))
    b, h = _handle(capi, "631.nh3")
    G0 = np.full(b.no2, 0.25)
    G = h.fock_rhf(g["P"], G0.copy())
    assert np.max(np.abs(G - 0.25 - g["G"])) < 1e-12 * max(1.0, np.max(np.abs(g["G"])))


def test_linearity_and_spin_consistency(capi):
    """size-independent properties: G is linear in P; UHF with PA=PB=P gives GA=GB=G_RHF(P)"""
    b, h = _handle(capi, "dh95.c2h2")
    rng = np.random.default_rng(11)
    P1 = rng.standard_normal(b.no2); P2 = rng.standard_normal(b.no2)
    G1 = h.fock_rhf(P1); G2 = h.fock_rhf(P2); G12 = h.fock_rhf(2.0 * P1 - 0.5 * P2)
    s = np.max(np.abs(G12))
    assert np.max(np.abs(G12 - (2.0 * G1 - 0.5 * G2))) < 1e-12 * s
    GA, GB = h.fock_uhf(P1, P1)
    assert np.max(np.abs(GA - G1)) < 1e-12 * s and np.max(np.abs(GB - G1)) < 1e-12 * s


def test_start_shell_matches_oracle(capi, oracle):
    """start_shell > 0 (reference TwoElectronInts.cpp:541; used by findPolarizationPotential)"""
    name = "631.h2o"
    from unomol_b200.basis import Basis
    b = Basis.from_patin(golden_input(name))
    ob = oracle.basis(golden_input(name))
    start = b.nshell - 3
    h = capi.Handle(b, start_shell=start)
    h.set_option("schwarz_tau", 0.0)
    rng = np.random.default_rng(2)
    P = rng.standard_normal(b.no2)
    vals, ijkl, _ = oracle.unique_eris(ob, start_shell=start)
    Gref = oracle.form_g_rhf(vals, ijkl, P)
    G = h.fock_rhf(P)
    assert np.max(np.abs(G - Gref)) < 1e-12 * max(1.0, np.max(np.abs(Gref)))


def test_two_rank_partials_sum_to_full(capi):
    """the multi-GPU split: partial G's of rank 0/2 and 1/2 on one device add up to the 1-rank G"""
    from unomol_b200.basis import Basis
    b = Basis.from_patin(golden_input("dh95.co2"))
    rng = np.random.default_rng(4)
    P = rng.standard_normal(b.no2)
    full = capi.Handle(b).fock_rhf(P)
    parts = [capi.Handle(b, rank=r, nranks=2).fock_rhf(P) for r in range(2)]
    assert np.max(np.abs(parts[0] + parts[1] - full)) < 1e-12 * np.max(np.abs(full))
    assert np.max(np.abs(parts[0])) > 0 and np.max(np.abs(parts[1])) > 0


def test_set_geometry_recalculates(capi, oracle):
    from unomol_b200.basis import Basis
    b = Basis.from_patin(golden_input("631.h2o"))
    h = capi.Handle(b)
    h.set_option("schwarz_tau", 0.0)
    xyz = b.xyz.copy(); xyz[0] += [0.1, -0.2, 0.05]
    h.set_geometry(xyz)
    ob = oracle.basis(golden_input("631.h2o"))
    oracle.lib.oracle_basis_set_center(ob.h, 0, *[float(v) for v in xyz[0]])
    rng = np.random.default_rng(9)
    P = rng.standard_normal(b.no2)
    vals, ijkl, _ = oracle.unique_eris(ob)
    Gref = oracle.form_g_rhf(vals, ijkl, P)
    assert np.max(np.abs(h.fock_rhf(P) - Gref)) < 1e-12 * max(1.0, np.max(np.abs(Gref)))


def test_schwarz_bounds_bound_the_integrals(capi, oracle):
    b, h = _handle(capi, "631.nh3")
    ob = oracle.basis(golden_input("631.nh3"))
    Q = h.schwarz()
    ns = b.nshell
    rng = np.random.default_rng(1)
    for _ in range(40):
        i, j, k, l = (int(x) for x in rng.integers(0, ns, 4))
        blk = oracle.quartet_block(ob, i, j, k, l)
        qa = Q[max(i, j) * (max(i, j) + 1) // 2 + min(i, j)]; qb = Q[max(k, l) * (max(k, l) + 1) // 2 + min(k, l)]
        assert np.max(np.abs(blk)) <= qa * qb * (1 + 1e-9) + 1e-13


@pytest.mark.parametrize("name", ["631.nh3", "dh95.co2", "tz2p.sf6"])
def test_register_kernels_agree_with_generic_kernel(capi, name):
    """the register-resident class kernels (eri_reg.cuh) against the generic shared-memory kernel"""
    b, h = _handle(capi, name)
    rng = np.random.default_rng(21)
    P = rng.standard_normal(b.no2); PB = rng.standard_normal(b.no2)
    h.set_option("reg_kernels", 1)
    G1 = h.fock_rhf(P); GA1, GB1 = h.fock_uhf(P, PB)
    h.set_option("reg_kernels", 0)
    G0 = h.fock_rhf(P); GA0, GB0 = h.fock_uhf(P, PB)
    s = np.max(np.abs(G0))
    assert np.max(np.abs(G1 - G0)) < 1e-13 * s
    assert np.max(np.abs(GA1 - GA0)) < 1e-13 * s and np.max(np.abs(GB1 - GB0)) < 1e-13 * s


def test_primitive_count_buckets_do_not_change_results(capi):
    """pair lists split by primitive-pair count (used for large systems) against unsplit lists"""
    b, h = _handle(capi, "dh95.co2")
    rng = np.random.default_rng(33)
    P = rng.standard_normal(b.no2)
    G0 = h.fock_rhf(P)
    h.set_option("bucket_min_pairs", 1)      # force bucketing on a small molecule
    G1 = h.fock_rhf(P)
    assert h.stats()["n_launches"] > 30
    assert np.max(np.abs(G1 - G0)) < 1e-13 * np.max(np.abs(G0))
    blk0 = h.eri_quartet(35, 2, 17, 30)
    h.set_option("bucket_min_pairs", 10 ** 9)
    assert np.max(np.abs(h.eri_quartet(35, 2, 17, 30) - blk0)) < 1e-14


# ---------------------------------------------------------------- edge cases
def test_h2_minimal_basis_two_functions(capi, oracle):
    """smallest input of the reference's test set: 2 basis functions, 6 unique integrals"""
    b, h = _handle(capi, "3g.h2")
    ob = oracle.basis(golden_input("3g.h2"))
    vals, ijkl, ncalc = oracle.unique_eris(ob)
    assert ncalc == 6
    gv, gi = h.dump_eris(1e-14)
    assert np.array_equal(gi, ijkl) and np.max(np.abs(gv - vals)) < ERI_TOL
    P = np.array([0.3, -0.2, 0.7])
    assert np.max(np.abs(h.fock_rhf(P) - oracle.form_g_rhf(vals, ijkl, P))) < 1e-13


def test_start_shell_past_the_end_gives_zero(capi):
    from unomol_b200.basis import Basis
    b = Basis.from_patin(golden_input("631.nh3"))
    h = capi.Handle(b, start_shell=b.nshell)
    G = h.fock_rhf(np.ones(b.no2))
    assert np.all(G == 0.0)
    assert h.stats()["n_quartets"] == 0


def test_everything_screened_out_gives_zero(capi):
    b, h = _handle(capi, "631.nh3")
    h.set_option("schwarz_tau", 1e30)
    assert np.all(h.fock_rhf(np.ones(b.no2)) == 0.0)


def test_zero_density_and_single_spin(capi):
    b, h = _handle(capi, "dh95.c2h2")
    Z = np.zeros(b.no2)
    assert np.all(h.fock_rhf(Z) == 0.0)
    rng = np.random.default_rng(8)
    PA = rng.standard_normal(b.no2)
    GA, GB = h.fock_uhf(PA, Z)
    # beta sees only the Coulomb field of alpha: GB = J[PA]; GA = J[PA] - K[PA]; RHF(PA) = 2J - K
    G = h.fock_rhf(PA)
    assert np.max(np.abs((GA + GB) - G)) < 1e-12 * np.max(np.abs(G))


def test_bad_arguments_fail_loudly(capi):
    from unomol_b200.basis import Basis
    b = Basis.from_patin(golden_input("3g.h2o"))
    with pytest.raises(capi.UnomolError):
        capi.Handle(b, rank=2, nranks=2)
    h = capi.Handle(b)
    with pytest.raises(capi.UnomolError):
        h.set_option("no_such_option", 1.0)
    import ctypes
    assert capi.lib.unomol_b200_eri_quartet(h.h, 0, 0, 0, 99, None) != 0


def test_h_shells_are_rejected_not_silently_wrong(capi):
    """l > 4 is outside the reference's own range (Basis.hpp:222): create must fail with UNOMOL_E_UNSUPPORTED (-3)"""
    from unomol_b200.basis import Basis
    b = Basis.from_patin(golden_input("3g.h2o"))
    b.lv = b.lv.copy(); b.lv[0] = 5
    with pytest.raises(capi.UnomolError, match="-3"):
        capi.Handle(b)


# ---------------------------------------------------------------- f / g shells
def test_high_l_quartets_vs_oracle_both_algorithms(capi, oracle):
    """fg.h2o: shells up to g.  Quartets with l_tot <= 8 follow the reference's Rys routine, l_tot > 8 its
    McMurchie-Davidson routine incl. the Fgamma t > 20 asymptotic branch (TwoElectronInts.cpp:661-665); one-, two- and
    multi-centre cases, every ordering of the four shells, plus the largest block (gg|gg) = 50 625 integrals."""
    b, h = _handle(capi, "fg.h2o")
    ob = oracle.basis(golden_input("fg.h2o"))
    rng = np.random.default_rng(17)
    cases = [(5, 5, 5, 5), (5, 4, 8, 8), (8, 8, 11, 11), (11, 11, 8, 8), (4, 5, 11, 8), (0, 5, 0, 0), (2, 11, 8, 3), (3, 3, 5, 5)]
    cases += [tuple(int(x) for x in rng.integers(0, b.nshell, 4)) for _ in range(60)]
    worst_rys = worst_md = 0.0
    for (i, j, k, l) in cases:
        d = float(np.max(np.abs(h.eri_quartet(i, j, k, l) - oracle.quartet_block(ob, i, j, k, l))))
        if b.lv[i] + b.lv[j] + b.lv[k] + b.lv[l] > 8:
            worst_md = max(worst_md, d)
        else:
            worst_rys = max(worst_rys, d)
    assert worst_rys < ERI_TOL and worst_md < ERI_TOL, (worst_rys, worst_md)


def test_all_rys_mode_six_to_nine_roots_on_the_gpu(capi, oracle):
    """SURVEY 8(a) row a8 on the device: option all_rys sends l_tot > 8 through the Rys quadrature with 6..9 roots (the range of
    the reference's Rys::rootN, which the MPI build relies on and which crashes there) instead of McMurchie-Davidson.  Blocks
    against the serial reference's values within 1e-10 (primitive cut of the Rys routine; see tests/test_highl_emulation.py for
    the one (ff|ff) block where the reference's McMurchie-Davidson value itself is off by 3.3e-7), the whole G of fg.h2o against
    the default build within 1e-7 relative, and the SCF energy moves by less than 1e-7 Eh."""
    b, h = _handle(capi, "fg.h2o")
    ob = oracle.basis(golden_input("fg.h2o"))
    h.set_option("all_rys", 1)
    rng = np.random.default_rng(19)
    cases = [(5, 5, 5, 5), (5, 4, 8, 8), (11, 11, 8, 8), (4, 5, 11, 8), (5, 5, 11, 11), (3, 3, 5, 5), (5, 5, 5, 4)]
    cases += [tuple(int(x) for x in rng.integers(0, b.nshell, 4)) for _ in range(40)]
    worst = 0.0
    nroots = set()
    for (i, j, k, l) in cases:
        lt = int(b.lv[i] + b.lv[j] + b.lv[k] + b.lv[l])
        if lt <= 8 or sorted((i, j, k, l)) == [8, 8, 11, 11]:
            continue
        nroots.add(lt // 2 + 1)
        worst = max(worst, float(np.max(np.abs(h.eri_quartet(i, j, k, l) - oracle.quartet_block(ob, i, j, k, l)))))
    assert 9 in nroots and 6 in nroots, nroots          # (gg|gg) needs nine roots
    assert worst < 1e-10, worst
    d88 = float(np.max(np.abs(h.eri_quartet(8, 8, 11, 11) - oracle.quartet_block(ob, 8, 8, 11, 11))))
    assert 1e-7 < d88 < 1e-6, d88
    P = rng.standard_normal(b.no2)
    G1 = h.fock_rhf(P)
    assert h.stats()["n_highl_launches"] > 0
    h0 = _handle(capi, "fg.h2o")[1]
    G0 = h0.fock_rhf(P)
    assert 0 < np.max(np.abs(G1 - G0)) < 1e-6 * np.max(np.abs(G0))


def test_high_l_unique_integral_list_vs_oracle(capi, oracle):
    """all 2.6 million function quartets of fg.h2o (2.3 million above the reference's 1e-14 storage threshold)"""
    b, h = _handle(capi, "fg.h2o")
    ob = oracle.basis(golden_input("fg.h2o"))
    vals, ijkl, ncalc = oracle.unique_eris(ob, thresh=-1.0)      # every computed record, reference loop order
    gv, gi = h.dump_eris(-1.0)
    assert len(gv) == ncalc == len(vals) and np.array_equal(gi, ijkl)
    assert np.max(np.abs(gv - vals)) < ERI_TOL


def test_high_l_start_shell_multirank_and_schwarz(capi, oracle):
    b, h = _handle(capi, "fg.h2o")
    ob = oracle.basis(golden_input("fg.h2o"))
    rng = np.random.default_rng(23)
    P = rng.standard_normal(b.no2)
    h.set_option("schwarz_tau", 0.0)
    full = h.fock_rhf(P)
    parts = []
    for r in range(2):
        hr = capi.Handle(b, rank=r, nranks=2)
        hr.set_option("schwarz_tau", 0.0)
        parts.append(hr.fock_rhf(P))
    assert np.max(np.abs(parts[0] + parts[1] - full)) < 1e-12 * np.max(np.abs(full))
    start = b.nshell - 2
    hs = capi.Handle(b, start_shell=start)
    hs.set_option("schwarz_tau", 0.0)
    vals, ijkl, _ = oracle.unique_eris(ob, start_shell=start)
    Gref = oracle.form_g_rhf(vals, ijkl, P)
    assert np.max(np.abs(hs.fock_rhf(P) - Gref)) < 1e-12 * max(1.0, np.max(np.abs(Gref)))
    Q = h.schwarz()
    for _ in range(25):
        i, j, k, l = (int(x) for x in rng.integers(0, b.nshell, 4))
        blk = oracle.quartet_block(ob, i, j, k, l)
        qa = Q[max(i, j) * (max(i, j) + 1) // 2 + min(i, j)]; qb = Q[max(k, l) * (max(k, l) + 1) // 2 + min(k, l)]
        assert np.max(np.abs(blk)) <= qa * qb * (1 + 1e-9) + 1e-13


def test_spatial_blocks_do_not_change_results(capi):
    """pair lists split into spatial blocks (used for N > ~2500 to keep a launch's P/J/K footprint in L2)"""
    b, h = _handle(capi, "tz2p.sf6")
    rng = np.random.default_rng(44)
    P = rng.standard_normal(b.no2)
    G0 = h.fock_rhf(P)
    h.set_option("col_blocks", 3)
    G1 = h.fock_rhf(P)
    assert h.stats()["n_launches"] > 60
    assert np.max(np.abs(G1 - G0)) < 1e-13 * np.max(np.abs(G0))
    GA, GB = h.fock_uhf(P, 0.5 * P)
    h.set_option("col_blocks", 1)
    GA0, GB0 = h.fock_uhf(P, 0.5 * P)
    assert np.max(np.abs(GA - GA0)) < 1e-13 * np.max(np.abs(GA0)) and np.max(np.abs(GB - GB0)) < 1e-13 * np.max(np.abs(GA0))


def test_multi_gpu_work_stealing_and_static_split():
    """needs >= 2 GPUs on the box: torchrun with one rank per GPU, partial G's all-reduced over NCCL"""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("single-GPU box")
    from conftest import ROOT
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multigpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "OK" in p.stdout, (p.stdout[-1500:], p.stderr[-1500:])


@pytest.mark.parametrize("two_devices", [False, True])
def test_in_process_handles_share_work_counters(capi, two_devices):
    """unomol_b200_steal_share: two handles of ONE process (rank 0/2 and 1/2, one host thread each) claim bras from the
    same work counters; their partial G's add up to the 1-rank G over repeated builds.  On one device the two handles
    simply run concurrently; with two devices the counters are reached through NVLink peer access (what the C++ shim does
    with UNOMOL_GPUS=N)."""
    import threading
    import torch
    if two_devices and torch.cuda.device_count() < 2:
        pytest.skip("single-GPU box")
    from unomol_b200.basis import Basis
    b = Basis.from_patin(golden_input("tz2p.sf6"))
    rng = np.random.default_rng(61)
    P = rng.standard_normal(b.no2)
    full = capi.Handle(b).fock_rhf(P)
    h0 = capi.Handle(b, device=0, rank=0, nranks=2)
    h1 = capi.Handle(b, device=1 if two_devices else 0, rank=1, nranks=2)
    h0.steal_share(h1)
    for rep in range(4):                     # alternating counter sets
        parts = [None, None]

        def run(i, h):
            parts[i] = h.fock_rhf(P)
        t = threading.Thread(target=run, args=(1, h1))
        t.start(); run(0, h0); t.join()
        assert np.max(np.abs(parts[0] + parts[1] - full)) < 1e-12 * np.max(np.abs(full)), rep
    n0, n1 = h0.stats()["n_quartets"], h1.stats()["n_quartets"]
    hf = capi.Handle(b); hf.fock_rhf(P)
    assert n0 > 0 and n1 > 0 and n0 + n1 == hf.stats()["n_quartets"]


@pytest.mark.parametrize("static_fraction", [0.0, 0.75, 1.0])
def test_static_plus_stealing_split(capi, static_fraction):
    """SURVEY 8(e) / north_star (4): the blocks of a launch are dealt to the ranks block-cyclically (static share) and the
    tail is stolen from the shared counter.  Two ranks as two handles of this process (one GPU is enough), every kernel family
    (tile and register kernels forced on the small cluster, generic kernel on SF6): the partial G's add up to the 1-rank G and
    every quartet is evaluated exactly once, for a purely stolen (0), mixed (0.75) and purely static (1) hand-out."""
    import threading
    from unomol_b200.basis import Basis, water_cluster
    for b, opts in ((Basis.from_patin(golden_input("tz2p.sf6")), {}), (water_cluster(8), {"tile_kernels": 2, "reg_kernels": 2})):
        rng = np.random.default_rng(67)
        P = rng.standard_normal(b.no2)
        hf = capi.Handle(b)
        hs = [capi.Handle(b, device=0, rank=r, nranks=2) for r in range(2)]
        for h in [hf] + hs:
            for k, v in opts.items():
                h.set_option(k, v)
        full = hf.fock_rhf(P)
        hs[0].steal_share(hs[1])
        for h in hs:
            h.set_option("static_fraction", static_fraction)
        for rep in range(3):
            parts = [None, None]

            def run(i):
                parts[i] = hs[i].fock_rhf(P)
            t = threading.Thread(target=run, args=(1,))
            t.start(); run(0); t.join()
            assert np.max(np.abs(parts[0] + parts[1] - full)) < 1e-12 * np.max(np.abs(full)), (static_fraction, rep)
        n0, n1 = hs[0].stats()["n_quartets"], hs[1].stats()["n_quartets"]
        assert n0 + n1 == hf.stats()["n_quartets"]
        if static_fraction > 0:
            assert n0 > 0 and n1 > 0      # a static share guarantees both ranks work, whoever starts first


def test_device_pair_tables_match_host_pair_tables(capi):
    """pair tables built on the GPU (pair_device.cu) against the threaded host path (engine.cu)"""
    for name in ("631.nh3", "tz2p.sf6"):
        b, h = _handle(capi, name)
        rng = np.random.default_rng(3)
        P = rng.standard_normal(b.no2)
        G1 = h.fock_rhf(P); Q1 = h.schwarz(); s1 = h.stats()
        h.set_option("device_pairs", 0)
        G0 = h.fock_rhf(P); Q0 = h.schwarz(); s0 = h.stats()
        assert s0["n_pairs_kept"] == s1["n_pairs_kept"] and s0["n_prim_pairs"] == s1["n_prim_pairs"]
        assert np.max(np.abs(Q1 - Q0)) < 1e-13 * np.max(Q0)
        assert np.max(np.abs(G1 - G0)) < 1e-13 * np.max(np.abs(G0))


# ---------------------------------------------------------------------------------------------------------------------
# The benchmarked workload: synthetic (H2O)_n / 6-31G clusters against the oracle's unscreened integral-direct G on the same
# P, with every code path that only engages at scale forced on: primitive-count buckets (>= 20 000 shell pairs by default),
# spatial blocks (N > ~2500), the one-bra register kernels with TMA-staged rows of P (tile_kernels = 0), host-built pair
# tables, static rank split.  Tolerance: 1e-12 relative to max|G|, RHF and UHF.
#
# Densities.  Schwarz screening (keep Q_ab Q_cd >= tau) drops integrals of 1e-14 < |v| < tau that the reference keeps (its only
# threshold is |v| > 1e-14 at storage, TwoElectronInts.cpp:513); what that costs G scales with max|P| times the number of
# far pairs.  So:
#   * DEFAULT options (tau = 1e-12) are tested with the densities SURVEY.md 8(d) names for Fock builds -- the superposition
#     of converged monomer densities and an SCF density two iterations downstream of it (dense, exponentially decaying);
#   * a dense standard-normal P (entries up to 4, no decay: nothing an SCF produces) is tested at tau = 1e-14, where the
#     screen is exact with respect to the reference's storage threshold -- this is the all-quartets arithmetic check;
#     measured at tau = 1e-12 it deviates by 1.2e-12 ((H2O)_8) to 1.5e-12 ((H2O)_12) of max|G| (profiles/exp_tau_err.py).
def _cluster(nw, tmp_path, oracle):
    from unomol_b200.basis import Basis, water_cluster
    path = str(tmp_path / ("patin.w%d" % nw))
    water_cluster(nw).write_patin(path)
    # both sides read the FILE: patin.dat carries 10 decimals, the in-memory generator full doubles (5e-11 bohr apart,
    # which is 1e-11 on an integral)
    return Basis.from_patin(path), oracle.basis(path)


_WATER_ORACLE = {}


def _water_oracle(oracle, tmp_path, nw):
    if nw not in _WATER_ORACLE:
        from unomol_b200 import driver
        b, ob = _cluster(nw, tmp_path, oracle)
        rng = np.random.default_rng(100 + nw)
        _, Psup = driver.cluster_superposition_density(nw)
        Pscf = driver.run_scf(b, pmatrix=Psup, maxits=2)["P"]
        dens = {"superposition": (Psup, 0.6 * Psup), "scf": (Pscf, 0.5 * Pscf + 0.25 * Psup),
                "normal": (rng.standard_normal(b.no2), rng.standard_normal(b.no2))}
        ref = {}
        for name, (P, PB) in dens.items():
            ref[name] = (P, PB, oracle.direct_g_threads(ob, P)) + tuple(oracle.direct_g_threads(ob, P, PB))
        _WATER_ORACLE[nw] = (b, ref)
    return _WATER_ORACLE[nw]


# tile_kernels / reg_kernels = 2 force those kernels on lists that the engine would give to the generic kernel because they are
# too short to fill the GPU (what (H2O)_154 runs by default is forced here on (H2O)_8 and (H2O)_12)
@pytest.mark.parametrize("nw,opts", [
    (8, {}),
    (8, {"tile_kernels": 2, "reg_kernels": 2}),
    (8, {"tile_kernels": 2, "reg_kernels": 2, "bucket_min_pairs": 1}),
    (8, {"tile_kernels": 2, "reg_kernels": 2, "bucket_min_pairs": 1, "col_blocks": 3}),
    (8, {"tile_kernels": 0, "reg_kernels": 2}),
    (8, {"tile_kernels": 0, "reg_kernels": 2, "bucket_min_pairs": 1, "col_blocks": 2}),
    (8, {"tile_kernels": 2, "reg_kernels": 2, "device_pairs": 0}),
    (12, {}),
    (12, {"tile_kernels": 2, "reg_kernels": 2}),
    (12, {"tile_kernels": 2, "reg_kernels": 2, "bucket_min_pairs": 1, "col_blocks": 3}),
])
def test_water_cluster_g_vs_oracle(capi, oracle, tmp_path, nw, opts):
    b, ref = _water_oracle(oracle, tmp_path, nw)
    h = capi.Handle(b)
    for k, v in opts.items():
        h.set_option(k, v)
    for name in ("superposition", "scf", "normal"):
        P, PB, G, GA, GB = ref[name]
        h.set_option("schwarz_tau", 1e-14 if name == "normal" else 1e-12)
        scale = np.max(np.abs(G))
        g = h.fock_rhf(P)
        st = h.stats()
        assert np.max(np.abs(g - G)) < 1e-12 * scale, (name, np.max(np.abs(g - G)) / scale, opts)
        if opts.get("tile_kernels") == 2:
            assert st["n_tile_launches"] > 0, st
        elif opts.get("tile_kernels") == 0:
            assert st["n_tile_launches"] == 0 and st["n_reg_launches"] > 0 and st["n_rows_launches"] > 0, st
        if "bucket_min_pairs" in opts:
            assert st["n_launches"] > 30, st          # buckets multiply the launches (6 classes -> 21 without them)
        ga, gb = h.fock_uhf(P, PB)
        scale_u = max(np.max(np.abs(GA)), np.max(np.abs(GB)))
        err_u = max(np.max(np.abs(ga - GA)), np.max(np.abs(gb - GB)))
        assert err_u < 1e-12 * scale_u, (name, err_u / scale_u, opts)
    h.close()


def test_water_cluster_two_rank_static_split_vs_oracle(capi, oracle, tmp_path):
    """the N > 1 static (snake) split on one GPU: two handles with rank 0/2 and 1/2, partial G's summed, against the oracle"""
    b, ref = _water_oracle(oracle, tmp_path, 8)
    P, PB, G, GA, GB = ref["scf"]
    parts = []
    for r in range(2):
        h = capi.Handle(b, rank=r, nranks=2)
        h.set_option("tile_kernels", 2); h.set_option("reg_kernels", 2)
        h.set_option("bucket_min_pairs", 1)
        parts.append(h.fock_rhf(P))
        h.close()
    assert np.max(np.abs(parts[0] + parts[1] - G)) < 1e-12 * np.max(np.abs(G))
    assert np.max(np.abs(parts[0])) > 0 and np.max(np.abs(parts[1])) > 0
