"""GPU (-m gpu): incremental geometry updates -- the hot-path part of the reference's polarisation-potential scan
(reference RHF.hpp:292-388): an extra centre carrying a few shells is moved over a grid, `xints.recalculate(basis)` is
called per point (:354) on a TwoElectronInts built with start_shell = number of molecular shells (:315), and G gets the
quartets that contain at least one of the extra shells.  Here unomol_b200_set_geometry() detects that one centre moved and
rebuilds only the shell pairs that contain its shells (engine.cu: update_pairs_incremental); every other pair keeps its
primitive pairs and Schwarz bound.  Checked per grid point against the oracle (same start_shell semantics,
TwoElectronInts.cpp:541), against a from-scratch build, and timed."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def probe_cluster(nw, tmp_path):
    """water_cluster(nw) plus one extra centre (charge 0) with an s(3), an s(1) and a p(1) shell, written to and read back
    from a patin.dat so that the engine and the oracle see the same numbers"""
    from unomol_b200.basis import Basis, water_cluster, O_631G, H_631G
    w = water_cluster(nw)
    shells = []
    for s in range(w.nshell):
        sl = slice(w.poff[s], w.poff[s] + w.npr[s])
        shells.append((int(w.lv[s]), int(w.cen[s]), w.alpha[sl].copy(), w.coef_raw[sl].copy()))
    pc = w.ncen
    shells.append((0, pc, np.array([4.5, 1.1, 0.35]), np.array([0.15, 0.55, 0.45])))
    shells.append((0, pc, np.array([0.09]), np.array([1.0])))
    shells.append((1, pc, np.array([0.22]), np.array([1.0])))
    b = Basis()
    b.ncen = w.ncen + 1
    b.charge = np.concatenate([w.charge, [0.0]])
    b.xyz = np.vstack([w.xyz, [[40.0, 0.0, 0.0]]])
    b.maxl = 1; b.nelec = w.nelec; b.maxits = 50; b.eps = 1e-10
    b.int_flag = [0, 0]; b.scf_flag = [2, 1, 0]; b.prt_flag = [0, 0, 0]
    b._set_shells(shells)
    path = str(tmp_path / ("patin.probe%d" % nw))
    b.write_patin(path)
    return Basis.from_patin(path), path, w.nshell, pc


def grid(b, npts):
    """a straight line that enters the cluster, passes close to several molecules and leaves again"""
    lo, hi = b.xyz[:-1].min(axis=0), b.xyz[:-1].max(axis=0)
    start = lo - np.array([6.0, 1.3, 0.7]); end = hi + np.array([5.0, -0.9, 1.1])
    t = np.linspace(0.0, 1.0, npts)[:, None]
    return np.round(start + t * (end - start), 10)       # 10 decimals: what a pos.grid.dat / patin.dat carries


def test_moving_centre_100_points_vs_oracle(oracle, tmp_path):
    from unomol_b200 import capi
    b, path, ns_mol, pc = probe_cluster(8, tmp_path)
    ob = oracle.basis(path)
    P = np.random.default_rng(3).uniform(-1.0, 1.0, b.no2)
    h = capi.Handle(b, start_shell=ns_mol)
    h.set_option("schwarz_tau", 1e-14)              # dense random P: exact-parity threshold (see test_gpu_parity.py)
    xyz = b.xyz.copy()
    worst = worst_fresh = 0.0
    pts = grid(b, 100)
    for ip, p in enumerate(pts):
        xyz[pc] = p
        h.set_geometry(xyz)
        G = h.fock_rhf(P)
        oracle.set_center(ob, pc, p)
        Gr = oracle.direct_g_start_threads(ob, P, ns_mol)
        scale = max(np.max(np.abs(Gr)), 1e-3)
        worst = max(worst, np.max(np.abs(G - Gr)) / scale)
        if ip % 25 == 7:
            b2 = b
            b2.xyz = xyz.copy()
            hf = capi.Handle(b2, start_shell=ns_mol)
            hf.set_option("schwarz_tau", 1e-14)
            worst_fresh = max(worst_fresh, np.max(np.abs(hf.fock_rhf(P) - G)) / scale)
            hf.close()
    st = h.stats()
    assert st["n_incremental_updates"] == len(pts), st
    assert worst < 1e-12, worst
    assert worst_fresh < 1e-13, worst_fresh


def test_moving_centre_setup_time_at_416_functions(tmp_path):
    """per-point set-up (pair records of the moved shells + their Schwarz bounds + plans) at (H2O)_32 + probe, 421 functions"""
    from unomol_b200 import capi
    b, path, ns_mol, pc = probe_cluster(32, tmp_path)
    P = np.random.default_rng(4).uniform(-1.0, 1.0, b.no2)
    h = capi.Handle(b, start_shell=ns_mol)
    t_full = h.stats()["precompute_ms"]
    xyz = b.xyz.copy()
    times = []
    pts = grid(b, 100)
    for ip, p in enumerate(pts):
        xyz[pc] = p
        h.set_geometry(xyz)
        times.append(h.stats()["precompute_ms"])
        if ip in (10, 50, 90):
            G = h.fock_rhf(P)
            b.xyz = xyz.copy()
            hf = capi.Handle(b, start_shell=ns_mol)
            Gf = hf.fock_rhf(P)
            hf.close()
            assert np.max(np.abs(G - Gf)) < 1e-13 * max(np.max(np.abs(Gf)), 1e-3)
    med = float(np.median(times))
    print("incremental set_geometry at %d functions: median %.3f ms, max %.3f ms; full rebuild %.3f ms" % (b.nbf, med, max(times), t_full))
    assert h.stats()["n_incremental_updates"] == len(pts)
    # measured on B200: 2.6-2.9 ms per point against 15 ms for a from-scratch build of the same tables (0.05 ms pair records,
    # 0.16 ms Schwarz bounds, 0.85 ms list merge + upload, 1.5 ms plans); the SCF that follows each point costs ~100x that
    assert med <= 4.0 and med < 0.4 * t_full, (med, t_full)


def test_full_rebuild_when_many_centres_move(tmp_path):
    from unomol_b200 import capi
    b, path, ns_mol, pc = probe_cluster(4, tmp_path)
    P = np.random.default_rng(5).uniform(-1.0, 1.0, b.no2)
    h = capi.Handle(b)
    xyz = b.xyz.copy()
    xyz[:6] += 0.05                                  # two molecules move: more than a quarter of the shells
    h.set_geometry(xyz)
    assert h.stats()["n_incremental_updates"] == 0
    G = h.fock_rhf(P)
    b.xyz = xyz.copy()
    hf = capi.Handle(b)
    assert np.max(np.abs(hf.fock_rhf(P) - G)) < 1e-13 * np.max(np.abs(G))
    # a different single centre after the probe: the reserved primitive-pair tail belongs to the first set -> full rebuild
    xyz[pc] += 0.3; h.set_geometry(xyz)
    n1 = h.stats()["n_incremental_updates"]
    xyz[0] += 0.1; h.set_geometry(xyz)
    assert h.stats()["n_incremental_updates"] == n1
    b.xyz = xyz.copy()
    hf2 = capi.Handle(b)
    G2 = h.fock_rhf(P)
    assert np.max(np.abs(hf2.fock_rhf(P) - G2)) < 1e-13 * np.max(np.abs(G2))
