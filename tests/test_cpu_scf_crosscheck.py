"""CPU: a small RHF in numpy assembled from the CHECKERS only -- oracle two-electron integrals (oracle/unomol_oracle.c),
the host one-electron matrices (unomol_b200/host/OneElectron.hpp through tests/host_emul/onee_check) -- with the
reference's iteration scheme (core-Hamiltonian guess, two plain updates, then mixing on dE >= 0; RHF.hpp:114-168, 536-569).
It pins two things without a GPU: the reference's own STO-3G water energy (test/short.dat.3g.h2o) through that chain, and
the monomer energy / superposition-of-monomers starting density behind profiles/scf_water_cluster.py."""
import json
import os
import subprocess
import numpy as np
import pytest
from conftest import GOLDEN, ROOT, golden_input


@pytest.fixture(scope="module")
def onee_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("onee") / "onee_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "host_emul", "onee_check.cpp")])
    return exe


def _one_electron(onee_check, path, tmp_path):
    out = tmp_path / "o.bin"
    subprocess.run([onee_check, path, str(out)], check=True, capture_output=True)
    return np.fromfile(out).reshape(3, -1)


def _full(packed, n):
    M = np.zeros((n, n))
    M[np.tril_indices(n)] = packed
    return M + M.T - np.diag(np.diag(M))


def _rhf(oracle, ob, S, H, nocc, enuc, maxits=200, eps=1e-11):
    n = ob.nbf
    tri = np.tril_indices(n)
    vals, ijkl, _ = oracle.unique_eris(ob)
    s, U = np.linalg.eigh(_full(S, n))
    X = U / np.sqrt(s)

    def diag(Fp):
        e, W = np.linalg.eigh(X.T @ _full(Fp, n) @ X)
        C = X @ W
        return (C[:, :nocc] @ C[:, :nocc].T)[tri]

    w = np.where(tri[0] == tri[1], 1.0, 2.0)
    P = diag(H)
    Pold = P.copy()
    eold, ediff, e_first = 0.0, 10.0, None
    for it in range(maxits):
        if it >= 2:
            if not ediff < 0.0:
                P = 0.5 * (P + Pold)
        G = oracle.form_g_rhf(vals, ijkl, P)
        e = float(np.sum(w * P * (2.0 * H + G)))
        ediff, eold = e - eold, e
        if e_first is None:
            e_first = e + enuc
        Pold = P.copy()
        P = diag(H + G)
        d = P - Pold
        pdiff = float(np.sqrt(np.sum(w * d * d)) / n)
        if it >= 2 and pdiff < eps and ediff < eps:
            break
    return e + enuc, e_first, P, it + 1


def _enuc(b):
    e = 0.0
    for i in range(b.ncen):
        for j in range(i):
            e += b.charge[i] * b.charge[j] / np.linalg.norm(b.xyz[i] - b.xyz[j])
    return e


def test_numpy_rhf_from_the_checkers_reproduces_the_reference_energy(oracle, onee_check, tmp_path):
    short = json.load(open(os.path.join(GOLDEN, "short_dat.json")))["3g.h2o"]
    path = golden_input("3g.h2o")
    ob = oracle.basis(path)
    S, T, H = _one_electron(onee_check, path, tmp_path)
    e, e_first, P, its = _rhf(oracle, ob, S, H, ob.nelec // 2, _enuc(ob))
    assert abs(e_first - short[0]) < 1e-9          # energy of the core-guess iteration
    assert abs(e - short[1]) < 1e-9, (e, short[1], its)


def test_monomer_energy_and_superposition_density(oracle, onee_check, tmp_path):
    """the 6-31G water monomer of the synthetic clusters: E = -75.9839964657 Eh is what the GPU driver converged to
    (profiles/r1_scf_water154.txt); the rotated block-diagonal starting density carries exactly 5 electron pairs per
    molecule against the cluster's own overlap matrix, i.e. the p blocks are rotated consistently with the geometry"""
    from unomol_b200.basis import water_cluster, water_monomer, superposition_density
    mono = water_monomer()
    pm = str(tmp_path / "mono.dat"); mono.write_patin(pm)
    ob = oracle.basis(pm)
    S, T, H = _one_electron(onee_check, pm, tmp_path)
    e, _, P, its = _rhf(oracle, ob, S, H, 5, _enuc(ob))
    assert abs(e - (-75.9839964657)) < 2e-9, e
    frames = []
    clu = water_cluster(3, frames=frames)
    pc = str(tmp_path / "clu.dat"); clu.write_patin(pc)
    Sc, _, _ = _one_electron(onee_check, pc, tmp_path)
    Pc = superposition_density(P, frames)
    n = clu.nbf
    Pf, Sf = _full(Pc, n), _full(Sc, n)
    nb = mono.nbf
    for m in range(3):
        sl = slice(m * nb, (m + 1) * nb)
        assert abs(np.trace(Pf[sl, sl] @ Sf[sl, sl]) - 5.0) < 1e-9
        # idempotency of the monomer density in the rotated frame: P S P = P
        assert np.max(np.abs(Pf[sl, sl] @ Sf[sl, sl] @ Pf[sl, sl] - Pf[sl, sl])) < 1e-8
