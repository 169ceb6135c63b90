"""CPU: the host-side dipole/quadrupole integrals (unomol_b200/host/Moments.hpp) against the reference's own raw
moment integrals (RMOM.DAT of fresh reference runs, tests/golden/momints_*.npz).  No GPU involved: the moment analysis
is O(N^2) post-processing of the converged density; tests/test_gpu_scf.py checks the
`moments.out` the SCF driver writes."""
import os
import subprocess
import numpy as np
import pytest
from conftest import GOLDEN, ROOT, golden_input


@pytest.fixture(scope="module")
def moments_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("mom") / "moments_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host_emul", "moments_check.cpp")])
    return exe


@pytest.mark.parametrize("name", ["631.nh3", "dh95.co2", "fg.h2o", "fg2.hf"])
def test_moment_integrals_match_reference(moments_check, tmp_path, name):
    ref = np.load(os.path.join(GOLDEN, "momints_%s.npz" % name.replace(".", "_")))["m"]
    out = tmp_path / "m.bin"
    subprocess.run([moments_check, golden_input(name), str(out)], check=True, capture_output=True)
    ours = np.fromfile(out).reshape(9, -1)
    assert ours.shape == ref.shape
    assert np.max(np.abs(ours - ref)) < 1e-12


@pytest.fixture(scope="module")
def onee_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("onee") / "onee_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "host_emul", "onee_check.cpp")])
    return exe


@pytest.mark.parametrize("name", ["631.nh3", "dh95.co2", "tz2p.sf6", "fg.h2o", "fg2.hf"])
@pytest.mark.parametrize("threads", ["1", "4"])
def test_one_electron_matrices_match_reference(onee_check, tmp_path, name, threads):
    """overlap, kinetic and core-Hamiltonian matrices of the host driver (threaded over shell-pair rows) against the
    reference's OneElectronInts (fixtures g_*.npz), incl. f/g shells; independent of the thread count"""
    g = np.load(os.path.join(GOLDEN, "g_%s.npz" % name.replace(".", "_")))
    out = tmp_path / "o.bin"
    subprocess.run([onee_check, golden_input(name), str(out)], check=True, capture_output=True,
                   env=dict(os.environ, UNOMOL_HOST_THREADS=threads))
    S, T, H = np.fromfile(out).reshape(3, -1)
    assert np.max(np.abs(S - g["S"])) < 1e-12
    assert np.max(np.abs(T - g["T"])) < 1e-11
    assert np.max(np.abs(H - g["H"])) < 1e-10 * max(1.0, np.max(np.abs(g["H"])))
