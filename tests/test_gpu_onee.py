"""GPU (-m gpu): one-electron and moment matrices from the device kernel (csrc/onee_device.cu, unomol_b200_one_electron)
against the reference's OneElectronInts / MomentInts outputs (fixtures g_*.npz: S, T, H; momints_*.npz: RMOM.DAT of fresh
reference runs), incl. d, f and g shells, and the SCF driver on them (energies at 1e-9 Eh are in test_gpu_scf.py, which runs
the driver with the device integrals by default)."""
import os
import numpy as np
import pytest
from conftest import GOLDEN, golden_input

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["3g.h2o", "631.nh3", "631.co", "b.dhdz", "dh95.co2", "dh95.c2h2", "tz2p.sf6", "fg.h2o", "fg2.hf"])
def test_s_t_h_vs_reference_fixture(name):
    from unomol_b200 import capi
    from unomol_b200.basis import Basis
    g = np.load(os.path.join(GOLDEN, "g_%s.npz" % name.replace(".", "_")))
    b = Basis.from_patin(golden_input(name))
    h = capi.Handle(b)
    S, T, H = h.one_electron(b.charge)
    assert np.max(np.abs(S - g["S"])) < 1e-13, np.max(np.abs(S - g["S"]))
    assert np.max(np.abs(T - g["T"])) < 1e-12 * max(1.0, np.max(np.abs(g["T"])))
    # H = T + V: V sums Z_C <a|1/r_C|b> over the nuclei; 1e-12 relative to the largest element (|H| up to ~500 for S 1s)
    assert np.max(np.abs(H - g["H"])) < 1e-12 * max(1.0, np.max(np.abs(g["H"]))), np.max(np.abs(H - g["H"]))


@pytest.mark.parametrize("name", ["631.nh3", "dh95.co2", "fg.h2o", "fg2.hf"])
def test_moment_integrals_vs_reference_rmom(name):
    from unomol_b200 import capi
    from unomol_b200.basis import Basis
    ref = np.load(os.path.join(GOLDEN, "momints_%s.npz" % name.replace(".", "_")))["m"]
    b = Basis.from_patin(golden_input(name))
    h = capi.Handle(b)
    S, T, H, M = h.one_electron(b.charge, moments=True)
    assert M.shape == ref.shape
    assert np.max(np.abs(M - ref)) < 1e-12, np.max(np.abs(M - ref))


def test_water_cluster_one_electron_time_and_host_agreement(tmp_path):
    """(H2O)_32: device kernel against the threaded host implementation (the driver's --onee dump), and its time"""
    import subprocess
    from unomol_b200 import capi, driver
    from unomol_b200.basis import Basis, water_cluster
    path = str(tmp_path / "patin.dat")
    water_cluster(32).write_patin(path)
    b = Basis.from_patin(path)
    h = capi.Handle(b)
    S, T, H = h.one_electron(b.charge)
    ms = h.stats()["onee_ms"]
    out = str(tmp_path / "onee.bin")
    subprocess.run([driver.BIN, "--onee", path, out], check=True, capture_output=True)
    ref = np.fromfile(out).reshape(3, -1)
    assert np.max(np.abs(S - ref[0])) < 1e-13 and np.max(np.abs(T - ref[1])) < 1e-12
    assert np.max(np.abs(H - ref[2])) < 1e-12 * np.max(np.abs(ref[2]))
    print("one-electron kernel at %d functions: %.2f ms" % (b.nbf, ms))
