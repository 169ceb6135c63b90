#!/usr/bin/env python3
"""profiles/scf_water_cluster.py [n] -- integral-direct RHF on the synthetic (H2O)_n / 6-31G cluster through the host
driver (unomol_b200/unomol_b200_scf = the reference's Unomol.cc on the GPU engine), started from a superposition of
monomer densities written as PMATRIX.DAT (the reference's own restart path, RHF.hpp:120-123), because the reference's
core-Hamiltonian guess + fixed damping does not converge a cluster of this size.  Prints energies, iterations, seconds."""
import os, re, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unomol_b200.basis import water_cluster, water_monomer, superposition_density

BIN = os.path.join(ROOT, "unomol_b200", "unomol_b200_scf")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 154


def run(basis, pmat=None):
    d = tempfile.mkdtemp()
    if pmat is not None:
        basis.scf_flag = [2, 1, 1]
        np.asarray(pmat, dtype=np.float64).tofile(os.path.join(d, "PMATRIX.DAT"))
    basis.write_patin(os.path.join(d, "patin.dat"))
    t0 = time.time()
    p = subprocess.run([BIN], cwd=d, capture_output=True, text=True)
    dt = time.time() - t0
    assert p.returncode == 0, p.stderr[-2000:]
    e0, e1, de = [float(x) for x in open(os.path.join(d, "short.gs.out")).read().split()]
    out = open(os.path.join(d, "scfout.gs.out")).read()
    its = int(re.search(r"Final Iteration\s*=\s*(\d+)", out).group(1))
    P = np.fromfile(os.path.join(d, "PMATRIX.DAT"))
    return dict(e_first=e0, e_final=e1, de=de, iterations=its, converged="NOT_ REACHED" not in out, seconds=dt, P=P,
                stderr=p.stderr)


mono = run(water_monomer())
print("monomer: E = %.10f Eh, %d iterations" % (mono["e_final"], mono["iterations"]))
frames = []
clu = water_cluster(n, frames=frames)
r = run(clu, superposition_density(mono["P"], frames))
m = re.search(r"SCF time = ([\d.]+)", r["stderr"])
print("(H2O)_%d, %d functions: E(first iteration) = %.8f, E(final) = %.8f Eh, dE = %.2e, %d iterations, converged %s, "
      "%.1f s wall (%s s in the SCF loop), E - n*E(monomer) = %.6f Eh"
      % (n, clu.nbf, r["e_first"], r["e_final"], r["de"], r["iterations"], r["converged"], r["seconds"],
         m.group(1) if m else "?", r["e_final"] - n * mono["e_final"]))
