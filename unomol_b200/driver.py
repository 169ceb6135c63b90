"""unomol_b200/driver.py -- run the host SCF driver (unomol_b200/unomol_b200_scf = the reference's Unomol.cc call surface on
the GPU engine) the way the reference's test script runs its binary: patin.dat in, short.gs.out / scfout.gs.out /
PMATRIX.DAT out (reference test/tstscr, RHF.hpp:120-123,165-176).  Used by bench.py and the GPU tests to obtain the
densities SURVEY.md 8(d) names for stand-alone Fock-build timing: a superposition of converged monomer densities and SCF
densities started from it."""
import os
import re
import subprocess
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "unomol_b200_scf")


def run_scf(basis, pmatrix=None, maxits=None, env=None, timeout=3600):
    """one driver run in a scratch directory; pmatrix = packed starting density (PMATRIX.DAT restart, scf_flag[2] = 1)"""
    if not os.path.exists(BIN):
        raise RuntimeError("unomol_b200_scf is not built (make -C unomol_b200/csrc)")
    d = tempfile.mkdtemp(prefix="unomol_drv_")
    saved = (list(basis.scf_flag), basis.maxits)
    try:
        if pmatrix is not None:
            basis.scf_flag = [basis.scf_flag[0], basis.scf_flag[1], 1]
            np.asarray(pmatrix, dtype=np.float64).tofile(os.path.join(d, "PMATRIX.DAT"))
        if maxits is not None:
            basis.maxits = int(maxits)
        basis.write_patin(os.path.join(d, "patin.dat"))
    finally:
        basis.scf_flag, basis.maxits = saved
    p = subprocess.run([BIN], cwd=d, capture_output=True, text=True, timeout=timeout,
                       env=dict(os.environ, UNOMOL_SKIP_FINITE_FIELD="1", **(env or {})))
    if p.returncode != 0:
        raise RuntimeError("unomol_b200_scf failed: " + p.stderr[-2000:])
    e0, e1, de = [float(x) for x in open(os.path.join(d, "short.gs.out")).read().split()]
    out = open(os.path.join(d, "scfout.gs.out")).read()
    m = re.search(r"Final Iteration\s*=\s*(\d+)", out)
    return {"e_first": e0, "e_final": e1, "de": de, "iterations": int(m.group(1)) if m else None,
            "converged": "NOT_ REACHED" not in out, "P": np.fromfile(os.path.join(d, "PMATRIX.DAT")), "dir": d,
            "stderr": p.stderr, "scfout": out}


_MONOMER = {}


def water_monomer_density():
    """converged RHF/6-31G density of one water molecule in the reference orientation (packed), cached per process"""
    if "P" not in _MONOMER:
        from .basis import water_monomer
        r = run_scf(water_monomer())
        if not r["converged"]:
            raise RuntimeError("monomer SCF did not converge")
        _MONOMER["P"], _MONOMER["E"] = r["P"], r["e_final"]
    return _MONOMER["P"]


def cluster_superposition_density(n, **kw):
    """(basis, packed P): water_cluster(n) and the block-diagonal superposition of rotated monomer densities"""
    from .basis import water_cluster, superposition_density
    frames = []
    b = water_cluster(n, frames=frames, **kw)
    return b, superposition_density(water_monomer_density(), frames)
