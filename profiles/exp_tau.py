"""profiles/exp_tau.py -- what the Schwarz threshold tau costs and what it loses.
(1) (H2O)_8: max |G - G_oracle| / max|G| against the unscreened oracle for tau in {1e-12, 1e-13, 1e-14, 0} and three
    densities: standard-normal P, the bench's synthetic P, a superposition of converged monomer densities;
(2) (H2O)_154: quartets and ms per build for the same thresholds."""
import sys, os, tempfile, subprocess; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from unomol_b200 import basis as B, capi
from oracle.oracle import Oracle
O = Oracle()
nw = 8
frames = []
b = B.water_cluster(nw, frames=frames)
d = tempfile.mkdtemp(); path = os.path.join(d, "patin.w"); b.write_patin(path); ob = O.basis(path)
# monomer density through the host driver
md = tempfile.mkdtemp(); B.water_monomer().write_patin(os.path.join(md, "patin.dat"))
subprocess.run([os.path.join(os.path.dirname(capi.LIB_PATH), "unomol_b200_scf")], cwd=md, capture_output=True)
Pm = np.fromfile(os.path.join(md, "PMATRIX.DAT"))
dens = {"normal": np.random.default_rng(1).standard_normal(b.no2), "bench": bench.synthetic_density(b),
        "superposition": B.superposition_density(Pm, frames)}
h = capi.Handle(b)
for name, P in dens.items():
    G = O.direct_g_threads(ob, P)
    for tau in (1e-12, 1e-13, 1e-14, 0.0):
        h.set_option("schwarz_tau", tau)
        g = h.fock_rhf(P)
        print("(H2O)_8 %-14s tau %-6g  max|dG|/max|G| = %.2e   (max|G| %.3g, %d quartets)" % (name, tau, np.max(np.abs(g - G)) / np.max(np.abs(G)), np.max(np.abs(G)), h.stats()["n_quartets"]))
h.close()
b = B.water_cluster(154); P = bench.synthetic_density(b); h = capi.Handle(b)
for tau in (1e-12, 1e-13, 1e-14):
    h.set_option("schwarz_tau", tau)
    for _ in range(3): h.fock_rhf(P)
    st = h.stats(); print("(H2O)_154 tau %-6g: %.1f ms, %d quartets, %d primitive quartets" % (tau, st["last_fock_ms"], st["n_quartets"], st["n_prim_quartets"]))
