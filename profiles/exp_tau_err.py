import sys, os, tempfile; sys.path.insert(0, '/root/repo')
import numpy as np
from unomol_b200 import basis as B, capi
from oracle.oracle import Oracle
O = Oracle()
for nw in (8, 12):
    d = tempfile.mkdtemp(); path = os.path.join(d, "patin.w"); B.water_cluster(nw).write_patin(path)
    b = B.Basis.from_patin(path); ob = O.basis(path)
    h = capi.Handle(b)
    for name, P in (("normal", np.random.default_rng(100 + nw).standard_normal(b.no2)), ("uniform", np.random.default_rng(3).uniform(-1, 1, b.no2))):
        G = O.direct_g_threads(ob, P)
        for tau in (1e-12, 5e-13, 2e-13, 1e-13, 1e-14):
            h.set_option("schwarz_tau", tau)
            g = h.fock_rhf(P)
            print("(H2O)_%d %-8s tau %-6g max|dG|/max|G| = %.2e" % (nw, name, tau, np.max(np.abs(g - G)) / np.max(np.abs(G))))
