"""profiles/exp_uhf.py -- UHF vs RHF Fock-build time on the bench workload ((H2O)_154 / 6-31G) and on SF6/TZ2P."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from unomol_b200 import basis as B, capi
for wl in ("sf6", "water154"):
    basis = bench.WORKLOADS[wl][1](B)
    h = capi.Handle(basis)
    P = bench.synthetic_density(basis)
    PB = bench.synthetic_density(basis, seed=7)
    for rep in range(3):
        h.fock_rhf(P); r = h.stats()["last_fock_ms"]
    for rep in range(3):
        h.fock_uhf(P, PB); u = h.stats()["last_fock_ms"]
    print("%s: RHF %.2f ms, UHF %.2f ms per Fock build (%d quartets)" % (wl, r, u, h.stats()["n_quartets"]))
    h.close()
