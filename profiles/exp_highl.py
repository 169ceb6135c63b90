"""profiles/exp_highl.py -- timing of the runtime-L kernel (f/g shells) on fg.h2o, and of the pair-table precompute.
Run on the GPU box: python profiles/exp_highl.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unomol_b200 import capi
from unomol_b200.basis import Basis, test_input, water_cluster

b = Basis.from_patin(test_input("fg.h2o"))
h = capi.Handle(b)
rng = np.random.default_rng(0)
P = rng.standard_normal(b.no2)
for tau in (0.0, 1e-12):
    h.set_option("schwarz_tau", tau)
    for rep in range(3):
        h.fock_rhf(P)
        s = h.stats()
    print("fg.h2o tau=%g: %.2f ms per Fock build, %d quartets, %d primitive quartets, %d launches" %
          (tau, s["last_fock_ms"], s["n_quartets"], s["n_prim_quartets"], s["n_launches"]))
t0 = time.time(); blk = h.eri_quartet(5, 5, 5, 5); print("(gg|gg) block: %.1f ms" % ((time.time() - t0) * 1e3))
for n in (32, 154):
    w = water_cluster(n)
    hw = capi.Handle(w)
    print("water%d create precompute_ms %.1f" % (n, hw.stats()["precompute_ms"]))
    for rep in range(2):
        hw.set_geometry(w.xyz)
        print("water%d set_geometry precompute_ms %.1f" % (n, hw.stats()["precompute_ms"]))
    hw.close()
