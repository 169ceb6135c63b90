"""profiles/exp_candidates.py -- primitive-quartet candidates tested vs survivors per shell quartet, (H2O)_154."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from unomol_b200 import basis as B, capi
basis = B.water_cluster(154); h = capi.Handle(basis); P = bench.synthetic_density(basis)
for _ in range(2): h.fock_rhf(P)
st = h.stats()
print("quartets %.3e  survivors %.3e (%.2f per quartet)  candidates tested %.3e (%.2f per quartet)  fock %.1f ms  precompute %.1f ms" % (
    st["n_quartets"], st["n_prim_quartets"], st["n_prim_quartets"] / st["n_quartets"], st["n_prim_candidates"],
    st["n_prim_candidates"] / st["n_quartets"], st["last_fock_ms"], st["precompute_ms"]))
