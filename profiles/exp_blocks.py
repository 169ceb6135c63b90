"""profiles/exp_blocks.py <n_waters> -- Fock-build time vs number of spatial blocks per pair list (L2 residency of P/J/K)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from unomol_b200 import basis as B, capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 154
basis = B.water_cluster(n); h = capi.Handle(basis); P = bench.synthetic_density(basis)
for nb in [1, 2, 3, 4]:
    h.set_option("col_blocks", nb)
    for _ in range(2): h.fock_rhf(P)
    st = h.stats(); print("(H2O)_%d col_blocks %d: fock %.1f ms, %d launches, quartets %.3e" % (n, nb, st["last_fock_ms"], st["n_launches"], st["n_quartets"]))
