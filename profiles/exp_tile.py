"""profiles/exp_tile.py [nwater] -- bra-tile kernels (eri_tile.cuh) against the one-bra register kernels (eri_reg.cuh) on a
water cluster: ms per Fock build, per-class plan statistics, and the G of the two paths against each other."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from unomol_b200 import basis as B, capi
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 154
basis = B.water_cluster(nw); P = bench.synthetic_density(basis)
out = {}
for tile in (1, 0):
    h = capi.Handle(basis)
    h.set_option("tile_kernels", tile)
    for _ in range(3): G = h.fock_rhf(P)
    st = h.stats(); out[tile] = G
    print("tile_kernels=%d: fock %.1f ms, kernels %.1f ms, %d quartets, %.1f model GF, launches %d (tile %d reg %d generic %d)" % (
        tile, st["last_fock_ms"], st["last_eri_kernel_ms"], st["n_quartets"], st["model_flops"] / 1e9, st["n_launches"],
        st["n_tile_launches"], st["n_reg_launches"], st["n_generic_launches"]))
    if nw >= 64:
        ga, gb = h.fock_uhf(P, 0.5 * P); ga, gb = h.fock_uhf(P, 0.5 * P)
        print("   UHF: fock %.1f ms" % h.stats()["last_fock_ms"])
    h.close()
print("max |G_tile - G_reg| / max|G| = %.2e" % (np.max(np.abs(out[1] - out[0])) / np.max(np.abs(out[0]))))
