"""profiles/exp_parts.py -- where does the per-quartet time of the register kernels go?  Times the (H2O)_154 build with
parts of the kernel disabled through the debug_flags option (results are invalid in those runs; timing only)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from unomol_b200 import basis as B, capi
basis = B.water_cluster(154); h = capi.Handle(basis); P = bench.synthetic_density(basis)
for flags, what in [(0, "full"), (1, "no exchange digestion"), (4, "no digestion at all"), (2, "no root evaluation (n=1 classes)"), (6, "neither"), (8, "no integral evaluation (register classes)"), (12, "no evaluation, no digestion: fetch + scan + prefactor only")]:
    h.set_option("debug_flags", flags)
    for _ in range(2): h.fock_rhf(P)
    st = h.stats(); print("%-36s fock %.1f ms" % (what, st["last_fock_ms"]))
