"""profiles/exp_variant.py [nwater] -- ms per Fock build (RHF, UHF) of the library selected by UNOMOL_B200_LIB (A/B kernel
variants built with `make lib OBJDIR=... LIB=... EXTRA_NVFLAGS=-D...`), plus per-class kernel shares when UNOMOL_CLASSES=1."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from unomol_b200 import basis as B, capi
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 154
opts = dict(kv.split("=") for kv in sys.argv[2:])
basis = B.water_cluster(nw); P = bench.synthetic_density(basis)
h = capi.Handle(basis)
for k, v in opts.items(): h.set_option(k, float(v))
for _ in range(3): G = h.fock_rhf(P)
st = h.stats()
msg = "%s %s: RHF %.1f ms (%d quartets)" % (os.path.basename(capi.LIB_PATH), opts, st["last_fock_ms"], st["n_quartets"])
for _ in range(2): h.fock_uhf(P, 0.5 * P)
print(msg + ", UHF %.1f ms" % h.stats()["last_fock_ms"], " checksum %.12e" % float(np.sum(G * np.arange(len(G)) % 7)))
