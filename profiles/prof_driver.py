#!/usr/bin/env python3
"""profiles/prof_driver.py <workload> <builds> -- minimal Fock-build loop to run under ncu (never a bench value)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from unomol_b200 import basis as B, capi
wl = sys.argv[1] if len(sys.argv) > 1 else "water154"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 2
basis = bench.WORKLOADS[wl][1](B)
h = capi.Handle(basis)
import os
if os.environ.get('UNOMOL_BUCKET_MIN'):
    h.set_option('bucket_min_pairs', float(os.environ['UNOMOL_BUCKET_MIN']))
for kv in filter(None, os.environ.get('UNOMOL_OPTS', '').split(',')):
    h.set_option(kv.split('=')[0], float(kv.split('=')[1]))
P = bench.synthetic_density(basis)
for _ in range(nb):
    G = h.fock_rhf(P)
st = h.stats()
print(wl, "fock ms %.2f kernel ms %.2f quartets %d model GF %.1f launches %d" % (st["last_fock_ms"], st["last_eri_kernel_ms"], st["n_quartets"], st["model_flops"] / 1e9, st["n_launches"]))
