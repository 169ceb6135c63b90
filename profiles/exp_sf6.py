import sys, os; sys.path.insert(0, '/root/repo')
import numpy as np, bench
from unomol_b200 import basis as B, capi
for wl in ("sf6", "co2"):
    basis = bench.WORKLOADS[wl][1](B); P = bench.synthetic_density(basis)
    for opts in ({}, {"tile_kernels": 0}, {"tile_kernels": 0, "reg_kernels": 0}):
        h = capi.Handle(basis)
        for k, v in opts.items(): h.set_option(k, v)
        for _ in range(5): h.fock_rhf(P)
        st = h.stats()
        print(wl, opts, "fock %.2f ms kernels %.2f ms launches %d (tile %d reg %d gen %d)" % (st["last_fock_ms"], st["last_eri_kernel_ms"], st["n_launches"], st["n_tile_launches"], st["n_reg_launches"], st["n_generic_launches"]))
        h.close()
