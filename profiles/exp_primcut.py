import sys; sys.path.insert(0,'/root/repo')
import numpy as np, bench
from unomol_b200 import basis as B, capi
basis=B.water_cluster(154); h=capi.Handle(basis); P=bench.synthetic_density(basis)
for cut in [1e-12,1e-9,1e-6,1e-3,1.0]:
    h.set_option('prim_cut',cut); 
    for _ in range(2): h.fock_rhf(P)
    st=h.stats(); print('prim_cut %.0e: fock %.1f ms quartets %.3e model GF %.1f'%(cut,st['last_fock_ms'],st['n_quartets'],st['model_flops']/1e9))
