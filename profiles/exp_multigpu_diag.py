"""profiles/exp_multigpu_diag.py -- under torchrun: N-rank G of (H2O)_nw against a 1-rank build of the same P, per engine option."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from unomol_b200 import capi
from unomol_b200.basis import water_cluster
from unomol_b200.multigpu import DistributedFock, env_rank
import bench
rank, world, local = env_rank()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 64
basis = water_cluster(nw); P = bench.synthetic_density(basis)
ref = capi.Handle(basis, device=local).fock_rhf(P)
for opts in ({}, {"tile_kernels": 0}, {"work_stealing": 0}, {"tile_kernels": 0, "work_stealing": 0}, {"reg_kernels": 0}):
    df = DistributedFock(basis)
    for k, v in opts.items(): df.h.set_option(k, v)
    errs = []
    for rep in range(3):
        G = df.fock_rhf(P)
        errs.append(float(np.max(np.abs(G - ref)) / np.max(np.abs(ref))))
    st = df.h.stats()
    if rank == 0:
        print("DIAG world=%d nw=%d %s: rel err per build %s (tile %d reg %d gen %d launches)" % (world, nw, opts, ["%.1e" % e for e in errs], st["n_tile_launches"], st["n_reg_launches"], st["n_generic_launches"]), flush=True)
    dist.barrier()
    df.close()
try:
    dn = DistributedFock(basis, in_library_allreduce=True)
    G = dn.fock_rhf(P)
    if rank == 0: print("DIAG in-library all-reduce: rel err %.1e" % (np.max(np.abs(G - ref)) / np.max(np.abs(ref))), flush=True)
    dn.close()
except Exception as e:
    print("DIAG in-library all-reduce failed on rank %d: %s" % (rank, e), flush=True)
dist.barrier()
dist.destroy_process_group()
