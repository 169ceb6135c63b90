"""tests/multigpu_check.py -- run under torchrun (one rank per GPU): the N-rank Fock build with shared work
counters (work stealing over NVLink) and with the static split must both reproduce the 1-rank G."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from unomol_b200 import capi
from unomol_b200.basis import Basis, water_cluster
from unomol_b200.multigpu import DistributedFock, env_rank

rank, world, local = env_rank()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
worst = 0.0
for basis in (Basis.from_patin(os.path.join(ROOT, "tests", "golden", "inputs", "patin.dat.tz2p.sf6")), water_cluster(8)):
    rng = np.random.default_rng(5)
    P = rng.standard_normal(basis.no2)
    ref = capi.Handle(basis, device=local).fock_rhf(P)          # every rank: the whole build on its own GPU
    df = DistributedFock(basis)
    assert df.stealing, "CUDA IPC work counters could not be shared"
    for rep in range(3):                                        # several builds: exercises the alternating counter sets
        G = df.fock_rhf(P)
        worst = max(worst, float(np.max(np.abs(G - ref)) / np.max(np.abs(ref))))
    for frac in (0.0, 1.0, 0.75):                               # share of the blocks dealt statically: none (the default), all, most
        df.h.set_option("static_fraction", frac)
        for rep in range(2):
            G = df.fock_rhf(P)
            worst = max(worst, float(np.max(np.abs(G - ref)) / np.max(np.abs(ref))))
    df.h.set_option("work_stealing", 0)                         # static snake-order split
    G = df.fock_rhf(P)
    worst = max(worst, float(np.max(np.abs(G - ref)) / np.max(np.abs(ref))))
    # off -> on across an ODD number of builds: the counter set used next must be clean (a dirty set made every CTA
    # break out at once and the build returned a partial G without any error)
    for nstatic in (0, 2):
        for _ in range(nstatic):
            df.fock_rhf(P)
        df.h.set_option("work_stealing", 1)
        for rep in range(3):
            G = df.fock_rhf(P)
            worst = max(worst, float(np.max(np.abs(G - ref)) / np.max(np.abs(ref))))
        df.h.set_option("work_stealing", 0)
    # the all-reduce inside the library (unomol_b200_attach_nccl) instead of torch.distributed
    dn = DistributedFock(basis, in_library_allreduce=True)
    for rep in range(2):
        G = dn.fock_rhf(P)
        worst = max(worst, float(np.max(np.abs(G - ref)) / np.max(np.abs(ref))))
    dn.close()
t = torch.tensor([worst], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("MULTIGPU_CHECK world=%d worst_rel_err=%.3e %s" % (world, t.item(), "OK" if t.item() < 1e-12 else "FAIL"))
dist.destroy_process_group()
sys.exit(0 if t.item() < 1e-12 else 1)
