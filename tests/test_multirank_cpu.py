"""CPU: the N>1 plumbing with world_size 2 over gloo (no GPU): two ranks digest disjoint shares of the unique
integrals with the oracle and the partial G's are summed by unomol_b200.multigpu.allreduce_packed, exactly the
bracket bench.py and DistributedFock put around the GPU build (reference: MPI_Reduce(G), RHF_MPI.hpp:108)."""
import os
import sys
import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle.oracle import Oracle
    from unomol_b200.multigpu import allreduce_packed, owner_of
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O = Oracle()
    b = O.basis(os.path.join(ROOT, "tests", "golden", "inputs", "patin.dat.631.nh3"))
    vals, ijkl, _ = O.unique_eris(b)
    rng = np.random.default_rng(17)
    P = rng.standard_normal(b.no2)
    mine = np.array([owner_of(i, world) == rank for i in range(len(vals))])
    G = O.form_g_rhf(np.ascontiguousarray(vals[mine]), np.ascontiguousarray(ijkl[mine]), P)
    allreduce_packed(G)
    full = O.form_g_rhf(vals, ijkl, P)
    q.put((rank, float(np.max(np.abs(G - full))), float(np.max(np.abs(full)))))
    dist.destroy_process_group()


def test_two_rank_partial_g_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, scale in res:
        assert err < 1e-12 * scale, (rank, err)


def test_round_robin_ownership_is_a_partition():
    from unomol_b200.multigpu import owner_of
    for n in (1, 2, 4, 8):
        owners = [owner_of(i, n) for i in range(1000)]
        assert set(owners) == set(range(n))
        counts = np.bincount(owners, minlength=n)
        assert counts.max() - counts.min() <= 1
