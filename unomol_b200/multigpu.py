"""unomol_b200/multigpu.py -- one process per GPU (torch.distributed, NCCL over NVLink) around the C ABI.

Replaces the reference's MPI layer for this path: round-robin ownership of work + MPI_Bcast(P) / MPI_Reduce(G)
around formGmatrix (reference TwoElectronIntsMPI.cpp:350-354, RHF_MPI.hpp:101-111).  Here every rank holds the
same P, builds the partial G of its share of the screened quartets on its own GPU (the share is dealt inside the
library: bra pairs of each Schwarz-sorted list round-robin over ranks, see eri_reg.cuh / eri_generic.cuh) and the
partials are summed with ONE all-reduce of the packed G (ncclAllReduce, FP64 sum) -- the only collective of the
path.  torch.distributed is plumbing only."""
import os
import numpy as np


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def owner_of(index, nranks):
    """rank that owns position `index` of a Schwarz-sorted bra list (blocks of nranks in snake order, as the kernels deal them)"""
    block, pos = divmod(index, nranks)
    return pos if block % 2 == 0 else nranks - 1 - pos


def allreduce_packed(G, group=None):
    """sum the ranks' partial packed G in place; G is a torch tensor (CUDA -> NCCL, CPU -> gloo) or a numpy array"""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return G
    if isinstance(G, np.ndarray):
        t = torch.from_numpy(G)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return G
    dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)
    return G


def enable_work_stealing(handle):
    """share rank 0's work counters with every rank (CUDA IPC handle broadcast); returns False when IPC is unavailable"""
    import torch.distributed as dist
    from . import capi
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return True
    box = [None]
    ok = True
    try:
        if dist.get_rank() == 0:
            box[0] = handle.steal_export()
    except capi.UnomolError:
        box[0] = b""
    dist.broadcast_object_list(box, src=0)
    if dist.get_rank() != 0:
        if not box[0]:
            ok = False
        else:
            try:
                handle.steal_import(box[0])
            except capi.UnomolError:
                ok = False
    flag = [ok]
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, ok)
    return all(gathered)


class DistributedFock:
    """RHF/UHF Fock build over all ranks: the Python-side equivalent of RHF_MPI::update's Bcast/Reduce bracket."""

    def __init__(self, basis, start_shell=0, tau=None):
        import torch
        from . import capi
        self.rank, self.world, self.local = env_rank()
        torch.cuda.set_device(self.local)
        self.h = capi.Handle(basis, start_shell=start_shell, device=self.local, rank=self.rank, nranks=self.world)
        if tau is not None:
            self.h.set_option("schwarz_tau", tau)
        self.stealing = enable_work_stealing(self.h) if self.world > 1 else False
        if self.world > 1 and not self.stealing:
            self.h.set_option("work_stealing", 0)
        self.no2 = basis.no2
        self.dP = torch.zeros(self.no2, dtype=torch.float64, device="cuda")
        self.dG = torch.zeros(self.no2, dtype=torch.float64, device="cuda")

    def fock_rhf(self, P):
        """P: packed numpy array (same on every rank).  Returns the full G = 2J-K as numpy on every rank."""
        import torch
        self.dP.copy_(torch.from_numpy(np.ascontiguousarray(P)))
        torch.cuda.synchronize()
        self.h.fock_rhf_device(self.dP.data_ptr(), self.dG.data_ptr())
        allreduce_packed(self.dG)
        return self.dG.cpu().numpy()
