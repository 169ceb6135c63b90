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


class NcclComm:
    """An NCCL communicator of our own over all ranks (ctypes on the libnccl that torch ships), for
    unomol_b200_attach_nccl: the library then all-reduces the packed G itself, on its own stream, at the end of every
    Fock build (the replacement of MPI_Reduce in RHF_MPI::update, reference RHF_MPI.hpp:108).  torch.distributed only
    carries the 128-byte unique id to the other ranks."""

    def __init__(self):
        import ctypes
        import glob
        import torch.distributed as dist
        cands = []
        try:
            import nvidia.nccl
            for base in list(getattr(nvidia.nccl, "__path__", [])):
                cands += glob.glob(os.path.join(base, "lib", "libnccl.so*"))
        except Exception:
            pass
        cands += ["libnccl.so.2", "libnccl.so"]
        self.lib = None
        for c in cands:
            try:
                self.lib = ctypes.CDLL(c, mode=ctypes.RTLD_GLOBAL)   # global: the engine finds ncclAllReduce with dlsym
                break
            except OSError:
                continue
        if self.lib is None:
            raise RuntimeError("libnccl not found")

        class UniqueId(ctypes.Structure):
            _fields_ = [("internal", ctypes.c_char * 128)]
        self.lib.ncclGetUniqueId.argtypes = [ctypes.POINTER(UniqueId)]
        self.lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
        self.lib.ncclCommDestroy.argtypes = [ctypes.c_void_p]
        uid = UniqueId()
        box = [None]
        if dist.get_rank() == 0:
            if self.lib.ncclGetUniqueId(ctypes.byref(uid)) != 0:
                raise RuntimeError("ncclGetUniqueId failed")
            box[0] = ctypes.string_at(ctypes.byref(uid), 128)      # all 128 bytes (the id contains NULs)
        dist.broadcast_object_list(box, src=0)
        ctypes.memmove(ctypes.byref(uid), box[0], 128)
        self.comm = ctypes.c_void_p()
        rc = self.lib.ncclCommInitRank(ctypes.byref(self.comm), dist.get_world_size(), uid, dist.get_rank())
        if rc != 0:
            raise RuntimeError("ncclCommInitRank failed (%d)" % rc)

    @property
    def ptr(self):
        return self.comm.value

    def close(self):
        if getattr(self, "comm", None) is not None and self.comm.value:
            self.lib.ncclCommDestroy(self.comm)
            self.comm = None


class DistributedFock:
    """RHF/UHF Fock build over all ranks: the Python-side equivalent of RHF_MPI::update's Bcast/Reduce bracket."""

    def __init__(self, basis, start_shell=0, tau=None, in_library_allreduce=False):
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
        # in_library_allreduce: hand the library its own NCCL communicator (unomol_b200_attach_nccl); fock_*_device then
        # returns the summed G and no torch collective is issued
        self.nccl = None
        if in_library_allreduce and self.world > 1:
            self.nccl = NcclComm()
            self.h.attach_nccl(self.nccl.ptr)
        self.no2 = basis.no2
        self.dP = torch.zeros(self.no2, dtype=torch.float64, device="cuda")
        self.dG = torch.zeros(self.no2, dtype=torch.float64, device="cuda")

    def fock_rhf(self, P):
        """P: packed numpy array (same on every rank).  Returns the full G = 2J-K as numpy on every rank."""
        import torch
        self.dP.copy_(torch.from_numpy(np.ascontiguousarray(P)))
        torch.cuda.synchronize()
        self.h.fock_rhf_device(self.dP.data_ptr(), self.dG.data_ptr())
        if self.nccl is None:
            allreduce_packed(self.dG)
        return self.dG.cpu().numpy()

    def close(self):
        self.h.attach_nccl(0)
        self.h.close()
        if self.nccl is not None:
            self.nccl.close()
