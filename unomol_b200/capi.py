"""unomol_b200/capi.py -- ctypes binding of libunomol_b200.so (include/unomol_b200.h).

This is the Python view of the C ABI that a maintainer of the reference would bind from C++ (see
INTEGRATION.md); the tests and bench.py call the CUDA path only through it.  There is no CPU fallback:
a missing library raises ImportError at import time, and every compute call raises UnomolError when the
library reports a failure (e.g. no CUDA device).
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# UNOMOL_B200_LIB: an alternative build of the same library (A/B kernel experiments, profiles/); never a fallback
LIB_PATH = os.environ.get("UNOMOL_B200_LIB") or os.path.join(_HERE, "libunomol_b200.so")

_D = ctypes.c_double
_I = ctypes.c_int
_P = ctypes.c_void_p
_pd = ctypes.POINTER(_D)
_pi = ctypes.POINTER(_I)


class UnomolError(RuntimeError):
    pass


class BasisDesc(ctypes.Structure):
    _fields_ = [("nshell", _I), ("nbf", _I), ("ncen", _I), ("maxl", _I),
                ("npr", _pi), ("lv", _pi), ("cen", _pi), ("off", _pi), ("poff", _pi),
                ("alpha", _pd), ("coef", _pd), ("xyz", _pd)]


class TwoInt(ctypes.Structure):
    _fields_ = [("val", _D), ("i", _I), ("j", _I), ("k", _I), ("l", _I)]


class Stats(ctypes.Structure):
    _fields_ = [("n_shell_pairs", ctypes.c_longlong), ("n_pairs_kept", ctypes.c_longlong),
                ("n_prim_pairs", ctypes.c_longlong), ("n_quartets", ctypes.c_longlong),
                ("n_quartets_total", ctypes.c_longlong), ("model_flops", _D), ("last_fock_ms", _D),
                ("last_eri_kernel_ms", _D), ("precompute_ms", _D), ("n_launches", _I), ("nbf", _I),
                ("nshell", _I), ("rank", _I), ("nranks", _I),
                ("n_prim_quartets", ctypes.c_longlong), ("n_prim_candidates", ctypes.c_longlong),
                ("n_tile_launches", _I), ("n_reg_launches", _I), ("n_rows_launches", _I), ("n_generic_launches", _I),
                ("n_highl_launches", _I), ("last_dump_kernel", _I), ("n_incremental_updates", _I), ("onee_ms", _D)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = [
    "unomol_b200_create", "unomol_b200_destroy", "unomol_b200_set_option", "unomol_b200_set_geometry",
    "unomol_b200_fock_rhf", "unomol_b200_fock_uhf", "unomol_b200_fock_rhf_device", "unomol_b200_fock_uhf_device",
    "unomol_b200_eri_quartet", "unomol_b200_dump_eris", "unomol_b200_schwarz", "unomol_b200_stats",
    "unomol_b200_attach_nccl", "unomol_b200_steal_export", "unomol_b200_steal_import", "unomol_b200_steal_share", "unomol_b200_device_buffers", "unomol_b200_scf_set_overlap", "unomol_b200_scf_diag",
    "unomol_b200_scf_load", "unomol_b200_scf_iterate_rhf", "unomol_b200_scf_iterate_rhf_begin", "unomol_b200_scf_iterate_rhf_finish", "unomol_b200_scf_fetch",
    "unomol_b200_one_electron", "unomol_b200_one_electron_dpm", "unomol_b200_scf_load_uhf", "unomol_b200_scf_iterate_uhf", "unomol_b200_scf_fetch_uhf",
    "unomol_b200_sample_quartets", "unomol_b200_fp64_peak", "unomol_b200_model_flops", "unomol_b200_strerror", "unomol_b200_version",
]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("libunomol_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C unomol_b200/csrc`); unomol_b200 has no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    L.unomol_b200_create.argtypes = [ctypes.POINTER(BasisDesc), _I, _I, _I, _I, ctypes.POINTER(_P)]
    L.unomol_b200_destroy.argtypes = [_P]; L.unomol_b200_destroy.restype = None
    L.unomol_b200_set_option.argtypes = [_P, ctypes.c_char_p, _D]
    L.unomol_b200_set_geometry.argtypes = [_P, _pd]
    L.unomol_b200_fock_rhf.argtypes = [_P, _pd, _pd]
    L.unomol_b200_fock_uhf.argtypes = [_P, _pd, _pd, _pd, _pd]
    L.unomol_b200_fock_rhf_device.argtypes = [_P, _P, _P, _I]
    L.unomol_b200_fock_uhf_device.argtypes = [_P, _P, _P, _P, _P, _I]
    L.unomol_b200_eri_quartet.argtypes = [_P, _I, _I, _I, _I, _pd]
    L.unomol_b200_dump_eris.argtypes = [_P, _D, ctypes.POINTER(TwoInt), ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    L.unomol_b200_schwarz.argtypes = [_P, _pd]
    L.unomol_b200_stats.argtypes = [_P, ctypes.POINTER(Stats)]
    L.unomol_b200_attach_nccl.argtypes = [_P, _P]
    L.unomol_b200_steal_export.argtypes = [_P, ctypes.c_char_p]
    L.unomol_b200_steal_import.argtypes = [_P, ctypes.c_char_p]
    L.unomol_b200_steal_share.argtypes = [_P, _P]
    L.unomol_b200_device_buffers.argtypes = [_P, ctypes.POINTER(_P), ctypes.POINTER(_P), ctypes.POINTER(_P)]
    L.unomol_b200_scf_set_overlap.argtypes = [_P, _pd]
    L.unomol_b200_scf_diag.argtypes = [_P, _pd, _I, _pd, _pd, _pd]
    L.unomol_b200_scf_load.argtypes = [_P, _pd, _pd]
    L.unomol_b200_scf_iterate_rhf.argtypes = [_P, _I, _I, _pd, _pd]
    L.unomol_b200_scf_iterate_rhf_begin.argtypes = [_P, _I]
    L.unomol_b200_scf_iterate_rhf_finish.argtypes = [_P, _I, _pd, _pd]
    L.unomol_b200_scf_fetch.argtypes = [_P, _pd, _pd, _pd]
    L.unomol_b200_one_electron.argtypes = [_P, _pd, _pd, _pd, _pd, _pd]
    L.unomol_b200_one_electron_dpm.argtypes = [_P, _pd, _I, _pd, _pd, _pd, _pd]
    L.unomol_b200_scf_load_uhf.argtypes = [_P, _pd, _pd, _pd]
    L.unomol_b200_scf_iterate_uhf.argtypes = [_P, _I, _I, _I, _pd, _pd]
    L.unomol_b200_scf_fetch_uhf.argtypes = [_P, _pd, _pd, _pd, _pd]
    L.unomol_b200_sample_quartets.argtypes = [_P, ctypes.c_longlong, ctypes.c_ulonglong, _pi, ctypes.POINTER(ctypes.c_longlong)]
    L.unomol_b200_fp64_peak.argtypes = [_I, _pd]
    L.unomol_b200_model_flops.restype = _D; L.unomol_b200_model_flops.argtypes = [_I, _I, _I, _I]
    L.unomol_b200_strerror.restype = ctypes.c_char_p; L.unomol_b200_strerror.argtypes = [_I]
    L.unomol_b200_version.restype = ctypes.c_char_p
    return L


lib = _load()


def _chk(rc, what):
    if rc != 0:
        raise UnomolError("%s failed: %s (%d)" % (what, lib.unomol_b200_strerror(rc).decode(), rc))


def _dp(a):
    return a.ctypes.data_as(_pd)


def _ip(a):
    return a.ctypes.data_as(_pi)


class Handle:
    """One libunomol_b200 handle = one GPU's share of the two-electron engine."""

    def __init__(self, basis, start_shell=0, device=0, rank=0, nranks=1):
        self.basis = basis
        # keep the arrays alive for the duration of the create call
        self._arr = dict(npr=np.ascontiguousarray(basis.npr, np.int32), lv=np.ascontiguousarray(basis.lv, np.int32),
                         cen=np.ascontiguousarray(basis.cen, np.int32), off=np.ascontiguousarray(basis.off, np.int32),
                         poff=np.ascontiguousarray(basis.poff, np.int32), alpha=np.ascontiguousarray(basis.alpha, float),
                         coef=np.ascontiguousarray(basis.coef, float), xyz=np.ascontiguousarray(basis.xyz, float))
        a = self._arr
        d = BasisDesc(basis.nshell, basis.nbf, basis.ncen, basis.maxl, _ip(a["npr"]), _ip(a["lv"]), _ip(a["cen"]),
                      _ip(a["off"]), _ip(a["poff"]), _dp(a["alpha"]), _dp(a["coef"]), _dp(a["xyz"]))
        self.h = _P()
        _chk(lib.unomol_b200_create(ctypes.byref(d), start_shell, device, rank, nranks, ctypes.byref(self.h)), "create")
        self.nbf = basis.nbf
        self.no2 = basis.nbf * (basis.nbf + 1) // 2

    def close(self):
        if getattr(self, "h", None):
            lib.unomol_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        _chk(lib.unomol_b200_set_option(self.h, name.encode(), float(value)), "set_option(%s)" % name)

    def set_geometry(self, xyz):
        xyz = np.ascontiguousarray(xyz, float)
        _chk(lib.unomol_b200_set_geometry(self.h, _dp(xyz)), "set_geometry")

    def fock_rhf(self, P, G=None):
        P = np.ascontiguousarray(P, float)
        G = np.zeros(self.no2) if G is None else G
        _chk(lib.unomol_b200_fock_rhf(self.h, _dp(P), _dp(G)), "fock_rhf")
        return G

    def fock_uhf(self, PA, PB, GA=None, GB=None):
        PA = np.ascontiguousarray(PA, float); PB = np.ascontiguousarray(PB, float)
        GA = np.zeros(self.no2) if GA is None else GA
        GB = np.zeros(self.no2) if GB is None else GB
        _chk(lib.unomol_b200_fock_uhf(self.h, _dp(PA), _dp(PB), _dp(GA), _dp(GB)), "fock_uhf")
        return GA, GB

    def fock_rhf_device(self, dP_ptr, dG_ptr, async_=False):
        _chk(lib.unomol_b200_fock_rhf_device(self.h, _P(dP_ptr), _P(dG_ptr), int(async_)), "fock_rhf_device")

    def fock_uhf_device(self, dPA, dPB, dGA, dGB, async_=False):
        _chk(lib.unomol_b200_fock_uhf_device(self.h, _P(dPA), _P(dPB), _P(dGA), _P(dGB), int(async_)), "fock_uhf_device")

    def eri_quartet(self, i, j, k, l):
        nc = lambda s: (int(self.basis.lv[s]) + 1) * (int(self.basis.lv[s]) + 2) // 2
        out = np.zeros((nc(i), nc(j), nc(k), nc(l)))
        _chk(lib.unomol_b200_eri_quartet(self.h, i, j, k, l, _dp(out)), "eri_quartet")
        return out

    def dump_eris(self, thresh=1e-14):
        n = ctypes.c_size_t(0)
        _chk(lib.unomol_b200_dump_eris(self.h, thresh, None, 0, ctypes.byref(n)), "dump_eris(count)")
        buf = (TwoInt * max(n.value, 1))()
        _chk(lib.unomol_b200_dump_eris(self.h, thresh, buf, n.value, ctypes.byref(n)), "dump_eris")
        rec = np.frombuffer(buf, dtype=np.dtype([("val", "f8"), ("i", "i4"), ("j", "i4"), ("k", "i4"), ("l", "i4")]),
                            count=n.value)
        vals = rec["val"].copy()
        ijkl = np.stack([rec["i"], rec["j"], rec["k"], rec["l"]], axis=1).astype(np.int32)
        return vals, ijkl

    def schwarz(self):
        ns = self.basis.nshell
        Q = np.zeros(ns * (ns + 1) // 2)
        _chk(lib.unomol_b200_schwarz(self.h, _dp(Q)), "schwarz")
        return Q

    def stats(self):
        s = Stats()
        _chk(lib.unomol_b200_stats(self.h, ctypes.byref(s)), "stats")
        return s.as_dict()

    def attach_nccl(self, comm_ptr):
        _chk(lib.unomol_b200_attach_nccl(self.h, _P(comm_ptr)), "attach_nccl")

    def steal_export(self):
        buf = ctypes.create_string_buffer(64)
        _chk(lib.unomol_b200_steal_export(self.h, buf), "steal_export")
        return buf.raw

    def steal_share(self, peer):
        """in-process variant: `peer` (another Handle of this process) claims work from this handle's counters"""
        _chk(lib.unomol_b200_steal_share(self.h, peer.h), "steal_share")

    def steal_import(self, handle64):
        _chk(lib.unomol_b200_steal_import(self.h, ctypes.create_string_buffer(handle64, 64)), "steal_import")

    def device_buffers(self):
        st = _P(); dP = (_P * 2)(); dG = (_P * 2)()
        _chk(lib.unomol_b200_device_buffers(self.h, ctypes.byref(st), dP, dG), "device_buffers")
        return st.value, [dP[0], dP[1]], [dG[0], dG[1]]

    def sample_quartets(self, nsample, seed=1):
        sh = np.zeros((max(nsample, 1), 4), np.int32); tot = ctypes.c_longlong(0)
        _chk(lib.unomol_b200_sample_quartets(self.h, nsample, seed, _ip(sh), ctypes.byref(tot)), "sample_quartets")
        return sh[:nsample], tot.value

    def scf_set_overlap(self, S):
        S = np.ascontiguousarray(S, float)
        _chk(lib.unomol_b200_scf_set_overlap(self.h, _dp(S)), "scf_set_overlap")

    def scf_load(self, H, P):
        _chk(lib.unomol_b200_scf_load(self.h, _dp(np.ascontiguousarray(H, float)), _dp(np.ascontiguousarray(P, float))), "scf_load")

    def scf_iterate_rhf(self, nocc, damp=False):
        e = _D(0.0); pd = _D(0.0)
        _chk(lib.unomol_b200_scf_iterate_rhf(self.h, nocc, int(damp), ctypes.byref(e), ctypes.byref(pd)), "scf_iterate_rhf")
        return e.value, pd.value

    def scf_iterate_rhf_begin(self, damp=False):
        _chk(lib.unomol_b200_scf_iterate_rhf_begin(self.h, int(damp)), "scf_iterate_rhf_begin")

    def scf_iterate_rhf_finish(self, nocc):
        e = _D(0.0); pd = _D(0.0)
        _chk(lib.unomol_b200_scf_iterate_rhf_finish(self.h, nocc, ctypes.byref(e), ctypes.byref(pd)), "scf_iterate_rhf_finish")
        return e.value, pd.value

    def scf_fetch(self, want_c=False):
        P = np.zeros(self.no2); ev = np.zeros(self.nbf)
        C = np.zeros((self.nbf, self.nbf)) if want_c else None
        _chk(lib.unomol_b200_scf_fetch(self.h, _dp(P), _dp(ev), _dp(C) if want_c else None), "scf_fetch")
        return (P, ev, C) if want_c else (P, ev)

    def one_electron(self, charge, moments=False, dpm_center=-1):
        """packed S, T, H (and the 9 moment matrices) from the device kernel (csrc/onee_device.cu); dpm_center >= 0 folds the
        positron charge model of the polarisation scan into H"""
        charge = np.ascontiguousarray(charge, float)
        S = np.zeros(self.no2); T = np.zeros(self.no2); H = np.zeros(self.no2)
        M = np.zeros((9, self.no2)) if moments else None
        _chk(lib.unomol_b200_one_electron_dpm(self.h, _dp(charge), int(dpm_center), _dp(S), _dp(T), _dp(H), _dp(M) if moments else None),
             "one_electron")
        return (S, T, H, M) if moments else (S, T, H)

    def scf_load_uhf(self, H, PA, PB):
        _chk(lib.unomol_b200_scf_load_uhf(self.h, _dp(np.ascontiguousarray(H, float)), _dp(np.ascontiguousarray(PA, float)),
                                          _dp(np.ascontiguousarray(PB, float))), "scf_load_uhf")

    def scf_iterate_uhf(self, nocc_a, nocc_b, damp=False):
        e = _D(0.0); pd = _D(0.0)
        _chk(lib.unomol_b200_scf_iterate_uhf(self.h, nocc_a, nocc_b, int(damp), ctypes.byref(e), ctypes.byref(pd)), "scf_iterate_uhf")
        return e.value, pd.value

    def scf_fetch_uhf(self):
        PA = np.zeros(self.no2); PB = np.zeros(self.no2); ea = np.zeros(self.nbf); eb = np.zeros(self.nbf)
        _chk(lib.unomol_b200_scf_fetch_uhf(self.h, _dp(PA), _dp(PB), _dp(ea), _dp(eb)), "scf_fetch_uhf")
        return PA, PB, ea, eb

    def scf_diag(self, F, nocc, want_c=False):
        F = np.ascontiguousarray(F, float)
        ev = np.zeros(self.nbf); P = np.zeros(self.no2)
        C = np.zeros((self.nbf, self.nbf)) if want_c else None
        _chk(lib.unomol_b200_scf_diag(self.h, _dp(F), nocc, _dp(ev), _dp(C) if want_c else None, _dp(P)), "scf_diag")
        return (ev, C, P) if want_c else (ev, P)


def fp64_peak(device=0):
    """measured DFMA throughput in TFLOP/s (the FP64 roofline denominator; not in MEASURED_PEAKS.json)"""
    t = _D(0.0)
    _chk(lib.unomol_b200_fp64_peak(device, ctypes.byref(t)), "fp64_peak")
    return t.value
