"""oracle/oracle.py -- ctypes front-ends for the two CPU checkers.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; the product package unomol_b200 never does.

* ``Oracle``    : liboracle.so, this repo's plain-C restatement (oracle/unomol_oracle.c).
* ``Reference`` : oracle/_ref/libunomol_ref.so, the UNMODIFIED reference compiled from /root/reference by
                  oracle/Makefile plus oracle/ref_harness.cc; present only where it was built
                  (``Reference.available()``).
"""
import ctypes
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_D = ctypes.c_double
_I = ctypes.c_int
_L = ctypes.c_long
_P = ctypes.c_void_p


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(_D))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(_I))


def build(ref=True):
    """compile liboracle.so (and oracle/_ref when /root/reference exists) -- building the checker is not using it"""
    target = "all" if ref else "oracle"
    subprocess.check_call(["make", "-s", "-C", HERE, target])


class OracleBasis:
    def __init__(self, lib, handle):
        self._lib, self.h = lib, handle
        dims = [_I() for _ in range(6)]
        lib.oracle_basis_dims(handle, *[ctypes.byref(d) for d in dims])
        self.nshell, self.nbf, self.ncen, self.maxl, self.nelec, self.nprim = [d.value for d in dims]
        ns = self.nshell
        self.npr = np.zeros(ns, np.int32); self.lv = np.zeros(ns, np.int32); self.cen = np.zeros(ns, np.int32)
        self.off = np.zeros(ns, np.int32); self.poff = np.zeros(ns, np.int32)
        self.alpha = np.zeros(self.nprim); self.coef = np.zeros(self.nprim)
        self.xyz = np.zeros((self.ncen, 3)); self.charge = np.zeros(self.ncen)
        lib.oracle_basis_copy(handle, _ip(self.npr), _ip(self.lv), _ip(self.cen), _ip(self.off), _ip(self.poff),
                              _dp(self.alpha), _dp(self.coef), _dp(self.xyz), _dp(self.charge))
        self.no2 = self.nbf * (self.nbf + 1) // 2

    def __del__(self):
        try:
            self._lib.oracle_basis_free(self.h)
        except Exception:
            pass


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = ctypes.CDLL(path)
        L.oracle_basis_read.restype = _P; L.oracle_basis_read.argtypes = [ctypes.c_char_p]
        L.oracle_basis_free.argtypes = [_P]
        L.oracle_basis_dims.argtypes = [_P] + [ctypes.POINTER(_I)] * 6
        L.oracle_basis_copy.argtypes = [_P] + [ctypes.POINTER(_I)] * 5 + [ctypes.POINTER(_D)] * 4
        L.oracle_rys_roots.argtypes = [_I, _D, ctypes.POINTER(_D), ctypes.POINTER(_D)]
        L.oracle_quartet_block.argtypes = [_P, _I, _I, _I, _I, ctypes.POINTER(_D)]
        L.oracle_unique_eris.restype = _L
        L.oracle_unique_eris.argtypes = [_P, _I, _D, ctypes.POINTER(_D), ctypes.POINTER(_I), _L, ctypes.POINTER(_L)]
        L.oracle_form_g_rhf.argtypes = [_L, ctypes.POINTER(_D), ctypes.POINTER(_I), ctypes.POINTER(_D), ctypes.POINTER(_D)]
        L.oracle_form_g_uhf.argtypes = [_L, ctypes.POINTER(_D), ctypes.POINTER(_I)] + [ctypes.POINTER(_D)] * 4
        L.oracle_direct_g_rhf.restype = _L
        L.oracle_direct_g_rhf.argtypes = [_P, _D, ctypes.POINTER(_D), ctypes.POINTER(_D), _L, _L, ctypes.POINTER(_L)]
        L.oracle_direct_g_uhf.restype = _L
        L.oracle_direct_g_uhf.argtypes = [_P, _D] + [ctypes.POINTER(_D)] * 4 + [_L, _L, ctypes.POINTER(_L)]
        L.oracle_g_elements_rhf.restype = _L
        L.oracle_g_elements_rhf.argtypes = [_P, _D, ctypes.POINTER(_D), _I, ctypes.POINTER(_I), ctypes.POINTER(_D), _I, _I]
        L.oracle_quartet_batch.restype = _L
        L.oracle_quartet_batch.argtypes = [_P, _L, ctypes.POINTER(_I), ctypes.POINTER(_D), ctypes.POINTER(_D), _I]
        L.oracle_direct_g_rhf_start.restype = _L
        L.oracle_direct_g_rhf_start.argtypes = [_P, _I, _D, ctypes.POINTER(_D), ctypes.POINTER(_D), _L, _L]
        L.oracle_cart_norm.restype = _D; L.oracle_cart_norm.argtypes = [_I, _I]
        L.oracle_basis_set_center.argtypes = [_P, _I, _D, _D, _D]

    def basis(self, patin_path):
        h = self.lib.oracle_basis_read(os.fsencode(patin_path))
        if not h:
            raise IOError("oracle could not read %s" % patin_path)
        return OracleBasis(self.lib, h)

    def rys_roots(self, n, x):
        r = np.zeros(5); w = np.zeros(5)
        if self.lib.oracle_rys_roots(n, float(x), _dp(r), _dp(w)) != 0:
            raise ValueError("nroots %d not supported" % n)
        return r[:n].copy(), w[:n].copy()

    def quartet_block(self, b, i, j, k, l):
        nc = lambda s: (b.lv[s] + 1) * (b.lv[s] + 2) // 2
        shape = (nc(i), nc(j), nc(k), nc(l))
        out = np.zeros(shape)
        n = self.lib.oracle_quartet_block(b.h, i, j, k, l, _dp(out))
        if n < 0:
            raise ValueError("l_tot > 8")
        return out

    def unique_eris(self, b, thresh=1e-14, start_shell=0):
        ncalc = _L(0)
        n = self.lib.oracle_unique_eris(b.h, start_shell, thresh, None, None, 0, ctypes.byref(ncalc))
        vals = np.zeros(n); ijkl = np.zeros((n, 4), np.int32)
        self.lib.oracle_unique_eris(b.h, start_shell, thresh, _dp(vals), _ip(ijkl), n, ctypes.byref(ncalc))
        return vals, ijkl, ncalc.value

    def form_g_rhf(self, vals, ijkl, P, G=None):
        G = np.zeros_like(P) if G is None else G
        self.lib.oracle_form_g_rhf(len(vals), _dp(vals), _ip(ijkl), _dp(np.ascontiguousarray(P)), _dp(G))
        return G

    def form_g_uhf(self, vals, ijkl, PA, PB):
        GA = np.zeros_like(PA); GB = np.zeros_like(PB)
        self.lib.oracle_form_g_uhf(len(vals), _dp(vals), _ip(ijkl), _dp(np.ascontiguousarray(PA)),
                                   _dp(np.ascontiguousarray(PB)), _dp(GA), _dp(GB))
        return GA, GB

    def direct_g_rhf(self, b, P, thresh=1e-14, sample_mod=1, sample_rem=0):
        G = np.zeros(b.no2); npq = _L(0)
        nq = self.lib.oracle_direct_g_rhf(b.h, thresh, _dp(np.ascontiguousarray(P)), _dp(G), sample_mod, sample_rem,
                                          ctypes.byref(npq))
        return G, nq, npq.value


    def set_center(self, b, icen, xyz):
        self.lib.oracle_basis_set_center(b.h, int(icen), float(xyz[0]), float(xyz[1]), float(xyz[2]))
        b.xyz[icen] = xyz

    def direct_g_start_threads(self, b, P, start_shell, thresh=1e-14, nthreads=None):
        """RHF G of the quartets with max shell index >= start_shell (a TwoElectronInts with start_shell > 0), all host cores"""
        import threading
        nthreads = nthreads or max(1, min(32, os.cpu_count() or 1))
        P = np.ascontiguousarray(P, float)
        parts = [np.zeros(b.no2) for _ in range(nthreads)]

        def run(t):
            self.lib.oracle_direct_g_rhf_start(b.h, int(start_shell), thresh, _dp(P), _dp(parts[t]), nthreads, t)
        th = [threading.Thread(target=run, args=(t,)) for t in range(nthreads)]
        for x in th: x.start()
        for x in th: x.join()
        return sum(parts)

    def direct_g_threads(self, b, P, PB=None, thresh=1e-14, nthreads=None):
        """direct G on all host cores: thread t digests the shell quartets with running index % nthreads == t into its own
        G (ctypes releases the GIL); RHF when PB is None, else (GA, GB).  Same arithmetic as direct_g_rhf, summed per thread."""
        import threading
        nthreads = nthreads or max(1, min(32, os.cpu_count() or 1))
        P = np.ascontiguousarray(P, float)
        PBc = None if PB is None else np.ascontiguousarray(PB, float)
        parts = [[np.zeros(b.no2), np.zeros(b.no2), 0] for _ in range(nthreads)]

        def run(t):
            npq = _L(0)
            if PBc is None:
                parts[t][2] = self.lib.oracle_direct_g_rhf(b.h, thresh, _dp(P), _dp(parts[t][0]), nthreads, t, ctypes.byref(npq))
            else:
                parts[t][2] = self.lib.oracle_direct_g_uhf(b.h, thresh, _dp(P), _dp(PBc), _dp(parts[t][0]), _dp(parts[t][1]),
                                                           nthreads, t, ctypes.byref(npq))
        th = [threading.Thread(target=run, args=(t,)) for t in range(nthreads)]
        for x in th: x.start()
        for x in th: x.join()
        GA = sum(p[0] for p in parts)
        if PBc is None:
            return GA
        return GA, sum(p[1] for p in parts)


    def g_elements(self, b, P, pairs, thresh=1e-14, nthreads=None):
        """G_ij = sum_kl P_kl [2(ij|kl) - (ik|jl)] for a few basis-function pairs of a system too large for a full G
        (bench.py parity block): all host cores, unscreened, with the reference's primitive cut and storage threshold."""
        import threading
        nthreads = nthreads or max(1, min(64, os.cpu_count() or 1))
        P = np.ascontiguousarray(P, float)
        ij = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))
        outs = [np.zeros(len(ij)) for _ in range(nthreads)]
        nblk = [0] * nthreads

        def run(t):
            nblk[t] = self.lib.oracle_g_elements_rhf(b.h, thresh, _dp(P), len(ij), _ip(ij), _dp(outs[t]), nthreads, t)
        th = [threading.Thread(target=run, args=(t,)) for t in range(nthreads)]
        for x in th: x.start()
        for x in th: x.join()
        return sum(outs), sum(nblk)


def timed_quartet_batches(fn, handle, shells, P, no2, digest, nthreads):
    """bench helper: split `shells` ([n,4] int32) over host threads, each calling the C batch routine `fn` on its share with
    its own G; returns (seconds of the slowest thread's wall clock for the whole batch, stored integrals, summed G)"""
    import threading
    import time
    shells = np.ascontiguousarray(shells, np.int32)
    P = np.ascontiguousarray(P, float)
    parts = [np.ascontiguousarray(shells[t::nthreads]) for t in range(nthreads)]
    Gs = [np.zeros(no2) for _ in range(nthreads)]
    stored = [0] * nthreads

    def run(t):
        stored[t] = fn(handle, len(parts[t]), _ip(parts[t]), _dp(P), _dp(Gs[t]), int(digest))
    th = [threading.Thread(target=run, args=(t,)) for t in range(nthreads)]
    t0 = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    return time.perf_counter() - t0, sum(stored), sum(Gs)


class Reference:
    """The unmodified reference behind oracle/ref_harness.cc."""
    PATH = os.path.join(HERE, "_ref", "libunomol_ref.so")
    UNOMOL = os.path.join(HERE, "_ref", "Unomol")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        L = self.lib = ctypes.CDLL(self.PATH)
        L.ref_rys_roots.argtypes = [_I, _D, ctypes.POINTER(_D), ctypes.POINTER(_D)]
        L.ref_basis_open.restype = _P; L.ref_basis_open.argtypes = [ctypes.c_char_p]
        L.ref_basis_close.argtypes = [_P]
        for f in ("nshell", "norb", "ncen", "nelec"):
            getattr(L, "ref_basis_" + f).argtypes = [_P]
        L.ref_basis_eps.restype = _D; L.ref_basis_eps.argtypes = [_P]
        L.ref_basis_shell.argtypes = [_P, _I] + [ctypes.POINTER(_I)] * 4 + [ctypes.POINTER(_D)] * 2
        L.ref_basis_center.argtypes = [_P, _I, ctypes.POINTER(_D), ctypes.POINTER(_D)]
        L.ref_quartet_block.argtypes = [_P, _I, _I, _I, _I, ctypes.POINTER(_D)]
        L.ref_tints_create.restype = _P; L.ref_tints_create.argtypes = [_P, _I, ctypes.c_char_p]
        L.ref_tints_destroy.argtypes = [_P]
        L.ref_tints_eri_seconds.restype = _D; L.ref_tints_eri_seconds.argtypes = [_P]
        L.ref_tints_count.restype = _L; L.ref_tints_count.argtypes = [_P]
        L.ref_tints_dump.restype = _L; L.ref_tints_dump.argtypes = [_P, ctypes.POINTER(_D), ctypes.POINTER(_I), _L]
        L.ref_tints_form_g_rhf.restype = _D; L.ref_tints_form_g_rhf.argtypes = [_P] + [ctypes.POINTER(_D)] * 2
        L.ref_tints_form_g_uhf.restype = _D; L.ref_tints_form_g_uhf.argtypes = [_P] + [ctypes.POINTER(_D)] * 4
        L.ref_one_electron.argtypes = [_P] + [ctypes.POINTER(_D)] * 3
        L.ref_quartet_batch.restype = _L
        L.ref_quartet_batch.argtypes = [_P, _L, ctypes.POINTER(_I), ctypes.POINTER(_D), ctypes.POINTER(_D), _I]

    def rys_roots(self, n, x):
        r = np.zeros(5); w = np.zeros(5)
        self.lib.ref_rys_roots(n, float(x), _dp(r), _dp(w))
        return r[:n].copy(), w[:n].copy()

    def basis(self, patin_path):
        return self.lib.ref_basis_open(os.fsencode(patin_path))

    def basis_close(self, h):
        self.lib.ref_basis_close(h)

    def basis_shells(self, h):
        out = []
        for s in range(self.lib.ref_basis_nshell(h)):
            npr, lv, cen, off = _I(), _I(), _I(), _I()
            al = np.zeros(64); co = np.zeros(64)
            self.lib.ref_basis_shell(h, s, ctypes.byref(npr), ctypes.byref(lv), ctypes.byref(cen), ctypes.byref(off),
                                     _dp(al), _dp(co))
            out.append((npr.value, lv.value, cen.value, off.value, al[:npr.value].copy(), co[:npr.value].copy()))
        return out

    def quartet_block(self, h, shape, i, j, k, l):
        out = np.zeros(shape)
        n = self.lib.ref_quartet_block(h, i, j, k, l, _dp(out))
        if n < 0:
            raise ValueError("l_tot > 8")
        return out

    def tints(self, h, start_shell=0, cache_name="/tmp/UNOMOL_REF_MINTS.DAT"):
        return self.lib.ref_tints_create(h, start_shell, os.fsencode(cache_name))

    def tints_destroy(self, t):
        self.lib.ref_tints_destroy(t)

    def tints_eri_seconds(self, t):
        return self.lib.ref_tints_eri_seconds(t)

    def tints_dump(self, t):
        n = self.lib.ref_tints_count(t)
        vals = np.zeros(n); ijkl = np.zeros((n, 4), np.int32)
        self.lib.ref_tints_dump(t, _dp(vals), _ip(ijkl), n)
        return vals, ijkl

    def form_g_rhf(self, t, P):
        G = np.zeros_like(P)
        sec = self.lib.ref_tints_form_g_rhf(t, _dp(np.ascontiguousarray(P)), _dp(G))
        return G, sec

    def form_g_uhf(self, t, PA, PB):
        GA = np.zeros_like(PA); GB = np.zeros_like(PB)
        sec = self.lib.ref_tints_form_g_uhf(t, _dp(np.ascontiguousarray(PA)), _dp(np.ascontiguousarray(PB)), _dp(GA), _dp(GB))
        return GA, GB, sec

    def one_electron(self, h):
        n = self.lib.ref_basis_norb(h); no2 = n * (n + 1) // 2
        S = np.zeros(no2); T = np.zeros(no2); H = np.zeros(no2)
        self.lib.ref_one_electron(h, _dp(S), _dp(T), _dp(H))
        return S, T, H


def run_reference_mpi(patin_path, nranks, timeout_s=600):
    """Run the reference's UNMODIFIED MPI driver (oracle/_ref/UnomolMPI: UnomolMPI.cc + TwoElectronIntsMPI.cpp + RHF_MPI.hpp /
    UHF_MPI.hpp compiled against oracle/mpi_shim/mpi.h) on `nranks` forked ranks of this host, in a scratch directory.
    Returns what the program itself reports: its integral pass ("Time for Two Electrons Integrals", slowest rank; reference
    TwoElectronIntsMPI.cpp:440-442), its SCF loop ("SCF time", RHF_MPI.hpp:229), iterations, final energy (short.gs.out).
    Test / bench infrastructure only."""
    import re
    import shutil
    import subprocess
    import tempfile
    import time
    exe = os.path.join(HERE, "_ref", "UnomolMPI")
    if not os.path.exists(exe):
        return None
    d = tempfile.mkdtemp(prefix="unomol_mpi_")
    try:
        lines = open(patin_path).read().split("\n")
        lines[3] = " 0 0"                      # no finite-field analysis, no polarisation scan: the SCF alone
        open(os.path.join(d, "patin.dat"), "w").write("\n".join(lines))
        env = dict(os.environ, UNOMOL_MPI_SHIM_NP=str(nranks))
        t0 = time.perf_counter()
        r = subprocess.run([exe], cwd=d, env=env, capture_output=True, text=True, timeout=timeout_s)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": "exit status %d: %s" % (r.returncode, r.stderr[-300:])}
        logs = [r.stderr] + [open(os.path.join(d, f)).read() for f in sorted(os.listdir(d)) if f.startswith("mpi_rank") and f.endswith(".err")]
        eri = [float(x) for log in logs for x in re.findall(r"Time for Two Electrons Integrals = ([0-9.eE+-]+)", log)]
        scf = [float(x) for log in logs for x in re.findall(r"SCF time = ([0-9.eE+-]+)", log)]
        head = r.stderr.split("SCF time")[0]
        nums = [int(x) for x in re.findall(r"Iteration\s+=\s+(\d+)", head)]
        its = nums[-1] if nums else 0
        short = open(os.path.join(d, "short.gs.out")).read().split()
        return {"ranks": nranks, "eri_pass_s_slowest_rank": max(eri) if eri else None, "eri_pass_s_fastest_rank": min(eri) if eri else None,
                "scf_s": max(scf) if scf else None, "iterations": its, "scf_s_per_iteration": (max(scf) / its) if scf and its else None,
                "wall_s": wall, "energy": float(short[1]) if len(short) > 1 else None}
    finally:
        shutil.rmtree(d, ignore_errors=True)
