#!/usr/bin/env python3
"""bench.py -- Fock-build benchmark of the B200-native two-electron engine (and of the reference on host cores).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload water154|water308|sf6|co2|...] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one integral-direct Fock build G = 2J[P]-K[P] (RHF) over one synthetic density P: every screened
contracted shell quartet of the workload is evaluated (Rys quadrature) and digested, nothing is stored.
metric  = contracted ERI shell quartets per second, whole job (BASELINE.json: "ERI quartets/s" + "s per Fock build",
          the latter is ms_per_step).
value   = device-resident (P, pair tables already in HBM), CUDA events on the library stream, max over ranks.
e2e     = the same metric through the reference-facing C-ABI call unomol_b200_fock_rhf() with HOST buffers
          (pinned), host->device copy of P and device->host copy of G inside the timed region.
roofline= FP64: algorithmic FLOPs of SURVEY.md 8(d) (reference Rys algorithm, per surviving primitive quartet)
          divided by the summed duration of the fused ERI+digestion class kernels, against the DFMA peak
          measured in this run (MEASURED_PEAKS.json carries no FP64 number).
cpu_baseline / --impl reference = the UNMODIFIED reference's calc_two_electron_ints_rys (oracle/_ref, built
          from /root/reference) timed on the host cores over a bounded, seeded sample of screened quartets.
reference_mpi = the reference's own MPI work distribution (UnomolMPI.cc / TwoElectronIntsMPI.cpp / RHF_MPI.hpp, unmodified,
          compiled against oracle/mpi_shim/mpi.h because the image has no MPI) run to convergence on SF6/TZ2P with one
          rank per host core: its integral pass and its seconds per SCF iteration, as the program reports them.  The
          >= 2000-function cluster cannot run there (its cache would hold ~10^12 integrals), see DESIGN.md.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "water154": ("synthetic (H2O)_154 / 6-31G, 2002 basis functions, 1386 shells", lambda B: B.water_cluster(154)),
    "water308": ("synthetic (H2O)_308 / 6-31G, 4004 basis functions, 2772 shells", lambda B: B.water_cluster(308)),
    "water32": ("synthetic (H2O)_32 / 6-31G, 416 basis functions", lambda B: B.water_cluster(32)),
    "sf6": ("SF6 / TZ2P (test/patin.dat.tz2p.sf6), 190 basis functions", lambda B: B.Basis.from_patin(B.test_input("tz2p.sf6"))),
    "co2": ("CO2 / DZP (test/patin.dat.dh95.co2), 90 basis functions", lambda B: B.Basis.from_patin(B.test_input("dh95.co2"))),
    "c2h2": ("C2H2 / DZP (test/patin.dat.dh95.c2h2), 90 basis functions", lambda B: B.Basis.from_patin(B.test_input("dh95.c2h2"))),
    "nh3": ("NH3 / 6-31G** (test/patin.dat.631.nh3), 30 basis functions", lambda B: B.Basis.from_patin(B.test_input("631.nh3"))),
    "fgh2o": ("H2O-like 3-centre system with s..g shells (tests/golden/inputs/patin.dat.fg.h2o), 67 basis functions: the runtime-L kernel",
              lambda B: B.Basis.from_patin(B.test_input("fg.h2o"))),
}


def synthetic_density(basis, seed=20261017):
    """seeded, symmetric, diagonally dominant stand-in for an SCF density (packed lower triangle)"""
    rng = np.random.default_rng(seed)
    n = basis.nbf
    A = rng.standard_normal((n, n)) * 0.05
    P = 0.5 * (A + A.T) + np.diag(rng.uniform(0.2, 1.0, n))
    return np.ascontiguousarray(P[np.tril_indices(n)])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_batch_fn(path):
    """(kind, C batch routine, its basis handle): the UNMODIFIED reference behind oracle/ref_harness.cc when oracle/_ref was
    built, else the oracle port"""
    from oracle.oracle import Oracle, Reference
    if Reference.available():
        R = Reference()
        return "reference", R.lib.ref_quartet_batch, R.basis(path), R
    O = Oracle()
    ob = O.basis(path)
    return "port", O.lib.oracle_quartet_batch, ob.h, (O, ob)


def cpu_quartet_rate(path, shells, P, no2, budget_s, nthreads):
    """quartets/s of the reference's ERI routine + formGMatrixKernel digestion over `shells`, looped in C
    (oracle/ref_harness.cc:ref_quartet_batch), chunk by chunk until the time budget is spent"""
    from oracle.oracle import timed_quartet_batches
    kind, fn, hd, keep = cpu_batch_fn(path)
    chunk = 100000 * nthreads
    timed_quartet_batches(fn, hd, shells[:2000 * nthreads], P, no2, 1, nthreads)     # warm-up: per-thread scratch objects
    n = 0
    dt = 0.0
    stored = 0
    while n < len(shells) and dt < budget_s:
        part = shells[n:n + chunk]
        t, st, _ = timed_quartet_batches(fn, hd, part, P, no2, 1, nthreads)
        dt += t; n += len(part); stored += st
    return n / dt, n, dt, kind, stored


def screened_sample_cpu(basis, nsample, tau, seed):
    """uniform sample of the screened canonical quartet list WITHOUT the GPU engine: sample shell pairs uniformly,
    get their Schwarz bounds from the reference's own diagonal quartets, keep quartets with Q_ab*Q_cd >= tau."""
    from oracle.oracle import Oracle, Reference
    rng = np.random.default_rng(seed)
    ns = basis.nshell
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "patin.dat"); basis.write_patin(path)
        if Reference.available():
            R = Reference(); hb = R.basis(path)
            blk = lambda i, j: R.quartet_block(hb, (nc(i), nc(j), nc(i), nc(j)), i, j, i, j)
        else:
            O = Oracle(); ob = O.basis(path)
            blk = lambda i, j: O.quartet_block(ob, i, j, i, j)
        nc = lambda s: (int(basis.lv[s]) + 1) * (int(basis.lv[s]) + 2) // 2
        npairs = 20000
        pi = rng.integers(0, ns, npairs); pj = rng.integers(0, ns, npairs)
        pi, pj = np.maximum(pi, pj), np.minimum(pi, pj)
        Q = np.zeros(npairs)
        cen = basis.xyz[basis.cen]
        far = np.sum((cen[pi] - cen[pj]) ** 2, axis=1) > 600.0     # > 24.5 bohr: negligible for these exponents
        for k in np.nonzero(~far)[0]:
            i, j = int(pi[k]), int(pj[k])
            b = blk(i, j)
            n1, n2 = nc(i), nc(j)
            Q[k] = np.sqrt(np.max(np.abs(b.reshape(n1 * n2, n1 * n2).diagonal())))
    # vectorised rejection sampling of (pair, pair) with Q_ab*Q_cd >= tau
    out = np.zeros((0, 4), np.int32)
    tries = 0
    while len(out) < nsample and tries < 400:
        a = rng.integers(0, npairs, 1 << 20); b2 = rng.integers(0, npairs, 1 << 20); tries += 1
        ok = (Q[a] * Q[b2] >= tau) & (Q[a] * Q[b2] > 0)
        a, b2 = a[ok], b2[ok]
        out = np.concatenate([out, np.stack([pi[a], pj[a], pi[b2], pj[b2]], axis=1).astype(np.int32)])
    out = out[:nsample]
    frac = float(np.mean((Q[:, None] * Q[None, :2000] >= tau))) if npairs else 0.0
    return out, frac


def load_workload(name, want_gpu_density=True):
    """(basis, packed P, density description, patin path).  The basis is round-tripped through a patin.dat file so that the
    GPU engine and the CPU checkers read the SAME numbers (the file carries 10 decimals)."""
    from unomol_b200 import basis as B
    desc, make = WORKLOADS[name]
    d = tempfile.mkdtemp(prefix="unomol_bench_")
    path = os.path.join(d, "patin.dat")
    make(B).write_patin(path)
    basis = B.Basis.from_patin(path)
    P, dens = None, None
    if name.startswith("water") and want_gpu_density:
        try:
            from unomol_b200 import driver
            _, P = driver.cluster_superposition_density(int(name[5:]))
            dens = "superposition of converged RHF/6-31G monomer densities (SURVEY 8(d)), block diagonal"
        except Exception as e:      # no GPU / no driver binary: the values of P do not change what is timed
            sys.stderr.write("bench: superposition density unavailable (%s), using the seeded synthetic P\n" % e)
    if P is None:
        P = synthetic_density(basis)
        dens = "seeded synthetic symmetric P (diagonally dominant)"
    return basis, np.ascontiguousarray(P), dens, path, desc


def reference_mpi_block(cores):
    """The reference's MPI path on the host cores (north_star: 'CPU (and MPI) path timed on the GPU box's own host cores')"""
    try:
        from oracle.oracle import run_reference_mpi
        from unomol_b200 import basis as B
        r = run_reference_mpi(B.test_input("tz2p.sf6"), cores)
    except Exception as e:      # the baseline must never take the bench line down
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}
    if r is None:
        return {"unavailable": "oracle/_ref/UnomolMPI is not built"}
    r["workload"] = WORKLOADS["sf6"][0]
    r["what"] = ("unmodified reference MPI driver over oracle/mpi_shim (fork + shared memory), RHF to convergence; integral pass = "
                 "TwoElectronIntsMPI::calculate (round-robin over the lsh loop, every rank stores its share), one SCF iteration = "
                 "MPI_Bcast(P) + digestion of the stored share + MPI_Reduce(G) + serial diagonalisation")
    return r


def oracle_parity(path, P, G_gpu, nelem=6, seed=20261017):
    """max relative deviation of a few G elements from the unscreened CPU oracle (sum over every shell quartet that touches
    the element, reference primitive cut and storage threshold), relative to max|G|"""
    from oracle.oracle import Oracle
    O = Oracle(); ob = O.basis(path)
    rng = np.random.default_rng(seed)
    n = ob.nbf
    pairs = [(0, 0), (n - 1, n - 1)]
    while len(pairs) < nelem:
        i, j = int(rng.integers(0, n)), int(rng.integers(0, n))
        if (max(i, j), min(i, j)) not in pairs:
            pairs.append((max(i, j), min(i, j)))
    # make half of the random picks near-diagonal (same or neighbouring molecule): that is where G is large
    for k in range(2, len(pairs), 2):
        i = pairs[k][0]
        pairs[k] = (i, max(0, i - int(rng.integers(0, 13))))
    t0 = time.perf_counter()
    g, nblk = O.g_elements(ob, P, pairs)
    dt = time.perf_counter() - t0
    got = np.array([G_gpu[i * (i + 1) // 2 + j] for i, j in pairs])
    scale = float(np.max(np.abs(G_gpu)))
    return {"elements": [list(p) for p in pairs], "max_rel_vs_oracle": float(np.max(np.abs(got - g)) / scale),
            "oracle": "oracle/unomol_oracle.c:oracle_g_elements_rhf, unscreened, %d shell-quartet blocks, %.1f s on the host cores"
                      % (nblk, dt), "max_abs_G": scale}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("UNOMOL_BENCH_WORKLOAD", "water154"))
    ap.add_argument("--tau", type=float, default=1e-12)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-extra", action="store_true", help="skip the SF6 / (H2O)_308 / UHF side measurements")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity block")
    ap.add_argument("--no-mpi", action="store_true", help="skip the run of the reference's MPI driver (SF6/TZ2P on all host cores)")
    ap.add_argument("--set", action="append", default=[], metavar="OPTION=VALUE",
                    help="engine option for experiments (unomol_b200_set_option), e.g. --set tile_kernels=0")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    have_cuda = False
    try:
        import torch
        have_cuda = torch.cuda.is_available()
    except Exception:
        pass
    if have_cuda:
        torch.cuda.set_device(local)
    basis, Pn, dens, patin, desc = load_workload(args.workload, want_gpu_density=have_cuda)
    config = {"workload": desc, "method": "RHF G=2J-K, integral-direct", "nbf": int(basis.nbf), "nshell": int(basis.nshell),
              "schwarz_tau": args.tau, "prim_cut": 1e-12, "density": dens,
              "l2": "pair tables + P/J/K working set exceeds L2 for water*, L2 flushed between steps otherwise"}

    # ------------------------------------------------------------------ reference arm (host cores)
    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count() or 1
        per_step = (250000 if basis.nbf > 300 else 60000) * cores
        sample_desc = None
        if have_cuda:
            # the GPU engine's own screened list, sampled uniformly (same list our arm processes)
            from unomol_b200 import capi
            h = capi.Handle(basis, device=local)
            h.set_option("schwarz_tau", args.tau)
            shells, ntot = h.sample_quartets(per_step, seed=7)
            h.close()
            sample_desc = "uniform seeded sample of the %d screened shell quartets the GPU arm processes" % ntot
        else:
            shells, frac = screened_sample_cpu(basis, per_step, args.tau, 7)
            sample_desc = "uniform seeded sample of the Schwarz-screened list (bounds from the reference's diagonal quartets)"
        from oracle.oracle import timed_quartet_batches
        kind, fn, hd, keep = cpu_batch_fn(patin)
        times = []
        for it in range(args.warmup + args.steps):
            dt, stored, _ = timed_quartet_batches(fn, hd, shells, Pn, basis.no2, 1, cores)
            if it >= args.warmup:
                times.append(dt)
        tot = sum(times)
        val = len(shells) * args.steps / tot
        line = {"impl": "reference", "metric": "eri_shell_quartets_per_s", "value": val, "unit": "quartets/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "quartets/s", "cores": cores, "kind": kind,
                                 "sample": "%d shell quartets per step (%s); per quartet the reference's calc_two_electron_ints_rys "
                                           "as calculate() calls it + storage threshold + formGMatrixKernel digestion into a per-thread G, "
                                           "looped in C (oracle/ref_harness.cc:ref_quartet_batch), %d host threads"
                                           % (len(shells), sample_desc, cores)},
                "e2e": {"value": val, "unit": "quartets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if not args.no_mpi:
            line["reference_mpi"] = reference_mpi_block(cores)
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch.distributed as dist
    from unomol_b200 import capi
    if not have_cuda:
        raise SystemExit("bench.py needs a CUDA device: unomol_b200 has no CPU fallback")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # ONE density for all ranks: the superposition density comes out of a monomer SCF converged to 1e-10, which every rank
        # would otherwise run for itself (rank-to-rank differences of that size showed up as 8e-11 in the N-rank G)
        tP = torch.from_numpy(Pn).cuda()
        dist.broadcast(tP, src=0)
        Pn = np.ascontiguousarray(tP.cpu().numpy())

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_handle(b, opts=()):
        hh = capi.Handle(b, device=local, rank=rank, nranks=world)
        hh.set_option("schwarz_tau", args.tau)
        for kv in opts:
            k, v = kv.split("=")
            hh.set_option(k, float(v))
        steal = False
        if world > 1 and not os.environ.get("UNOMOL_NO_STEAL"):
            from unomol_b200.multigpu import enable_work_stealing
            steal = enable_work_stealing(hh)
            if not steal:
                hh.set_option("work_stealing", 0)     # CUDA IPC unavailable: every rank falls back to the static split
        if world > 1 and nccl_state["comm"] is None and not os.environ.get("UNOMOL_TORCH_ALLREDUCE"):
            try:
                from unomol_b200.multigpu import NcclComm
                nccl_state["comm"] = NcclComm()
            except Exception as e:
                sys.stderr.write("bench: own NCCL communicator unavailable (%s); torch.distributed all-reduce\n" % e)
                nccl_state["comm"] = False
        if world > 1 and nccl_state["comm"]:
            hh.attach_nccl(nccl_state["comm"].ptr)   # the library all-reduces the packed G itself at the end of every build
        return hh, steal

    nccl_state = {"comm": None}
    h, stealing = make_handle(basis, args.set)
    lib_allreduce = bool(world > 1 and nccl_state["comm"])
    config["allreduce"] = ("inside the library (ncclAllReduce on the build stream, unomol_b200_attach_nccl)" if lib_allreduce else
                           ("torch.distributed all_reduce after the build" if world > 1 else "none"))
    for kv in args.set:
        k, v = kv.split("=")
        config.setdefault("options", {})[k] = float(v)
    config["multi_gpu_split"] = ("dynamic: shared work counters over NVLink (work stealing)" if stealing else
                                 ("static snake-order split" if world > 1 else "single GPU, dynamic CTA scheduling"))
    no2 = basis.no2
    P_host = torch.from_numpy(Pn).pin_memory()
    G_host = torch.zeros(no2, dtype=torch.float64).pin_memory()
    dP = P_host.cuda(); dG = torch.zeros(no2, dtype=torch.float64, device="cuda")
    stream_ptr, _, _ = h.device_buffers()
    ext = torch.cuda.ExternalStream(stream_ptr, device=torch.device("cuda", local))
    small = basis.nbf < 1000
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None

    def timed(hh, step_fn, nsteps, stream):
        """total milliseconds over nsteps (device events on the library stream), max over ranks"""
        sync_all()
        tot = 0.0
        with torch.cuda.stream(stream):
            for _ in range(nsteps):
                if flush is not None:
                    flush.fill_(1)
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream); step_fn(); e1.record(stream)
                e1.synchronize()
                tot += e0.elapsed_time(e1)
        sync_all()
        t = torch.tensor([tot], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def device_step():
        h.fock_rhf_device(dP.data_ptr(), dG.data_ptr(), async_=True)
        if world > 1 and not lib_allreduce:
            dist.all_reduce(dG)

    with torch.cuda.stream(ext):
        for _ in range(args.warmup):
            device_step()
    sampler = ClockSampler(local) if rank == 0 else None
    total_ms = timed(h, device_step, args.steps, ext)
    clocks = sampler.stop() if sampler else None
    ms_per_step = total_ms / args.steps

    # Counters and the kernel window of ONE build that every rank starts together (barrier first: with shared work counters the
    # first rank to start would otherwise take most of the work), then the all-reduce alone.
    # The window is taken from three such builds and the one with the shortest slowest-rank window is reported (a single build
    # carries +-0.5 % of noise, enough to put a window above the mean step time at N = 8).
    kmax = kmin = None
    for _ in range(3):
        sync_all()
        h.fock_rhf_device(dP.data_ptr(), dG.data_ptr(), async_=False)
        st = h.stats()
        k1 = torch.tensor([st["last_eri_kernel_ms"]], dtype=torch.float64, device="cuda"); k0 = k1.clone()
        if world > 1:
            dist.all_reduce(k1, op=dist.ReduceOp.MAX); dist.all_reduce(k0, op=dist.ReduceOp.MIN)
        if kmax is None or float(k1[0]) < float(kmax[0]):
            kmax, kmin = k1, k0
    nq = torch.tensor([float(st["n_quartets"]), float(st["model_flops"]), float(st["n_prim_quartets"])], dtype=torch.float64, device="cuda")
    allreduce_ms = 0.0
    if world > 1:
        dist.all_reduce(nq)
        scratch = torch.zeros_like(dG)
        with torch.cuda.stream(ext):
            for _ in range(3):
                dist.all_reduce(scratch)     # warm-up: torch's communicator has not moved this size yet
        allreduce_ms = timed(h, lambda: dist.all_reduce(scratch), 5, ext) / 5.0     # the same 8*no2 bytes, alone
        with torch.cuda.stream(ext):     # a clean all-reduced G for the parity block
            device_step()
        ext.synchronize()
    n_quartets, model_flops, n_primq = float(nq[0]), float(nq[1]), float(nq[2])
    kernel_ms, kernel_ms_min = float(kmax[0]), float(kmin[0])
    value = n_quartets / (ms_per_step * 1e-3)
    G_final = dG.cpu().numpy().copy()

    # parity: (1) N > 1: the all-reduced G against a ONE-rank build of the same P on rank 0; (2) a few elements of G against
    # the unscreened CPU oracle
    parity = {}
    if world > 1:
        if rank == 0:
            h1 = capi.Handle(basis, device=local, rank=0, nranks=1)
            h1.set_option("schwarz_tau", args.tau)
            G1 = h1.fock_rhf(Pn)
            h1.close()
            parity["n_rank_vs_1_rank_max_rel"] = float(np.max(np.abs(G_final - G1)) / np.max(np.abs(G1)))
        sync_all()
    if rank == 0 and not args.no_parity:
        parity.update(oracle_parity(patin, Pn, G_final))
        parity["max_rel"] = max(parity["max_rel_vs_oracle"], parity.get("n_rank_vs_1_rank_max_rel", 0.0))
    sync_all()

    # e2e: host buffers through the C ABI (N=1), or pinned H2D + device build + all-reduce + D2H (N>1)
    Gn = G_host.numpy()

    def e2e_step():
        if world == 1:
            Gn[:] = 0.0
            h.fock_rhf(Pn, Gn)
        else:
            with torch.cuda.stream(ext):      # copies, build and all-reduce ordered on the library's stream
                dP.copy_(P_host, non_blocking=True)
                device_step()
                G_host.copy_(dG, non_blocking=True)
            ext.synchronize()

    e2e_step()
    sync_all(); t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    sync_all(); e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_val = n_quartets * args.steps / e2e_s

    # One full SCF iteration, device-resident (unomol_b200_scf_iterate_rhf = reference RHF.hpp:87-112 with P, G, F, H on the
    # GPU): Fock build -> energy -> F = H + G -> X^T F X -> eigen-decomposition -> C -> P -> |dP|.  N > 1: every rank builds
    # its partial G, the packed device G is all-reduced over NCCL on the library stream, and every rank runs the (replicated)
    # cuSOLVER/cuBLAS algebra on its own GPU -- the counterpart of RHF_MPI::update (reference RHF_MPI.hpp:101-131).  H and S
    # are synthetic (identity overlap): the O(N^3) algebra does not depend on their values.
    scf_ms = None
    if basis.nbf <= 6000:
        n = basis.nbf
        Sd = np.zeros(no2); Sd[np.cumsum(np.arange(1, n + 1)) - 1] = 1.0
        h.scf_set_overlap(Sd)
        nocc = max(1, getattr(basis, "nelec", 2) // 2)
        Hn = -np.abs(Pn)
        h.scf_load(Hn, Pn)
        _, _, dGlib = h.device_buffers()

        class _DevView:                      # zero-copy torch view of the library's packed device G
            __cuda_array_interface__ = {"shape": (no2,), "typestr": "<f8", "data": (int(dGlib[0]), False), "version": 2}
        Gview = torch.as_tensor(_DevView(), device=torch.device("cuda", local)) if world > 1 else None

        def scf_iteration():
            if world == 1:
                return h.scf_iterate_rhf(nocc)
            with torch.cuda.stream(ext):
                h.scf_iterate_rhf_begin()
                if not lib_allreduce:
                    dist.all_reduce(Gview)
            return h.scf_iterate_rhf_finish(nocc)

        scf_iteration()                      # warm-up: cuSOLVER workspace, first-use allocations
        sync_all(); t0 = time.perf_counter()
        for _ in range(2):
            scf_iteration()
        sync_all()
        t = torch.tensor([(time.perf_counter() - t0) / 2 * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        scf_ms = float(t.item())

    peak = capi.fp64_peak(local)

    # ------------------------------------------------------------------ side measurements (configs 3-5 of BASELINE.json)
    def side_build(name, nspin=1, nsteps=3, opts=()):
        """ms per Fock build (all-reduce included at N > 1) and model-flop rate of another workload, same rank layout"""
        b2, P2, _, _, d2 = load_workload(name, want_gpu_density=False)     # seeded synthetic P: identical on every rank
        h2, _ = make_handle(b2, opts)
        sp, _, _ = h2.device_buffers()
        s2 = torch.cuda.ExternalStream(sp, device=torch.device("cuda", local))
        dP2 = torch.from_numpy(P2).cuda(); dPB = (0.6 * dP2).contiguous()
        dG2 = torch.zeros(b2.no2, dtype=torch.float64, device="cuda"); dGB = torch.zeros_like(dG2)

        def step():
            if nspin == 1:
                h2.fock_rhf_device(dP2.data_ptr(), dG2.data_ptr(), async_=True)
            else:
                h2.fock_uhf_device(dP2.data_ptr(), dPB.data_ptr(), dG2.data_ptr(), dGB.data_ptr(), async_=True)
            if world > 1 and not lib_allreduce:
                dist.all_reduce(dG2)
                if nspin == 2:
                    dist.all_reduce(dGB)
        with torch.cuda.stream(s2):
            for _ in range(2):
                step()
        ms = timed(h2, step, nsteps, s2) / nsteps
        sync_all()
        if nspin == 1:
            h2.fock_rhf_device(dP2.data_ptr(), dG2.data_ptr(), async_=False)
        else:
            h2.fock_uhf_device(dP2.data_ptr(), dPB.data_ptr(), dG2.data_ptr(), dGB.data_ptr(), async_=False)
        s = h2.stats()
        v = torch.tensor([float(s["n_quartets"]), float(s["model_flops"])], dtype=torch.float64, device="cuda")
        km = torch.tensor([s["last_eri_kernel_ms"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v); dist.all_reduce(km, op=dist.ReduceOp.MAX)
        h2.close()
        tf = float(v[1]) / (float(km[0]) * 1e-3) / 1e12 / world if float(km[0]) > 0 else 0.0
        return {"workload": d2, "method": "RHF" if nspin == 1 else "UHF", "ms_per_build": ms, "quartets_per_build": float(v[0]),
                "quartets_per_s": float(v[0]) / (ms * 1e-3), "kernel_ms_per_build_max_rank": float(km[0]),
                "model_tflops_per_gpu": tf, "frac_fp64_peak": tf / peak if peak else None, "precompute_ms": s["precompute_ms"]}

    extra = None
    if not args.no_extra and args.workload == "water154":
        extra = {"sf6_tz2p_rhf": side_build("sf6", 1, 10), "co2_dzp_uhf": side_build("co2", 2, 10), "fg_h2o_rhf": side_build("fgh2o", 1, 5),
                 "water154_uhf": side_build("water154", 2, 2), "water308_rhf": side_build("water308", 1, 2)}

    if rank == 0:
        achieved = model_flops / (kernel_ms * 1e-3) / 1e12 / max(world, 1) if kernel_ms > 0 else 0.0
        # DRAM / L2 bytes of one build from the committed ncu pass of the same workload (profiles/*traffic*.json), not measured live
        traffic = l2_bytes = None
        for tname in ("r2_traffic_%s.json", "r1_traffic_%s.json"):
            tpath = os.path.join(ROOT, "profiles", tname % args.workload)
            if os.path.exists(tpath) and world == 1:
                tj = json.load(open(tpath))["total"]
                traffic, l2_bytes = tj["dram_bytes"], tj["l2_bytes"]
                break
        line = {"metric": "eri_shell_quartets_per_s", "value": value, "unit": "quartets/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "fock_build_s": ms_per_step * 1e-3, "scf_iteration_s": (scf_ms * 1e-3 if scf_ms else None),
                "precompute_ms": st["precompute_ms"], "prim_quartets_per_build": n_primq,
                "quartets_per_build": n_quartets,
                "quartets_unscreened": float(st["n_quartets_total"]),
                "gpu_launches": int(st["n_launches"]) * args.steps,
                "launches_per_build": {"tile": st["n_tile_launches"], "reg": st["n_reg_launches"], "generic": st["n_generic_launches"],
                                       "highl": st["n_highl_launches"]},
                "clocks": clocks, "parity": parity,
                "e2e": {"value": e2e_val, "unit": "quartets/s", "h2d_bytes_per_step": int(no2 * 8),
                        "d2h_bytes_per_step": int(no2 * 8), "s_per_step": e2e_s / args.steps},
                "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": achieved / peak if peak else None, "traffic": traffic, "l2_bytes": l2_bytes,
                             "kernel": "eri_tile_kernel<*> (s/p classes) + eri_reg_kernel<*> / eri_class_kernel<*>: fused ERI + J/K digestion, "
                                       "all class launches of one build",
                             "kernel_ms_per_build": kernel_ms, "kernel_ms_per_build_min_rank": kernel_ms_min,
                             "allreduce_ms": allreduce_ms, "model_gflop_per_build": model_flops / 1e9,
                             "window": "first class launch to last class launch of one build started by all ranks together, slowest rank; "
                                       "per GPU: model flops of all ranks / n_gpus over that window",
                             "peak_source": "DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)"}}
        if extra:
            line["extra"] = extra
        if world == 1:
            shells, ntot = h.sample_quartets(6000000, seed=20261017)
            rate, n, dt, kind, stored = cpu_quartet_rate(patin, shells, Pn, no2, args.cpu_seconds, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "quartets/s", "cores": 1, "kind": kind,
                                    "sample": "%d of %d screened shell quartets (uniform seeded sample of the GPU's list), %.1f s; per quartet "
                                              "the reference's calc_two_electron_ints_rys as calculate() calls it + storage threshold + "
                                              "formGMatrixKernel digestion (%d stored integrals), looped in C (oracle/ref_harness.cc:"
                                              "ref_quartet_batch)" % (n, ntot, dt, stored)}
            if not args.no_mpi:
                line["reference_mpi"] = reference_mpi_block(os.cpu_count() or 1)
        print(json.dumps(line))
    sync_all()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
