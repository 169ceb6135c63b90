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


def cpu_quartet_rate(basis, shells, budget_s, kind_pref="reference"):
    """time the reference's (or the oracle port's) per-quartet ERI routine over the given shell quartets"""
    from oracle.oracle import Oracle, Reference
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "patin.dat")
        basis.write_patin(path)
        if kind_pref == "reference" and Reference.available():
            R = Reference(); hb = R.basis(path); kind = "reference"
            nc = lambda s: (int(basis.lv[s]) + 1) * (int(basis.lv[s]) + 2) // 2
            out = np.zeros(1296)
            import ctypes
            fn = R.lib.ref_quartet_block; outp = out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
            call = lambda q: fn(hb, int(q[0]), int(q[1]), int(q[2]), int(q[3]), outp)
        else:
            O = Oracle(); ob = O.basis(path); kind = "port"
            import ctypes
            out = np.zeros(1296)
            fn = O.lib.oracle_quartet_block; outp = out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
            call = lambda q: fn(ob.h, int(q[0]), int(q[1]), int(q[2]), int(q[3]), outp)
        t0 = time.perf_counter(); n = 0
        for q in shells:
            call(q); n += 1
            if (n & 255) == 0 and time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    return n / dt, n, dt, kind


def _ref_worker(args):
    path, shells, lv = args
    from oracle.oracle import Oracle, Reference
    import ctypes
    out = np.zeros(1296); outp = out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    if Reference.available():
        R = Reference(); hb = R.basis(path); fn = R.lib.ref_quartet_block
        call = lambda q: fn(hb, int(q[0]), int(q[1]), int(q[2]), int(q[3]), outp)
    else:
        O = Oracle(); ob = O.basis(path); fn = O.lib.oracle_quartet_block
        call = lambda q: fn(ob.h, int(q[0]), int(q[1]), int(q[2]), int(q[3]), outp)
    t0 = time.perf_counter()
    for q in shells:
        call(q)
    return time.perf_counter() - t0


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("UNOMOL_BENCH_WORKLOAD", "water154"))
    ap.add_argument("--tau", type=float, default=1e-12)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--set", action="append", default=[], metavar="OPTION=VALUE",
                    help="engine option for experiments (unomol_b200_set_option), e.g. --set ket_runs=0")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from unomol_b200 import basis as B
    desc, make = WORKLOADS[args.workload]
    basis = make(B)
    config = {"workload": desc, "method": "RHF G=2J-K, integral-direct", "nbf": int(basis.nbf), "nshell": int(basis.nshell),
              "schwarz_tau": args.tau, "prim_cut": 1e-12, "density": "seeded synthetic symmetric P",
              "l2": "pair tables + P/J/K working set exceeds L2 for water*, L2 flushed between steps otherwise"}

    # ------------------------------------------------------------------ reference arm (host cores)
    if args.impl == "reference":
        if rank != 0:
            return 0
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        from oracle.oracle import Reference
        per_step = 200000 * cores if basis.nbf > 300 else 50000 * cores
        shells, frac = screened_sample_cpu(basis, per_step, args.tau, 7)
        d = tempfile.mkdtemp(); path = os.path.join(d, "patin.dat"); basis.write_patin(path)
        chunks = [shells[i::cores] for i in range(cores)]
        times = []
        with mp.get_context("fork").Pool(cores) as pool:
            for it in range(args.warmup + args.steps):
                t0 = time.perf_counter()
                pool.map(_ref_worker, [(path, c, None) for c in chunks])
                dt = time.perf_counter() - t0
                if it >= args.warmup:
                    times.append(dt)
        tot = sum(times)
        val = len(shells) * args.steps / tot
        kind = "reference" if Reference.available() else "port"
        line = {"impl": "reference", "metric": "eri_shell_quartets_per_s", "value": val, "unit": "quartets/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "quartets/s", "cores": cores, "kind": kind,
                                 "sample": "%d screened shell quartets per step (uniform seeded sample, Q_ab*Q_cd>=%g), "
                                           "reference calc_two_electron_ints_rys only (no digestion), %d forked workers"
                                           % (len(shells), args.tau, cores)},
                "e2e": {"value": val, "unit": "quartets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from unomol_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: unomol_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h = capi.Handle(basis, device=local, rank=rank, nranks=world)
    h.set_option("schwarz_tau", args.tau)
    for kv in args.set:
        k, v = kv.split("=")
        h.set_option(k, float(v))
        config.setdefault("options", {})[k] = float(v)
    stealing = False
    if world > 1 and not os.environ.get("UNOMOL_NO_STEAL"):
        from unomol_b200.multigpu import enable_work_stealing
        stealing = enable_work_stealing(h)
        if not stealing:
            h.set_option("work_stealing", 0)     # CUDA IPC unavailable: every rank falls back to the static split
    config["multi_gpu_split"] = ("dynamic: shared work counters over NVLink (work stealing)" if stealing else
                                 ("static snake-order split" if world > 1 else "single GPU, dynamic CTA scheduling"))
    no2 = basis.no2
    P_host = torch.from_numpy(synthetic_density(basis)).pin_memory()
    G_host = torch.zeros(no2, dtype=torch.float64).pin_memory()
    dP = P_host.cuda(); dG = torch.zeros(no2, dtype=torch.float64, device="cuda")
    stream_ptr, _, _ = h.device_buffers()
    ext = torch.cuda.ExternalStream(stream_ptr, device=torch.device("cuda", local))
    small = basis.nbf < 1000
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        h.fock_rhf_device(dP.data_ptr(), dG.data_ptr(), async_=True)
        if world > 1:
            dist.all_reduce(dG)

    def run_timed(step_fn, nsteps):
        """returns total milliseconds over nsteps (device events on the library stream), max over ranks"""
        sync_all()
        tot = 0.0
        with torch.cuda.stream(ext):
            for _ in range(nsteps):
                if flush is not None:
                    flush.fill_(1)
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(ext); step_fn(); e1.record(ext)
                e1.synchronize()
                tot += e0.elapsed_time(e1)
        sync_all()
        t = torch.tensor([tot], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.cuda.stream(ext):
        for _ in range(args.warmup):
            device_step()
    sampler = ClockSampler(local) if rank == 0 else None
    total_ms = run_timed(device_step, args.steps)
    clocks = sampler.stop() if sampler else None

    # one synchronous build for the counters (quartets, primitive quartets, kernel time)
    h.fock_rhf_device(dP.data_ptr(), dG.data_ptr(), async_=False)
    st = h.stats()
    nq = torch.tensor([float(st["n_quartets"]), float(st["model_flops"])], dtype=torch.float64, device="cuda")
    kms = torch.tensor([st["last_eri_kernel_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(nq); dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    n_quartets, model_flops, kernel_ms = float(nq[0]), float(nq[1]), float(kms[0])
    ms_per_step = total_ms / args.steps
    value = n_quartets / (ms_per_step * 1e-3)

    # e2e: host buffers through the C ABI (N=1), or pinned H2D + device build + all-reduce + D2H (N>1)
    Pn = P_host.numpy(); Gn = G_host.numpy()

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

    # one full SCF iteration on the device path: Fock build (host P -> host G through the C ABI) + F = H + G ->
    # X^T F X -> eigen-decomposition -> C -> P on cuSOLVER/cuBLAS (reference RHF.hpp:87-112).  H and S are synthetic
    # (identity-like overlap): the O(N^3) algebra does not depend on their values.
    scf_ms = None
    if basis.nbf <= 6000:
        # One full SCF iteration, device-resident (unomol_b200_scf_iterate_rhf = reference RHF.hpp:87-112 with P, G, F, H on
        # the GPU): Fock build -> energy -> F = H + G -> X^T F X -> eigen-decomposition -> C -> P -> |dP|.  N > 1: every rank
        # builds its partial G, the packed device G is all-reduced over NCCL on the library stream, and every rank runs the
        # (replicated) cuSOLVER/cuBLAS algebra on its own GPU -- the counterpart of RHF_MPI::update, where rank 0 does the
        # algebra and broadcasts P (reference RHF_MPI.hpp:101-131).  H and S are synthetic (identity overlap): the O(N^3)
        # algebra does not depend on their values.
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

    if rank == 0:
        peak = capi.fp64_peak(local)
        achieved = model_flops / (kernel_ms * 1e-3) / 1e12 / max(world, 1) if kernel_ms > 0 else 0.0
        # per-GPU roofline: model flops of all ranks / world over the slowest rank's kernel time
        # DRAM / L2 bytes of one build from the committed ncu pass of the same workload (profiles/r1_traffic_*.json;
        # `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum`), not measured live
        traffic = l2_bytes = None
        tpath = os.path.join(ROOT, "profiles", "r1_traffic_%s.json" % args.workload)
        if os.path.exists(tpath) and world == 1:
            tj = json.load(open(tpath))["total"]
            traffic, l2_bytes = tj["dram_bytes"], tj["l2_bytes"]
        line = {"metric": "eri_shell_quartets_per_s", "value": value, "unit": "quartets/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "fock_build_s": ms_per_step * 1e-3, "scf_iteration_s": (scf_ms * 1e-3 if scf_ms else None),
                "precompute_ms": st["precompute_ms"], "prim_quartets_per_build": float(st["n_prim_quartets"]),
                "quartets_per_build": n_quartets,
                "quartets_unscreened": float(st["n_quartets_total"]),
                "gpu_launches": int(st["n_launches"]) * args.steps,
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "quartets/s", "h2d_bytes_per_step": int(no2 * 8),
                        "d2h_bytes_per_step": int(no2 * 8), "s_per_step": e2e_s / args.steps},
                "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": achieved / peak if peak else None, "traffic": traffic, "l2_bytes": l2_bytes,
                             "kernel": "eri_reg_kernel<*> / eri_class_kernel<*> (fused ERI + J/K digestion, all class launches of one build)",
                             "kernel_ms_per_build": kernel_ms, "model_gflop_per_build": model_flops / 1e9,
                             "peak_source": "DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)"}}
        if world == 1:
            ns_cap = 4000000
            shells, ntot = h.sample_quartets(ns_cap, seed=20261017)
            rate, n, dt, kind = cpu_quartet_rate(basis, shells, args.cpu_seconds)
            line["cpu_baseline"] = {"value": rate, "unit": "quartets/s", "cores": 1, "kind": kind,
                                    "sample": "%d of %d screened shell quartets (uniform seeded sample of the GPU's list), "
                                              "%.1f s, ERI evaluation only (reference calc_two_electron_ints_rys, no digestion)"
                                              % (n, ntot, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
