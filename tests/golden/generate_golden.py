#!/usr/bin/env python3
"""tests/golden/generate_golden.py -- regenerates the committed fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and oracle/_ref built by oracle/Makefile):
    python tests/golden/generate_golden.py [--with-sf6]
Writes, next to this script:
  inputs/patin.dat.*      verbatim copies of the reference's test INPUT files (data, not source)
  short_dat.json          the reference's own golden energies test/short.dat.* (E_iter0, E_final, dE)
  ref_runs.json           fresh runs of oracle/_ref/Unomol (finite-field flag off): E_iter0, E_final, iterations
  rys_grid.npz            roots/weights of the reference's Rys::root1..5 on a fixed X grid
  rys_rootn_grid.npz      roots/weights of the reference's Rys::rootN (6..9 roots) at the X where it returns
  eri_3g_h2o.npz          every stored unique integral of H2O/STO-3G (reference cache dump)
  eri_631_nh3.npz, eri_631_co.npz   ditto (108 345 computed each)
  g_<input>.npz           reference formGmatrix for seeded random P (RHF and UHF), several inputs
  quartets_<input>.npz    reference calc_two_electron_ints_rys blocks for seeded random ordered shell quartets
  *_fg_h2o.npz            the same for OUR f/g-shell input inputs/patin.dat.fg.h2o (see highl_input)
  moments/moments.out.*   moments.out of fresh reference runs; moments/moments.dat.* the reference's checked-in goldens
  momints_<input>.npz     the reference's raw dipole/quadrupole integrals (RMOM.DAT)
"""
import json, os, re, shutil, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle import Reference  # noqa: E402

REFT = "/root/reference/test"
INPUTS = ["3g.h2", "3g.h2o", "3g.hf", "3g.nh3", "3g.ch4", "3g.co", "3g.n2", "431.h2o", "431.nh3", "631.h2", "631.hf",
          "631.h2o", "631.nh3", "631.co", "631.n2", "631.ch4", "b.dhdz", "c.dhdz", "o.dhdz", "f.dhdz", "d6s3p.h2",
          "dh95.co2", "dh95.c2h2", "tz2p.sf6"]


def run_unomol(name):
    d = tempfile.mkdtemp()
    try:
        txt = open(os.path.join(REFT, "patin.dat." + name)).read().split("\n")
        # 4th data line is 'int_flag[0] int_flag[1]' -> finite-field off so only the ground-state SCF runs
        k = [i for i, l in enumerate(txt) if l.strip()][3]
        txt[k] = " 0 0"
        open(os.path.join(d, "patin.dat"), "w").write("\n".join(txt))
        p = subprocess.run([Reference.UNOMOL], cwd=d, capture_output=True, text=True, timeout=3600)
        s = [float(x) for x in open(os.path.join(d, "short.gs.out")).read().split()]
        out = open(os.path.join(d, "scfout.gs.out")).read()
        its = int(re.search(r"Final Iteration\s*=\s*(\d+)", out).group(1))
        conv = "NOT_ REACHED" not in out
        m = re.search(r"Time for Two Electrons Integrals = ([\d.e+-]+)", p.stderr)
        m2 = re.search(r"SCF time = ([\d.e+-]+)", p.stderr)
        ev = [float(x.split()[1]) for x in re.findall(r"^\s+\d+\s+[-\d.e+]+\s+\d+\s*$", out, flags=re.M)]
        return dict(e_init=s[0], e_final=s[1], de=s[2], iterations=its, converged=conv,
                    eri_seconds=float(m.group(1)) if m else None, scf_seconds=float(m2.group(1)) if m2 else None,
                    orbital_energies=ev)
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    with_sf6 = "--with-sf6" in sys.argv
    R = Reference()
    os.makedirs(os.path.join(HERE, "inputs"), exist_ok=True)
    for n in INPUTS:
        shutil.copyfile(os.path.join(REFT, "patin.dat." + n), os.path.join(HERE, "inputs", "patin.dat." + n))
    short = {}
    for f in sorted(os.listdir(REFT)):
        if f.startswith("short.dat."):
            short[f[len("short.dat."):]] = [float(x) for x in open(os.path.join(REFT, f)).read().split()]
    json.dump(short, open(os.path.join(HERE, "short_dat.json"), "w"), indent=1)
    # Rys grid
    xs = np.concatenate([np.linspace(0.0, 60.0, 1201), [1e-9, 3e-7, 3.1e-7, 15.05, 16.0, 18.0, 20.0, 25.0, 28.0, 33.0,
                                                        33.01, 40.0, 40.01, 47.01, 53.01, 59.01, 80.0, 500.0]])
    rys = {"x": xs}
    for n in range(1, 6):
        rr = np.zeros((len(xs), n)); ww = np.zeros((len(xs), n))
        for i, x in enumerate(xs):
            rr[i], ww[i] = R.rys_roots(n, x)
        rys["r%d" % n] = rr; rys["w%d" % n] = ww
    np.savez_compressed(os.path.join(HERE, "rys_grid.npz"), **rys)
    # stored integral lists
    for n in ["3g.h2o", "631.nh3", "631.co"]:
        h = R.basis(os.path.join(REFT, "patin.dat." + n)); t = R.tints(h)
        vals, ijkl = R.tints_dump(t)
        np.savez_compressed(os.path.join(HERE, "eri_%s.npz" % n.replace(".", "_")), vals=vals, ijkl=ijkl)
        R.tints_destroy(t); R.basis_close(h)
    # G matrices for seeded random P, and random quartet blocks
    for n in ["3g.h2o", "631.nh3", "631.co", "b.dhdz", "dh95.co2", "dh95.c2h2"] + (["tz2p.sf6"] if with_sf6 else []):
        h = R.basis(os.path.join(REFT, "patin.dat." + n))
        nbf = R.lib.ref_basis_norb(h); no2 = nbf * (nbf + 1) // 2
        rng = np.random.default_rng(12345)
        P = rng.standard_normal(no2); PB = rng.standard_normal(no2)
        t = R.tints(h)
        G, _ = R.form_g_rhf(t, P)
        GA, GB, _ = R.form_g_uhf(t, P, PB)
        S, T, H = R.one_electron(h)
        np.savez_compressed(os.path.join(HERE, "g_%s.npz" % n.replace(".", "_")), P=P, PB=PB, G=G, GA=GA, GB=GB, S=S, T=T, H=H)
        R.tints_destroy(t)
        shells = R.basis_shells(h)
        ns = len(shells)
        nq = 60 if nbf > 50 else 120
        quart = rng.integers(0, ns, size=(nq, 4))
        blocks = []
        for (i, j, k, l) in quart:
            nc = lambda s: (shells[s][1] + 1) * (shells[s][1] + 2) // 2
            blocks.append(R.quartet_block(h, (nc(i), nc(j), nc(k), nc(l)), int(i), int(j), int(k), int(l)).ravel())
        np.savez_compressed(os.path.join(HERE, "quartets_%s.npz" % n.replace(".", "_")), quartets=quart,
                            offsets=np.cumsum([0] + [len(b) for b in blocks]), values=np.concatenate(blocks))
        R.basis_close(h)
    # fresh reference SCF runs
    runs = {}
    path = os.path.join(HERE, "ref_runs.json")
    if os.path.exists(path):
        runs = json.load(open(path))
    for n in INPUTS:
        if n == "tz2p.sf6" and not with_sf6:
            continue
        if n in runs:
            continue
        runs[n] = run_unomol(n)
        print(n, runs[n]["e_final"], runs[n]["iterations"], flush=True)
        json.dump(runs, open(path, "w"), indent=1)


def cation_variants():
    """BASELINE config 3 (UHF on the DZP molecules): the reference picks UHF only by the parity of nelec
    (Unomol.cc:13), so the cation inputs are the shipped files with nelec lowered by one."""
    runs = json.load(open(os.path.join(HERE, "ref_runs.json")))
    global REFT
    for name, ne in [("dh95.co2", 21), ("dh95.c2h2", 13)]:
        lines = open(os.path.join(HERE, "inputs", "patin.dat." + name)).read().split("\n")
        k = [i for i, l in enumerate(lines) if l.strip()][1]
        f = lines[k].split(); lines[k] = "     %d    %s" % (ne, f[1])
        open(os.path.join(HERE, "inputs", "patin.dat.%s.cation" % name), "w").write("\n".join(lines))
    REFT = os.path.join(HERE, "inputs")
    for n in ["dh95.co2.cation", "dh95.c2h2.cation"]:
        runs[n] = run_unomol(n)
    json.dump(runs, open(os.path.join(HERE, "ref_runs.json"), "w"), indent=1)


def highl_input():
    """f/g shells: none of the reference's shipped inputs has l > 2, so inputs/patin.dat.fg.h2o is OURS
    (a water-like 3-centre system with s..g shells on O, s p f / s d f on the hydrogens, 67 functions, exponents chosen so
    that one-, two- and multi-centre quartets with l_tot > 8 occur and some reach t > 20 in the reference's Fgamma).
    Fixtures come from the unmodified reference: stored-list G matrices, shell-quartet blocks through
    calc_two_electron_ints_rys (l_tot <= 8) / calc_two_electron_ints_md (l_tot > 8), and a full SCF run."""
    global REFT
    R = Reference()
    runs = json.load(open(os.path.join(HERE, "ref_runs.json")))
    REFT = os.path.join(HERE, "inputs")
    fixed = {"fg.h2o": [(5, 4, 3, 3), (5, 4, 8, 8), (8, 8, 11, 11), (5, 5, 11, 11), (8, 5, 11, 4), (4, 8, 5, 11), (11, 5, 8, 3),
                        (5, 0, 0, 0), (4, 4, 0, 0), (5, 3, 2, 1), (8, 2, 11, 10), (5, 5, 2, 2), (4, 4, 4, 4), (3, 5, 4, 2), (5, 5, 5, 4)],
             # fg2.hf (ours): CONTRACTED f and g shells, so the same-shell primitive triangle with doubled off-diagonal
             # coefficients (TwoElectronInts.cpp:293-299) and primitive sums are exercised on the McMurchie-Davidson path
             "fg2.hf": [(3, 3, 3, 3), (4, 4, 7, 7), (4, 3, 3, 2), (7, 7, 7, 7), (4, 3, 7, 7), (7, 4, 3, 7), (3, 3, 2, 2),
                        (4, 0, 0, 0), (3, 1, 2, 0), (7, 5, 6, 5), (4, 4, 4, 3)]}
    for n in ["fg.h2o", "fg2.hf"]:
        path = os.path.join(HERE, "inputs", "patin.dat." + n)
        h = R.basis(path)
        nbf = R.lib.ref_basis_norb(h); no2 = nbf * (nbf + 1) // 2
        rng = np.random.default_rng(12345)
        P = rng.standard_normal(no2); PB = rng.standard_normal(no2)
        t = R.tints(h)
        G, _ = R.form_g_rhf(t, P)
        GA, GB, _ = R.form_g_uhf(t, P, PB)
        S, T, H = R.one_electron(h)
        np.savez_compressed(os.path.join(HERE, "g_%s.npz" % n.replace(".", "_")), P=P, PB=PB, G=G, GA=GA, GB=GB, S=S, T=T, H=H)
        R.tints_destroy(t)
        shells = R.basis_shells(h)
        ns = len(shells)
        nc = lambda s: (shells[s][1] + 1) * (shells[s][1] + 2) // 2
        quart = list(fixed[n])
        for q in rng.integers(0, ns, size=(400, 4)):
            q = tuple(int(x) for x in q)
            if max(shells[s][1] for s in q) >= 3 and nc(q[0]) * nc(q[1]) * nc(q[2]) * nc(q[3]) <= 6000 and len(quart) < (75 if n == "fg.h2o" else 45):
                quart.append(q)
        blocks = [R.quartet_block(h, (nc(i), nc(j), nc(k), nc(l)), i, j, k, l).ravel() for (i, j, k, l) in quart]
        np.savez_compressed(os.path.join(HERE, "quartets_%s.npz" % n.replace(".", "_")), quartets=np.array(quart),
                            offsets=np.cumsum([0] + [len(b) for b in blocks]), values=np.concatenate(blocks))
        R.basis_close(h)
        runs[n] = run_unomol(n)
        print(n, runs[n]["e_final"], runs[n]["iterations"], flush=True)
    json.dump(runs, open(os.path.join(HERE, "ref_runs.json"), "w"), indent=1)


def moments_fixtures():
    """moments.out of fresh runs of the unmodified reference (the checked-in test/moments.dat.* carry a stale electronic
    qyy) and the reference's raw moment integrals (RMOM.DAT records, Structs.hpp:17-20)."""
    os.makedirs(os.path.join(HERE, "moments"), exist_ok=True)
    for n in ["3g.h2o", "631.h2o", "631.nh3", "631.co", "dh95.co2", "b.dhdz", "dh95.co2.cation", "fg.h2o", "fg2.hf"]:
        d = tempfile.mkdtemp()
        try:
            txt = open(os.path.join(HERE, "inputs", "patin.dat." + n)).read().split("\n")
            k = [i for i, l in enumerate(txt) if l.strip()][3]
            txt[k] = " 0 0"
            open(os.path.join(d, "patin.dat"), "w").write("\n".join(txt))
            subprocess.run([Reference.UNOMOL], cwd=d, capture_output=True, text=True, timeout=3600)
            shutil.copyfile(os.path.join(d, "moments.out"), os.path.join(HERE, "moments", "moments.out." + n))
            if n in ("631.nh3", "dh95.co2", "fg.h2o", "fg2.hf"):
                rec = np.fromfile(os.path.join(d, "RMOM.DAT"), dtype=np.dtype([("v", "f8", 9), ("ijr", "u4"), ("pad", "u4")]))
                m = np.zeros((9, len(rec))); m[:, rec["ijr"]] = rec["v"].T
                np.savez_compressed(os.path.join(HERE, "momints_%s.npz" % n.replace(".", "_")), m=m)
        finally:
            shutil.rmtree(d, ignore_errors=True)
    for f in sorted(os.listdir(REFT0)):
        if f.startswith("moments.dat."):
            shutil.copyfile(os.path.join(REFT0, f), os.path.join(HERE, "moments", f))


REFT0 = "/root/reference/test"


def polscan_fixture(uhf=False):
    """polarisation-potential scan (reference RHF.hpp:292-388, UHF.hpp:293-383): the reference ships no input for it, so this one
    is ours -- water/6-31G (uhf: its cation, 9 electrons) with int_flag[0] = 1, three extra shells (s2 s1 p1) in posin.bas, six
    grid points in pos.grid.dat -- and the unmodified reference is run on it; vpol.out / spol.out are the goldens
    (tests/test_gpu_scf.py::test_polarisation_scan...)."""
    d = os.path.join(HERE, "polscan_uhf" if uhf else "polscan")
    os.makedirs(d, exist_ok=True)
    txt = open(os.path.join(REFT0, "patin.dat.631.h2o")).read().split("\n")
    nb = [i for i, l in enumerate(txt) if l.strip()]
    txt[nb[3]] = " 1 0"
    if uhf:
        t1 = txt[nb[1]].split()
        assert int(t1[0]) == 10
        txt[nb[1]] = "      9 " + " ".join(t1[1:])
    open(os.path.join(d, "patin.dat"), "w").write("\n".join(txt))
    open(os.path.join(d, "posin.bas"), "w").write(" 3 5 1\n 2 0\n   1.6000000000   0.4000000000\n   0.4500000000   0.7000000000\n"
                                                  " 1 0\n   0.1200000000   1.0000000000\n 1 1\n   0.2500000000   1.0000000000\n")
    open(os.path.join(d, "pos.grid.dat"), "w").write(" 6\n 0.0 0.0 6.0\n 0.0 0.0 4.5\n 0.3 0.2 3.5\n 0.6 0.4 2.8\n 1.5 -0.5 2.0\n 3.0 1.0 -1.5\n")
    t = tempfile.mkdtemp()
    for f in ("patin.dat", "posin.bas", "pos.grid.dat"):
        shutil.copyfile(os.path.join(d, f), os.path.join(t, f))
    subprocess.run([Reference.UNOMOL], cwd=t, capture_output=True, text=True, timeout=3600, check=True)
    for f in ("vpol.out", "spol.out"):
        shutil.copyfile(os.path.join(t, f), os.path.join(d, f))
    shutil.rmtree(t, ignore_errors=True)

def finitefield_fixtures():
    """finitefield/<input>.out: finitefield.out of the unmodified reference run on the shipped inputs as they are (int_flag[1] = 1:
    Unomol.cc:16-17 -> RHF.hpp:235-271 / UHF.hpp:239-272, three SCFs in a field of 5e-3 a.u. along x, y, z)."""
    d = os.path.join(HERE, "finitefield")
    os.makedirs(d, exist_ok=True)
    # open shell: the water cation (non-degenerate hole; the CO2 / C2H2 cations have a degenerate pi hole and their UHF solutions in
    # a field depend on which component the ground state happened to pick -- the reference's own x and y values differ there)
    for name in ["3g.h2o", "631.nh3", "631.co", "631.h2o.cation"]:
        t = tempfile.mkdtemp()
        txt = open(os.path.join(HERE, "inputs", "patin.dat." + name.replace(".cation", ""))).read().split("\n")
        nb = [i for i, l in enumerate(txt) if l.strip()]
        k = nb[3]
        if name.endswith(".cation"):
            f = txt[nb[1]].split(); txt[nb[1]] = "     %d    %s" % (int(f[0]) - 1, f[1])
        txt[k] = " 0 1"
        open(os.path.join(t, "patin.dat"), "w").write("\n".join(txt))
        subprocess.run([Reference.UNOMOL], cwd=t, capture_output=True, text=True, timeout=3600, check=True)
        shutil.copyfile(os.path.join(t, "finitefield.out"), os.path.join(d, name + ".out"))
        shutil.rmtree(t, ignore_errors=True)


def rootn_fixture():
    """rys_rootn_grid.npz: the reference's general routine Rys::rootN (Rys.cpp:231-312) for 6..9 roots wherever it returns.  It
    smashes its stack or hangs for 2 <~ X <~ 15, so every point is probed in a child process with a
    timeout; the fixture keeps the points that came back (r = t^2/(1-t^2) and weights as the reference stores them)."""
    probe = (
        "import ctypes, sys\n"
        "L = ctypes.CDLL(sys.argv[1])\n"
        "L.ref_rys_rootN.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]\n"
        "n = int(sys.argv[2]); r = (ctypes.c_double * 9)(); w = (ctypes.c_double * 9)()\n"
        "for x in sys.argv[3:]:\n"
        "    rc = L.ref_rys_rootN(n, float(x), r, w)\n"
        "    print('OK', x, rc, ' '.join(repr(v) for v in list(r)[:n] + list(w)[:n]), flush=True)\n")
    lib = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "libunomol_ref.so")
    xs = [0.0, 1e-6, 0.01, 0.1, 0.25, 0.5, 0.75, 1.0, 1.25, 1.5, 1.75, 2.0, 2.5, 3.0, 5.0, 8.0, 12.0, 14.0, 15.0, 15.5] + \
         [16.0 + 0.5 * i for i in range(60)] + [47.0, 50.0, 55.0, 60.0, 66.0, 70.0, 75.0, 80.0, 88.0, 95.0, 110.0, 150.0, 300.0]
    out = {}
    for n in range(6, 10):
        good_x, good_v = [], []
        for x in xs:                      # one child per point: a crash must not take its neighbours down
            try:
                p = subprocess.run([sys.executable, "-c", probe, lib, str(n), repr(x)], capture_output=True, text=True, timeout=10)
            except subprocess.TimeoutExpired:
                continue
            for line in p.stdout.splitlines():
                t = line.split()
                if t and t[0] == "OK" and int(t[2]) == 0 and p.returncode == 0:
                    v = [float(u) for u in t[3:]]
                    if all(np.isfinite(v)):
                        good_x.append(float(t[1])); good_v.append(v)
        out["x%d" % n] = np.array(good_x)
        out["r%d" % n] = np.array([v[:n] for v in good_v]); out["w%d" % n] = np.array([v[n:] for v in good_v])
        sys.stderr.write("rootN n=%d: %d of %d points returned (X = %s ...)\n" % (n, len(good_x), len(xs), good_x[:14]))
    np.savez_compressed(os.path.join(HERE, "rys_rootn_grid.npz"), **out)


if __name__ == "__main__":
    if "--finitefield-only" in sys.argv:
        finitefield_fixtures()
        sys.exit(0)
    if "--rootn-only" in sys.argv:
        rootn_fixture()
        sys.exit(0)
    if "--moments-only" in sys.argv:
        moments_fixtures()
        sys.exit(0)
    if "--highl-only" in sys.argv:
        highl_input()
        sys.exit(0)
    if "--polscan-only" in sys.argv:
        polscan_fixture()
        polscan_fixture(uhf=True)
        sys.exit(0)
    main()
    cation_variants()
    highl_input()
    moments_fixtures()
    polscan_fixture()
    polscan_fixture(uhf=True)
    finitefield_fixtures()
    rootn_fixture()
