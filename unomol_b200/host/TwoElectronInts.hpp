// unomol_b200/host/TwoElectronInts.hpp -- drop-in for the reference class unomol::TwoElectronInts
// (reference TwoElectronInts.hpp:79-110): same constructor and public methods, forwarding to the C ABI of
// libunomol_b200.so (include/unomol_b200.h).  Works with the reference's own Basis as well as with
// unomol_b200/host/Basis.hpp: it only uses the accessors both provide.
//
//   TwoElectronInts(const Basis&, int start_shell, const std::string& base_str)   reference :83-88
//   calculate / recalculate(const Basis&)                                         reference :94-98
//   formGmatrix(P, G)               G += 2J - K        reference TwoElectronInts.cpp:822-844
//   formGmatrix(PA, PB, GA, GB)     G^s += J - K^s     reference TwoElectronInts.cpp:846-869
//   directFormGMatrix(P, G, basis)  integral-direct RHF, reference :871-1055 (here every build is direct)
// base_str named the MINTS.DAT integral cache file (reference Unomol.cc:9); there is no cache any more, the
// argument is accepted and ignored.  Errors terminate the process like the reference's fatal_error
// (Util.cpp:5-8), after printing unomol_b200_strerror().
//
// Multi-GPU from C++ (replaces `mpirun -n N UnomolMPI`, reference UnomolMPI.cc / TwoElectronIntsMPI.cpp:350-354 /
// RHF_MPI.hpp:101-111): UNOMOL_GPUS=N makes this object drive N GPUs of the box from one process -- one C-ABI handle
// (rank r of N) and one host thread per GPU for the duration of a Fock build, the screened quartets handed out through
// work counters shared over NVLink peer access (unomol_b200_steal_share; static snake split if the GPUs cannot
// access each other), partial G's summed on the host (N * no2 doubles; 16 MB per GPU at 2002 functions).
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>
#include "../../include/unomol_b200.h"

namespace unomol {

class TwoElectronInts {
  public:
    TwoElectronInts() = delete;
    TwoElectronInts(const TwoElectronInts &) = delete;

    template <class BasisT>
    TwoElectronInts(const BasisT &basis, int start_shell, const std::string &base_str) : start(start_shell) {
        (void)base_str;
        const char *dev = getenv("UNOMOL_DEVICE");
        device = dev ? atoi(dev) : 0;
        const char *ng = getenv("UNOMOL_GPUS");
        ngpu = ng ? std::max(1, atoi(ng)) : 1;
        calculate(basis);
    }

    ~TwoElectronInts() {
        for (auto *x : extra) unomol_b200_destroy(x);
        unomol_b200_destroy(h);
    }

    template <class BasisT>
    void calculate(const BasisT &basis) {
        // flatten Basis/Shell/Center (reference Basis.hpp) into the C-ABI descriptor
        const int ns = basis.number_of_shells(), nc = basis.number_of_centers();
        std::vector<int> npr(ns), lv(ns), cen(ns), off(ns), poff(ns);
        std::vector<double> alpha, coef, xyz(3 * nc);
        for (int s = 0; s < ns; ++s) {
            const auto &sh = basis.shell_ptr()[s];
            npr[s] = sh.number_of_prims(); lv[s] = sh.Lvalue(); cen[s] = sh.center();
            off[s] = basis.offset(s); poff[s] = (int)alpha.size();
            for (int k = 0; k < npr[s]; ++k) { alpha.push_back(sh.alf(k)); coef.push_back(sh.cof(k)); }
        }
        for (int c = 0; c < nc; ++c)
            for (int x = 0; x < 3; ++x) xyz[3 * c + x] = basis.center_ptr()[c].position(x);
        if (h) {   // same basis, new geometry: recalculate()
            check(unomol_b200_set_geometry(h, xyz.data()), "set_geometry");
            for (auto *x : extra) check(unomol_b200_set_geometry(x, xyz.data()), "set_geometry");
            return;
        }
        unomol_basis_desc d;
        d.nshell = ns; d.nbf = basis.number_of_orbitals(); d.ncen = nc; d.maxl = basis.maxLvalue();
        d.npr = npr.data(); d.lv = lv.data(); d.cen = cen.data(); d.off = off.data(); d.poff = poff.data();
        d.alpha = alpha.data(); d.coef = coef.data(); d.xyz = xyz.data();
        check(unomol_b200_create(&d, start, device, 0, ngpu, &h), "create");
        for (int r = 1; r < ngpu; ++r) {
            unomol_b200_t *x = nullptr;
            check(unomol_b200_create(&d, start, device + r, r, ngpu, &x), "create (additional GPU)");
            extra.push_back(x);
        }
        bool shared = ngpu > 1 && !getenv("UNOMOL_NO_STEAL");
        for (auto *x : extra)
            if (shared && unomol_b200_steal_share(h, x) != UNOMOL_OK) shared = false;
        if (ngpu > 1 && !shared) {   // static snake-order split on every handle
            check(unomol_b200_set_option(h, "work_stealing", 0), "set_option");
            for (auto *x : extra) check(unomol_b200_set_option(x, "work_stealing", 0), "set_option");
        }
        // UNOMOL_RYS2_EXACT=1: the exact two-root Rys quadrature instead of the reference-compatible band 15 < X <= 40
        // (reference Rys.cpp:614-624, off by up to 8.7e-7 there); energies then follow the reference's -DUNOMOL_MD_INTS build
        const char *r2 = getenv("UNOMOL_RYS2_EXACT");
        if (r2 && atoi(r2)) {
            check(unomol_b200_set_option(h, "rys2_exact", 1), "set_option");
            for (auto *x : extra) check(unomol_b200_set_option(x, "rys2_exact", 1), "set_option");
        }
        // UNOMOL_EXACT=1: production mode without the reference's numerical defects -- the exact two-root quadrature AND the Rys
        // quadrature with 6..9 roots for l_tot > 8 (option all_rys) instead of the reference's McMurchie-Davidson routine, whose
        // Boys function switches to the bare asymptote at t = 20 (7e-11) and whose two-centre variant is off by up to 3e-7 on
        // (ff|ff) blocks (tests/test_highl_emulation.py).  Every quartet is then evaluated by a quadrature that is exact to rounding.
        const char *ex = getenv("UNOMOL_EXACT");
        if (ex && atoi(ex)) {
            for (const char *opt : {"rys2_exact", "all_rys"}) {
                check(unomol_b200_set_option(h, opt, 1), "set_option");
                for (auto *x : extra) check(unomol_b200_set_option(x, opt, 1), "set_option");
            }
        }
        const char *tau = getenv("UNOMOL_SCHWARZ_TAU");
        if (tau) {
            check(unomol_b200_set_option(h, "schwarz_tau", atof(tau)), "set_option");
            for (auto *x : extra) check(unomol_b200_set_option(x, "schwarz_tau", atof(tau)), "set_option");
        }
        no2 = (size_t)d.nbf * (d.nbf + 1) / 2;
        if (ngpu > 1) fprintf(stderr, "unomol_b200: %d GPUs, %s\n", ngpu, shared ? "shared work counters (work stealing over NVLink)" : "static split");
        unomol_b200_stats_t st;
        unomol_b200_stats(h, &st);
        fprintf(stderr, "unomol_b200: %lld shell pairs kept of %lld, %lld primitive pairs, set-up %.3f ms\n", st.n_pairs_kept,
                st.n_shell_pairs, st.n_prim_pairs, st.precompute_ms);
    }

    template <class BasisT>
    void recalculate(const BasisT &basis) { calculate(basis); }

    void formGmatrix(const double *Pmat, double *Gmat) {
        if (extra.empty()) { check(unomol_b200_fock_rhf(h, Pmat, Gmat), "fock_rhf"); return; }
        multi(Pmat, nullptr, Gmat, nullptr);
    }

    void formGmatrix(const double *PmatA, const double *PmatB, double *GmatA, double *GmatB) {
        if (extra.empty()) { check(unomol_b200_fock_uhf(h, PmatA, PmatB, GmatA, GmatB), "fock_uhf"); return; }
        multi(PmatA, PmatB, GmatA, GmatB);
    }

    template <class BasisT>
    void directFormGMatrix(const double *Pmat, double *Gmat, const BasisT &) { formGmatrix(Pmat, Gmat); }

    unomol_b200_t *handle() const noexcept { return h; }
    int number_of_gpus() const noexcept { return ngpu; }

  private:
    static void check(int rc, const char *what) {
        if (rc != UNOMOL_OK) {
            fprintf(stderr, "unomol_b200 %s: %s\n", what, unomol_b200_strerror(rc));
            exit(EXIT_FAILURE);
        }
    }
    // one host thread per additional GPU for the duration of the build; rank 0 runs on the calling thread and accumulates
    // straight into the caller's G, the other partial G's are added afterwards (the join is the barrier the shared work
    // counters need between builds)
    void multi(const double *PA, const double *PB, double *GA, double *GB) {
        const size_t nx = extra.size();
        part.assign(nx * no2 * (PB ? 2 : 1), 0.0);
        std::vector<int> rc(nx, UNOMOL_OK);
        std::vector<std::thread> pool;
        for (size_t r = 0; r < nx; ++r)
            pool.emplace_back([&, r]() {
                double *ga = part.data() + r * no2 * (PB ? 2 : 1);
                rc[r] = PB ? unomol_b200_fock_uhf(extra[r], PA, PB, ga, ga + no2) : unomol_b200_fock_rhf(extra[r], PA, ga);
            });
        const int rc0 = PB ? unomol_b200_fock_uhf(h, PA, PB, GA, GB) : unomol_b200_fock_rhf(h, PA, GA);
        for (auto &t : pool) t.join();
        check(rc0, "fock (GPU 0)");
        for (size_t r = 0; r < nx; ++r) {
            check(rc[r], "fock (additional GPU)");
            const double *ga = part.data() + r * no2 * (PB ? 2 : 1);
            for (size_t i = 0; i < no2; ++i) GA[i] += ga[i];
            if (PB) for (size_t i = 0; i < no2; ++i) GB[i] += ga[no2 + i];
        }
    }
    unomol_b200_t *h = nullptr;
    std::vector<unomol_b200_t *> extra;
    std::vector<double> part;
    size_t no2 = 0;
    int start = 0, device = 0, ngpu = 1;
};

}  // namespace unomol
