// unomol_b200/host/OneElectron.hpp -- overlap, kinetic and core-Hamiltonian matrices on the host (O(N^2), <1 % of
// the run; SURVEY.md section 2 #17: stays on the CPU).  Same call surface as the reference's
// OneElectronInts(bas, S, T, H) (reference OneElectronInts.cpp:127-202): packed lower-triangular S, T and
// H = T + V with V = -sum_C Z_C <a|1/r_C|b>.  Clean-room McMurchie-Davidson implementation (Hermite E
// coefficients + R_tuv Coulomb tensor); the Boys function follows the reference's Fgamma switch at t = 20
// (reference MD_Rfunction.hpp:2184-2206) so that H agrees with the reference's to rounding level.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace unomol {

namespace onee {

inline void boys(double *fm, double t, int m) {
    const double half_sqrt_pi = 0.88622692545275801365;
    if (t > 20.0) {   // asymptotic branch, as the reference
        fm[0] = half_sqrt_pi / std::sqrt(t);
        for (int i = 1; i <= m; ++i) fm[i] = fm[i - 1] * (i - 0.5) / t;
        return;
    }
    const double mph = m + 0.5;
    double term = 0.5 / mph, sum = term;
    for (int i = 1; i <= 400; ++i) {
        term *= t / (mph + i);
        sum += term;
        if (term < 1e-17 * sum) break;
    }
    const double et = std::exp(-t);
    fm[m] = sum * et;
    for (int i = m - 1; i >= 0; --i) fm[i] = (fm[i + 1] * 2.0 * t + et) / (2 * i + 1.0);
}

constexpr int LM = 4 + 2;   // max l per shell (+2 for the kinetic-energy shift on the second index)

// Hermite expansion coefficients E[i][j][t] for one axis (E_0^{00} = 1; the Gaussian prefactor is applied outside)
struct ECoef {
    double e[LM + 1][LM + 1][2 * LM + 2];
    void build(int la, int lb, double pa, double pb, double inv2p) {
        std::memset(e, 0, sizeof(e));
        e[0][0][0] = 1.0;
        for (int i = 0; i <= la; ++i)
            for (int j = 0; j <= lb; ++j) {
                if (i == 0 && j == 0) continue;
                for (int t = 0; t <= i + j; ++t) {
                    double v;
                    if (i > 0) {
                        v = pa * e[i - 1][j][t] + (t + 1) * e[i - 1][j][t + 1];
                        if (t > 0) v += inv2p * e[i - 1][j][t - 1];
                    } else {
                        v = pb * e[i][j - 1][t] + (t + 1) * e[i][j - 1][t + 1];
                        if (t > 0) v += inv2p * e[i][j - 1][t - 1];
                    }
                    e[i][j][t] = v;
                }
            }
    }
};

}  // namespace onee

template <class BasisT>
void OneElectronInts(const BasisT &bas, double *Smat, double *Tmat, double *Hmat) {
    using namespace onee;
    const int ns = bas.number_of_shells(), ncen = bas.number_of_centers();
    auto ncart = [](int l) { return (l + 1) * (l + 2) / 2; };
    auto comp = [](int l, int c, int *lmn) {
        int k = 0;
        for (int lx = l; lx >= 0; --lx)
            for (int ly = l - lx; ly >= 0; --ly, ++k)
                if (k == c) { lmn[0] = lx; lmn[1] = ly; lmn[2] = l - lx - ly; return; }
    };
    // per-component norms with the reference's own recurrence (AuxFunctions.hpp:49-64)
    double df[8];
    df[0] = 1.0;
    { double dx = 1.0; for (int i = 1; i < 8; ++i) { df[i] = df[i - 1] * dx; dx *= (2 * i + 1); } }
    auto cnorm = [&](const int *lmn) { return 1.0 / std::sqrt(df[lmn[0]] * df[lmn[1]] * df[lmn[2]]); };
    const int RD = 2 * 4 + 1;
    // The nuclear-attraction part is O(N^2 * ncen): 0.8 s at 416 functions, ~90 s at 2002 on one core.  Rows of the shell-pair
    // triangle are independent, so they are dealt to host threads (UNOMOL_HOST_THREADS, default = hardware threads, <= 32),
    // largest rows first; every (i, j) element is written by exactly one thread and the arithmetic per element does not
    // depend on the thread count.  Primitive pairs whose Gaussian-product prefactor |c_a c_b| exp(-ab|AB|^2/p) is below
    // 1e-20 are skipped (the reference keeps them; their contribution to any matrix element is < 1e-17).
    auto do_row = [&](int ish, ECoef &ex, ECoef &ey, ECoef &ez, std::vector<double> &R) {
    auto Rat = [&](int n, int t, int u, int v) -> double & { return R[(((size_t)n * RD + t) * RD + u) * RD + v]; };
    {
        const auto &A = bas.shell_ptr()[ish];
        const double *ra = bas.center_ptr()[A.center()].r_vec();
        const int la = A.Lvalue();
        for (int jsh = 0; jsh <= ish; ++jsh) {
            const auto &B = bas.shell_ptr()[jsh];
            const double *rb = bas.center_ptr()[B.center()].r_vec();
            const int lb = B.Lvalue(), L = la + lb;
            const double ab2 = (ra[0] - rb[0]) * (ra[0] - rb[0]) + (ra[1] - rb[1]) * (ra[1] - rb[1]) + (ra[2] - rb[2]) * (ra[2] - rb[2]);
            const int na = ncart(la), nb = ncart(lb);
            std::vector<double> sv(na * nb, 0.0), tv(na * nb, 0.0), vv(na * nb, 0.0);
            for (int ip = 0; ip < A.number_of_prims(); ++ip)
                for (int jp = 0; jp < B.number_of_prims(); ++jp) {
                    const double a = A.alf(ip), b = B.alf(jp), p = a + b, ip2 = 0.5 / p;
                    const double c12 = A.cof(ip) * B.cof(jp);
                    const double kab = std::exp(-a * b / p * ab2);
                    if (std::fabs(c12) * kab < 1e-20) continue;
                    double P[3];
                    for (int x = 0; x < 3; ++x) P[x] = (a * ra[x] + b * rb[x]) / p;
                    ex.build(la, lb + 2, P[0] - ra[0], P[0] - rb[0], ip2);
                    ey.build(la, lb + 2, P[1] - ra[1], P[1] - rb[1], ip2);
                    ez.build(la, lb + 2, P[2] - ra[2], P[2] - rb[2], ip2);
                    const double s3 = std::pow(M_PI / p, 1.5) * kab;   // 3-D overlap prefactor
                    // Coulomb tensor summed over nuclei: Rsum[t][u][v] = sum_C -Z_C R^0_tuv(p, P-C)
                    std::vector<double> Rsum((size_t)RD * RD * RD, 0.0);
                    for (int ic = 0; ic < ncen; ++ic) {
                        const double Z = bas.center_ptr()[ic].charge();
                        const double *rc = bas.center_ptr()[ic].r_vec();
                        const double pc[3] = {P[0] - rc[0], P[1] - rc[1], P[2] - rc[2]};
                        double fm[RD + 2];
                        boys(fm, p * (pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2]), L);
                        double m2p = 1.0;
                        for (int n = 0; n <= L; ++n) { Rat(n, 0, 0, 0) = m2p * fm[n]; m2p *= -2.0 * p; }
                        for (int tot = 1; tot <= L; ++tot)
                            for (int t = 0; t <= tot; ++t)
                                for (int u = 0; u <= tot - t; ++u) {
                                    const int v = tot - t - u;
                                    for (int n = 0; n <= L - tot; ++n) {
                                        double val;
                                        if (t > 0) val = (t > 1 ? (t - 1) * Rat(n + 1, t - 2, u, v) : 0.0) + pc[0] * Rat(n + 1, t - 1, u, v);
                                        else if (u > 0) val = (u > 1 ? (u - 1) * Rat(n + 1, t, u - 2, v) : 0.0) + pc[1] * Rat(n + 1, t, u - 1, v);
                                        else val = (v > 1 ? (v - 1) * Rat(n + 1, t, u, v - 2) : 0.0) + pc[2] * Rat(n + 1, t, u, v - 1);
                                        Rat(n, t, u, v) = val;
                                    }
                                }
                        for (int t = 0; t <= L; ++t)
                            for (int u = 0; u <= L - t; ++u)
                                for (int v = 0; v <= L - t - u; ++v) Rsum[((size_t)t * RD + u) * RD + v] -= Z * Rat(0, t, u, v);
                    }
                    const double vpref = 2.0 * M_PI / p * kab;
                    for (int ia = 0; ia < na; ++ia) {
                        int l1[3];
                        comp(la, ia, l1);
                        for (int ib = 0; ib < nb; ++ib) {
                            int l2[3];
                            comp(lb, ib, l2);
                            const double nf = cnorm(l1) * cnorm(l2) * c12;
                            const double sx = ex.e[l1[0]][l2[0]][0], sy = ey.e[l1[1]][l2[1]][0], sz = ez.e[l1[2]][l2[2]][0];
                            sv[ia * nb + ib] += nf * s3 * sx * sy * sz;
                            // kinetic: -1/2 d^2/dx^2 acting on the second function, per axis
                            auto kin = [&](const ECoef &E, int i, int j) {
                                double v = -2.0 * b * b * E.e[i][j + 2][0] + b * (2 * j + 1) * E.e[i][j][0];
                                if (j >= 2) v -= 0.5 * j * (j - 1) * E.e[i][j - 2][0];
                                return v;
                            };
                            const double tx = kin(ex, l1[0], l2[0]) * sy * sz, ty = sx * kin(ey, l1[1], l2[1]) * sz,
                                         tz = sx * sy * kin(ez, l1[2], l2[2]);
                            tv[ia * nb + ib] += nf * s3 * (tx + ty + tz);
                            double sum = 0.0;
                            for (int t = 0; t <= l1[0] + l2[0]; ++t)
                                for (int u = 0; u <= l1[1] + l2[1]; ++u)
                                    for (int v = 0; v <= l1[2] + l2[2]; ++v)
                                        sum += ex.e[l1[0]][l2[0]][t] * ey.e[l1[1]][l2[1]][u] * ez.e[l1[2]][l2[2]][v] *
                                               Rsum[((size_t)t * RD + u) * RD + v];
                            vv[ia * nb + ib] += nf * vpref * sum;
                        }
                    }
                }
            for (int ia = 0; ia < na; ++ia) {
                const int ir = bas.offset(ish) + ia;
                for (int ib = 0; ib < nb; ++ib) {
                    const int jr = bas.offset(jsh) + ib;
                    if (jr > ir) continue;
                    const size_t ij = (size_t)ir * (ir + 1) / 2 + jr;
                    Smat[ij] = sv[ia * nb + ib];
                    Tmat[ij] = tv[ia * nb + ib];
                    Hmat[ij] = tv[ia * nb + ib] + vv[ia * nb + ib];
                }
            }
        }
    }
    };
    int nthr = (int)std::thread::hardware_concurrency();
    if (const char *e = std::getenv("UNOMOL_HOST_THREADS")) nthr = std::atoi(e);
    nthr = std::max(1, std::min(std::min(nthr, 32), ns / 8));
    std::atomic<int> next(0);
    auto worker = [&]() {
        std::vector<ECoef> E(3);
        std::vector<double> R((size_t)(RD + 1) * RD * RD * RD);
        for (int k = next.fetch_add(1); k < ns; k = next.fetch_add(1)) do_row(ns - 1 - k, E[0], E[1], E[2], R);
    };
    if (nthr <= 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthr; ++t) pool.emplace_back(worker);
        for (auto &t : pool) t.join();
    }
}

}  // namespace unomol
