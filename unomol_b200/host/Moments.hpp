// unomol_b200/host/Moments.hpp -- dipole / quadrupole moment analysis on the host (O(N^2) post-processing of the
// converged density; SURVEY.md section 2 #18: stays on the CPU).  It exists because the parity contract names
// moments.dat next to scfout.dat: the drivers write `moments.out` in the layout of the reference's AnalyzeMoments
// (reference Moments.cpp:189-273 RHF, :276-363 UHF) and `mol_dipmom.out` in that of AnalyzeMOMoments (:365-404).
// The integrals <a| x^k y^l z^m |b>, k+l+m <= 2, come from Hermite E coefficients (same quantities as the reference's
// calc_moments, Moments.cpp:5-92): per axis  <1> = E0,  <x> = E1 + Px E0,  <x^2> = 2 E2 + 2 Px E1 + (Px^2 + 1/2p) E0.
// The reference streams the integrals through MOMINTS.DAT / RMOM.DAT; here they stay in memory.
#pragma once
#include <cmath>
#include <cstdio>
#include <vector>
#include "OneElectron.hpp"

namespace unomol {

struct MomentMatrices {
    // packed lower-triangular, idx = i(i+1)/2 + j; order dx dy dz qxx qxy qxz qyy qyz qzz (reference Moments.hpp MomInts)
    std::vector<double> m[9];
};

template <class BasisT>
void MomentInts(const BasisT &bas, MomentMatrices &M) {
    using namespace onee;
    const int ns = bas.number_of_shells(), no = bas.number_of_orbitals();
    for (auto &v : M.m) v.assign((size_t)no * (no + 1) / 2, 0.0);
    auto ncart = [](int l) { return (l + 1) * (l + 2) / 2; };
    auto comp = [](int l, int c, int *lmn) {
        int k = 0;
        for (int lx = l; lx >= 0; --lx)
            for (int ly = l - lx; ly >= 0; --ly, ++k)
                if (k == c) { lmn[0] = lx; lmn[1] = ly; lmn[2] = l - lx - ly; return; }
    };
    double df[8];
    df[0] = 1.0;
    { double dx = 1.0; for (int i = 1; i < 8; ++i) { df[i] = df[i - 1] * dx; dx *= (2 * i + 1); } }   // AuxFunctions.hpp:49-64
    auto cnorm = [&](const int *lmn) { return 1.0 / std::sqrt(df[lmn[0]] * df[lmn[1]] * df[lmn[2]]); };
    static ECoef E[3];
    for (int ish = 0; ish < ns; ++ish) {
        const auto &A = bas.shell_ptr()[ish];
        const double *ra = bas.center_ptr()[A.center()].r_vec();
        const int la = A.Lvalue(), na = ncart(la);
        for (int jsh = 0; jsh <= ish; ++jsh) {
            const auto &B = bas.shell_ptr()[jsh];
            const double *rb = bas.center_ptr()[B.center()].r_vec();
            const int lb = B.Lvalue(), nb = ncart(lb);
            double ab2 = 0.0;
            for (int x = 0; x < 3; ++x) ab2 += (ra[x] - rb[x]) * (ra[x] - rb[x]);
            std::vector<double> val(9 * (size_t)na * nb, 0.0);
            for (int ip = 0; ip < A.number_of_prims(); ++ip)
                for (int jp = 0; jp < B.number_of_prims(); ++jp) {
                    const double a = A.alf(ip), b = B.alf(jp), p = a + b, ip2 = 0.5 / p;
                    const double s3 = A.cof(ip) * B.cof(jp) * std::pow(M_PI / p, 1.5) * std::exp(-a * b / p * ab2);
                    double P[3];
                    for (int x = 0; x < 3; ++x) {
                        P[x] = (a * ra[x] + b * rb[x]) / p;
                        E[x].build(la, lb, P[x] - ra[x], P[x] - rb[x], ip2);
                    }
                    for (int ia = 0; ia < na; ++ia) {
                        int l1[3];
                        comp(la, ia, l1);
                        for (int ib = 0; ib < nb; ++ib) {
                            int l2[3];
                            comp(lb, ib, l2);
                            const double nf = cnorm(l1) * cnorm(l2) * s3;
                            double m0[3], m1[3], m2[3];   // per axis: <1>, <x>, <x^2>
                            for (int x = 0; x < 3; ++x) {
                                const double *e = E[x].e[l1[x]][l2[x]];
                                m0[x] = e[0];
                                m1[x] = e[1] + P[x] * e[0];
                                m2[x] = 2.0 * e[2] + 2.0 * P[x] * e[1] + (P[x] * P[x] + ip2) * e[0];
                            }
                            double *v = &val[9 * ((size_t)ia * nb + ib)];
                            v[0] += nf * m1[0] * m0[1] * m0[2];
                            v[1] += nf * m0[0] * m1[1] * m0[2];
                            v[2] += nf * m0[0] * m0[1] * m1[2];
                            v[3] += nf * m2[0] * m0[1] * m0[2];
                            v[4] += nf * m1[0] * m1[1] * m0[2];
                            v[5] += nf * m1[0] * m0[1] * m1[2];
                            v[6] += nf * m0[0] * m2[1] * m0[2];
                            v[7] += nf * m0[0] * m1[1] * m1[2];
                            v[8] += nf * m0[0] * m0[1] * m2[2];
                        }
                    }
                }
            for (int ia = 0; ia < na; ++ia) {
                const int ir = bas.offset(ish) + ia;
                for (int ib = 0; ib < nb; ++ib) {
                    const int jr = bas.offset(jsh) + ib;
                    if (jr > ir) continue;
                    for (int k = 0; k < 9; ++k) M.m[k][(size_t)ir * (ir + 1) / 2 + jr] = val[9 * ((size_t)ia * nb + ib) + k];
                }
            }
        }
    }
}

// moments.out (reference Moments.cpp:189-273; UHF :276-363 uses P = (PA + PB)/2).  P = C_occ C_occ^T without the factor 2,
// so the electronic moment is sum_{i>j} 4 P_ij M_ij + sum_i 2 P_ii M_ii (the reference's `factors`, Moments.cpp:151-154).
template <class CenterT>
void AnalyzeMoments(const MomentMatrices &M, const double *PA, const double *PB, const CenterT *center, int ncen, int no) {
    double nM[9] = {0}, eM[9] = {0}, tM[9];
    for (int ic = 0; ic < ncen; ++ic) {
        const double q = center[ic].charge();
        const double *r = center[ic].r_vec();
        nM[0] += q * r[0]; nM[1] += q * r[1]; nM[2] += q * r[2];
        nM[3] += q * r[0] * r[0]; nM[4] += q * r[0] * r[1]; nM[5] += q * r[0] * r[2];
        nM[6] += q * r[1] * r[1]; nM[7] += q * r[1] * r[2]; nM[8] += q * r[2] * r[2];
    }
    size_t ij = 0;
    for (int i = 0; i < no; ++i)
        for (int j = 0; j <= i; ++j, ++ij) {
            const double pij = (PB ? 0.5 * (PA[ij] + PB[ij]) : PA[ij]) * (i == j ? 2.0 : 4.0);
            for (int k = 0; k < 9; ++k) eM[k] += pij * M.m[k][ij];
        }
    for (int k = 0; k < 9; ++k) tM[k] = nM[k] - eM[k];
    const double q00 = tM[8] - 0.5 * (tM[3] + tM[6]);
    FILE *out = fopen("moments.out", "w");
    if (!out) return;
    fprintf(out, " MULTIPOLE MOMENT ANALYSIS \n units in bohr - hartree atomic units \n\n DIPOLE MOMENTS \n\n");
    fprintf(out, "          Total         Electronic        Nuclear \n");
    const char *dn[3] = {"x", "y", "z"}, *qn[6] = {"xx", "xy", "xz", "yy", "yz", "zz"};
    for (int k = 0; k < 3; ++k) fprintf(out, " %s  %15.7le  %15.7le  %15.7le \n", dn[k], tM[k], eM[k], nM[k]);
    fprintf(out, "\n QUADRUPOLE MOMENTS \n\n          Total         Electronic        Nuclear \n");
    for (int k = 0; k < 6; ++k) fprintf(out, " %s  %15.7le  %15.7le  %15.7le \n", qn[k], tM[3 + k], eM[3 + k], nM[3 + k]);
    fprintf(out, "\n Dipole moment     = %25.15le \n Quadrupole moment = %25.15le \n", tM[2], q00);
    fclose(out);
}

// mol_dipmom.out (reference Moments.cpp:365-404): C^T D C for the three dipole matrices, lower triangle over MO pairs.
// C row-major, eigenvectors in columns (RHF.hpp:178-190).
inline void AnalyzeMOMoments(const MomentMatrices &M, const double *C, int no, FILE *out, const char *title) {
    std::vector<double> full((size_t)no * no), tmp((size_t)no * no), mo[3];
    for (int k = 0; k < 3; ++k) {
        size_t ij = 0;
        for (int i = 0; i < no; ++i)
            for (int j = 0; j <= i; ++j, ++ij) full[(size_t)i * no + j] = full[(size_t)j * no + i] = M.m[k][ij];
        for (int i = 0; i < no; ++i)          // tmp = D C
            for (int q = 0; q < no; ++q) {
                double s = 0.0;
                for (int j = 0; j < no; ++j) s += full[(size_t)i * no + j] * C[(size_t)j * no + q];
                tmp[(size_t)i * no + q] = s;
            }
        mo[k].assign((size_t)no * (no + 1) / 2, 0.0);
        ij = 0;
        for (int p = 0; p < no; ++p)
            for (int q = 0; q <= p; ++q, ++ij) {
                double s = 0.0;
                for (int i = 0; i < no; ++i) s += C[(size_t)i * no + p] * tmp[(size_t)i * no + q];
                mo[k][ij] = s;
            }
    }
    fprintf(out, " %s \n\n orbital 1       orbital2      dx-dy-dz\n", title);
    size_t ij = 0;
    for (int i = 0; i < no; ++i)
        for (int j = 0; j <= i; ++j, ++ij) fprintf(out, " %12d %12d %15.6le %15.6le %15.6le\n", i, j, mo[0][ij], mo[1][ij], mo[2][ij]);
}

}  // namespace unomol
