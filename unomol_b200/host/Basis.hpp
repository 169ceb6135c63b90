// unomol_b200/host/Basis.hpp -- host-side data model with the reference's names and accessors
// (reference Basis.hpp:17-356: Shell, Center, Basis).  Clean-room: flat std::vector storage, same file
// format (Basis.hpp:181-255), same contraction normalisation (Basis.hpp:56-75), same eps floor (:249-254).
// int_flag[0] == 1: the extra shells of posin.bas on one extra centre (charge +1, the positron of the polarisation scan) are read
// too and become active with dpm_augment(), exactly as in the reference (Basis.hpp:208-243, 341-345).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

namespace unomol {

[[noreturn]] inline void fatal_error(const char *message) {   // reference Util.cpp:5-8
    fprintf(stderr, "%s\n", message);
    exit(EXIT_FAILURE);
}

class Shell {
    int npr = 0, lsh = 0, cen = 0;
    std::vector<double> al, co;

  public:
    double alf(int i) const noexcept { return al[i]; }
    double cof(int i) const noexcept { return co[i]; }
    const double *alf_ptr() const noexcept { return al.data(); }
    const double *cof_ptr() const noexcept { return co.data(); }
    int number_of_prims() const noexcept { return npr; }
    int Lvalue() const noexcept { return lsh; }
    int center() const noexcept { return cen; }
    void setCenter(int icen) { cen = icen; }

    void normalize() noexcept {   // reference Basis.hpp:56-75 (double sum over the raw coefficients)
        const double twofact = 2.8284271247461903, piterm = 5.568327996831707;
        const double lpow = 1.5 + lsh;
        double sum = 0.0;
        for (int i = 0; i < npr; i++)
            for (int j = 0; j < npr; j++) sum += co[i] * co[j] * pow(sqrt(al[i] * al[j]) / (al[i] + al[j]), lpow);
        sum *= twofact;
        sum = 1.0 / sqrt(sum);
        for (int i = 0; i < npr; i++) co[i] = co[i] * sum * sqrt(pow(2 * al[i], lpow) / piterm);
    }

    std::istream &read_shell(std::istream &is, int cen_in = -1) {   // reference Basis.hpp:77-96
        is >> npr >> lsh;
        if (cen_in == -1) is >> cen;
        else cen = cen_in;
        al.assign(npr, 0.0);
        co.assign(npr, 0.0);
        for (int i = 0; i < npr; ++i) is >> al[i] >> co[i];
        return is;
    }
};

class Center {
    double chg = 0.0, r[3] = {0.0, 0.0, 0.0};

  public:
    double charge() const noexcept { return chg; }
    const double *r_vec() const noexcept { return r; }
    double position(int i) const noexcept { return r[i]; }
    void setPosition(double x, double y, double z) noexcept { r[0] = x; r[1] = y; r[2] = z; }
    void setCharge(double q) noexcept { chg = q; }
};

class Basis {
    std::vector<Shell> shells;
    std::vector<Center> centers;
    std::vector<int> offsets;
    double eps = 0.0;
    int nshell = 0, ncen = 0, norb = 0, maxl = 0, nelec = 0, maxits = 0;
    int tnshell = 0, tncen = 0, tnorb = 0, skipcen = 0;
    int scf_flag[3] = {0, 0, 0}, int_flag[3] = {0, 0, 0}, prt_flag[4] = {0, 0, 0, 0};

  public:
    explicit Basis(const std::string &infile = std::string("patin.dat")) {
        std::ifstream in(infile.c_str());
        if (!in) fatal_error(("could not open file " + infile).c_str());
        in >> nshell >> norb >> ncen >> maxl >> nelec >> maxits >> eps;
        in >> int_flag[0] >> int_flag[1] >> scf_flag[0] >> scf_flag[1] >> scf_flag[2];
        in >> prt_flag[0] >> prt_flag[1] >> prt_flag[2];
        if (!in) fatal_error("malformed patin.dat header");
        tnshell = nshell; tncen = ncen; tnorb = norb;
        std::ifstream ain;
        if (int_flag[0] == 1) {   // reference Basis.hpp:208-222
            ain.open("posin.bas");
            if (!ain) fatal_error("could not open posin.bas");
            int xsh = 0, xno = 0, xmaxl = 0;
            ain >> xsh >> xno >> xmaxl;
            tnshell += xsh; tnorb += xno; ++tncen;
            if (xmaxl > maxl) maxl = xmaxl;
        }
        if (maxl > 4) fatal_error("Angular Momentum is too large for present program\n");
        centers.resize(tncen);
        shells.resize(tnshell);
        for (int i = 0; i < ncen; ++i) {
            double q, x, y, z;
            in >> q >> x >> y >> z;
            centers[i].setCharge(q);
            centers[i].setPosition(x, y, z);
        }
        int off = 0;
        for (int ish = 0; ish < nshell; ++ish) {
            shells[ish].read_shell(in);
            const int lv = shells[ish].Lvalue();
            offsets.push_back(off);
            off += (lv + 1) * (lv + 2) / 2;
        }
        if (!in) fatal_error("malformed patin.dat body");
        if (int_flag[0] == 1) {   // reference Basis.hpp:235-243
            centers[ncen].setCharge(1.0);
            centers[ncen].setPosition(0.0, 0.0, 0.0);
            for (int ish = nshell; ish < tnshell; ++ish) {
                shells[ish].read_shell(ain, ncen);
                const int lv = shells[ish].Lvalue();
                offsets.push_back(off);
                off += (lv + 1) * (lv + 2) / 2;
            }
            if (!ain) fatal_error("malformed posin.bas");
        }
        skipcen = ncen;
        for (auto &s : shells) s.normalize();
        const double xeps = DBL_EPSILON * norb * norb * 0.5;
        if (eps < xeps) {
            fprintf(stderr, "SCF convergence of %15.6le is too low!\n", eps);
            eps = xeps;
            fprintf(stderr, "SCF Convergence set at %15.6le\n", eps);
        }
    }

    int offset(int ish) const noexcept { return offsets[ish]; }
    int maxLvalue() const noexcept { return maxl; }
    int number_of_electrons() const noexcept { return nelec; }
    int number_of_shells() const noexcept { return nshell; }
    int number_of_centers() const noexcept { return ncen; }
    int number_of_orbitals() const noexcept { return norb; }
    int maximum_iterations() const noexcept { return maxits; }
    int scf_flags(int i) const noexcept { return scf_flag[i]; }
    int prt_flags(int i) const noexcept { return prt_flag[i]; }
    int int_flags(int i) const noexcept { return int_flag[i]; }
    const Shell *shell_ptr() const noexcept { return shells.data(); }
    const Center *center_ptr() const noexcept { return centers.data(); }
    double scf_eps() const noexcept { return eps; }
    void SetCenterPosition(double x, double y, double z, int n) noexcept { centers[n].setPosition(x, y, z); }
    int total_number_of_shells() const noexcept { return tnshell; }
    int total_number_of_centers() const noexcept { return tncen; }
    int total_number_of_orbitals() const noexcept { return tnorb; }
    int skip_center() const noexcept { return skipcen; }
    constexpr double FiniteFieldValue() const noexcept { return 5.e-3; }   // reference Basis.hpp:233,349-351
    void dpm_augment() noexcept {   // reference Basis.hpp:341-345
        nshell = tnshell;
        norb = tnorb;
        ++ncen;
    }
};

}  // namespace unomol
