// unomol_b200/host/SCF.hpp -- RestrictedHartreeFock / UnRestrictedHartreeFock drivers with the reference's call
// surface and iteration logic (reference RHF.hpp:22-685, UHF.hpp:23-790), on top of the GPU two-electron engine.
//
// Mirrors, member for member where it matters for parity:
//   update()         zero G -> tints.formGmatrix -> energy -> F = H + G -> diagonalise -> C -> P -> |dP|
//                    (RHF.hpp:87-112, UHF.hpp:101-134)
//   findEnergy()     nuclear repulsion, one-electron matrices, X = S^-1/2, core guess, two plain updates, then
//                    scf_converger(); update(); is_converged()   (RHF.hpp:114-176, UHF.hpp:137-177)
//   scf_converger()  shift history if the last dE < 0, else P <- (P + Pold)/2   (RHF.hpp:536-569; the
//                    extrapolation branch is dead in the serial reference: extrap is never set)
//   is_converged()   signed tests on ediff / pdiff per cflag   (RHF.hpp:390-402)
//   final_output()   short.gs.out (E_iter0, E_final, dE) and scfout.gs.out energies + orbital energies
//                    (RHF.hpp:404-459); PMATRIX.DAT checkpoint (RHF.hpp:172-174)
// What changes: the packed EISPACK-style eigensolver / transforms (SymmPack.cpp:272-348) are replaced by
// cuSOLVER / cuBLAS on the device through unomol_b200_scf_set_overlap / unomol_b200_scf_diag (no CPU fallback).
//                    moments.out / mol_dipmom.out through host/Moments.hpp (reference RHF.hpp:457-458, UHF.hpp:483);
//                    the one-electron and moment INTEGRALS come from the device (unomol_b200_one_electron)
// The polarisation-potential scan (RHF.hpp:292-388, UHF.hpp:293-383) is findPolarizationPotential of both classes below, the
// finite-field analysis (RHF.hpp:235-271, UHF.hpp:239-272) FiniteFieldAnalysis.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <vector>
#include "Basis.hpp"
#include "OneElectron.hpp"
#include "Moments.hpp"
#include "TwoElectronInts.hpp"

namespace unomol {

// One-electron and moment matrices: on the device through the C ABI (unomol_b200_one_electron, csrc/onee_device.cu) unless
// UNOMOL_HOST_ONEE=1 asks for the threaded host versions (host/OneElectron.hpp, host/Moments.hpp), which are the cross-check.
template <class BasisT>
inline void OneElectronIntsAuto(const BasisT &basis, TwoElectronInts &t, double *S, double *T, double *H) {
    if (std::getenv("UNOMOL_HOST_ONEE")) { OneElectronInts(basis, S, T, H); return; }
    std::vector<double> z(basis.number_of_centers());
    for (int c = 0; c < basis.number_of_centers(); ++c) z[c] = basis.center_ptr()[c].charge();
    const int rc = unomol_b200_one_electron(t.handle(), z.data(), S, T, H, nullptr);
    if (rc != UNOMOL_OK) {
        fprintf(stderr, "unomol_b200 one_electron: %s\n", unomol_b200_strerror(rc));
        exit(EXIT_FAILURE);
    }
}
template <class BasisT>
inline void MomentIntsAuto(const BasisT &basis, TwoElectronInts &t, MomentMatrices &M) {
    if (std::getenv("UNOMOL_HOST_ONEE")) { MomentInts(basis, M); return; }
    const size_t no = basis.number_of_orbitals(), no2 = no * (no + 1) / 2;
    std::vector<double> z(basis.number_of_centers()), S(no2), T(no2), H(no2), buf(9 * no2);
    for (int c = 0; c < basis.number_of_centers(); ++c) z[c] = basis.center_ptr()[c].charge();
    const int rc = unomol_b200_one_electron(t.handle(), z.data(), S.data(), T.data(), H.data(), buf.data());
    if (rc != UNOMOL_OK) {
        fprintf(stderr, "unomol_b200 one_electron (moments): %s\n", unomol_b200_strerror(rc));
        exit(EXIT_FAILURE);
    }
    for (int m = 0; m < 9; ++m) M.m[m].assign(buf.begin() + m * no2, buf.begin() + (m + 1) * no2);
}

namespace SymmPack {
inline double TraceSymmPackProduct(const double *a, const double *b, int n) noexcept {   // SymmPack.cpp:7-18
    double sum = 0.0;
    int ij = 0;
    for (int i = 0; i < n; ++ij, ++i) {
        for (int j = 0; j < i; ++ij, ++j) sum += a[ij] * b[ij];
        sum += 0.5 * a[ij] * b[ij];
    }
    return sum + sum;
}
inline double SymmPackDiffNorm(const double *a, const double *b, int n) noexcept {   // SymmPack.cpp:20-36
    double sum = 0.0;
    int ij = 0;
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < i; ++ij, ++j) { const double t = a[ij] - b[ij]; sum += t * t; }
        const double t = a[ij] - b[ij];
        sum += 0.5 * t * t;
        ++ij;
    }
    sum = sum + sum;
    return std::sqrt(sum) / n;
}
}  // namespace SymmPack

inline double nuclear_repulsion_energy(int ncen, const Center *center) {   // RHF.hpp:273-290
    double sum = 0.0;
    for (int i = 0; i < ncen; ++i)
        for (int j = i + 1; j < ncen; ++j) {
            const double *r1 = center[i].r_vec(), *r2 = center[j].r_vec();
            const double r12 = std::sqrt((r1[0] - r2[0]) * (r1[0] - r2[0]) + (r1[1] - r2[1]) * (r1[1] - r2[1]) + (r1[2] - r2[2]) * (r1[2] - r2[2]));
            sum += center[i].charge() * center[j].charge() / r12;
        }
    return sum;
}

inline void scf_check(int rc, const char *what) {
    if (rc != UNOMOL_OK) {
        fprintf(stderr, "unomol_b200 %s: %s\n", what, unomol_b200_strerror(rc));
        exit(EXIT_FAILURE);
    }
}

class RestrictedHartreeFock {
  public:
    RestrictedHartreeFock() = delete;
    RestrictedHartreeFock(Basis *b, TwoElectronInts *t) : basis(*b), tints(*t) {
        no = basis.number_of_orbitals();
        no2 = no * (no + 1) / 2;
        ncen = basis.number_of_centers();
        nocc = basis.number_of_electrons();
        if (nocc % 2) fatal_error("Odd number of electrons in RHF");
        nocc /= 2;
        maxits = basis.maximum_iterations();
        eps = basis.scf_eps();
        scf_accel = basis.scf_flags(1);
        cflag = basis.scf_flags(0);
        for (auto *v : {&Pold2, &Pold, &Pmat, &Gmat, &Hmat, &Fock, &Tmat, &Smat}) v->assign(no2, 0.0);
        Evals.assign(no, 0.0);
        Cmat.assign((size_t)no * no, 0.0);
    }

    void update() noexcept {
        if (on_device) {
            // device-resident iteration: P, G, F, H stay on the GPU; mixing (scf_converger) is done there too
            scf_check(unomol_b200_scf_iterate_rhf(tints.handle(), nocc, mix_next ? 1 : 0, &energy, &pdiff), "scf_iterate_rhf");
            mix_next = false;
            ediff = energy - eold;
            eold = energy;
            ++iteration;
            return;
        }
        for (int i = 0; i < no2; ++i) Gmat[i] = 0.0;
        // UNOMOL_DIRECT_G=1: through the reference's directFormGMatrix entry (TwoElectronInts.hpp:106; uncalled in the reference)
        if (std::getenv("UNOMOL_DIRECT_G")) tints.directFormGMatrix(Pmat.data(), Gmat.data(), basis);
        else tints.formGmatrix(Pmat.data(), Gmat.data());
        const double e1 = SymmPack::TraceSymmPackProduct(Pmat.data(), Hmat.data(), no) * 2.0;
        const double e2 = SymmPack::TraceSymmPackProduct(Pmat.data(), Gmat.data(), no);
        energy = e1 + e2;
        ediff = energy - eold;
        eold = energy;
        for (int i = 0; i < no2; ++i) Fock[i] = Hmat[i] + Gmat[i];
        if (scf_accel == 1) Pold2 = Pold;
        Pold = Pmat;
        scf_check(unomol_b200_scf_diag(tints.handle(), Fock.data(), nocc, Evals.data(), Cmat.data(), Pmat.data()), "scf_diag");
        pdiff = SymmPack::SymmPackDiffNorm(Pmat.data(), Pold.data(), no);
        ++iteration;
    }

    void findEnergy() noexcept {
        nucrep = nuclear_repulsion_energy(ncen, basis.center_ptr());
        OneElectronIntsAuto(basis, tints, Smat.data(), Tmat.data(), Hmat.data());
        scf_check(unomol_b200_scf_set_overlap(tints.handle(), Smat.data()), "scf_set_overlap");   // formXmatrix
        if (basis.scf_flags(2)) {
            FILE *in = fopen("PMATRIX.DAT", "r");
            if (!in || fread(Pmat.data(), sizeof(double), no2, in) != (size_t)no2) fatal_error("Could not read PMATRIX.DAT");
            fclose(in);
        } else {
            PmatrixGuess();
        }
        // One GPU: keep the whole iteration on the device (UNOMOL_HOST_SCF=1 forces the host-side bookkeeping of the
        // reference's update()); several GPUs in this process: partial G's are summed on the host, so stay on the host path.
        on_device = tints.number_of_gpus() == 1 && !std::getenv("UNOMOL_HOST_SCF");
        if (on_device) scf_check(unomol_b200_scf_load(tints.handle(), Hmat.data(), Pmat.data()), "scf_load");
        iteration = 0;
        eold = 0.0;
        update();
        report();
        init_energy = energy + nucrep;
        iteration = 1;
        update();
        report();
        iteration = 2;
        while (iteration < maxits) {
            scf_converger();
            update();
            if (is_converged()) break;
            report();
        }
        if (on_device) scf_check(unomol_b200_scf_fetch(tints.handle(), Pmat.data(), Evals.data(), Cmat.data()), "scf_fetch");
        PmatGs = Pmat;                       // RHF.hpp:169-171: the ground state seeds every point of the polarisation scan
        energyGs = energy + nucrep;
        FILE *fp = fopen("PMATRIX.DAT", "w");
        if (fp) { fwrite(Pmat.data(), sizeof(double), no2, fp); fclose(fp); }
        final_output(init_energy);
    }

    // Finite-field analysis (reference RHF.hpp:235-271, FField.cpp:5-19): three SCFs restarted from the ground-state density with
    // H - E.(dx, dy, dz), E = 5e-3 a.u. along x, y, z; polarisation energies and alpha = -2 dE / E^2 into finitefield.out.
    // The dipole matrices are the device one-electron kernel's (the reference reads the same numbers back from RMOM.DAT).
    void FiniteFieldAnalysis() {
        double polnrg[3], alfpol[3];
        const double Emag = basis.FiniteFieldValue();
        MomentMatrices mom;
        MomentIntsAuto(basis, tints, mom);
        for (int ix = 0; ix < 3; ++ix) {
            OneElectronIntsAuto(basis, tints, Smat.data(), Tmat.data(), Hmat.data());
            for (int i = 0; i < no2; ++i) Hmat[i] -= Emag * mom.m[ix][i];
            scf_check(unomol_b200_scf_set_overlap(tints.handle(), Smat.data()), "scf_set_overlap");
            Pmat = PmatGs;
            if (on_device) scf_check(unomol_b200_scf_load(tints.handle(), Hmat.data(), Pmat.data()), "scf_load");
            iteration = 0;
            eold = 0.0;
            mix_next = false;
            update();
            while (iteration < maxits) {
                scf_converger();
                update();
                if (is_converged()) break;
            }
            polnrg[ix] = energy + nucrep - energyGs;
            alfpol[ix] = -polnrg[ix] * 2.0 / Emag / Emag;
        }
        FILE *out = fopen("finitefield.out", "w");
        fprintf(out, "        Unomol Finite Field Analysis\n");
        fprintf(out, "        Electric field magnitude = %20.12le\n", Emag);
        fprintf(out, "     alfa                pol energy\n");
        fprintf(out, " x   %20.12le %20.12le\n", alfpol[0], polnrg[0]);
        fprintf(out, " y   %20.12le %20.12le\n", alfpol[1], polnrg[1]);
        fprintf(out, " z   %20.12le %20.12le\n", alfpol[2], polnrg[2]);
        fprintf(out, " alf isotropic = %20.12le\n", (alfpol[0] + alfpol[1] + alfpol[2]) / 3.0);
        fprintf(out, " alf aniso     = %20.12le\n", (2. * alfpol[2] - alfpol[1] - alfpol[0]) / 3.0);
        fclose(out);
    }

    // The polarisation-potential scan (reference RHF.hpp:292-388): the basis is augmented by the posin.bas shells on one extra
    // centre (the positron, charge +1), that centre is moved over the points of pos.grid.dat, and at every point an SCF restarted
    // from the ground-state density gives V_pol = E_final - E_first and V_stat = E_first - E_ground; vpol.out / spol.out as the
    // reference writes them.  Per point here: one incremental pair-table update (only the pairs of the moved centre are rebuilt,
    // engine.cu: update_pairs_incremental), S/T/H with the positron charge model on the device (unomol_b200_one_electron_dpm =
    // OneElectronInts + GDPMInts), X = S^-1/2, and the device-resident RHF iteration.  The reference adds the integrals of the new
    // shells (xints, start_shell = old shell count) to the cached molecular ones (tints); an integral-direct build has no cache
    // to add to, so ONE engine over the augmented basis computes both parts in one pass (same G).
    void findPolarizationPotential() {
        const int pcen = basis.skip_center();
        basis.dpm_augment();
        no = basis.number_of_orbitals();
        no2 = no * (no + 1) / 2;
        ncen = basis.number_of_centers();
        eps = 1.e-12;
        PmatGs.resize(no2, 0.0);             // new functions start empty (RHF.hpp:170)
        for (auto *v : {&Hmat, &Tmat, &Smat}) v->assign(no2, 0.0);
        std::ifstream in("pos.grid.dat");
        if (!in) fatal_error("could not open pos.grid.dat");
        int npts = 0;
        in >> npts;
        FILE *vout = fopen("vpol.out", "w"), *sout = fopen("spol.out", "w");
        std::vector<double> z(ncen);
        for (int c = 0; c < ncen; ++c) z[c] = (c == pcen) ? 0.0 : basis.center_ptr()[c].charge();   // OneElectronInts.cpp:43 skips it
        TwoElectronInts *x = nullptr;
        for (int i = 0; i < npts; ++i) {
            double px, py, pz;
            in >> px >> py >> pz;
            basis.SetCenterPosition(px, py, pz, pcen);
            nucrep = nuclear_repulsion_energy(ncen, basis.center_ptr());
            if (!x) x = new TwoElectronInts(basis, 0, std::string("XINTS.DAT"));
            else x->recalculate(basis);
            scf_check(unomol_b200_one_electron_dpm(x->handle(), z.data(), pcen, Smat.data(), Tmat.data(), Hmat.data(), nullptr), "one_electron_dpm");
            scf_check(unomol_b200_scf_set_overlap(x->handle(), Smat.data()), "scf_set_overlap");
            scf_check(unomol_b200_scf_load(x->handle(), Hmat.data(), PmatGs.data()), "scf_load");
            iteration = 0;
            eold = 0.0;
            auto upd = [&](bool mix) {
                scf_check(unomol_b200_scf_iterate_rhf(x->handle(), nocc, mix ? 1 : 0, &energy, &pdiff), "scf_iterate_rhf");
                ediff = energy - eold;
                eold = energy;
                ++iteration;
            };
            upd(false);
            const double e_first = energy + nucrep;
            while (iteration < maxits) {
                upd(!(ediff < 0.0));         // scf_converger: mix unless the last step lowered the energy
                if (is_converged()) break;
            }
            const double e_final = energy + nucrep;
            const double vpol = e_final - e_first;
            // the reference adds the nuclear repulsion a second time at the first point only (RHF.hpp:337 vs :374); kept as is
            const double vstat = (i == 0) ? e_first - energyGs + nucrep : e_first - energyGs;
            const double r2 = px * px + py * py + pz * pz;
            const double alfa = -2.0 * vpol * r2 * r2;
            fprintf(vout, "%15.10lf %15.10lf %15.10lf %25.15le %25.15le %25.15le %25.15le\n", px, py, pz, vpol, alfa, vstat, (vstat + vpol));
            fflush(vout);
            fprintf(sout, "%3d %20.10le %20.10le %20.10le %20.10le %25.15le\n", i, energyGs, e_first, e_final, vstat, ediff);
            fflush(sout);
            unomol_b200_stats_t st;
            unomol_b200_stats(x->handle(), &st);
            fprintf(stderr, "polarisation scan point %d: %d iterations, pair-table update %.3f ms (%d incremental so far), one-electron kernel %.3f ms\n",
                    i, iteration, st.precompute_ms, st.n_incremental_updates, st.onee_ms);
        }
        fclose(vout);
        fclose(sout);
        delete x;
    }

    void PmatrixGuess() {   // core-Hamiltonian guess, RHF.hpp:205-211
        scf_check(unomol_b200_scf_diag(tints.handle(), Hmat.data(), nocc, Evals.data(), Cmat.data(), Pmat.data()), "scf_diag");
    }

    bool is_converged() const noexcept {
        switch (cflag) {
            case 0: return ediff < eps;
            case 1: return pdiff < eps;
            default: return pdiff < eps && ediff < eps;
        }
    }

    void scf_converger() {
        if (on_device) {          // the mixing itself happens on the device at the start of the next iteration
            mix_next = !(ediff < 0.0);
            return;
        }
        if (ediff < 0.0) {
            Pold2 = Pold;
            Pold = Pmat;
            return;
        }
        for (int i = 0; i < no2; ++i) Pmat[i] = (Pmat[i] + Pold[i]) * 0.5;
    }

    void final_output(double e0) {
        FILE *out = fopen("short.gs.out", "w");
        fprintf(out, "%25.16le\n%25.16le\n%25.16le\n", e0, energy + nucrep, ediff);
        fclose(out);
        out = fopen("scfout.gs.out", "w");
        const double trace_t = 2.0 * SymmPack::TraceSymmPackProduct(Pmat.data(), Tmat.data(), no);
        const double virial = std::fabs((energy + nucrep - trace_t) / (trace_t) / 2.0);
        if (maxits <= iteration) fprintf(out, "WARNING CONVERGENCE _NOT_ REACHED!\n");
        fprintf(out, "Final Iteration              = %12u\n", iteration);
        fprintf(out, "Hartree Fock Energy          = %25.15le Hartree\n", energy + nucrep);
        fprintf(out, "Electronic Energy            = %25.15le Hartree\n", energy);
        fprintf(out, "Nuclear Rep. Energy          = %25.15le Hartree\n", nucrep);
        fprintf(out, "Kinetic Energy               = %25.15le Hartree\n", trace_t);
        fprintf(out, "Difference in Final Energies = %25.16le Hartree\n", ediff);
        fprintf(out, "Pmatrix diff norm            = %25.16le\n", pdiff);
        fprintf(out, "virial                       = %25.16le\n", virial);
        fprintf(out, "xxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxx\n               Orbital Energies\nOrbital Energy          Occupancy\n");
        for (int i = 0; i < no; i++) fprintf(out, "%7u %25.16le %12u\n", i + 1, Evals[i], i < nocc ? 2 : 0);
        fprintf(out, "xxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxx\n");
        PopulationAnalysis(out);
        fclose(out);
        // reference RHF.hpp:457-458
        MomentMatrices mom;
        MomentIntsAuto(basis, tints, mom);
        AnalyzeMoments(mom, Pmat.data(), (const double *)nullptr, basis.center_ptr(), ncen, no);
        // O(N^3) on the host and N(N+1)/2 text lines: written like the reference up to 1000 functions, beyond that only on
        // request (UNOMOL_MOL_DIPMOM=1) -- at 2002 functions it is a 100 MB file and a minute of host time
        out = (no <= 1000 || std::getenv("UNOMOL_MOL_DIPMOM")) ? fopen("mol_dipmom.out", "w") : nullptr;
        if (out) {
            fprintf(out, " MULTIPOLE MOMENT ANALYSIS \n units in bohr - hartree atomic units \n\n");
            AnalyzeMOMoments(mom, Cmat.data(), no, out, "MO TRANSISITION DIPOLE MOMENTS");
            fclose(out);
        }
    }

    // Mulliken populations (2 P S)_ii and atomic charges, same layout as the reference (RHF.hpp:571-660)
    void PopulationAnalysis(FILE *out) const {
        auto at = [&](const std::vector<double> &m, int i, int j) { return i >= j ? m[(size_t)i * (i + 1) / 2 + j] : m[(size_t)j * (j + 1) / 2 + i]; };
        std::vector<double> pop(no, 0.0), net(ncen);
        for (int i = 0; i < no; ++i) {
            double sum = 0.0;
            for (int k = 0; k < no; ++k) sum += at(Pmat, i, k) * at(Smat, k, i);
            pop[i] = sum + sum;
        }
        for (int c = 0; c < ncen; ++c) net[c] = basis.center_ptr()[c].charge();
        int ir = 0;
        for (int s = 0; s < basis.number_of_shells(); ++s) {
            const int lv = basis.shell_ptr()[s].Lvalue(), cn = basis.shell_ptr()[s].center();
            for (int k = 0; k < (lv + 1) * (lv + 2) / 2; ++k, ++ir) net[cn] -= pop[ir];
        }
        fprintf(out, "xxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxx\n           Mulliken Populations \nOrbital  Net Population\n");
        for (int i = 0; i < no; ++i) fprintf(out, "%7u %25.16le\n", i + 1, pop[i]);
        fprintf(out, "xxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxx\nMulliken Atomic Charges \nOrbital   Nuclear Charge   Net Charge \n");
        for (int c = 0; c < ncen; ++c) fprintf(out, "%7u %15.10lf %15.10lf \n", c + 1, basis.center_ptr()[c].charge(), net[c]);
    }

    double total_energy() const { return energy + nucrep; }
    int iterations() const { return iteration; }

  private:
    void report() const {
        fprintf(stderr, "Iteration     =  %6d\nEnergy        =  %25.16le\nDelta Energy  =  %25.15le\nDelta Density =  %25.15le\n\n", iteration,
                eold + nucrep, ediff, pdiff);
    }
    Basis &basis;
    TwoElectronInts &tints;
    int no = 0, no2 = 0, ncen = 0, nocc = 0, maxits = 0, scf_accel = 0, cflag = 0, iteration = 0;
    double eps = 0, ediff = 10.0, pdiff = 10.0, eold = 0, nucrep = 0, energy = 0, init_energy = 0, energyGs = 0;
    std::vector<double> Pold2, Pold, Pmat, Gmat, Hmat, Fock, Tmat, Smat, Evals, Cmat, PmatGs;
    bool on_device = false, mix_next = false;
};

class UnRestrictedHartreeFock {
  public:
    UnRestrictedHartreeFock() = delete;
    UnRestrictedHartreeFock(Basis *b, TwoElectronInts *t) : basis(*b), tints(*t) {
        no = basis.number_of_orbitals();
        no2 = no * (no + 1) / 2;
        ncen = basis.number_of_centers();
        noccB = basis.number_of_electrons();
        noccA = noccB - noccB / 2;   // UHF.hpp:38-40
        noccB /= 2;
        maxits = basis.maximum_iterations();
        eps = basis.scf_eps();
        scf_accel = basis.scf_flags(1);
        cflag = basis.scf_flags(0);
        for (auto *v : {&Pold2A, &PoldA, &PmatA, &GmatA, &Pold2B, &PoldB, &PmatB, &GmatB, &Hmat, &Fock, &Tmat, &Smat}) v->assign(no2, 0.0);
        EvalsA.assign(no, 0.0);
        EvalsB.assign(no, 0.0);
    }

    void update() noexcept {   // UHF.hpp:101-134
        if (on_device) {
            // device-resident iteration: both densities, both G, F and H stay on the GPU; the mixing of scf_converger too
            scf_check(unomol_b200_scf_iterate_uhf(tints.handle(), noccA, noccB, mix_next ? 1 : 0, &energy, &pdiff), "scf_iterate_uhf");
            mix_next = false;
            ediff = energy - eold;
            eold = energy;
            ++iteration;
            return;
        }
        for (int i = 0; i < no2; ++i) GmatA[i] = GmatB[i] = 0.0;
        tints.formGmatrix(PmatA.data(), PmatB.data(), GmatA.data(), GmatB.data());
        energy = SymmPack::TraceSymmPackProduct(PmatA.data(), Hmat.data(), no) + SymmPack::TraceSymmPackProduct(PmatB.data(), Hmat.data(), no) +
                 0.5 * (SymmPack::TraceSymmPackProduct(PmatA.data(), GmatA.data(), no) + SymmPack::TraceSymmPackProduct(PmatB.data(), GmatB.data(), no));
        ediff = energy - eold;
        eold = energy;
        for (int i = 0; i < no2; ++i) Fock[i] = Hmat[i] + GmatA[i];
        if (scf_accel == 1) Pold2A = PoldA;
        PoldA = PmatA;
        scf_check(unomol_b200_scf_diag(tints.handle(), Fock.data(), noccA, EvalsA.data(), nullptr, PmatA.data()), "scf_diag");
        pdiff = SymmPack::SymmPackDiffNorm(PmatA.data(), PoldA.data(), no);
        for (int i = 0; i < no2; ++i) Fock[i] = Hmat[i] + GmatB[i];
        if (scf_accel == 1) Pold2B = PoldB;
        PoldB = PmatB;
        scf_check(unomol_b200_scf_diag(tints.handle(), Fock.data(), noccB, EvalsB.data(), nullptr, PmatB.data()), "scf_diag");
        pdiff += SymmPack::SymmPackDiffNorm(PmatB.data(), PoldB.data(), no);
        ++iteration;
    }

    void findEnergy() noexcept {   // UHF.hpp:137-177
        nucrep = nuclear_repulsion_energy(ncen, basis.center_ptr());
        OneElectronIntsAuto(basis, tints, Smat.data(), Tmat.data(), Hmat.data());
        scf_check(unomol_b200_scf_set_overlap(tints.handle(), Smat.data()), "scf_set_overlap");
        if (basis.scf_flags(2)) {
            FILE *in = fopen("PMATRIX.DAT", "r");
            if (!in || fread(PmatA.data(), sizeof(double), no2, in) != (size_t)no2 || fread(PmatB.data(), sizeof(double), no2, in) != (size_t)no2)
                fatal_error("Could not read PMATRIX.DAT");
            fclose(in);
        } else {   // PmatrixGuess, UHF.hpp:206-215: both spins from the core Hamiltonian
            scf_check(unomol_b200_scf_diag(tints.handle(), Hmat.data(), noccA, EvalsA.data(), nullptr, PmatA.data()), "scf_diag");
            scf_check(unomol_b200_scf_diag(tints.handle(), Hmat.data(), noccB, EvalsB.data(), nullptr, PmatB.data()), "scf_diag");
        }
        // One GPU: the whole iteration stays on the device (UNOMOL_HOST_SCF=1 forces the host bookkeeping of the reference's update())
        on_device = tints.number_of_gpus() == 1 && !std::getenv("UNOMOL_HOST_SCF");
        if (on_device) scf_check(unomol_b200_scf_load_uhf(tints.handle(), Hmat.data(), PmatA.data(), PmatB.data()), "scf_load_uhf");
        iteration = 0;
        eold = 0.0;
        update();
        const double init_energy = energy + nucrep;
        iteration = 1;
        eold = 0.0;   // the reference resets eold here (UHF.hpp:152-154)
        update();
        while (iteration < maxits) {
            scf_converger();
            update();
            if (is_converged()) break;
            fprintf(stderr, "Iteration    =     %5d\nDelta Energy =  %25.15le\n", iteration, ediff);
        }
        if (on_device)
            scf_check(unomol_b200_scf_fetch_uhf(tints.handle(), PmatA.data(), PmatB.data(), EvalsA.data(), EvalsB.data()), "scf_fetch_uhf");
        PmatGsA = PmatA;                     // UHF.hpp:166-170: the ground state seeds every point of the polarisation scan
        PmatGsB = PmatB;
        energyGs = energy + nucrep;
        FILE *fp = fopen("PMATRIX.DAT", "w");
        if (fp) { fwrite(PmatA.data(), sizeof(double), no2, fp); fwrite(PmatB.data(), sizeof(double), no2, fp); fclose(fp); }
        FILE *out = fopen("short.gs.out", "w");
        fprintf(out, "%25.16le\n%25.16le\n%25.16le\n", init_energy, energy + nucrep, ediff);
        fclose(out);
        out = fopen("scfout.gs.out", "w");
        if (maxits <= iteration) fprintf(out, "WARNING CONVERGENCE _NOT_ REACHED!\n");
        fprintf(out, "Final Iteration              = %12u\n", iteration);
        fprintf(out, "Hartree Fock Energy          = %25.15le Hartree\n", energy + nucrep);
        fprintf(out, "Electronic Energy            = %25.15le Hartree\n", energy);
        fprintf(out, "Nuclear Rep. Energy          = %25.15le Hartree\n", nucrep);
        fprintf(out, "Difference in Final Energies = %25.16le Hartree\n", ediff);
        fprintf(out, "Pmatrix diff norm            = %25.16le\n", pdiff);
        fprintf(out, "               Alpha Orbital Energies\n");
        for (int i = 0; i < no; i++) fprintf(out, "%7u %25.16le %12u\n", i + 1, EvalsA[i], i < noccA ? 1 : 0);
        fprintf(out, "               Beta Orbital Energies\n");
        for (int i = 0; i < no; i++) fprintf(out, "%7u %25.16le %12u\n", i + 1, EvalsB[i], i < noccB ? 1 : 0);
        fclose(out);
        // reference UHF.hpp:483 (moments.out from (PA + PB)/2; the MO transition dipoles need the eigenvectors, which this
        // driver does not bring back from the device for UHF)
        MomentMatrices mom;
        MomentIntsAuto(basis, tints, mom);
        AnalyzeMoments(mom, PmatA.data(), PmatB.data(), basis.center_ptr(), ncen, no);
    }

    // Finite-field analysis for open shells (reference UHF.hpp:239-272); see RestrictedHartreeFock::FiniteFieldAnalysis
    void FiniteFieldAnalysis() {
        double polnrg[3], alfpol[3];
        const double Emag = basis.FiniteFieldValue();
        MomentMatrices mom;
        MomentIntsAuto(basis, tints, mom);
        for (int ix = 0; ix < 3; ++ix) {
            OneElectronIntsAuto(basis, tints, Smat.data(), Tmat.data(), Hmat.data());
            for (int i = 0; i < no2; ++i) Hmat[i] -= Emag * mom.m[ix][i];
            scf_check(unomol_b200_scf_set_overlap(tints.handle(), Smat.data()), "scf_set_overlap");
            PmatA = PmatGsA;
            PmatB = PmatGsB;
            if (on_device) scf_check(unomol_b200_scf_load_uhf(tints.handle(), Hmat.data(), PmatA.data(), PmatB.data()), "scf_load_uhf");
            iteration = 0;
            eold = 0.0;
            mix_next = false;
            update();
            while (iteration < maxits) {
                scf_converger();
                update();
                if (is_converged()) break;
            }
            polnrg[ix] = energy + nucrep - energyGs;
            alfpol[ix] = -polnrg[ix] * 2.0 / Emag / Emag;
        }
        FILE *out = fopen("finitefield.out", "w");
        fprintf(out, "        Unomol Finite Field Analysis\n");
        fprintf(out, "        Electric field magnitude = %20.12le\n", Emag);
        fprintf(out, "     alfa                pol energy\n");
        fprintf(out, " x   %20.12le %20.12le\n", alfpol[0], polnrg[0]);
        fprintf(out, " y   %20.12le %20.12le\n", alfpol[1], polnrg[1]);
        fprintf(out, " z   %20.12le %20.12le\n", alfpol[2], polnrg[2]);
        fclose(out);
    }

    // The polarisation-potential scan for open shells (reference UHF.hpp:293-383); same construction as
    // RestrictedHartreeFock::findPolarizationPotential above -- ONE engine over the augmented basis, incremental pair tables per
    // grid point, S/T/H with the positron charge model on the device -- with the device-resident UHF iteration.  vpol.out has the
    // reference's five columns here (UHF.hpp:337-338 writes no V_stat columns).
    void findPolarizationPotential() {
        const int pcen = basis.skip_center();
        basis.dpm_augment();
        no = basis.number_of_orbitals();
        no2 = no * (no + 1) / 2;
        ncen = basis.number_of_centers();
        eps = 1.e-12;
        PmatGsA.resize(no2, 0.0);            // new functions start empty (UHF.hpp:167,169)
        PmatGsB.resize(no2, 0.0);
        for (auto *v : {&Hmat, &Tmat, &Smat}) v->assign(no2, 0.0);
        std::ifstream in("pos.grid.dat");
        if (!in) fatal_error("could not open pos.grid.dat");
        int npts = 0;
        in >> npts;
        FILE *vout = fopen("vpol.out", "w"), *sout = fopen("spol.out", "w");
        std::vector<double> z(ncen);
        for (int c = 0; c < ncen; ++c) z[c] = (c == pcen) ? 0.0 : basis.center_ptr()[c].charge();
        TwoElectronInts *x = nullptr;
        for (int i = 0; i < npts; ++i) {
            double px, py, pz;
            in >> px >> py >> pz;
            basis.SetCenterPosition(px, py, pz, pcen);
            nucrep = nuclear_repulsion_energy(ncen, basis.center_ptr());
            if (!x) x = new TwoElectronInts(basis, 0, std::string("XINTS.DAT"));
            else x->recalculate(basis);
            scf_check(unomol_b200_one_electron_dpm(x->handle(), z.data(), pcen, Smat.data(), Tmat.data(), Hmat.data(), nullptr), "one_electron_dpm");
            scf_check(unomol_b200_scf_set_overlap(x->handle(), Smat.data()), "scf_set_overlap");
            scf_check(unomol_b200_scf_load_uhf(x->handle(), Hmat.data(), PmatGsA.data(), PmatGsB.data()), "scf_load_uhf");
            iteration = 0;
            eold = 0.0;
            auto upd = [&](bool mix) {
                scf_check(unomol_b200_scf_iterate_uhf(x->handle(), noccA, noccB, mix ? 1 : 0, &energy, &pdiff), "scf_iterate_uhf");
                ediff = energy - eold;
                eold = energy;
                ++iteration;
            };
            upd(false);
            const double e_first = energy + nucrep;
            while (iteration < maxits) {
                upd(!(ediff < 0.0));         // scf_converger: mix unless the last step lowered the energy
                if (is_converged()) break;
            }
            const double e_final = energy + nucrep;
            const double vpol = e_final - e_first;
            // the reference adds the nuclear repulsion a second time at the first point only (UHF.hpp:334 vs :371); kept as is
            const double vstat = (i == 0) ? e_first - energyGs + nucrep : e_first - energyGs;
            const double r2 = px * px + py * py + pz * pz;
            const double alfa = -2.0 * vpol * r2 * r2;
            fprintf(vout, "%15.10lf %15.10lf %15.10lf %25.15le %25.15le\n", px, py, pz, vpol, alfa);
            fflush(vout);
            fprintf(sout, "%3d %20.10le %20.10le %20.10le %20.10le %25.15le\n", i, energyGs, e_first, e_final, vstat, ediff);
            fflush(sout);
            unomol_b200_stats_t st;
            unomol_b200_stats(x->handle(), &st);
            fprintf(stderr, "polarisation scan point %d (UHF): %d iterations, pair-table update %.3f ms (%d incremental so far), one-electron kernel %.3f ms\n",
                    i, iteration, st.precompute_ms, st.n_incremental_updates, st.onee_ms);
        }
        fclose(vout);
        fclose(sout);
        delete x;
    }

    bool is_converged() const noexcept {
        switch (cflag) {
            case 0: return ediff < eps;
            case 1: return pdiff < eps;
            default: return pdiff < eps && ediff < eps;
        }
    }

    void scf_converger() {   // UHF.hpp:690-740, serial path (extrap never set)
        if (on_device) {      // the device keeps the previous densities; mixing happens at the start of the next iteration
            mix_next = !(ediff < 0.0);
            return;
        }
        if (ediff < 0.0) {
            Pold2A = PoldA; PoldA = PmatA;
            Pold2B = PoldB; PoldB = PmatB;
            return;
        }
        for (int i = 0; i < no2; ++i) {
            PmatA[i] = (PmatA[i] + PoldA[i]) * 0.5;
            PmatB[i] = (PmatB[i] + PoldB[i]) * 0.5;
        }
    }

    double total_energy() const { return energy + nucrep; }
    int iterations() const { return iteration; }

  private:
    Basis &basis;
    TwoElectronInts &tints;
    int no = 0, no2 = 0, ncen = 0, noccA = 0, noccB = 0, maxits = 0, scf_accel = 0, cflag = 0, iteration = 0;
    bool on_device = false, mix_next = false;
    double eps = 0, ediff = 10.0, pdiff = 10.0, eold = 0, nucrep = 0, energy = 0, energyGs = 0;
    std::vector<double> Pold2A, PoldA, PmatA, GmatA, Pold2B, PoldB, PmatB, GmatB, Hmat, Fock, Tmat, Smat, EvalsA, EvalsB, PmatGsA, PmatGsB;
};

}  // namespace unomol
