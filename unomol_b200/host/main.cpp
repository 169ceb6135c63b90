// unomol_b200/host/main.cpp -- the reference's serial driver (reference Unomol.cc:8-26) on the GPU engine:
// reads ./patin.dat (or argv[1]), builds the two-electron engine, picks UHF iff the electron count is odd,
// runs the ground-state SCF and writes short.gs.out / scfout.gs.out / PMATRIX.DAT in the working directory.
// The polarisation-potential scan (Unomol.cc:23) follows for closed shells when int_flag[0] is set; the finite-field analysis
// (Unomol.cc:16,22) is not implemented and is refused loudly.
#include <chrono>
#include <cstdlib>
#include <string>
#include "SCF.hpp"

int main(int argc, char **argv) {
    if (argc > 3 && std::string(argv[1]) == "--onee") {
        // host-only: dump packed S, T, H (no GPU needed) -- used by the CPU tests against the reference fixtures
        unomol::Basis b(argv[2]);
        const int no2 = b.number_of_orbitals() * (b.number_of_orbitals() + 1) / 2;
        std::vector<double> S(no2), T(no2), H(no2);
        unomol::OneElectronInts(b, S.data(), T.data(), H.data());
        FILE *f = fopen(argv[3], "wb");
        fwrite(S.data(), 8, no2, f); fwrite(T.data(), 8, no2, f); fwrite(H.data(), 8, no2, f);
        fclose(f);
        return EXIT_SUCCESS;
    }
    std::string label("MINTS.DAT");
    unomol::Basis bas(argc > 1 ? std::string(argv[1]) : std::string("patin.dat"));
    auto t0 = std::chrono::steady_clock::now();
    unomol::TwoElectronInts t(bas, 0, label);
    auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "Time for Two Electrons Integrals set-up = %g seconds\n", std::chrono::duration<double>(t1 - t0).count());
    const int nelec = bas.number_of_electrons();
    // UNOMOL_SKIP_FINITE_FIELD=1: leave out the finite-field analysis patin.dat asks for (three more SCFs); the driver says so
    const char *skip_env = std::getenv("UNOMOL_SKIP_FINITE_FIELD");
    const bool skip_ff = skip_env && skip_env[0] == '1';
    if (bas.int_flags(1) && skip_ff) std::fprintf(stderr, "unomol_b200_scf: finite-field analysis skipped (UNOMOL_SKIP_FINITE_FIELD=1)\n");
    if (nelec % 2) {
        unomol::UnRestrictedHartreeFock uhf(&bas, &t);
        uhf.findEnergy();
        std::fprintf(stderr, "UHF energy %.15f after %d iterations\n", uhf.total_energy(), uhf.iterations());
        if (bas.int_flags(1) && !skip_ff) uhf.FiniteFieldAnalysis();   // reference Unomol.cc:17
        if (bas.int_flags(0)) uhf.findPolarizationPotential();       // reference Unomol.cc:18
    } else {
        unomol::RestrictedHartreeFock rhf(&bas, &t);
        rhf.findEnergy();
        std::fprintf(stderr, "RHF energy %.15f after %d iterations\n", rhf.total_energy(), rhf.iterations());
        if (bas.int_flags(1) && !skip_ff) rhf.FiniteFieldAnalysis();   // reference Unomol.cc:22
        if (bas.int_flags(0)) rhf.findPolarizationPotential();       // reference Unomol.cc:23
    }
    auto t2 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "SCF time = %g s\n", std::chrono::duration<double>(t2 - t1).count());
    return EXIT_SUCCESS;
}
