// tests/host_emul/onee_check.cpp -- TEST INFRASTRUCTURE: dumps the packed S, T, H matrices of
// unomol_b200/host/OneElectron.hpp for one patin.dat (no CUDA needed), for tests/test_host_moments.py.
// usage: onee_check <patin.dat> <out.bin>   (3 * no2 doubles: S, T, H)
#include <cstdio>
#include <string>
#include <vector>
#include "../../unomol_b200/host/Basis.hpp"
#include "../../unomol_b200/host/OneElectron.hpp"

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    unomol::Basis basis{std::string(argv[1])};
    const int no = basis.number_of_orbitals();
    const size_t no2 = (size_t)no * (no + 1) / 2;
    std::vector<double> S(no2), T(no2), H(no2);
    unomol::OneElectronInts(basis, S.data(), T.data(), H.data());
    FILE *f = fopen(argv[2], "wb");
    if (!f) return 3;
    fwrite(S.data(), sizeof(double), no2, f);
    fwrite(T.data(), sizeof(double), no2, f);
    fwrite(H.data(), sizeof(double), no2, f);
    fclose(f);
    return 0;
}
