// tests/host_emul/moments_check.cpp -- TEST INFRASTRUCTURE: dumps the nine packed moment matrices of
// unomol_b200/host/Moments.hpp for one patin.dat (no CUDA needed), for tests/test_host_moments.py.
// usage: moments_check <patin.dat> <out.bin>   (9 * no2 doubles, order dx dy dz qxx qxy qxz qyy qyz qzz)
#include <cstdio>
#include <cstdlib>
#include <string>
#include "../../unomol_b200/host/Basis.hpp"
#include "../../unomol_b200/host/Moments.hpp"

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    unomol::Basis basis{std::string(argv[1])};
    unomol::MomentMatrices M;
    unomol::MomentInts(basis, M);
    FILE *f = fopen(argv[2], "wb");
    if (!f) return 3;
    for (int k = 0; k < 9; ++k) fwrite(M.m[k].data(), sizeof(double), M.m[k].size(), f);
    fclose(f);
    return 0;
}
