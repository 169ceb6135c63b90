// tests/host_emul/rys_host.cpp -- TEST INFRASTRUCTURE: the product's Rys evaluator (unomol_b200/csrc/rys_roots.cuh, a
// host + device header) compiled as plain C++ so that the CPU suite can check it against the reference's roots and
// weights (tests/golden/rys_grid.npz) and against its own defining moments.
#include "../../unomol_b200/csrc/rys_roots.cuh"

using namespace ub200;

extern "C" int unomol_rys_host(int n, double x, int exact, double *r, double *w) {
    const RysTables T = rys_host_tables(exact);
    switch (n) {
        case 1: rys_roots<1>(x, r, w, T); return 0;
        case 2: rys_roots<2>(x, r, w, T); return 0;
        case 3: rys_roots<3>(x, r, w, T); return 0;
        case 4: rys_roots<4>(x, r, w, T); return 0;
        case 5: rys_roots<5>(x, r, w, T); return 0;
        case 6: rys_roots<6>(x, r, w, T); return 0;
        case 7: rys_roots<7>(x, r, w, T); return 0;
        case 8: rys_roots<8>(x, r, w, T); return 0;
        case 9: rys_roots<9>(x, r, w, T); return 0;
    }
    return -1;
}

// F_0(x) .. F_3(x) through the two-root moment path (x < 46): Taylor row of F_3 + downward recursion (boys_poly03)
extern "C" void unomol_boys_host(double x, double *F) {
    const RysTables T = rys_host_tables(0);
    boys_poly03(x, T.f3poly, F);
}
// ... and through the recursion-based grid of round 2's first evaluator (still used by nothing but this cross-check)
extern "C" void unomol_boys_grid_host(double x, double *F) {
    const RysTables T = rys_host_tables(0);
    boys_grid<3, RYS_BOYS_MTOP>(x, T.boys, F);
}

// F_0(x) alone through its own grid (the (ss|ss) kernels)
extern "C" double unomol_f0_host(double x) {
    const RysTables T = rys_host_tables(0);
    return rys1_f0(x, T);
}
