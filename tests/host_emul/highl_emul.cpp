// tests/host_emul/highl_emul.cpp -- TEST INFRASTRUCTURE: compiles the runtime-L kernel body
// (unomol_b200/csrc/eri_highl.cuh) as plain single-threaded C++ (HL_NT = 1, HL_SYNC a no-op) so that its
// arithmetic can be checked against the oracle on machines without a GPU.  It never ships: the product library
// compiles the same header with nvcc for sm_100a.  The shell-pair / primitive-pair set-up below mirrors
// engine.cu (build_pairs, host path) with the up-front prune disabled, as the engine does when f/g shells exist.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
struct int2 { int x, y; };
#include "../../unomol_b200/csrc/eri_highl.cuh"

using namespace ub200;

namespace {
struct Basis {
    int ns, nbf;
    const int *npr, *lv, *cen, *off, *poff;
    const double *alpha, *coef, *xyz;
};

struct Pairs {
    std::vector<ShellPair> sp;      // canonical id i(i+1)/2 + j
    std::vector<PrimPair> prims;
};

Pairs build_pairs(const Basis &B) {
    Pairs P;
    for (int i = 0; i < B.ns; ++i)
        for (int j = 0; j <= i; ++j) {
            int a = i, b = j;
            if (B.lv[i] < B.lv[j]) { a = j; b = i; }
            const double *A = B.xyz + 3 * B.cen[a], *Bc = B.xyz + 3 * B.cen[b];
            ShellPair sp{};
            double ab2 = 0.0;
            for (int x = 0; x < 3; ++x) { sp.AB[x] = A[x] - Bc[x]; ab2 += sp.AB[x] * sp.AB[x]; }
            const bool same = (a == b);
            std::vector<PrimPair> keep;
            for (int ia = 0; ia < B.npr[a]; ++ia) {
                const double axp = B.alpha[B.poff[a] + ia], c1 = B.coef[B.poff[a] + ia];
                const int jend = same ? ia + 1 : B.npr[b];
                for (int ib = 0; ib < jend; ++ib) {
                    const double bxp = B.alpha[B.poff[b] + ib], c2 = B.coef[B.poff[b] + ib];
                    PrimPair pp;
                    pp.p = axp + bxp;
                    pp.ip = 1.0 / pp.p;
                    pp.u = std::exp(-axp * bxp * ab2 * pp.ip) * pp.ip;
                    for (int x = 0; x < 3; ++x) {
                        pp.P[x] = (axp * A[x] + bxp * Bc[x]) * pp.ip;
                        pp.PA[x] = pp.P[x] - A[x];
                    }
                    pp.c = c1 * c2 * ((same && ia != ib) ? 2.0 : 1.0);
                    keep.push_back(pp);
                }
            }
            std::stable_sort(keep.begin(), keep.end(), [](const PrimPair &x, const PrimPair &y) { return x.u > y.u; });
            sp.umax = keep.front().u;
            sp.pmin = keep.front().p;
            for (auto &pp : keep) sp.pmin = std::min(sp.pmin, pp.p);
            sp.offa = B.off[a]; sp.offb = B.off[b];
            sp.sha = a; sp.shb = b;
            sp.pairid = i * (i + 1) / 2 + j;
            sp.prim_off = (int)P.prims.size();
            sp.nprim = (int)keep.size();
            P.prims.insert(P.prims.end(), keep.begin(), keep.end());
            P.sp.push_back(sp);
        }
    return P;
}

bool one_centre(const ShellPair &s) { return s.AB[0] == 0.0 && s.AB[1] == 0.0 && s.AB[2] == 0.0; }
}  // namespace

static int g_all_rys = 0;

extern "C" {

// 1: quartets with l_tot > 8 through the Rys quadrature (6..9 roots) instead of McMurchie-Davidson (engine option "all_rys")
void hl_emul_set_all_rys(int v) { g_all_rys = v; }

// (ish jsh | ksh lsh), every Cartesian component, out[((i*n2 + j)*n3 + k)*n4 + l]
int hl_emul_quartet(int ns, int nbf, const int *npr, const int *lv, const int *cen, const int *off, const int *poff,
                    const double *alpha, const double *coef, const double *xyz, double prim_cut, int ish, int jsh, int ksh,
                    int lsh, double *out) {
    Basis B{ns, nbf, npr, lv, cen, off, poff, alpha, coef, xyz};
    Pairs P = build_pairs(B);
    auto pid = [](int i, int j) { int hi = std::max(i, j), lo = std::min(i, j); return hi * (hi + 1) / 2 + lo; };
    const ShellPair &bra = P.sp[pid(ish, jsh)], &ket = P.sp[pid(ksh, lsh)];
    HighLArgs hl;
    hl.la = lv[bra.sha]; hl.lb = lv[bra.shb]; hl.lc = lv[ket.sha]; hl.ld = lv[ket.shb];
    const int NA = hl_ncart(hl.la), NB = hl_ncart(hl.lb), NC = hl_ncart(hl.lc), ND = hl_ncart(hl.ld);
    std::vector<double> V((size_t)NA * NB * NC * ND);
    std::vector<double> sm(HL_SMEM_DOUBLES + 4 * HL_NC);
    hl.scratch = V.data(); hl.slab = (long long)V.size();
    hl.rys = rys_host_tables();
    hl.all_rys = g_all_rys;
    hl_init_tables(hl, sm.data());
    hl_quartet_block(hl, bra, ket, P.prims.data(), prim_cut, one_centre(bra), one_centre(ket), sm.data(), V.data());
    const bool sw1 = (ish != jsh) && bra.sha != ish, sw2 = (ksh != lsh) && ket.sha != ksh;
    const int n1 = hl_ncart(lv[ish]), n2 = hl_ncart(lv[jsh]), n3 = hl_ncart(lv[ksh]), n4 = hl_ncart(lv[lsh]);
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j)
            for (int k = 0; k < n3; ++k)
                for (int l = 0; l < n4; ++l) {
                    const int a = sw1 ? j : i, b = sw1 ? i : j, c = sw2 ? l : k, d = sw2 ? k : l;
                    out[((i * n2 + j) * n3 + k) * n4 + l] = V[((a * NB + b) * NC + c) * ND + d];
                }
    return n1 * n2 * n3 * n4;
}

// Whole Fock build through the kernel body: every canonical pair of shell pairs (bra id >= ket id), RHF (nspin = 1,
// G = 2J - K) or UHF (nspin = 2).  P*/G* packed lower-triangular; G is overwritten.  only_highl != 0 restricts the sum to
// quartets containing an f or g shell (what the product routes to this kernel).
int hl_emul_fock(int ns, int nbf, const int *npr, const int *lv, const int *cen, const int *off, const int *poff,
                 const double *alpha, const double *coef, const double *xyz, double prim_cut, double value_cut, int nspin,
                 const double *PA, const double *PB, double *GA, double *GB, int only_highl) {
    Basis B{ns, nbf, npr, lv, cen, off, poff, alpha, coef, xyz};
    Pairs P = build_pairs(B);
    const size_t n = nbf, nn = n * n;
    std::vector<double> PJ(nn), PK0(nn), PK1(nn), J(nn, 0.0), K0(nn, 0.0), K1(nn, 0.0);
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < n; ++j) {
            const size_t a = std::max(i, j), b = std::min(i, j), p = a * (a + 1) / 2 + b;
            PK0[i * n + j] = PA[p];
            if (nspin == 2) { PK1[i * n + j] = PB[p]; PJ[i * n + j] = PA[p] + PB[p]; }
        }
    ClassTask task{};
    task.nbf = nbf; task.nspin = nspin;
    task.PJ = nspin == 1 ? PK0.data() : PJ.data();
    task.PK[0] = PK0.data(); task.PK[1] = PK1.data();
    task.jscale = nspin == 1 ? 4.0 : 2.0;
    task.J = J.data(); task.K[0] = K0.data(); task.K[1] = K1.data();
    std::vector<double> V(50625), sm(HL_SMEM_DOUBLES + 4 * HL_NC);
    const int np = (int)P.sp.size();
    for (int ib = 0; ib < np; ++ib)
        for (int ik = 0; ik <= ib; ++ik) {
            const ShellPair &bra = P.sp[ib], &ket = P.sp[ik];
            HighLArgs hl;
            hl.la = lv[bra.sha]; hl.lb = lv[bra.shb]; hl.lc = lv[ket.sha]; hl.ld = lv[ket.shb];
            if (only_highl && std::max(std::max(hl.la, hl.lb), std::max(hl.lc, hl.ld)) <= 2) continue;
            hl.scratch = V.data(); hl.slab = 50625;
            hl.rys = rys_host_tables();
            hl.all_rys = g_all_rys;
            hl_init_tables(hl, sm.data());
            hl_quartet_block(hl, bra, ket, P.prims.data(), prim_cut, one_centre(bra), one_centre(ket), sm.data(), V.data());
            double sym = 1.0;
            if (bra.sha == bra.shb) sym *= 0.5;
            if (ket.sha == ket.shb) sym *= 0.5;
            if (ib == ik) sym *= 0.5;
            const int nint = hl_ncart(hl.la) * hl_ncart(hl.lb) * hl_ncart(hl.lc) * hl_ncart(hl.ld);
            double mx = 0.0;
            for (int o = 0; o < nint; ++o) mx = std::fmax(mx, std::fabs(V[o]));
            if (mx > value_cut) hl_digest(hl, task, bra, ket, V.data(), sym);
        }
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j <= i; ++j) {
            const size_t p = i * (i + 1) / 2 + j;
            const double jj = J[i * n + j] + J[j * n + i];
            GA[p] = jj - (K0[i * n + j] + K0[j * n + i]);
            if (nspin == 2) GB[p] = jj - (K1[i * n + j] + K1[j * n + i]);
        }
    return 0;
}

}  // extern "C"
