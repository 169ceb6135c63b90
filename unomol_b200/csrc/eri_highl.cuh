// unomol_b200/csrc/eri_highl.cuh -- runtime-L ERI + J/K digestion kernel for shell quartets that contain an f or g
// shell (l = 3, 4), FP64.  SURVEY.md section 8 rows a4-a7 for l > 2 and row a9.
//
// The reference dispatches on the total angular momentum of the quartet (TwoElectronInts.cpp:661-665):
//   l_tot <= 8 : Rys quadrature, calc_two_electron_ints_rys (TwoElectronInts.cpp:420-509; Rys.hpp:85-212), <= 5 roots,
//                with the primitive cut sr < 1e-12 (:479);
//   l_tot >  8 : McMurchie-Davidson, calc_two_electron_ints_md and its one-/two-centre variants
//                (TwoElectronInts.cpp:9-418; MD_Dfunction.hpp:39-72; MD_Rfunction.hpp:49-70, 2185-2319), NO primitive
//                cut, Boys function by MD_Rfunction::Fgamma -- whose t > 20 branch is the bare asymptotic value
//                (relative error up to 2.5e-10 just above 20).  That branch is reproduced here on purpose: parity
//                with the reference to 1e-12 per integral is the contract, and an exact Boys function would miss it.
// Both branches live in one kernel so that a launch is defined by its (bra list, ket list) like every other launch.
//
// Work decomposition: one CTA per contracted shell quartet at a time (a CTA claims single quartets of the launch's
// Schwarz-screened (bra, ket) list from the launch's work counter).  All per-primitive tables (2-D Rys tables and their shifted
// per-axis integrals; Hermite E coefficients and the R_tuv tensor) live in shared memory; the Cartesian block
// V[a][b][c][d] (up to 15^4 = 50 625 doubles for (gg|gg)) lives in a per-CTA global scratch slab, each thread
// owning a fixed strided subset so that no atomics are needed.  f/g shells are rare and almost always uncontracted;
// this kernel is written for coverage and parity, the s/p/d classes keep their specialised kernels.
//
// The body is written against four macros (HL_TID, HL_NT, HL_SYNC, HL_ATOMIC_ADD) with every work-sharing loop in
// strided form, so that tests/host_emul/ can compile the same source as plain single-threaded C++ (HL_NT = 1) and check
// it against the oracle on machines without a GPU.  That emulation build is test infrastructure only.
#pragma once
#include "rys_roots.cuh"
#include "unomol_types.h"

#ifdef __CUDACC__
#define HL_FN __device__ __forceinline__
#define HL_TID ((int)threadIdx.x)
#define HL_NT ((int)blockDim.x)
#define HL_SYNC() __syncthreads()
#define HL_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define HL_FN static inline
#define HL_TID 0
#define HL_NT 1
#define HL_SYNC() ((void)0)
#define HL_ATOMIC_ADD(p, v) (*(p) += (v))
#endif

namespace ub200 {

constexpr int HL_MAXL = 4;                       // reference Basis.hpp:222
constexpr int HL_NC = 15;                        // Cartesian components of a g shell
constexpr int HL_LD = HL_MAXL + 1;               // 5
constexpr int HL_TD = 2 * HL_MAXL + 1;           // 9: Hermite index of one pair
constexpr int HL_RD = 4 * HL_MAXL + 1;           // 17: R_tuv index
constexpr int HL_THREADS = 128;
// shared-memory layout (doubles).  The Rys tables alias the head of the McMurchie-Davidson region.
constexpr int HL_OFF_E12 = 0;                                   // [3][5][5][9]
constexpr int HL_OFF_E34 = HL_OFF_E12 + 3 * HL_LD * HL_LD * HL_TD;
constexpr int HL_OFF_AY = HL_OFF_E34 + 3 * HL_LD * HL_LD * HL_TD;   // [17][17][18]: R^m_{0,ly,lz}
constexpr int HL_OFF_R = HL_OFF_AY + HL_RD * HL_RD * (HL_RD + 1);   // [17][17][17]
constexpr int HL_OFF_END = HL_OFF_R + HL_RD * HL_RD * HL_RD;
constexpr int HL_OFF_G = 0;                                     // Rys: [3*5][<=25] 2-D tables
constexpr int HL_OFF_S = HL_OFF_G + 15 * 25;                    // Rys: [3*5][<=81] shifted per-axis integrals
constexpr int HL_RYS_BATCH = 5;                                 // roots whose tables are in shared memory at a time
constexpr int HL_OFF_S_HI = HL_OFF_G + 3 * HL_RYS_BATCH * HL_TD * HL_TD;   // l_tot > 8 through Rys: G [15][<=81], S [15][<=625]
constexpr int HL_OFF_NRM = HL_OFF_END;                          // [4][15] per-component norms
constexpr int HL_OFF_RED = HL_OFF_NRM + 4 * HL_NC;              // [HL_THREADS] reduction scratch
constexpr int HL_SMEM_DOUBLES = HL_OFF_RED + HL_THREADS;
static_assert(HL_OFF_S_HI + 3 * HL_RYS_BATCH * HL_LD * HL_LD * HL_LD * HL_LD <= HL_OFF_END, "the all-Rys tables fit the McMurchie-Davidson region");
constexpr size_t HL_SMEM_BYTES = sizeof(double) * HL_SMEM_DOUBLES + sizeof(int) * 4 * HL_NC;

struct HighLArgs {
    int la, lb, lc, ld;          // class of the bra list (la >= lb) and of the ket list (lc >= ld)
    double *scratch;             // per-CTA slabs for the Cartesian block
    long long slab;              // doubles per slab (>= ncart(la) ncart(lb) ncart(lc) ncart(ld))
    RysTables rys;               // Boys grid / piecewise root tables (device pointers; host arrays in the emulation build)
    int all_rys;                 // 1: Rys quadrature also for l_tot > 8 (6..9 roots), as the reference's MPI build does through
                                 // Rys::rootN (TwoElectronIntsMPI.cpp has no McMurchie-Davidson dispatch); 0: the serial reference's rule
};

HL_FN int hl_ncart(int l) { return (l + 1) * (l + 2) / 2; }
// Cartesian components in AuxFunctions order (reference AuxFunctions.hpp:35-41): lx = l..0, ly = l-lx..0; packed lx | ly<<4 | lz<<8
HL_FN int hl_cart_pack(int l, int c) {
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= c) ++i;
    const int j = c - i * (i + 1) / 2;
    return (l - i) | ((i - j) << 4) | (j << 8);
}
// per-component factor with the reference's own recurrence (AuxFunctions.hpp:49-64): df = 1, 1, 3, 45, 4725
HL_FN double hl_cart_norm(int pk) {
    const double df[5] = {1.0, 1.0, 3.0, 45.0, 4725.0};
    return 1.0 / sqrt(df[pk & 15] * df[(pk >> 4) & 15] * df[(pk >> 8) & 15]);
}
HL_FN double hl_binom(int n, int k) {
    const double B[5][5] = {{1, 0, 0, 0, 0}, {1, 1, 0, 0, 0}, {1, 2, 1, 0, 0}, {1, 3, 3, 1, 0}, {1, 4, 6, 4, 1}};
    return B[n][k];
}

// MD_Rfunction::Fgamma, reference MD_Rfunction.hpp:2185-2206 (semantics kept, see the header comment)
HL_FN void hl_fgamma(double *fm, double t, int m) {
    const double sqrtpi = 0.88622692545275801365;
    if (t > 20.0) {
        fm[0] = sqrtpi / sqrt(t);
        for (int i = 1; i <= m; i++) fm[i] = fm[i - 1] * (i - 0.5) / t;
        return;
    }
    const double mphalf = m + 0.5;
    double term = 0.5 / mphalf, sum = term;
    for (int i = 1; i <= 200; i++) {
        term *= (t / (mphalf + i));
        sum += term;
        if (term < 1.e-12) break;
    }
    const double twot = 2.0 * t, expt = exp(-t);
    fm[m] = sum * expt;
    for (int i = m - 1; i >= 0; i--) fm[i] = (fm[i + 1] * twot + expt) / (i + i + 1.0);
}

// MD_Dfunction::eval, reference MD_Dfunction.hpp:39-72: E[i][j][n], i <= l1 (first shell), j <= l2, n <= i + j
HL_FN void hl_ecoef(double *E, double abi, double ax, double bx, int l1, int l2) {
#define EE(i, j, n) E[((i) * HL_LD + (j)) * HL_TD + (n)]
    for (int i = 0; i <= l1; i++)
        for (int j = 0; j <= l2; j++)
            for (int n = 0; n < HL_TD; n++) EE(i, j, n) = 0.0;
    EE(0, 0, 0) = 1.0;
    for (int j = 1; j <= l2; j++) {
        EE(0, j, 0) = bx * EE(0, j - 1, 0) + EE(0, j - 1, 1);
        for (int n = 1; n < j; n++) EE(0, j, n) = abi * EE(0, j - 1, n - 1) + bx * EE(0, j - 1, n) + (n + 1) * EE(0, j - 1, n + 1);
        EE(0, j, j) = abi * EE(0, j - 1, j - 1);
    }
    for (int i = 1; i <= l1; i++)
        for (int j = 0; j <= l2; j++) {
            const int ipj = i + j;
            EE(i, j, 0) = ax * EE(i - 1, j, 0) + EE(i - 1, j, 1);
            for (int n = 1; n < ipj; n++) EE(i, j, n) = abi * EE(i - 1, j, n - 1) + ax * EE(i - 1, j, n) + (n + 1) * EE(i - 1, j, n + 1);
            EE(i, j, ipj) = abi * EE(i - 1, j, ipj - 1);
        }
#undef EE
}

// 2-D Rys recurrence for one (root, axis), reference Rys.hpp:194-212; G[i*GJ + j], i <= La, j <= Lb; G[0][0] = scale
HL_FN void hl_vrr(double *G, int La, int Lb, double B00, double B1, double B1p, double C, double Cp, double scale) {
    const int GJ = Lb + 1;
    G[0] = scale;
    for (int j = 1; j <= Lb; ++j) G[j] = (j > 1 ? (j - 1) * B1p * G[j - 2] : 0.0) + Cp * G[j - 1];
    for (int i = 1; i <= La; ++i) {
        G[i * GJ] = (i > 1 ? (i - 1) * B1 * G[(i - 2) * GJ] : 0.0) + C * G[(i - 1) * GJ];
        for (int j = 1; j <= Lb; ++j)
            G[i * GJ + j] = (j > 1 ? (j - 1) * B1p * G[i * GJ + j - 2] : 0.0) + i * B00 * G[(i - 1) * GJ + j - 1] + Cp * G[i * GJ + j - 1];
    }
}

template <int NR>
HL_FN void hl_roots_n(double X, double *rt, double *wt, const RysTables &T) { rys_t2<NR>(X, rt, wt, T); }   // rt[] = t^2
HL_FN void hl_roots(int nr, double X, double *rt, double *wt, const RysTables &T) {
    switch (nr) {
        case 1: hl_roots_n<1>(X, rt, wt, T); break;
        case 2: hl_roots_n<2>(X, rt, wt, T); break;
        case 3: hl_roots_n<3>(X, rt, wt, T); break;
        case 4: hl_roots_n<4>(X, rt, wt, T); break;
        case 5: hl_roots_n<5>(X, rt, wt, T); break;
        // all-Rys mode only (l_tot > 8): the range of the reference's Rys::rootN
        case 6: hl_roots_n<6>(X, rt, wt, T); break;
        case 7: hl_roots_n<7>(X, rt, wt, T); break;
        case 8: hl_roots_n<8>(X, rt, wt, T); break;
        default: hl_roots_n<9>(X, rt, wt, T); break;
    }
}

// max over the CTA (every thread gets the result)
HL_FN double hl_block_max(double v, double *red) {
    const int tid = HL_TID, nt = HL_NT;
    HL_SYNC();
    red[tid] = v;
    HL_SYNC();
    for (int s = nt / 2; s > 0; s >>= 1) {
        if (tid < s) red[tid] = fmax(red[tid], red[tid + s]);
        HL_SYNC();
    }
    const double r = red[0];
    HL_SYNC();
    return r;
}

// Contracted Cartesian block of one shell quartet into V[((a*NB + b)*NC + c)*ND + d] (norms applied, no symmetry
// factor).  sm = the CTA's shared memory (HL_SMEM_DOUBLES doubles followed by 4*15 ints).  Returns the number of
// primitive quartets evaluated (uniform over the CTA).
HL_FN unsigned long long hl_quartet_block(const HighLArgs &hl, const ShellPair &bra, const ShellPair &ket, const PrimPair *prims,
                                          double prim_cut, bool one12, bool one34, double *sm, double *V) {
    const int tid = HL_TID, nt = HL_NT;
    const int la = hl.la, lb = hl.lb, lc = hl.lc, ld = hl.ld;
    const int La = la + lb, Lb = lc + ld, ltot = La + Lb;
    const int NA = hl_ncart(la), NB = hl_ncart(lb), NC = hl_ncart(lc), ND = hl_ncart(ld);
    const int NINT = NA * NB * NC * ND;
    const int *cexp = reinterpret_cast<const int *>(sm + HL_SMEM_DOUBLES);   // [4][15] packed exponents
    const double *cnrm = sm + HL_OFF_NRM;
    const PrimPair *bp = prims + bra.prim_off, *kp = prims + ket.prim_off;
    unsigned long long nprimq = 0;
    for (int o = tid; o < NINT; o += nt) V[o] = 0.0;
    HL_SYNC();
    if (ltot <= 8 || hl.all_rys) {
        // ------------------------------------------------ Rys branch
        const int nr = ltot / 2 + 1, GJ = Lb + 1, GSZ = (La + 1) * GJ;
        const int nS = (la + 1) * (lb + 1) * (lc + 1) * (ld + 1);
        // l_tot > 8 (all-Rys mode, 6..9 roots): the tables of at most HL_RYS_BATCH roots at a time, in the larger layout
        double *G = sm + HL_OFF_G, *S = sm + (ltot <= 8 ? HL_OFF_S : HL_OFF_S_HI);
        const double cut2 = prim_cut * prim_cut;
        for (int ib = 0; ib < bra.nprim; ++ib) {
            const PrimPair b = bp[ib];
            for (int ik = 0; ik < ket.nprim; ++ik) {
                const PrimPair k = kp[ik];
                const double txp = b.p + k.p;
                const double t0 = SR_TERM * b.u * k.u;
                if (t0 * t0 < cut2 * txp) continue;    // sr < 1e-12 BEFORE the contraction coefficients (:478-480)
                ++nprimq;
                const double itx = 1.0 / txp;
                const double sr = t0 * sqrt(itx) * (b.c * k.c);
                const double pq[3] = {b.P[0] - k.P[0], b.P[1] - k.P[1], b.P[2] - k.P[2]};
                const double X = b.p * k.p * itx * (pq[0] * pq[0] + pq[1] * pq[1] + pq[2] * pq[2]);
                double rt[9], wt[9];
                hl_roots(nr, X, rt, wt, hl.rys);
                for (int r0 = 0; r0 < nr; r0 += HL_RYS_BATCH) {
                const int nrb = (nr - r0 < HL_RYS_BATCH) ? nr - r0 : HL_RYS_BATCH;      // roots of this batch (all of them when nr <= 5)
                for (int tsk = tid; tsk < 3 * nrb; tsk += nt) {
                    const int ir = tsk / 3, ax = tsk - 3 * ir;
                    const double dr = rt[r0 + ir];
                    const double fff = dr * itx;
                    const double B00 = 0.5 * fff;
                    const double B1 = (0.5 - B00 * k.p) * b.ip;
                    const double B1p = (0.5 - B00 * b.p) * k.ip;
                    const double Cc = b.PA[ax] - k.p * pq[ax] * fff;
                    const double Cp = k.PA[ax] + b.p * pq[ax] * fff;
                    hl_vrr(G + tsk * GSZ, La, Lb, B00, B1, B1p, Cc, Cp, ax == 2 ? wt[r0 + ir] * sr : 1.0);   // weight and prefactor ride on z
                }
                HL_SYNC();
                // horizontal transfer per axis (reference Rys.hpp:173-192): exponents (ia, ib | ic, id) of one axis
                for (int o = tid; o < 3 * nrb * nS; o += nt) {
                    const int tsk = o / nS;
                    int r = o - tsk * nS;
                    const int ax = tsk % 3;
                    const int id = r % (ld + 1); r /= (ld + 1);
                    const int ic = r % (lc + 1); r /= (lc + 1);
                    const int ibb = r % (lb + 1);
                    const int ia = r / (lb + 1);
                    const double *Gt = G + tsk * GSZ;
                    const double abx = bra.AB[ax], cdx = ket.AB[ax];
                    double sum = 0.0, x12 = 1.0;
                    for (int i = 0; i <= ibb; ++i) {
                        double x34 = hl_binom(ibb, i) * x12;
                        for (int j = 0; j <= id; ++j) {
                            sum += hl_binom(id, j) * x34 * Gt[(ia + ibb - i) * GJ + (ic + id - j)];
                            x34 *= cdx;
                        }
                        x12 *= abx;
                    }
                    S[o] = sum;
                }
                HL_SYNC();
                for (int o = tid; o < NINT; o += nt) {
                    int r = o;
                    const int d = r % ND; r /= ND;
                    const int c = r % NC; r /= NC;
                    const int bb = r % NB;
                    const int a = r / NB;
                    const int pa = cexp[a], pb = cexp[HL_NC + bb], pc = cexp[2 * HL_NC + c], pd = cexp[3 * HL_NC + d];
                    int idx[3];
                    for (int ax = 0; ax < 3; ++ax) {
                        const int sh = 4 * ax;
                        idx[ax] = ((((pa >> sh) & 15) * (lb + 1) + ((pb >> sh) & 15)) * (lc + 1) + ((pc >> sh) & 15)) * (ld + 1) + ((pd >> sh) & 15);
                    }
                    double s = 0.0;
                    for (int ir = 0; ir < nrb; ++ir) {
                        const double *Sr = S + 3 * ir * nS;
                        s = fma(Sr[idx[0]] * Sr[nS + idx[1]], Sr[2 * nS + idx[2]], s);
                    }
                    V[o] += s;
                }
                HL_SYNC();
                }
            }
        }
    } else {
        // ------------------------------------------------ McMurchie-Davidson branch (no primitive cut)
        double *E12 = sm + HL_OFF_E12, *E34 = sm + HL_OFF_E34, *Ay = sm + HL_OFF_AY, *R = sm + HL_OFF_R;
        const int ESZ = HL_LD * HL_LD * HL_TD;
#define AY(ly, lz, m) Ay[((ly) * HL_RD + (lz)) * (HL_RD + 1) + (m)]
#define RR(lx, ly, lz) R[((lx) * HL_RD + (ly)) * HL_RD + (lz)]
        for (int ib = 0; ib < bra.nprim; ++ib) {
            const PrimPair b = bp[ib];
            const double abi = 0.5 * b.ip;
            // P - A and P - B; exactly zero for a one-centre pair (the reference's one-/two-centre variants)
            for (int ax = tid; ax < 3; ax += nt)
                hl_ecoef(E12 + ax * ESZ, abi, one12 ? 0.0 : b.PA[ax], one12 ? 0.0 : b.PA[ax] + bra.AB[ax], la, lb);
            for (int ik = 0; ik < ket.nprim; ++ik) {
                const PrimPair k = kp[ik];
                ++nprimq;
                const double cdi = 0.5 * k.ip;
                const double txp = b.p + k.p;
                // sr = SRterm * c12 e^{-..} * c34 e^{-..} / (p q sqrt(p+q))  (:326-334 with abi, cdi halved)
                const double sr = SR_TERM * (b.c * b.u) * (k.c * k.u) / sqrt(txp);
                const double w = b.p * k.p / txp;
                const double pq[3] = {b.P[0] - k.P[0], b.P[1] - k.P[1], b.P[2] - k.P[2]};
                const double tt = w * (pq[0] * pq[0] + pq[1] * pq[1] + pq[2] * pq[2]);
                HL_SYNC();   // previous primitive's readers of E34 / R are done
                for (int ax = tid; ax < 3; ax += nt)
                    hl_ecoef(E34 + ax * ESZ, cdi, one34 ? 0.0 : k.PA[ax], one34 ? 0.0 : k.PA[ax] + ket.AB[ax], lc, ld);
                // R^m_{00lz} (MD_Rfunction.hpp:49-66 + loop_eval z stage): one thread, <= 153 entries
                if (tid == nt - 1) {
                    double fm[HL_RD];
                    hl_fgamma(fm, tt, ltot);
                    double sterm = sr;
                    const double term = -(w + w);
                    for (int m = 0; m <= ltot; ++m) {
                        AY(0, 0, m) = sterm * fm[m];
                        sterm *= term;
                    }
                    const double z = pq[2];
                    for (int lz = 1; lz <= ltot; ++lz)
                        for (int m = 0; m <= ltot - lz; ++m)
                            AY(0, lz, m) = z * AY(0, lz - 1, m + 1) + (lz > 1 ? (lz - 1) * AY(0, lz - 2, m + 1) : 0.0);
                }
                HL_SYNC();
                // y stage: independent per lz
                for (int lz = tid; lz <= ltot; lz += nt) {
                    const double y = pq[1];
                    for (int ly = 1; ly <= ltot - lz; ++ly)
                        for (int m = 0; m <= ltot - ly - lz; ++m)
                            AY(ly, lz, m) = y * AY(ly - 1, lz, m + 1) + (ly > 1 ? (ly - 1) * AY(ly - 2, lz, m + 1) : 0.0);
                }
                HL_SYNC();
                // x stage: independent per (ly, lz); two rolling rows in thread-private storage, only m = 0 is kept
                for (int o = tid; o < (ltot + 1) * (ltot + 1); o += nt) {
                    const int ly = o / (ltot + 1), lz = o - ly * (ltot + 1);
                    if (ly + lz > ltot) continue;
                    const int M = ltot - ly - lz;
                    const double x = pq[0];
                    double r0[HL_RD], r1[HL_RD];
                    for (int m = 0; m <= M; ++m) r0[m] = AY(ly, lz, m);
                    RR(0, ly, lz) = r0[0];
                    if (M >= 1) {
                        for (int m = 0; m <= M - 1; ++m) r1[m] = x * r0[m + 1];
                        RR(1, ly, lz) = r1[0];
                    }
                    for (int lx = 2; lx <= M; ++lx) {
                        // row(lx)[m] = x row(lx-1)[m+1] + (lx-1) row(lx-2)[m+1], written over row(lx-2)
                        double *older = (lx & 1) ? r1 : r0, *newer = (lx & 1) ? r0 : r1;
                        for (int m = 0; m <= M - lx; ++m) older[m] = x * newer[m + 1] + (lx - 1) * older[m + 1];
                        RR(lx, ly, lz) = older[0];
                    }
                }
                HL_SYNC();
                // Six-index contraction (:388-411), factorised: the ket pair's Hermite expansion is folded into R first,
                //   W[cd][tuv] = sum_{t'u'v'} (-1)^{t'+u'+v'} E34x[t'] E34y[u'] E34z[v'] R[t+t'][u+u'][v+v'],
                // then the bra pair's: V[ab][cd] += sum_{tuv} E12x[t] E12y[u] E12z[v] W[cd][tuv].  Same terms as the
                // reference's nested loops, summed in a different order (~50x fewer operations for (gg|gg)).  W is built
                // for a chunk of ket function pairs at a time in the shared memory that held the y-stage table.
                {
                    const int D = La + 1, NHd = D * D * D, NAB = NA * NB, NCD = NC * ND;
                    const int CH = (HL_RD * HL_RD * (HL_RD + 1)) / NHd;   // >= 7
                    double *W = Ay;
                    for (int cd0 = 0; cd0 < NCD; cd0 += CH) {
                        const int cha = (NCD - cd0 < CH) ? NCD - cd0 : CH;
                        for (int o = tid; o < cha * NHd; o += nt) {
                            const int cdl = o / NHd, h = o - cdl * NHd;
                            const int t = h / (D * D), u = (h / D) % D, v = h % D;
                            if (t + u + v > La) continue;
                            const int cd = cd0 + cdl, c = cd / ND, d = cd - c * ND;
                            const int pc = cexp[2 * HL_NC + c], pd = cexp[3 * HL_NC + d];
                            const int l3 = pc & 15, m3 = (pc >> 4) & 15, n3 = (pc >> 8) & 15;
                            const int l4 = pd & 15, m4 = (pd >> 4) & 15, n4 = (pd >> 8) & 15;
                            const double *ex34 = E34 + (l3 * HL_LD + l4) * HL_TD, *ey34 = E34 + ESZ + (m3 * HL_LD + m4) * HL_TD,
                                         *ez34 = E34 + 2 * ESZ + (n3 * HL_LD + n4) * HL_TD;
                            const int l34 = l3 + l4, m34 = m3 + m4, n34 = n3 + n4;
                            double sum = 0.0;
                            for (int ix = 0; ix <= l34; ++ix)
                                for (int iy = 0; iy <= m34; ++iy) {
                                    const double v34 = ex34[ix] * ey34[iy];
                                    const double *rzp = &RR(t + ix, u + iy, v);
                                    double sx = ((ix + iy) & 1) ? -1.0 : 1.0;
                                    for (int iz = 0; iz <= n34; ++iz) {
                                        sum += sx * v34 * ez34[iz] * rzp[iz];
                                        sx = -sx;
                                    }
                                }
                            W[o] = sum;
                        }
                        HL_SYNC();
                        for (int o = tid; o < NAB * cha; o += nt) {
                            const int ab = o / cha, cdl = o - ab * cha;
                            const int a = ab / NB, bb = ab - a * NB;
                            const int pa = cexp[a], pb = cexp[HL_NC + bb];
                            const int l1 = pa & 15, m1 = (pa >> 4) & 15, n1 = (pa >> 8) & 15;
                            const int l2 = pb & 15, m2 = (pb >> 4) & 15, n2 = (pb >> 8) & 15;
                            const double *ex12 = E12 + (l1 * HL_LD + l2) * HL_TD, *ey12 = E12 + ESZ + (m1 * HL_LD + m2) * HL_TD,
                                         *ez12 = E12 + 2 * ESZ + (n1 * HL_LD + n2) * HL_TD;
                            const int l12 = l1 + l2, m12 = m1 + m2, n12 = n1 + n2;
                            const double *Wc = W + cdl * NHd;
                            double sum = 0.0;
                            for (int t = 0; t <= l12; ++t)
                                for (int u = 0; u <= m12; ++u) {
                                    const double v12 = ex12[t] * ey12[u];
                                    const double *wz = Wc + (t * D + u) * D;
                                    for (int v = 0; v <= n12; ++v) sum += v12 * ez12[v] * wz[v];
                                }
                            V[ab * NCD + cd0 + cdl] += sum;
                        }
                        HL_SYNC();
                    }
                }
            }
            HL_SYNC();   // E12 is rewritten by the next bra primitive
        }
#undef AY
#undef RR
    }
    HL_SYNC();
    for (int o = tid; o < NINT; o += nt) {
        int r = o;
        const int d = r % ND; r /= ND;
        const int c = r % NC; r /= NC;
        const int bb = r % NB;
        const int a = r / NB;
        V[o] *= cnrm[a] * cnrm[HL_NC + bb] * cnrm[2 * HL_NC + c] * cnrm[3 * HL_NC + d];
    }
    HL_SYNC();
    return nprimq;
}

// component tables of the launch's four shell types, built once per CTA
HL_FN void hl_init_tables(const HighLArgs &hl, double *sm) {
    int *cexp = reinterpret_cast<int *>(sm + HL_SMEM_DOUBLES);
    double *cnrm = sm + HL_OFF_NRM;
    const int ls[4] = {hl.la, hl.lb, hl.lc, hl.ld};
    for (int o = HL_TID; o < 4 * HL_NC; o += HL_NT) {
        const int s = o / HL_NC, c = o - s * HL_NC;
        if (c < hl_ncart(ls[s])) {
            const int pk = hl_cart_pack(ls[s], c);
            cexp[o] = pk;
            cnrm[o] = hl_cart_norm(pk);
        } else {
            cexp[o] = 0;
            cnrm[o] = 0.0;
        }
    }
    HL_SYNC();
}

// J/K digestion of one block (reference TwoElectronInts.cpp:699-820 in shell-block form; same contractions as
// eri_generic.cuh).  V already carries the norms; sym = the shell quartet's symmetry factor.
HL_FN void hl_digest(const HighLArgs &hl, const ClassTask &task, const ShellPair &bra, const ShellPair &ket, const double *V, double sym) {
    const int NA = hl_ncart(hl.la), NB = hl_ncart(hl.lb), NC = hl_ncart(hl.lc), ND = hl_ncart(hl.ld);
    const int NCD = NC * ND;
    const int n = task.nbf;
    const int oa = bra.offa, ob = bra.offb, oc = ket.offa, od = ket.offb;
    const int N_JAB = NA * NB, N_JCD = NCD, N_K = NA * NC + NA * ND + NB * NC + NB * ND;
    const int nout = N_JAB + N_JCD + task.nspin * N_K;
    for (int o = HL_TID; o < nout; o += HL_NT) {
        double s = 0.0;
        if (o < N_JAB) {
            const int a = o / NB, b = o - a * NB;
            for (int c = 0; c < NC; ++c)
                for (int d = 0; d < ND; ++d) s = fma(V[o * NCD + c * ND + d], task.PJ[(size_t)(oc + c) * n + od + d], s);
            HL_ATOMIC_ADD(task.J + (size_t)(oa + a) * n + ob + b, task.jscale * sym * s);
        } else if (o < N_JAB + N_JCD) {
            const int cd = o - N_JAB, c = cd / ND, d = cd - c * ND;
            for (int a = 0; a < NA; ++a)
                for (int b = 0; b < NB; ++b) s = fma(V[(a * NB + b) * NCD + cd], task.PJ[(size_t)(oa + a) * n + ob + b], s);
            HL_ATOMIC_ADD(task.J + (size_t)(oc + c) * n + od + d, task.jscale * sym * s);
        } else {
            int r = o - N_JAB - N_JCD;
            const int sp = r / N_K;
            r -= sp * N_K;
            const double *P = task.PK[sp];
            double *K = task.K[sp];
            if (r < NA * NC) {
                const int a = r / NC, c = r - a * NC;
                for (int b = 0; b < NB; ++b)
                    for (int d = 0; d < ND; ++d) s = fma(V[((a * NB + b) * NC + c) * ND + d], P[(size_t)(ob + b) * n + od + d], s);
                HL_ATOMIC_ADD(K + (size_t)(oa + a) * n + oc + c, sym * s);
            } else if (r < NA * NC + NA * ND) {
                r -= NA * NC;
                const int a = r / ND, d = r - a * ND;
                for (int b = 0; b < NB; ++b)
                    for (int c = 0; c < NC; ++c) s = fma(V[((a * NB + b) * NC + c) * ND + d], P[(size_t)(ob + b) * n + oc + c], s);
                HL_ATOMIC_ADD(K + (size_t)(oa + a) * n + od + d, sym * s);
            } else if (r < NA * NC + NA * ND + NB * NC) {
                r -= NA * NC + NA * ND;
                const int b = r / NC, c = r - b * NC;
                for (int a = 0; a < NA; ++a)
                    for (int d = 0; d < ND; ++d) s = fma(V[((a * NB + b) * NC + c) * ND + d], P[(size_t)(oa + a) * n + od + d], s);
                HL_ATOMIC_ADD(K + (size_t)(ob + b) * n + oc + c, sym * s);
            } else {
                r -= NA * NC + NA * ND + NB * NC;
                const int b = r / ND, d = r - b * ND;
                for (int a = 0; a < NA; ++a)
                    for (int c = 0; c < NC; ++c) s = fma(V[((a * NB + b) * NC + c) * ND + d], P[(size_t)(oa + a) * n + oc + c], s);
                HL_ATOMIC_ADD(K + (size_t)(ob + b) * n + od + d, sym * s);
            }
        }
    }
}

}  // namespace ub200
