// unomol_b200/csrc/rys_roots.cuh -- Rys quadrature roots and weights for 1..5 roots, FP64, host + device.
//
// Replaces the reference's Rys::root1..root5 (reference Rys.cpp:314-2197, dispatch Rys.hpp:145-164): same
// contract -- r[i] = t_i^2 / (1 - t_i^2), w[i] = weights of the n-point quadrature for the weight function
// exp(-X t^2) on [0,1], roots ascending -- but an independent evaluator.  Nothing here is taken from the
// reference's piecewise fits; all table data comes from tools/gen_rys_tables.py (mpmath, 60 digits):
//
//   * Boys function F_m(X) from ONE 16-byte table entry {F_10(X_i), exp(-X_i)} on the grid X_i = i/16: downward
//     recursion at X_i (all terms positive: stable), an 8-term Taylor series for the highest order needed
//     (dF_m/dX = -F_{m+1}), exp(-X) = exp(-X_i) exp(X_i - X) by the same 8 terms, downward recursion at X for the
//     lower orders.  No exp(), sqrt() or division on this path and no X-dependent branch below the asymptotic
//     limit, so the lanes of a warp stay together whatever their X (the reference's seven-band fits ran with 17 of
//     32 lanes on average: profiles/r1_ncu_water154_top2_final.txt).
//   * one root:  w = F_0, r = F_1 / (F_0 - F_1).
//   * two roots: closed form from the moments F_0..F_3 (2x2 Hankel system -> monic orthogonal polynomial ->
//     quadratic), differences of products evaluated with the fma correction.
//   * 3..5 roots: degree-12 polynomials on unit intervals of X (coefficients in global memory, Horner), good to
//     2.2e-16; beyond X = 52 / 58 / 64 the Gauss-Hermite limit.
//
// PARITY MODE (default, RysTables::rys2_exact == 0).  For two roots and 15 < X <= 40 the reference evaluates
//     r_i = (a_i X + b_i) exp(-X) + R_i/(X - R_i),  w_1 = (a_w X + b_w) exp(-X) + W_1 sqrt(pi/4X),  w_0 = sqrt(pi/4X) - w_1
// (reference Rys.cpp:614-624): the classic (33,40] form, which the reference ALSO applies to 15 < X <= 33 where its
// dedicated fit is missing, so its two-root quadrature is off by up to 8.7e-7 there.  Per-quartet
// parity at 1e-12 and SCF energies at 1e-9 Eh need that behaviour, so rys2_compat_band_t2() below restates exactly
// that formula with its six fit constants; everything else in the two-root routine is ours.  rys2_exact = 1 (engine
// option "rys2_exact") uses the moment formula up to X = 46 instead and is the mathematically correct quadrature.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define UNOMOL_HD __host__ __device__ __forceinline__
#else
#define UNOMOL_HD inline
#endif

namespace ub200 {

// grid parameters (macros) and the Gauss-Hermite limits; the arrays live inside a constexpr function so that device
// code can fold them after unrolling
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr double rys_herm(int weights, int i) {
#define RYS_CONST(name, n) constexpr double name[n]
#include "rys_consts.inc"
#undef RYS_CONST
    return weights ? rys_herm_w[i] : rys_herm_r[i];
}
// ... and of 6..9 roots (rys_consts_hi.inc): row n - 6 holds n entries
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr double rys_herm_hi(int weights, int i) {
#define RYS_CONST(name, n) constexpr double name[n]
#include "rys_consts_hi.inc"
#undef RYS_CONST
    return weights ? rys_herm_hi_w[i] : rys_herm_hi_r[i];
}
template <int N>
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr double rys_herm_n(int weights, int i) { return N <= 5 ? rys_herm(weights, 5 * (N - 1) + i) : rys_herm_hi(weights, 9 * (N - 6) + i); }

// Table pointers in the caller's address space: device global memory (rys_tables.cu) or a shared-memory copy of the
// Boys grid inside kernels, the host arrays of rys_host_tables() in host code.
struct RysTables {
    const double *boys;        // [RYS_BOYS_NPTS][2] = {F_10(X_i), exp(-X_i)}, X_i = i / RYS_BOYS_HINV, X_i <= 46: two roots
    const double *boys1;       // [RYS_BOYS1_NPTS][2] = {F_8(X_i), exp(-X_i)}, X_i <= 35: one root (two recursion steps less)
    const double *boys0;       // [RYS_BOYS1_NPTS][2] = {F_7(X_i), exp(-X_i)}: F_0 alone (superseded by f0poly; kept for the tests)
    const double *f3poly;      // [RYS_F3POLY_TAB_NPTS][RYS_FP_STRIDE]: Taylor rows F_{3+j}(X_i) / j! + exp(-X_i), X_i = i / 4 <= 46: the moments
                               // F_0..F_3 of the two-root classes (boys_poly03); kernels may point this at a shared-memory copy
    const double *f3poly_glob; // the same table in global memory, whole range
    const double *expcol;      // exp(-X_i), X_i = i / 4, as expcol[i * exp_stride]: the exp(-X) of the two-root band 15 < X <= 40.  The last
                               // column of f3poly_glob (stride RYS_FP_STRIDE) or a kernel's compact shared-memory copy (stride 1)
    const double *f0poly;      // [RYS_F0POLY_TAB_NPTS][RYS_FP_STRIDE]: Taylor rows F_k(X_i) / k!, X_i = i / 4 <= 35: F_0 and F_1 of
                               // the one-root classes (boys_poly01)
    const double *piece[3];    // 3, 4, 5 roots: [interval][k = 0..12][y_0..y_{n-1}, w_0..w_{n-1}], y = t^2
    const double *piece_hi[4]; // 6..9 roots, same layout (rys_tables_hi.inc): all-Rys mode of the runtime-L kernel, the range of
                               // the reference's Rys::rootN (Rys.cpp:231-312)
    int rys2_exact;            // 0: reference-compatible two-root band (see above); 1: exact two-root quadrature
    int exp_stride;
};

constexpr double RYS_SQRT_PI_4 = 0.88622692545275801;   // sqrt(pi/4)
constexpr double RYS_X_ASYM1 = (double)RYS_BOYS1_XMAX;    // F_0 = sqrt(pi/4X) to 6e-17 beyond (exp(-35)/70 = 9e-18)

// 1/sqrt(x) and 1/x for POSITIVE NORMAL x (sums of exponents, X >= 35, discriminants): the hardware seed (MUFU.RSQ64H /
// MUFU.RCP64H, 2^-22) refined by the same steps the CUDA library routines use -- one cubic step for the square root (2^-66), two
// Newton steps for the reciprocal -- without their tests for zero / infinity / subnormal arguments, which cost seven non-FP64
// instructions and a divergence point per call (SASS of the tile kernels: ISETP + BRA + BSSY / BSYNC + three moves around the
// five DFMA / DMUL).  UNOMOL_FAST_RSQRT=0 restores the library calls.
#ifndef UNOMOL_FAST_RSQRT
#define UNOMOL_FAST_RSQRT 1
#endif
UNOMOL_HD double rys_rsqrt(double x) {
#ifdef __CUDA_ARCH__
#if UNOMOL_FAST_RSQRT
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(x, -(y * y), 1.0);
    return fma(fma(e, 0.375, 0.5), y * e, y);
#else
    return rsqrt(x);
#endif
#else
    return 1.0 / sqrt(x);
#endif
}

// a*b - c*d with one rounding of the dominant error (Kahan)
UNOMOL_HD double rys_dop(double a, double b, double c, double d) {
    const double cd = c * d;
    const double err = fma(-c, d, cd);
    return fma(a, b, -cd) + err;
}

// compile-time constants of the scaled downward recursion (see boys_grid)
constexpr double rys_tscale(int m, int mtab) {          // T_m = prod_{k=m+1..mtab} (2k-1);  f_m = h_m / T_m
    double t = 1.0;
    for (int k = m + 1; k <= mtab; ++k) t *= (double)(2 * k - 1);
    return t;
}
constexpr double rys_ifact(int k) {
    double f = 1.0;
    for (int i = 2; i <= k; ++i) f *= (double)i;
    return 1.0 / f;
}

// Constants of boys_grid<MT, MTAB> as one 32-entry block: [0, MTAB - MT) = T_{MT+1+i}, [12, 20) = q_k = 1 / (T_{MT+k} k!),
// [20, 29) = 1 / k!.  Device code reads them from a __constant__ table so that they are constant-bank operands of the
// DFMA / DMUL; as 64-bit immediates every use costs two UMOV (ncu, (ps|ss) tile kernel: 6.8 % of the issued instructions).
constexpr double boys_cval(int mt, int mtab, int i) {
    if (i < 12) return (mt + 1 + i <= mtab) ? rys_tscale(mt + 1 + i, mtab) : 0.0;
    if (i < 20) return rys_ifact(i - 12) / rys_tscale(mt + i - 12, mtab);
    if (i < 29) return rys_ifact(i - 20);
    return 0.0;
}
constexpr int boys_cblock(int mt) { return mt == 0 ? 0 : (mt == 1 ? 1 : 2); }   // the three instances in use
#ifdef __CUDACC__
#define RYS_B8(mt, mtab, o) boys_cval(mt, mtab, o), boys_cval(mt, mtab, o + 1), boys_cval(mt, mtab, o + 2), boys_cval(mt, mtab, o + 3), \
                            boys_cval(mt, mtab, o + 4), boys_cval(mt, mtab, o + 5), boys_cval(mt, mtab, o + 6), boys_cval(mt, mtab, o + 7)
#define RYS_B32(mt, mtab) RYS_B8(mt, mtab, 0), RYS_B8(mt, mtab, 8), RYS_B8(mt, mtab, 16), RYS_B8(mt, mtab, 24)
static __constant__ double rys_boys_ctab[96] = {RYS_B32(0, RYS_BOYS0_MTOP), RYS_B32(1, RYS_BOYS1_MTOP), RYS_B32(3, RYS_BOYS_MTOP)};
#undef RYS_B32
#undef RYS_B8
#endif
#ifdef __CUDA_ARCH__
#define RYS_BC(mt, mtab, i) rys_boys_ctab[32 * boys_cblock(mt) + (i)]
#else
#define RYS_BC(mt, mtab, i) boys_cval(mt, mtab, i)
#endif

// F[0..MT] = F_0(x) .. F_MT(x) for 0 <= x < (grid end); tab = {F_MTAB(X_i), exp(-X_i)} on X_i = i/16.
// The downward recursion f_{m-1} = (2 X_i f_m + e) / (2m-1) is run on h_m = T_m f_m (T_m = (2m+1)(2m+3)..(2 MTAB - 1)),
// which turns a step into ONE fma on the dependent chain, h_{m-1} = 2 X_i h_m + e T_m; the Taylor sum
// F_MT(x) = sum_k f_{MT+k} d^k / k!, d = X_i - x, runs as a Horner chain in lockstep with it (it consumes h_{MT+k} in
// the order the recursion produces them), so the critical path is MTAB - MT + 2 fma long instead of 2 (MTAB - MT) + 8:
// the FP64 pipe of these kernels was waiting on dependent results (ncu: "wait" was the top stall reason; a dependent DFMA
// issues every 8.5 cycles, profiles/exp_fp64_latency.cu).  MT = 0 needs neither exp(-x) nor the recursion at x.
template <int MT, int MTAB>
UNOMOL_HD void boys_grid(double x, const double *tab, double *F) {
    static_assert(MT + 7 <= MTAB && MTAB - MT <= 12, "the Taylor series of the top order reaches F_{MT+7}");
    static_assert((MT == 0 && MTAB == RYS_BOYS0_MTOP) || (MT == 1 && MTAB == RYS_BOYS1_MTOP) || (MT == 3 && MTAB == RYS_BOYS_MTOP),
                  "constant blocks exist for these instances");
    const int i = (int)fma(x, (double)RYS_BOYS_HINV, 0.5);
    const double xi = (double)i * (1.0 / RYS_BOYS_HINV);
    const double d = xi - x;                       // |d| <= 1/32
#ifdef __CUDA_ARCH__
    const double2 te = *reinterpret_cast<const double2 *>(tab + 2 * i);
    double h = te.x;
    const double e = te.y;
#else
    double h = tab[2 * i];
    const double e = tab[2 * i + 1];
#endif
    const double x2 = xi + xi;
#pragma unroll
    for (int m = MTAB; m > MT + 7; --m) h = fma(x2, h, e * RYS_BC(MT, MTAB, m - MT - 1));
    // h = h_{MT+7}; Horner: t <- t d + h_{MT+k} q_k with q_k = 1 / (T_{MT+k} k!)
    double t = h * RYS_BC(MT, MTAB, 12 + 7);
    double ex = fma(d, RYS_BC(MT, MTAB, 20 + 8), RYS_BC(MT, MTAB, 20 + 7));   // exp(d) = sum_k d^k / k!, same Horner direction
#pragma unroll
    for (int k = 7; k > 0; --k) {
        h = fma(x2, h, e * RYS_BC(MT, MTAB, k - 1));                          // h_{MT+k-1} = 2 X_i h_{MT+k} + e T_{MT+k}
        t = fma(t, d, h * RYS_BC(MT, MTAB, 12 + k - 1));
        if (MT > 0) ex = fma(ex, d, RYS_BC(MT, MTAB, 20 + k - 1));
    }
    F[MT] = t;
    if (MT > 0) {
        ex *= e;                                   // exp(-x)
        const double xx = x + x;
#pragma unroll
        for (int m = MT; m > 0; --m) F[m - 1] = (m == 1) ? fma(xx, F[m], ex) : fma(xx, F[m], ex) * (1.0 / (2 * m - 1));
    }
}

// F_0(x) and, with WITH_F1, F_1(x) for 0 <= x < 35 from the Taylor row of the nearest point of a coarse grid (X_i = i/4):
//   F_0(x) = sum_k a_k d^k,  a_k = F_k(X_i) / k!,  d = X_i - x, |d| <= 1/8, degree 10 ((1/8)^11 / 11! = 3e-18);
//   F_1(x) = -dF_0/dx = sum_k k a_k d^(k-1): the derivative comes out of the same Horner pass (q <- q d + p before p <- p d + a_k).
// 17 (F_0: even / odd halves in d^2, 7 dependent FP64 operations) or 31 instructions against the 48 / 65 of the recursion-based
// boys_grid<0, 7> / boys_grid<1, 8>.  This is the branch of the root evaluation that the few lanes of a warp whose ket is NEAR the
// bra take while the others wait (ncu: 17..23 % of the warp instructions of the one-root tile kernels ran there with <= 4 lanes),
// so its length is paid almost in full per primitive quartet.
#include "rys_consts_poly.inc"
// CLAMP: x may be anything (NaN included); the row index is clamped to the table and the result is then garbage for x beyond the
// grid -- for callers that evaluate this branch-free beside the asymptotic form and select afterwards (eri_tile.cuh).
template <bool WITH_F1, bool CLAMP = false>
UNOMOL_HD void boys_poly01(double x, const double *tab, double &f0, double &f1) {
    static_assert(RYS_FP_DEG == 10 && RYS_FP_STRIDE == 12, "coefficient rows of 11 + 1 doubles");
    int i = (int)fma(x, (double)RYS_FP_HINV, 0.5);
    if (CLAMP) i = i < 0 ? 0 : (i > RYS_F0POLY_TAB_NPTS - 1 ? RYS_F0POLY_TAB_NPTS - 1 : i);
    const double d = (double)i * (1.0 / RYS_FP_HINV) - x;
    double a[12];
#ifdef __CUDA_ARCH__
    const double2 *c2 = reinterpret_cast<const double2 *>(tab + i * RYS_FP_STRIDE);
#pragma unroll
    for (int k = 0; k < 6; ++k) { const double2 v = c2[k]; a[2 * k] = v.x; a[2 * k + 1] = v.y; }
#else
    for (int k = 0; k < 12; ++k) a[k] = tab[i * RYS_FP_STRIDE + k];
#endif
    if (WITH_F1) {
        double p = a[10], q = 0.0;
#pragma unroll
        for (int k = 9; k >= 0; --k) { q = fma(q, d, p); p = fma(p, d, a[k]); }
        f0 = p;
        f1 = q;
    } else {
        const double d2 = d * d;
        double e = a[10], o = a[9];
#pragma unroll
        for (int k = 8; k >= 0; k -= 2) e = fma(e, d2, a[k]);
#pragma unroll
        for (int k = 7; k >= 1; k -= 2) o = fma(o, d2, a[k]);
        f0 = fma(o, d, e);
        f1 = 0.0;
    }
}

// exp(d) for |d| <= 1/8: Taylor series of degree 10 ((1/8)^11 / 11! = 3e-18), even / odd halves in d^2
UNOMOL_HD double rys_exp_small(double d) {
    const double d2 = d * d;
    double e = 1.0 / 3628800.0, o = 1.0 / 362880.0;
    e = fma(e, d2, 1.0 / 40320.0); o = fma(o, d2, 1.0 / 5040.0);
    e = fma(e, d2, 1.0 / 720.0);   o = fma(o, d2, 1.0 / 120.0);
    e = fma(e, d2, 1.0 / 24.0);    o = fma(o, d2, 1.0 / 6.0);
    e = fma(e, d2, 0.5);           o = fma(o, d2, 1.0);
    e = fma(e, d2, 1.0);
    return fma(o, d, e);
}

// exp(-x) for 0 <= x < RYS_BOYS_XMAX from the exp(-X_i) column of the F_3 rows: exp(-x) = exp(-X_i) exp(X_i - x)
UNOMOL_HD double rys_exp_neg(double x, const double *expcol, int stride) {
    const int i = (int)fma(x, (double)RYS_FP_HINV, 0.5);
    const double d = (double)i * (1.0 / RYS_FP_HINV) - x;
    return expcol[i * stride] * rys_exp_small(d);
}

// F[0..3] = F_0(x) .. F_3(x) for 0 <= x < RYS_BOYS_XMAX: F_3 from its Taylor row (degree 10, even / odd halves), exp(-x) from the
// row's last slot, F_2, F_1, F_0 by the downward recursion at x (positive terms: stable).  About 40 instructions against the
// 100 of boys_grid<3, 10>; see boys_poly01 for why the length of this branch matters.
UNOMOL_HD void boys_poly03(double x, const double *tab, double *F) {
    const int i = (int)fma(x, (double)RYS_FP_HINV, 0.5);
    const double d = (double)i * (1.0 / RYS_FP_HINV) - x;
    double a[12];
#ifdef __CUDA_ARCH__
    const double2 *c2 = reinterpret_cast<const double2 *>(tab + i * RYS_FP_STRIDE);
#pragma unroll
    for (int k = 0; k < 6; ++k) { const double2 v = c2[k]; a[2 * k] = v.x; a[2 * k + 1] = v.y; }
#else
    for (int k = 0; k < 12; ++k) a[k] = tab[i * RYS_FP_STRIDE + k];
#endif
    const double d2 = d * d;
    double e = a[10], o = a[9];
#pragma unroll
    for (int k = 8; k >= 0; k -= 2) e = fma(e, d2, a[k]);
#pragma unroll
    for (int k = 7; k >= 1; k -= 2) o = fma(o, d2, a[k]);
    const double ex = a[11] * rys_exp_small(d);
    const double xx = x + x;
    F[3] = fma(o, d, e);
    F[2] = fma(xx, F[3], ex) * (1.0 / 5.0);
    F[1] = fma(xx, F[2], ex) * (1.0 / 3.0);
    F[0] = fma(xx, F[1], ex);
}

// ---- nodes as t^2 --------------------------------------------------------------------------------------------------
// The two-dimensional recurrences need t_i^2 (reference Rys.hpp:127: t2 = r/(1+r)); the reference's root routines return
// r = t^2/(1-t^2) only to divide it back.  rys_t2<N> returns t2[i] = t_i^2 (ascending) and w[i] directly -- the kernels
// call it -- and rys_roots<N> converts to the reference's r for the parity tests of the evaluator itself.

UNOMOL_HD double rys_rcp(double x) {
#if defined(__CUDA_ARCH__) && UNOMOL_FAST_RSQRT
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(fma(-x, y, 1.0), y, y);
    return fma(fma(-x, y, 1.0), y, y);
#else
    return 1.0 / x;
#endif
}

// Gauss-Hermite limit (all exp(-X) terms below double precision): t_i^2 = R_i / X, w_i = W_i sqrt(pi/4X)
template <int N>
UNOMOL_HD void rys_hermite_limit_t2(double x, double *t2, double *w) {
    const double rx = rys_rsqrt(x);
    const double s = RYS_SQRT_PI_4 * rx, ix = rx * rx;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        t2[i] = rys_herm_n<N>(0, i) * ix;
        w[i] = rys_herm_n<N>(1, i) * s;
    }
}

// One root as moments: w = F_0(x), f1 = F_1(x) = w t^2.  The (ps|ss) kernels use these directly.
UNOMOL_HD void rys1_f0f1(double x, double &w, double &f1, const RysTables &T) {
    if (x < RYS_X_ASYM1) {
        boys_poly01<true>(x, T.f0poly, w, f1);
    } else {
        const double rx = rys_rsqrt(x);
        w = RYS_SQRT_PI_4 * rx;
        f1 = (0.5 * w) * (rx * rx);
    }
}

UNOMOL_HD double rys1_f0(double x, const RysTables &T) {
    if (x < RYS_X_ASYM1) {
        double f0, f1;
        boys_poly01<false>(x, T.f0poly, f0, f1);
        return f0;
    }
    return RYS_SQRT_PI_4 * rys_rsqrt(x);
}

// Reference-compatible two-root band, 15 < X <= 40: restates reference Rys.cpp:614-624 (see the header comment),
//     r_i = (a_i X + b_i) exp(-X) + R_i / (X - R_i),   w_1 = (a_w X + b_w) exp(-X) + W_1 sqrt(pi/4X),   w_0 = sqrt(pi/4X) - w_1,
// evaluated for t_i^2 = r_i / (1 + r_i) = N_i / (N_i + u_i) with u_i = X - R_i, N_i = (a_i X + b_i) exp(-X) u_i + R_i: one
// reciprocal for both nodes instead of four divisions (same numbers to rounding).
UNOMOL_HD void rys2_compat_band_t2(double x, double *t2, double *w, const double *expcol, int exp_stride) {
    const double g = rys_exp_neg(x, expcol, exp_stride);      // exp(-x) to 2e-16 without the library routine's range reduction (x <= 40 here)
    const double wsum = 0.886226925452758 * rys_rsqrt(x);           // sqrt(.785398163397448 / x)
    const double u0 = x - 0.275255128608411, u1 = x - 2.72474487139158;
    const double n0 = fma(fma(-0.87894730749888, x, 10.9243702330261) * g, u0, 0.275255128608411);
    const double n1 = fma(fma(-9.28903924275977, x, 81.0642367843811) * g, u1, 2.72474487139158);
    const double d0 = n0 + u0, d1 = n1 + u1;
    const double inv = rys_rcp(d0 * d1);
    t2[0] = n0 * (d1 * inv);
    t2[1] = n1 * (d0 * inv);
    w[1] = fma(fma(4.468573893084, x, -77.9250653461045), g, 0.0917517095361369 * wsum);
    w[0] = wsum - w[1];
}

template <int N>
UNOMOL_HD void rys_t2(double x, double *t2, double *w, const RysTables &T);

template <>
UNOMOL_HD void rys_t2<1>(double x, double *t2, double *w, const RysTables &T) {
    double f1;
    rys1_f0f1(x, w[0], f1, T);
    t2[0] = f1 / w[0];
}

// Two roots below the Gauss-Hermite limit (x <= 40 in parity mode, x < RYS_BOYS_XMAX in exact mode; callers that decide "far" on
// their own -- eri_tile.cuh -- may be a rounding error beyond: the table rows and the band formula extend that far)
UNOMOL_HD void rys2_near_t2(double x, double *t2, double *w, const RysTables &T) {
    if (T.rys2_exact || x <= 15.0) {
        double m[4];
        boys_poly03(x, T.f3poly, m);
        // orthogonal polynomial D y^2 + n1 y + n0 in y = t^2 from the Hankel system [m0 m1; m1 m2] (c0, c1)^T = -(m2, m3)^T:
        // D = m0 m2 - m1^2 > 0, n0 = m1 m3 - m2^2 > 0, n1 = m1 m2 - m0 m3 < 0.  With S = sqrt(n1^2 - 4 D n0):
        // y1 = (S - n1) / (2 D), y0 = 2 n0 / (S - n1), y1 - y0 = S / D, w1 = (m1 - y0 m0) D / S.  One rsqrt (S and 1/S)
        // and one reciprocal (of D (S - n1)) instead of a square root and three divisions.
        const double D = rys_dop(m[0], m[2], m[1], m[1]);
        const double n0 = rys_dop(m[1], m[3], m[2], m[2]);
        const double n1 = rys_dop(m[1], m[2], m[0], m[3]);
        const double disc = fma(n1, n1, -4.0 * D * n0);
        const double iS = rys_rsqrt(disc), S = disc * iS;
        const double q = S - n1;                        // both terms positive
        const double inv = rys_rcp(D * q);              // 1/D = inv q, 1/q = inv D
        const double y1 = 0.5 * q * (inv * q);
        const double y0 = 2.0 * n0 * (inv * D);
        t2[0] = y0;
        t2[1] = y1;
        w[1] = fma(-y0, m[0], m[1]) * D * iS;
        w[0] = m[0] - w[1];
    } else {
        rys2_compat_band_t2(x, t2, w, T.expcol, T.exp_stride);
    }
}

template <>
UNOMOL_HD void rys_t2<2>(double x, double *t2, double *w, const RysTables &T) {
    if (T.rys2_exact ? x < (double)RYS_BOYS_XMAX : x <= 40.0) rys2_near_t2(x, t2, w, T);
    else rys_hermite_limit_t2<2>(x, t2, w);
}

template <int N>
UNOMOL_HD void rys_piecewise_t2(double x, double *t2, double *w, const double *tab, double xa) {
    if (x >= xa) { rys_hermite_limit_t2<N>(x, t2, w); return; }
    const int iv = (int)x;                                 // unit intervals
    const double s = fma(2.0, x - (double)iv, -1.0);       // [-1, 1]
    const double *c = tab + (size_t)iv * ((RYS_P3_DEG + 1) * 2 * N);
    double acc[2 * N];
#ifdef __CUDA_ARCH__
    // rows of 2N doubles start 16-byte aligned (even offsets into a 16-byte aligned table): N double2 loads per degree
    const double2 *c2 = reinterpret_cast<const double2 *>(c);
#pragma unroll
    for (int f = 0; f < N; ++f) {
        const double2 v = __ldg(c2 + RYS_P3_DEG * N + f);
        acc[2 * f] = v.x;
        acc[2 * f + 1] = v.y;
    }
#pragma unroll
    for (int k = RYS_P3_DEG - 1; k >= 0; --k)
#pragma unroll
        for (int f = 0; f < N; ++f) {
            const double2 v = __ldg(c2 + k * N + f);
            acc[2 * f] = fma(acc[2 * f], s, v.x);
            acc[2 * f + 1] = fma(acc[2 * f + 1], s, v.y);
        }
#else
#pragma unroll
    for (int f = 0; f < 2 * N; ++f) acc[f] = c[RYS_P3_DEG * 2 * N + f];
#pragma unroll
    for (int k = RYS_P3_DEG - 1; k >= 0; --k)
#pragma unroll
        for (int f = 0; f < 2 * N; ++f) acc[f] = fma(acc[f], s, c[k * 2 * N + f]);
#endif
#pragma unroll
    for (int i = 0; i < N; ++i) { t2[i] = acc[i]; w[i] = acc[N + i]; }
}
static_assert(RYS_P3_DEG == RYS_P4_DEG && RYS_P4_DEG == RYS_P5_DEG, "one degree for all piecewise tables");

template <>
UNOMOL_HD void rys_t2<3>(double x, double *t2, double *w, const RysTables &T) { rys_piecewise_t2<3>(x, t2, w, T.piece[0], RYS_P3_XA); }
template <>
UNOMOL_HD void rys_t2<4>(double x, double *t2, double *w, const RysTables &T) { rys_piecewise_t2<4>(x, t2, w, T.piece[1], RYS_P4_XA); }
template <>
UNOMOL_HD void rys_t2<5>(double x, double *t2, double *w, const RysTables &T) { rys_piecewise_t2<5>(x, t2, w, T.piece[2], RYS_P5_XA); }

// 6..9 roots: the range of the reference's general routine Rys::rootN (Rys.cpp:231-312: Boys moments -> orthogonal polynomials ->
// bracketed roots -> Christoffel weights, which crashes or hangs for 2 <~ X <~ 15).  Here the same kind of
// table as for 3..5 roots, generated from the definition of the quadrature at 80 digits.
static_assert(RYS_P6_DEG == RYS_P3_DEG && RYS_P7_DEG == RYS_P3_DEG && RYS_P8_DEG == RYS_P3_DEG && RYS_P9_DEG == RYS_P3_DEG, "one degree for all piecewise tables");
template <>
UNOMOL_HD void rys_t2<6>(double x, double *t2, double *w, const RysTables &T) { rys_piecewise_t2<6>(x, t2, w, T.piece_hi[0], RYS_P6_XA); }
template <>
UNOMOL_HD void rys_t2<7>(double x, double *t2, double *w, const RysTables &T) { rys_piecewise_t2<7>(x, t2, w, T.piece_hi[1], RYS_P7_XA); }
template <>
UNOMOL_HD void rys_t2<8>(double x, double *t2, double *w, const RysTables &T) { rys_piecewise_t2<8>(x, t2, w, T.piece_hi[2], RYS_P8_XA); }
template <>
UNOMOL_HD void rys_t2<9>(double x, double *t2, double *w, const RysTables &T) { rys_piecewise_t2<9>(x, t2, w, T.piece_hi[3], RYS_P9_XA); }

// the reference's form: r[i] = t_i^2 / (1 - t_i^2)  (Rys.hpp:145-164); used by the evaluator's own tests
template <int N>
UNOMOL_HD void rys_roots(double x, double *r, double *w, const RysTables &T) {
    rys_t2<N>(x, r, w, T);
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = r[i] / (1.0 - r[i]);
}

#if !defined(__CUDACC__) || defined(UNOMOL_RYS_HOST_TABLES)
// host copies of the tables (host emulation of the kernels, CPU tests of this evaluator)
namespace rys_host {
#define RYS_TABLE(name, n) static const double name[n]
#include "rys_tables.inc"
#include "rys_tables_hi.inc"
#include "rys_tables_poly.inc"
#undef RYS_TABLE
}  // namespace rys_host
inline RysTables rys_host_tables(int rys2_exact = 0) {
    RysTables T;
    T.boys = rys_host::rys_boys_tab;
    T.boys1 = rys_host::rys_boys1_tab;
    T.boys0 = rys_host::rys_boys0_tab;
    T.f0poly = rys_host::rys_f0poly_tab;
    T.f3poly = T.f3poly_glob = rys_host::rys_f3poly_tab;
    T.piece[0] = rys_host::rys_piece3_tab;
    T.piece[1] = rys_host::rys_piece4_tab;
    T.piece[2] = rys_host::rys_piece5_tab;
    T.piece_hi[0] = rys_host::rys_piece6_tab;
    T.piece_hi[1] = rys_host::rys_piece7_tab;
    T.piece_hi[2] = rys_host::rys_piece8_tab;
    T.piece_hi[3] = rys_host::rys_piece9_tab;
    T.expcol = T.f3poly_glob + RYS_FP_DEG + 1;
    T.exp_stride = RYS_FP_STRIDE;
    T.rys2_exact = rys2_exact;
    return T;
}
#endif

}  // namespace ub200
