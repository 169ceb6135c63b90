// unomol_b200/csrc/eri_generic.cuh -- class-templated, integral-direct ERI + J/K digestion kernel (FP64).
//
// One template instantiation per quartet class (la>=lb | lc>=ld), any class up to (dd|dd) (<= 5 Rys roots).
// Replaces, fused into one kernel and without ever storing an integral:
//   calc_two_electron_ints_rys        reference TwoElectronInts.cpp:420-509
//   Rys::Recur / RecurKernel          reference Rys.hpp:113-143,194-212
//   Rys::Shift / ShiftKernel          reference Rys.hpp:85-111,173-192
//   formGMatrixKernel / Kernel2       reference TwoElectronInts.cpp:699-820
//
// Work decomposition: a CTA owns one bra shell pair at a time (grid-stride over the Schwarz-sorted bra list);
// inside it, GROUPS of T lanes (T = 1..32, a power of two chosen per class so that a lane holds <= 32
// accumulators) each take one ket shell pair, i.e. one contracted shell quartet:
//   per primitive quartet   : every lane of the group evaluates the prefactor, the primitive cut and the Rys
//                             roots (uniform inside the group, no divergence), the 3*NROOTS (root,axis) 2-D
//                             recurrence tables are split over the group's lanes and written to shared memory,
//                             then each lane accumulates its slice of the contracted [e0|f0] integrals
//                             (e = a+b .. , f = c+d ..) in registers: 2 DFMA per element per root;
//   per contracted quartet  : horizontal transfer ket then bra through shared memory (the transfer is linear,
//                             so it is applied ONCE after contraction, not per primitive as in the reference),
//                             Cartesian norms and the 1/2 symmetry factors are applied, and the six J/K
//                             block contractions with the density are done from shared memory; results go to
//                             the square J/K accumulators with FP64 red.global.add.
// Shared memory is laid out group-interleaved (element idx of group g of a warp at idx*GPW+g) so that
// same-idx accesses of a warp are bank-conflict free.
#pragma once
#include <cuda_runtime.h>
#include "rys_roots.cuh"
#include "unomol_types.h"

namespace ub200 {

__host__ __device__ constexpr int ncart(int l) { return (l + 1) * (l + 2) / 2; }
__host__ __device__ constexpr int ncart_range(int lo, int hi) {
    int s = 0;
    for (int t = lo; t <= hi; ++t) s += ncart(t);
    return s;
}
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }

template <int LA, int LB, int LC, int LD>
struct QC {
    static_assert(LA >= LB && LC >= LD, "pairs are ordered l_first >= l_second");
    static constexpr int La = LA + LB, Lb = LC + LD, NR = (La + Lb) / 2 + 1;
    static constexpr int NE = ncart_range(LA, La), NF = ncart_range(LC, Lb);
    static constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC), ND = ncart(LD);
    static constexpr int NAB = NA * NB, NCD = NC * ND, NINT = NAB * NCD, NEF = NE * NF;
    static constexpr int T = NEF <= 32 ? 1 : NEF <= 64 ? 2 : NEF <= 128 ? 4 : NEF <= 256 ? 8 : NEF <= 512 ? 16 : 32;
    static constexpr int NACC = (NEF + T - 1) / T;
    static constexpr int GI = La + 1, GJ = Lb + 1, GSZ = GI * GJ;
    static constexpr int BUFA = cmax(NEF, NINT), BUFB = NE * NCD;
    static constexpr int GROUP_DOUBLES = NR * 3 * GSZ + BUFA + BUFB;
    static constexpr int THREADS = (T == 1) ? 64 : 128;
    static constexpr int GPW = 32 / T;                    // groups per warp
    static constexpr int GROUPS = THREADS / T;            // groups per CTA
    // + element table (NEF ints) + horizontal-transfer term tables (4 packed terms per ket / bra function pair)
    //   + per-pair Cartesian norm products
    static constexpr size_t SMEM = sizeof(double) * (GROUP_DOUBLES * GROUPS + NAB + NCD) + sizeof(int) * (NEF + 4 * NCD + 4 * NAB);
};

// Cartesian components, AuxFunctions order (lx = L..0, ly = L-lx..0), packed lx | ly<<4 | lz<<8
__device__ __forceinline__ int cart_pack(int l, int c) {
    // c -> (lx,ly,lz): row i = l - lx holds i+1 entries
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= c) ++i;
    int j = c - i * (i + 1) / 2;  // = i - ly
    int lx = l - i, ly = i - j, lz = j;
    return lx | (ly << 4) | (lz << 8);
}
__device__ __forceinline__ int cart_index(int lx, int ly, int lz) {
    int i = ly + lz;  // = l - lx
    return i * (i + 1) / 2 + lz;
}
// per-component norm 1/sqrt((2lx-1)!!(2ly-1)!!(2lz-1)!!), exact for l <= 2 (reference AuxFunctions.hpp:49-64)
__device__ __forceinline__ double cart_norm(int pk) {
    int lx = pk & 15, ly = (pk >> 4) & 15, lz = (pk >> 8) & 15;
    double d = 1.0;
    if (lx == 2) d *= 3.0;
    if (ly == 2) d *= 3.0;
    if (lz == 2) d *= 3.0;
    return d == 1.0 ? 1.0 : (d == 3.0 ? 0.5773502691896258 /* 1/sqrt(3) */ : rsqrt(d));
}
__device__ __forceinline__ double binom_small(int n, int k) {
    // n <= 2
    return (n == 2 && k == 1) ? 2.0 : 1.0;
}
__device__ __forceinline__ double ipow_small(double x, int n) { return n == 0 ? 1.0 : (n == 1 ? x : x * x); }

// index of hermite-free cartesian component (lx,ly,lz) inside the degree range [LO, ..]
template <int LO>
__device__ __forceinline__ int range_index(int lx, int ly, int lz) {
    int t = lx + ly + lz, base = 0;
    for (int s = LO; s < t; ++s) base += ncart(s);
    return base + cart_index(lx, ly, lz);
}
template <int LO, int HI>
__device__ __forceinline__ int range_unpack(int idx) {
    int t = LO;
    while (idx >= ncart(t)) { idx -= ncart(t); ++t; }
    return cart_pack(t, idx);
}

template <int LA, int LB, int LC, int LD, int MODE>
__global__ void __launch_bounds__(QC<LA, LB, LC, LD>::THREADS)
eri_class_kernel(const ClassTask task) {
    using C = QC<LA, LB, LC, LD>;
    constexpr int T = C::T, NR = C::NR, GI = C::GI, GJ = C::GJ, GSZ = C::GSZ, GPW = C::GPW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    double *nrm_ab = smem + (size_t)C::GROUP_DOUBLES * C::GROUPS;
    double *nrm_cd = nrm_ab + C::NAB;
    int *elem_tab = reinterpret_cast<int *>(nrm_cd + C::NCD);
    int *ket_terms = elem_tab + C::NEF;      // [NCD][4]: source f index | jx<<8 | jy<<10 | jz<<12 | binom<<14 ; -1 = unused
    int *bra_terms = ket_terms + 4 * C::NCD; // [NAB][4]: source e index | ix<<8 | ...

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lig = lane & (T - 1);               // lane in group
    const int gw = lane / T;                      // group in warp
    const int group = tid / T;                    // group in CTA
    const unsigned gmask = (T == 32) ? 0xffffffffu : (((1u << T) - 1u) << (lane & ~(T - 1)));
    double *sw = smem + (size_t)warp * C::GROUP_DOUBLES * GPW;   // this warp's region
#define SG(idx) sw[(idx) * GPW + gw]
    constexpr int OFF_G = 0, OFF_A = NR * 3 * GSZ, OFF_B = OFF_A + C::BUFA;

    // element table: k -> offsets of (ex,fx),(ey,fy),(ez,fz) inside a GI x GJ table
    for (int k = tid; k < C::NEF; k += blockDim.x) {
        int e = k / C::NF, f = k % C::NF;
        int pe = range_unpack<LA, C::La>(e), pf = range_unpack<LC, C::Lb>(f);
        int ox = (pe & 15) * GJ + (pf & 15), oy = ((pe >> 4) & 15) * GJ + ((pf >> 4) & 15),
            oz = ((pe >> 8) & 15) * GJ + ((pf >> 8) & 15);
        elem_tab[k] = ox | (oy << 10) | (oz << 20);
    }
    // horizontal-transfer plans, built once per CTA (reference Rys.hpp:173-192: shift from the second shell of a
    // pair onto the first): target pair (c,d) = sum over j <= d of binom(d,j) CD^j [f = c + d - j]
    for (int cd = tid; cd < C::NCD; cd += blockDim.x) {
        const int c = cd / C::ND, d = cd - c * C::ND;
        const int pc = cart_pack(LC, c), pd = cart_pack(LD, d);
        const int cx = pc & 15, cy = (pc >> 4) & 15, cz = pc >> 8, dx = pd & 15, dy = (pd >> 4) & 15, dz = pd >> 8;
        int nt = 0;
        for (int jx = 0; jx <= dx; ++jx)
            for (int jy = 0; jy <= dy; ++jy)
                for (int jz = 0; jz <= dz; ++jz) {
                    const int f = range_index<LC>(cx + dx - jx, cy + dy - jy, cz + dz - jz);
                    const int bn = (int)(binom_small(dx, jx) * binom_small(dy, jy) * binom_small(dz, jz));
                    ket_terms[cd * 4 + nt++] = f | (jx << 8) | (jy << 10) | (jz << 12) | (bn << 14);
                }
        for (; nt < 4; ++nt) ket_terms[cd * 4 + nt] = -1;
        nrm_cd[cd] = cart_norm(pc) * cart_norm(pd);
    }
    for (int ab = tid; ab < C::NAB; ab += blockDim.x) {
        const int a = ab / C::NB, b = ab - a * C::NB;
        const int pa = cart_pack(LA, a), pb = cart_pack(LB, b);
        const int ax = pa & 15, ay = (pa >> 4) & 15, az = pa >> 8, bx = pb & 15, by = (pb >> 4) & 15, bz = pb >> 8;
        int nt = 0;
        for (int ix = 0; ix <= bx; ++ix)
            for (int iy = 0; iy <= by; ++iy)
                for (int iz = 0; iz <= bz; ++iz) {
                    const int e = range_index<LA>(ax + bx - ix, ay + by - iy, az + bz - iz);
                    const int bn = (int)(binom_small(bx, ix) * binom_small(by, iy) * binom_small(bz, iz));
                    bra_terms[ab * 4 + nt++] = e | (ix << 8) | (iy << 10) | (iz << 12) | (bn << 14);
                }
        for (; nt < 4; ++nt) bra_terms[ab * 4 + nt] = -1;
        nrm_ab[ab] = cart_norm(pa) * cart_norm(pb);
    }
    __syncthreads();

    unsigned long long n_quart = 0, n_primq = 0;

    // DIGEST: every WARP takes bras one at a time -- from the shared work counter when there is one (dynamic
    // self-scheduling, heaviest bras first; across ranks when the counter is IPC-mapped peer memory), else in the
    // static snake order over ranks -- and its groups stride over that bra's Schwarz-surviving kets.
    // DUMP / SCHWARZ: outer = chunk of GROUPS explicit (bra,ket) tasks, one per group.
    // A work item is (bra, slice): the kets of one bra are dealt to bra_split warps.  Lists of small molecules have a few
    // hundred bras with thousands of kets each (no screening to speak of), far fewer bras than the GPU has warps; one warp
    // per bra left most of the machine idle there (SF6/TZ2P: 5.6 -> 9.6 ms per build when bras became per-warp work items).
    const int warps_per_cta = C::THREADS / 32;
    const int nsplit = (MODE == MODE_DIGEST && task.bra_split > 1) ? task.bra_split : 1;
    const int nitems = (MODE == MODE_DIGEST) ? task.nbra * nsplit : 0;
    const int nouter = (MODE == MODE_DIGEST) ? nitems : (task.ntask + C::GROUPS - 1) / C::GROUPS;
    int wseq = blockIdx.x * warps_per_cta + warp;   // static sequence position of this warp
    bool static_done = false;                       // lane 0's state of claim_block
    for (int outer = (MODE == MODE_DIGEST) ? 0 : blockIdx.x;; outer += (MODE == MODE_DIGEST) ? 1 : gridDim.x) {
        int bi = 0, kfirst = 0, kcount = 0, kstep = 1;
        if (MODE == MODE_DIGEST) {
            int item;
            if (task.work_counter) {
                if (lane == 0) item = (int)min(claim_block(task, static_done), (long long)INT_MAX);
                item = __shfl_sync(0xffffffffu, item, 0);
            } else {
                item = task.nranks * wseq + ((wseq & 1) ? task.nranks - 1 - task.rank : task.rank);
                wseq += gridDim.x * warps_per_cta;
            }
            if (item >= nitems) break;
            bi = item / nsplit;
            const int slice = item - bi * nsplit;
            kfirst = slice * GPW + gw; kcount = task.ket_count[bi]; kstep = nsplit * GPW;
        } else {
            if (outer >= nouter) break;
            const int t = outer * C::GROUPS + group;
            if (t < task.ntask) { bi = task.task_list[t].x; kfirst = task.task_list[t].y; kcount = kfirst + 1; }
        }
        const ShellPair bra = task.bra[bi];
        for (int ki = kfirst; ki < kcount; ki += kstep) {
            const ShellPair ket = task.ket[ki];
            if (MODE == MODE_DIGEST) {
                int imax = max(max(bra.sha, bra.shb), max(ket.sha, ket.shb));
                if (imax < task.start_shell) continue;
            }
            // ---------------- contracted [e0|f0] accumulation over primitive quartets
            double acc[C::NACC];
#pragma unroll
            for (int m = 0; m < C::NACC; ++m) acc[m] = 0.0;
            const PrimPair *bp = task.prims + bra.prim_off;
            const PrimPair *kp = task.prims + ket.prim_off;
            // primitive cut with early exits: see eri_reg.cuh (lists sorted by u descending)
            const double cut2 = task.prim_cut * task.prim_cut;
            for (int ib = 0; ib < bra.nprim; ++ib) {
                const PrimPair b = bp[ib];
                const double tb0 = SR_TERM * b.u;
                {
                    const double tq = tb0 * ket.umax;
                    if (tq * tq < cut2 * ket.pmin) break;
                    if (tq * tq < cut2 * (ket.pmin + b.p)) continue;
                }
                for (int ik = 0; ik < ket.nprim; ++ik) {
                    const PrimPair k = kp[ik];
                    const double txp = b.p + k.p;
                    const double t = tb0 * k.u;
                    if (t * t < cut2 * txp) {
                        if (t * t < cut2 * (ket.pmin + b.p)) break;
                        continue;
                    }
                    const double itx = 1.0 / txp;
                    double sr = t * sqrt(itx);
                    sr *= b.c * k.c;
                    if (lig == 0) ++n_primq;
                    const double pq0 = b.P[0] - k.P[0], pq1 = b.P[1] - k.P[1], pq2 = b.P[2] - k.P[2];
                    const double X = b.p * k.p * itx * (pq0 * pq0 + pq1 * pq1 + pq2 * pq2);
                    double rt[NR], wt[NR];
                    rys_t2<NR>(X, rt, wt, task.rys);      // rt[] = t^2
                    // 2-D recurrences: (root, axis) tables split over the lanes of the group
                    if (T > 1) __syncwarp(gmask);   // previous iteration's readers are done
                    for (int tsk = lig; tsk < 3 * NR; tsk += T) {
                        const int ir = tsk / 3, ax = tsk - 3 * ir;
                        const double dr = rt[ir];
                        const double fff = dr * itx;
                        const double B00 = 0.5 * fff;
                        const double B1 = (0.5 - B00 * k.p) * b.ip;
                        const double B1p = (0.5 - B00 * b.p) * k.ip;
                        const double pq = ax == 0 ? pq0 : (ax == 1 ? pq1 : pq2);
                        const double Cc = b.PA[ax] - k.p * pq * fff;
                        const double Cp = k.PA[ax] + b.p * pq * fff;
                        const double scale = (ax == 2) ? wt[ir] * sr : 1.0;   // weight and prefactor ride on z
                        const int g0 = OFF_G + tsk * GSZ;
                        // G[i][j] at g0 + i*GJ + j ; reference Rys.hpp:194-212
                        double gi0_prev2 = 0.0, gi0_prev = scale;   // G[i-2][0], G[i-1][0]
                        SG(g0) = scale;
                        if (GJ > 1) {
                            // row 0
                            double a0 = scale, a1 = Cp * scale;
                            SG(g0 + 1) = a1;
                            for (int j = 1; j < GJ - 1; ++j) {
                                double a2 = j * B1p * a0 + Cp * a1;
                                SG(g0 + j + 1) = a2;
                                a0 = a1; a1 = a2;
                            }
                        }
                        for (int i = 1; i < GI; ++i) {
                            // G[i][0] = (i-1) B1 G[i-2][0] + C G[i-1][0]
                            double gi0 = (i - 1) * B1 * gi0_prev2 + Cc * gi0_prev;
                            SG(g0 + i * GJ) = gi0;
                            if (GJ > 1) {
                                // G[i][1] = i B00 G[i-1][0] + C' G[i][0]
                                double a0 = gi0, a1 = i * B00 * gi0_prev + Cp * gi0;
                                SG(g0 + i * GJ + 1) = a1;
                                for (int j = 1; j < GJ - 1; ++j) {
                                    double a2 = j * B1p * a0 + i * B00 * SG(g0 + (i - 1) * GJ + j) + Cp * a1;
                                    SG(g0 + i * GJ + j + 1) = a2;
                                    a0 = a1; a1 = a2;
                                }
                            }
                            gi0_prev2 = gi0_prev; gi0_prev = gi0;
                        }
                    }
                    if (T > 1) __syncwarp(gmask);
                    // products
#pragma unroll
                    for (int m = 0; m < C::NACC; ++m) {
                        const int kel = lig + m * T;
                        if (C::NEF % T == 0 || kel < C::NEF) {
                            const int et = elem_tab[kel];
                            const int ox = et & 1023, oy = (et >> 10) & 1023, oz = et >> 20;
                            double s = 0.0;
#pragma unroll
                            for (int ir = 0; ir < NR; ++ir) {
                                const int g0 = OFF_G + ir * 3 * GSZ;
                                s = fma(SG(g0 + ox) * SG(g0 + GSZ + oy), SG(g0 + 2 * GSZ + oz), s);
                            }
                            acc[m] += s;
                        }
                    }
                }
            }
            if (lig == 0) ++n_quart;
            // ---------------- horizontal transfer, ket then bra (reference Rys.hpp:173-192, applied once)
            if (T > 1) __syncwarp(gmask);
#pragma unroll
            for (int m = 0; m < C::NACC; ++m) {
                const int kel = lig + m * T;
                if (C::NEF % T == 0 || kel < C::NEF) SG(OFF_A + kel) = acc[m];
            }
            if (T > 1) __syncwarp(gmask);
            // step 1: H1[e][c,d] = sum_j binom(d,j) CD^j E[e][f(c+d-j)]   (term plans from shared memory)
            for (int o = lig; o < C::NE * C::NCD; o += T) {
                const int e = o / C::NCD, cd = o - e * C::NCD;
                double v;
                if (LD == 0) {
                    v = SG(OFF_A + e * C::NF + cd);       // f == c, no shift
                } else {
                    v = 0.0;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int tm = ket_terms[cd * 4 + t];
                        if (tm < 0) break;
                        const double cf = (double)(tm >> 14) * ipow_small(ket.AB[0], (tm >> 8) & 3) * ipow_small(ket.AB[1], (tm >> 10) & 3) *
                                          ipow_small(ket.AB[2], (tm >> 12) & 3);
                        v = fma(cf, SG(OFF_A + e * C::NF + (tm & 255)), v);
                    }
                }
                SG(OFF_B + o) = v;
            }
            if (T > 1) __syncwarp(gmask);
            // symmetry factor of the shell quartet (digestion only)
            double sym = 1.0;
            if (MODE == MODE_DIGEST) {
                if (bra.sha == bra.shb) sym *= 0.5;
                if (ket.sha == ket.shb) sym *= 0.5;
                if (task.same_class && bra.pairid == ket.pairid) sym *= 0.5;
            }
            // step 2: V[a,b][c,d] = norm * sum_i binom(b,i) AB^i H1[e(a+b-i)][cd]
            for (int o = lig; o < C::NINT; o += T) {
                const int ab = o / C::NCD, cd = o - ab * C::NCD;
                double v;
                if (LB == 0) {
                    v = SG(OFF_B + ab * C::NCD + cd);     // NB == 1: e == a
                } else {
                    v = 0.0;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int tm = bra_terms[ab * 4 + t];
                        if (tm < 0) break;
                        const double cf = (double)(tm >> 14) * ipow_small(bra.AB[0], (tm >> 8) & 3) * ipow_small(bra.AB[1], (tm >> 10) & 3) *
                                          ipow_small(bra.AB[2], (tm >> 12) & 3);
                        v = fma(cf, SG(OFF_B + (tm & 255) * C::NCD + cd), v);
                    }
                }
                SG(OFF_A + o) = v * (nrm_ab[ab] * nrm_cd[cd] * sym);
            }
            if (T > 1) __syncwarp(gmask);

            if (MODE == MODE_DUMP) {
                const int t = outer * C::GROUPS + group;
                double *dst = task.out + task.task_out[t];
                for (int o = lig; o < C::NINT; o += T) dst[o] = SG(OFF_A + o);
            } else if (MODE == MODE_SCHWARZ) {
                const int t = outer * C::GROUPS + group;
                double mx = 0.0;
                for (int ab = lig; ab < C::NAB; ab += T) mx = fmax(mx, fabs(SG(OFF_A + ab * C::NCD + ab)));
                for (int s = T / 2; s > 0; s >>= 1) mx = fmax(mx, __shfl_xor_sync(gmask, mx, s));
                if (lig == 0) task.out[t] = sqrt(mx);
            } else {
                // blocks entirely below the reference's storage threshold |val| <= 1e-14 never reach its G (see eri_reg.cuh)
                {
                    double mx = 0.0;
                    for (int o = lig; o < C::NINT; o += T) mx = fmax(mx, fabs(SG(OFF_A + o)));
                    for (int s = T / 2; s > 0; s >>= 1) mx = fmax(mx, __shfl_xor_sync(gmask, mx, s));
                    if (mx <= task.value_cut * sym) continue;
                }
                // ---------------- J/K digestion (reference TwoElectronInts.cpp:699-820, shell-block form)
                // V at OFF_A: [a][b][c][d].  Outputs are dealt to the lanes of the group.
                const int n = task.nbf;
                const int oa = bra.offa, ob = bra.offb, oc = ket.offa, od = ket.offb;
                constexpr int NA = C::NA, NB = C::NB, NC = C::NC, ND = C::ND;
                constexpr int N_JAB = C::NAB, N_JCD = C::NCD, N_K = NA * NC + NA * ND + NB * NC + NB * ND;
                const int nout = N_JAB + N_JCD + task.nspin * N_K;
                for (int o = lig; o < nout; o += T) {
                    if (o < N_JAB) {
                        // J[a,b] += sum_cd V PJ[c,d]
                        const int a = o / NB, b = o - a * NB;
                        double s = 0.0;
                        for (int c = 0; c < NC; ++c)
                            for (int d = 0; d < ND; ++d)
                                s = fma(SG(OFF_A + o * C::NCD + c * ND + d), task.PJ[(size_t)(oc + c) * n + od + d], s);
                        atomicAdd(task.J + (size_t)(oa + a) * n + ob + b, task.jscale * s);
                    } else if (o < N_JAB + N_JCD) {
                        const int cd = o - N_JAB, c = cd / ND, d = cd - c * ND;
                        double s = 0.0;
                        for (int a = 0; a < NA; ++a)
                            for (int b = 0; b < NB; ++b)
                                s = fma(SG(OFF_A + (a * NB + b) * C::NCD + cd), task.PJ[(size_t)(oa + a) * n + ob + b], s);
                        atomicAdd(task.J + (size_t)(oc + c) * n + od + d, task.jscale * s);
                    } else {
                        int r = o - N_JAB - N_JCD;
                        const int sp = r / N_K;
                        r -= sp * N_K;
                        const double *P = task.PK[sp];
                        double *K = task.K[sp];
                        double s = 0.0;
                        if (r < NA * NC) {
                            // K[a,c] += sum_bd V P[b,d]
                            const int a = r / NC, c = r - a * NC;
                            for (int b = 0; b < NB; ++b)
                                for (int d = 0; d < ND; ++d)
                                    s = fma(SG(OFF_A + ((a * NB + b) * NC + c) * ND + d), P[(size_t)(ob + b) * n + od + d], s);
                            atomicAdd(K + (size_t)(oa + a) * n + oc + c, s);
                        } else if (r < NA * NC + NA * ND) {
                            // K[a,d] += sum_bc V P[b,c]
                            r -= NA * NC;
                            const int a = r / ND, d = r - a * ND;
                            for (int b = 0; b < NB; ++b)
                                for (int c = 0; c < NC; ++c)
                                    s = fma(SG(OFF_A + ((a * NB + b) * NC + c) * ND + d), P[(size_t)(ob + b) * n + oc + c], s);
                            atomicAdd(K + (size_t)(oa + a) * n + od + d, s);
                        } else if (r < NA * NC + NA * ND + NB * NC) {
                            // K[b,c] += sum_ad V P[a,d]
                            r -= NA * NC + NA * ND;
                            const int b = r / NC, c = r - b * NC;
                            for (int a = 0; a < NA; ++a)
                                for (int d = 0; d < ND; ++d)
                                    s = fma(SG(OFF_A + ((a * NB + b) * NC + c) * ND + d), P[(size_t)(oa + a) * n + od + d], s);
                            atomicAdd(K + (size_t)(ob + b) * n + oc + c, s);
                        } else {
                            // K[b,d] += sum_ac V P[a,c]
                            r -= NA * NC + NA * ND + NB * NC;
                            const int b = r / ND, d = r - b * ND;
                            for (int a = 0; a < NA; ++a)
                                for (int c = 0; c < NC; ++c)
                                    s = fma(SG(OFF_A + ((a * NB + b) * NC + c) * ND + d), P[(size_t)(oa + a) * n + oc + c], s);
                            atomicAdd(K + (size_t)(ob + b) * n + od + d, s);
                        }
                    }
                }
            }
            if (T > 1) __syncwarp(gmask);
        }
    }
#undef SG
    if (MODE == MODE_DIGEST && task.counters) {
        // one atomic per warp
        for (int s = 16; s > 0; s >>= 1) {
            n_quart += __shfl_xor_sync(0xffffffffu, n_quart, s);
            n_primq += __shfl_xor_sync(0xffffffffu, n_primq, s);
        }
        if (lane == 0) {
            atomicAdd(task.counters, n_quart);
            atomicAdd(task.counters + 1, n_primq);
        }
    }
}

template <int LA, int LB, int LC, int LD>
cudaError_t launch_class(const ClassTask &task, int mode, int grid, cudaStream_t stream);

#define UNOMOL_INSTANTIATE_CLASS(LA, LB, LC, LD)                                                                   \
    template <>                                                                                                    \
    cudaError_t launch_class<LA, LB, LC, LD>(const ClassTask &task, int mode, int grid, cudaStream_t stream) {      \
        using C = QC<LA, LB, LC, LD>;                                                                              \
        static bool attr_done_dev[64] = {}; /* per device: one process may drive several GPUs */                   \
        int attr_dev = 0;                                                                                          \
        cudaGetDevice(&attr_dev);                                                                                  \
        bool &attr_done = attr_done_dev[attr_dev & 63];                                                            \
        if (!attr_done) {                                                                                          \
            cudaFuncSetAttribute(eri_class_kernel<LA, LB, LC, LD, MODE_DIGEST>,                                    \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);                       \
            cudaFuncSetAttribute(eri_class_kernel<LA, LB, LC, LD, MODE_DUMP>,                                      \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);                       \
            cudaFuncSetAttribute(eri_class_kernel<LA, LB, LC, LD, MODE_SCHWARZ>,                                   \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);                       \
            attr_done = true;                                                                                      \
        }                                                                                                          \
        if (grid <= 0) return cudaSuccess;                                                                         \
        if (mode == MODE_DIGEST)                                                                                   \
            eri_class_kernel<LA, LB, LC, LD, MODE_DIGEST><<<grid, C::THREADS, C::SMEM, stream>>>(task);            \
        else if (mode == MODE_DUMP)                                                                                \
            eri_class_kernel<LA, LB, LC, LD, MODE_DUMP><<<grid, C::THREADS, C::SMEM, stream>>>(task);              \
        else                                                                                                       \
            eri_class_kernel<LA, LB, LC, LD, MODE_SCHWARZ><<<grid, C::THREADS, C::SMEM, stream>>>(task);           \
        return cudaGetLastError();                                                                                 \
    }

}  // namespace ub200
