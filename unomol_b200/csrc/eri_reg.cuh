// unomol_b200/csrc/eri_reg.cuh -- register-resident fused ERI + J/K kernel for the low-angular-momentum
// quartet classes (NE*NF <= 32 contracted [e0|f0] intermediates: (ss|ss) (ps|ss) (ps|ps) (pp|ss) (pp|ps)
// (ds|ss) (ds|ps) (dp|ss) (dd|ss)).  These classes carry >95 % of the work of s/p basis sets
// ((H2O)_n/6-31G: profiles/r1_launches_water154_generic.csv) where the generic shared-memory kernel
// (eri_generic.cuh) is bound by table traffic and index decoding.
//
// Same arithmetic as eri_generic.cuh (see the reference citations there); what changes is the mapping:
//   * one THREAD per contracted shell quartet: a CTA owns one bra pair, its threads stride over the bra's
//     Schwarz-surviving kets;
//   * the bra's ShellPair + PrimPair records are staged into shared memory by TMA bulk copies
//     (cp.async.bulk + mbarrier, double buffered so the next bra's data lands while this one is computed)
//     and read by broadcast; the ket's primitive pair is held in registers across the bra-primitive loop;
//   * 2-D recurrence tables, [e0|f0] accumulators, the horizontal transfer and the six J/K block
//     contractions are fully unrolled at compile time (component tables are constexpr functions), so they
//     live in registers;
//   * J_ab is accumulated per thread over all kets of the bra and reduced once per bra with warp shuffles
//     (one FP64 red per element per warp instead of one per quartet).
//   * one-root classes use w = F0 and w*t^2 = F1 straight from the fit (no division).
#pragma once
#include <cuda_runtime.h>
#include <utility>
#include "eri_generic.cuh"

namespace ub200 {

// ---- compile-time helpers -----------------------------------------------------------------------------
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

__host__ __device__ constexpr int c_row(int c) { int i = 0; while ((i + 1) * (i + 2) / 2 <= c) ++i; return i; }
__host__ __device__ constexpr int c_lx(int l, int c) { return l - c_row(c); }
__host__ __device__ constexpr int c_lz(int l, int c) { return c - c_row(c) * (c_row(c) + 1) / 2; }
__host__ __device__ constexpr int c_ly(int l, int c) { return c_row(c) - c_lz(l, c); }
__host__ __device__ constexpr int c_index(int lx, int ly, int lz) { return (ly + lz) * (ly + lz + 1) / 2 + lz; }
// degree / component of index idx in the concatenated range of degrees LO..HI
__host__ __device__ constexpr int r_deg(int lo, int idx) { int t = lo; while (idx >= ncart(t)) { idx -= ncart(t); ++t; } return t; }
__host__ __device__ constexpr int r_cmp(int lo, int idx) { int t = lo; while (idx >= ncart(t)) { idx -= ncart(t); ++t; } return idx; }
__host__ __device__ constexpr int r_index(int lo, int lx, int ly, int lz) {
    int base = 0;
    for (int s = lo; s < lx + ly + lz; ++s) base += ncart(s);
    return base + c_index(lx, ly, lz);
}
__host__ __device__ constexpr double c_norm(int l, int c) {
    // 1/sqrt((2lx-1)!!(2ly-1)!!(2lz-1)!!) for l <= 2
    return (c_lx(l, c) == 2 || c_ly(l, c) == 2 || c_lz(l, c) == 2) ? 0.5773502691896258 : 1.0;
}
__host__ __device__ constexpr double c_binom(int n, int k) { return (n == 2 && k == 1) ? 2.0 : 1.0; }

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a) --------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase) {
    unsigned done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(phase)
                     : "memory");
    }
}

// streaming (evict-first) 16-byte loads for data every thread reads exactly once per quartet (ket records, ket
// primitives): keeps L1 for the rows of P that the whole CTA gathers from (ncu: L1 hit rate 22 %, L2 at 73-85 % of
// its throughput on the large launches)
template <class T>
__device__ __forceinline__ T load_streaming(const T *p) {
    static_assert(sizeof(T) % 16 == 0, "16-byte multiples only");
    T out;
    const int4 *src = reinterpret_cast<const int4 *>(p);
    int4 *dst = reinterpret_cast<int4 *>(&out);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) dst[i] = __ldcs(src + i);
    return out;
}

constexpr int REG_MAX_BRA_PRIMS = 36;   // 6 x 6 primitives; larger contractions fall back to the generic kernel
constexpr int REG_THREADS = 128;           // CTA size without row staging
constexpr int REG_THREADS_ROWS = 256;      // with the bra's P rows staged in shared memory (shared by twice the threads)
constexpr int REG_ROWS_MIN_KETS = 512;     // stage rows only for bras with at least this many kets (else gather from L2)
constexpr size_t REG_ROWS_MAX_BYTES = 66 * 1024;   // rows of shells a and b: (NA+NB) * ld * 8 bytes must fit this

struct __align__(16) BraStage {
    ShellPair pair;
    PrimPair prims[REG_MAX_BRA_PRIMS];
};

// ROWS: the NA+NB rows of the exchange density P that belong to the bra's two shells are staged once per bra in
// dynamic shared memory by one TMA bulk copy each (rows are contiguous in the square matrix), so the four exchange
// gathers of every quartet (P_bd, P_bc, P_ad, P_ac) become shared-memory reads instead of L2 sectors.  ncu on the
// largest launches of (H2O)_154 showed L2 at 73-85 % of its throughput with 22-25 % L1 hit rate: those gathers and
// the exchange reds are most of the traffic.  RHF only (one spin), and only when the rows fit REG_ROWS_MAX_BYTES.
template <int LA, int LB, int LC, int LD, bool ROWS>
__global__ void __launch_bounds__(ROWS ? REG_THREADS_ROWS : REG_THREADS) eri_reg_kernel(const ClassTask task) {
    using C = QC<LA, LB, LC, LD>;
    constexpr int NR = C::NR, GI = C::GI, GJ = C::GJ, NE = C::NE, NF = C::NF, NEF = C::NEF;
    constexpr int NA = C::NA, NB = C::NB, NC = C::NC, ND = C::ND, NAB = C::NAB, NCD = C::NCD, NINT = C::NINT;
    __shared__ BraStage stage[2];
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ double jab_red[REG_THREADS_ROWS / 32][NAB];
    __shared__ __align__(8) unsigned long long row_bar;
    extern __shared__ __align__(16) double srows[];   // [(NA+NB) * ld] when ROWS
    const int nthreads = ROWS ? REG_THREADS_ROWS : REG_THREADS;
    unsigned row_phase = 0u;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&row_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    // this rank owns bras rank, rank+nranks, ...; this CTA walks them with stride gridDim.x
    // Bras are dealt to ranks in blocks of nranks, alternating direction (snake order): the lists are sorted by
    // Schwarz bound, so plain round-robin would hand rank 0 the heavier bra of every block.
    auto bra_of = [&](int j) { return task.nranks * j + ((j & 1) ? task.nranks - 1 - task.rank : task.rank); };
    auto issue = [&](int b, int s) {
        // thread 0: TMA the bra's pair record and its primitive pairs
        const ShellPair *gp = task.bra + b;
        const int np = gp->nprim;   // small scalar read; the bulk copies carry the data
        const unsigned bytes = (unsigned)(sizeof(ShellPair) + sizeof(PrimPair) * np);
        mbar_expect_tx(&bars[s], bytes);
        tma_bulk_g2s(&stage[s].pair, gp, (unsigned)sizeof(ShellPair), &bars[s]);
        tma_bulk_g2s(&stage[s].prims[0], task.prims + gp->prim_off, (unsigned)(sizeof(PrimPair) * np), &bars[s]);
    };

    // Bra sequence of this CTA.  Dynamic (work_counter != null): thread 0 claims chunks of consecutive bras with a
    // system-scope atomic on a counter shared by every CTA of every rank (the lists are sorted by Schwarz bound, so
    // work is handed out heaviest first and whoever is free takes the next chunk: work stealing inside a GPU and,
    // when the counter lives in IPC-mapped peer memory, across the GPUs of the box).  Static fallback: snake order.
    __shared__ int s_next;
    int jb = blockIdx.x, cpos = 0, cend = 0;   // thread 0's sequencer state
    bool static_done = false;
    auto advance = [&]() -> int {             // thread 0 only
        if (task.work_counter) {
            if (cpos >= cend) {
                const long long blk = claim_block(task, static_done) * task.chunk;
                cpos = (int)min(blk, (long long)INT_MAX - task.chunk);
                cend = cpos + task.chunk;
            }
            return cpos < task.nbra ? cpos++ : task.nbra;
        }
        const int b = bra_of(jb);
        jb += gridDim.x;
        return b < task.nbra ? b : task.nbra;
    };
    unsigned phase[2] = {0u, 0u};
    int s = 0;
    if (tid == 0) {
        const int b0 = advance();
        s_next = b0;
        if (b0 < task.nbra) issue(b0, 0);
    }
    __syncthreads();
    int bi = s_next;
    __syncthreads();
    unsigned long long n_quart = 0, n_primq = 0, n_cand = 0;

    while (bi < task.nbra) {
        if (tid == 0) {
            const int b1 = advance();
            s_next = b1;
            if (b1 < task.nbra) issue(b1, s ^ 1);     // prefetch the next bra while this one is computed
        }
        mbar_wait(&bars[s], phase[s]);
        phase[s] ^= 1u;
        const ShellPair &bra = stage[s].pair;
        const PrimPair *bp = stage[s].prims;
        const int nbp = bra.nprim;
        const int kcount = task.ket_count[bi];
        const double abx = bra.AB[0], aby = bra.AB[1], abz = bra.AB[2];
        const double bumax = bra.umax, pminb = bra.pmin;
        // stage the rows P[a-shell][:] and P[b-shell][:] for this bra (CTA-uniform decision)
        const bool use_rows = ROWS && kcount >= REG_ROWS_MIN_KETS;
        bool rows_ready = !use_rows;
        if (ROWS && use_rows && tid == 0) {
            const unsigned bytes_a = (unsigned)(sizeof(double) * NA * task.nbf), bytes_b = (unsigned)(sizeof(double) * NB * task.nbf);
            mbar_expect_tx(&row_bar, bytes_a + bytes_b);
            tma_bulk_g2s(srows, task.PK[0] + (size_t)bra.offa * task.nbf, bytes_a, &row_bar);
            tma_bulk_g2s(srows + (size_t)NA * task.nbf, task.PK[0] + (size_t)bra.offb * task.nbf, bytes_b, &row_bar);
        }
        const double cut2 = task.prim_cut * task.prim_cut;

        double jab[NAB];
#pragma unroll
        for (int i = 0; i < NAB; ++i) jab[i] = 0.0;

        for (int ki = tid; ki < kcount; ki += nthreads) {
            const KetHot ket = load_streaming(task.ket_hot + ki);   // 32 B per thread, coalesced
            {
                const int imax = max(max(bra.sha, bra.shb), max(ket.sha, ket.shb));
                if (imax < task.start_shell) continue;
            }
            double acc[NEF];
#pragma unroll
            for (int m = 0; m < NEF; ++m) acc[m] = 0.0;
            const PrimPair *kp = task.prims + ket.prim_off;
            // Primitive quartets.  The reference skips one when sr = SR*u_b*u_k/sqrt(p+q) < cut
            // (TwoElectronInts.cpp:478-479); tested here as (SR*u_b*u_k)^2 < cut^2*(p+q) (no rsqrt).  Both
            // primitive lists are sorted by u descending, so SR*u_b*u_k/sqrt(pmin_bra+q) bounds every later bra
            // primitive and SR*umax_bra*u_k/sqrt(pmin_bra) every later ket primitive: the scan leaves early.
            // SCAN-THEN-EVALUATE: each lane first advances (cheap, divergent) to its next surviving primitive
            // quartet, then the warp reconverges for the expensive evaluation, so the FP64 work runs with the
            // lanes that have a survivor instead of one lane at a time (ncu: 7 of 32 lanes active on the DFMAs
            // of the interleaved form).
            int ik = 0, ib = 0;
            bool have_k = false;
            PrimPair k;
            double tk = 0.0;
            const int nkp = ket.nprim;
            for (;;) {
                bool found = false;
                while (ik < nkp) {
                    if (!have_k) {
                        {   // scan fields only (first 16 bytes of the record)
                            const double2 up = *reinterpret_cast<const double2 *>(kp + ik);
                            k.u = up.x; k.p = up.y;
                        }
                        tk = SR_TERM * k.u;
                        const double tb = tk * bumax;
                        if (tb * tb < cut2 * pminb) { ik = nkp; break; }
                        if (tb * tb < cut2 * (pminb + k.p)) { ++ik; continue; }
                        have_k = true;
                        ib = 0;
                        {   // this ket primitive has at least a chance: fetch the rest of its hot 48 bytes, and the cold
                            // 32 bytes only when the ket carries angular momentum (P-C and 1/q are unused otherwise)
                            const double2 *src = reinterpret_cast<const double2 *>(kp + ik);
                            const double2 c1 = src[1], c2 = src[2];
                            k.c = c1.x; k.P[0] = c1.y; k.P[1] = c2.x; k.P[2] = c2.y;
                            if constexpr (GJ > 1) {
                                const double2 c3 = src[3], c4 = src[4];
                                k.ip = c3.x; k.PA[0] = c3.y; k.PA[1] = c4.x; k.PA[2] = c4.y;
                            }
                        }
                    }
                    while (ib < nbp) {
                        const double t = tk * bp[ib].u;
                        ++n_cand;
                        if (t * t >= cut2 * (bp[ib].p + k.p)) { found = true; break; }
                        if (t * t < cut2 * (pminb + k.p)) { ib = nbp; break; }
                        ++ib;
                    }
                    if (found) break;
                    have_k = false;
                    ++ik;
                }
                if (!found) break;
                {
                    const double bpv = bp[ib].p;
                    const double txp = bpv + k.p;
                    const double rtx = rys_rsqrt(txp);
                    const double itx = rtx * rtx;
                    double sr = tk * bp[ib].u * rtx;
                    sr *= bp[ib].c * k.c;
                    ++n_primq;
                    const double pq0 = bp[ib].P[0] - k.P[0], pq1 = bp[ib].P[1] - k.P[1], pq2 = bp[ib].P[2] - k.P[2];
                    const double X = bpv * k.p * itx * (pq0 * pq0 + pq1 * pq1 + pq2 * pq2);
                    if (task.debug_flags & 8) {
                        acc[0] += sr * X;   // timing experiment: no integral evaluation at all
                    } else if constexpr (NR == 1 && GI * GJ == 1) {
                        double w = 1.0, f1 = 0.0;
                        if (!(task.debug_flags & 2)) rys1_f0f1(X, w, f1, task.rys);
                        acc[0] = fma(sr, w, acc[0]);
                    } else if constexpr (NR == 1) {
                        // (ps|ss): one root, G[1][0] = C per axis: sr*w*C = sr*(PA*w - q/(p+q)*PQ*F1)
                        double w = 1.0, f1 = 0.0;
                        if (!(task.debug_flags & 2)) rys1_f0f1(X, w, f1, task.rys);
                        const double a = sr * w, bq = sr * f1 * k.p * itx;
                        acc[0] = fma(a, bp[ib].PA[0], fma(-bq, pq0, acc[0]));
                        acc[1] = fma(a, bp[ib].PA[1], fma(-bq, pq1, acc[1]));
                        acc[2] = fma(a, bp[ib].PA[2], fma(-bq, pq2, acc[2]));
                    } else {
                        double rt[NR], wt[NR];
                        rys_t2<NR>(X, rt, wt, task.rys);      // rt[] = t^2
#pragma unroll
                        for (int ir = 0; ir < NR; ++ir) {
                            const double dr = rt[ir];
                            const double fff = dr * itx;
                            const double B00 = 0.5 * fff;
                            const double B1 = (0.5 - B00 * k.p) * bp[ib].ip;
                            const double B1p = (0.5 - B00 * bpv) * k.ip;
                            double g[3][GI][GJ];
#pragma unroll
                            for (int ax = 0; ax < 3; ++ax) {
                                const double pq = ax == 0 ? pq0 : (ax == 1 ? pq1 : pq2);
                                const double Cc = bp[ib].PA[ax] - k.p * pq * fff;
                                const double Cp = k.PA[ax] + bpv * pq * fff;
                                const double scale = (ax == 2) ? wt[ir] * sr : 1.0;
                                g[ax][0][0] = scale;
                                if constexpr (GJ > 1) {
                                    g[ax][0][1] = Cp * scale;
#pragma unroll
                                    for (int j = 1; j < GJ - 1; ++j) g[ax][0][j + 1] = j * B1p * g[ax][0][j - 1] + Cp * g[ax][0][j];
                                }
#pragma unroll
                                for (int i = 1; i < GI; ++i) {
                                    g[ax][i][0] = (i > 1 ? (i - 1) * B1 * g[ax][i > 1 ? i - 2 : 0][0] : 0.0) + Cc * g[ax][i - 1][0];
                                    if constexpr (GJ > 1) {
                                        g[ax][i][1] = i * B00 * g[ax][i - 1][0] + Cp * g[ax][i][0];
#pragma unroll
                                        for (int j = 1; j < GJ - 1; ++j)
                                            g[ax][i][j + 1] = j * B1p * g[ax][i][j - 1] + i * B00 * g[ax][i - 1][j] + Cp * g[ax][i][j];
                                    }
                                }
                            }
                            static_for<NEF>([&](auto kk) {
                                constexpr int K = decltype(kk)::value;
                                constexpr int e = K / NF, f = K % NF;
                                constexpr int te = r_deg(LA, e), ce = r_cmp(LA, e), tf = r_deg(LC, f), cf = r_cmp(LC, f);
                                constexpr int ex = c_lx(te, ce), ey = c_ly(te, ce), ez = c_lz(te, ce);
                                constexpr int fx = c_lx(tf, cf), fy = c_ly(tf, cf), fz = c_lz(tf, cf);
                                acc[K] = fma(g[0][ex][fx] * g[1][ey][fy], g[2][ez][fz], acc[K]);
                            });
                        }
                    }
                }
                ++ib;
            }
            ++n_quart;
            // ---- horizontal transfer in registers: ket, then bra (reference Rys.hpp:173-192, once per quartet)
            double cdx = 0.0, cdy = 0.0, cdz = 0.0;
            if constexpr (LD > 0) {
                const ShellPair *kfull = task.ket + ki;
                cdx = kfull->AB[0]; cdy = kfull->AB[1]; cdz = kfull->AB[2];
            }
            double h1[NE * NCD];
            static_for<NE * NCD>([&](auto oo) {
                constexpr int O = decltype(oo)::value;
                constexpr int e = O / NCD, cd = O % NCD, c = cd / ND, d = cd % ND;
                constexpr int cx = c_lx(LC, c), cy = c_ly(LC, c), cz = c_lz(LC, c);
                constexpr int dx = c_lx(LD, d), dy = c_ly(LD, d), dz = c_lz(LD, d);
                double v = 0.0;
                static_for<dx + 1>([&](auto jx_) {
                    constexpr int jx = decltype(jx_)::value;
                    static_for<dy + 1>([&](auto jy_) {
                        constexpr int jy = decltype(jy_)::value;
                        static_for<dz + 1>([&](auto jz_) {
                            constexpr int jz = decltype(jz_)::value;
                            constexpr int f = r_index(LC, cx + dx - jx, cy + dy - jy, cz + dz - jz);
                            constexpr double bn = c_binom(dx, jx) * c_binom(dy, jy) * c_binom(dz, jz);
                            double fac = bn;
                            if constexpr (jx >= 1) fac *= cdx;
                            if constexpr (jx >= 2) fac *= cdx;
                            if constexpr (jy >= 1) fac *= cdy;
                            if constexpr (jy >= 2) fac *= cdy;
                            if constexpr (jz >= 1) fac *= cdz;
                            if constexpr (jz >= 2) fac *= cdz;
                            v = fma(fac, acc[e * NF + f], v);
                        });
                    });
                });
                h1[O] = v;
            });
            double sym = 1.0;
            if (!task.out) {
                if (bra.sha == bra.shb) sym *= 0.5;
                if (ket.sha == ket.shb) sym *= 0.5;
                if (task.same_class && bra.pairid == ket.pairid) sym *= 0.5;
            }
            double V[NINT];
            static_for<NINT>([&](auto oo) {
                constexpr int O = decltype(oo)::value;
                constexpr int ab = O / NCD, cd = O % NCD, a = ab / NB, b = ab % NB, c = cd / ND, d = cd % ND;
                constexpr int ax = c_lx(LA, a), ay = c_ly(LA, a), az = c_lz(LA, a);
                constexpr int bx = c_lx(LB, b), by = c_ly(LB, b), bz = c_lz(LB, b);
                double v = 0.0;
                static_for<bx + 1>([&](auto ix_) {
                    constexpr int ix = decltype(ix_)::value;
                    static_for<by + 1>([&](auto iy_) {
                        constexpr int iy = decltype(iy_)::value;
                        static_for<bz + 1>([&](auto iz_) {
                            constexpr int iz = decltype(iz_)::value;
                            constexpr int e = r_index(LA, ax + bx - ix, ay + by - iy, az + bz - iz);
                            constexpr double bn = c_binom(bx, ix) * c_binom(by, iy) * c_binom(bz, iz);
                            double fac = bn;
                            if constexpr (ix >= 1) fac *= abx;
                            if constexpr (ix >= 2) fac *= abx;
                            if constexpr (iy >= 1) fac *= aby;
                            if constexpr (iy >= 2) fac *= aby;
                            if constexpr (iz >= 1) fac *= abz;
                            if constexpr (iz >= 2) fac *= abz;
                            v = fma(fac, h1[e * NCD + cd], v);
                        });
                    });
                });
                constexpr double nrm = c_norm(LA, a) * c_norm(LB, b) * c_norm(LC, c) * c_norm(LD, d);
                V[O] = v * (nrm * sym);
            });
            if (task.out) {
                // dump mode (test hook): the contracted block as computed by THIS kernel, no symmetry factor
                double *dst = task.out + task.task_out[bi] + (size_t)ki * NINT;
#pragma unroll
                for (int o = 0; o < NINT; ++o) dst[o] = V[o];
                continue;
            }
            // The reference never stores an integral with |val| <= 1e-14 (TwoElectronInts.cpp:513,667-671), so such
            // integrals never reach its G.  A quartet whose whole block is below that threshold is therefore skipped
            // here before any gather / red (most Schwarz survivors of a large cluster are of this kind: the bound
            // Q_ab*Q_cd >= tau says nothing about the 1/R decay between distant charge distributions).
            {
                double vmax = 0.0;
#pragma unroll
                for (int o = 0; o < NINT; ++o) vmax = fmax(vmax, fabs(V[o]));
                if (vmax <= task.value_cut * sym) continue;
            }
            // ---- J/K digestion (reference TwoElectronInts.cpp:699-820, shell-block form)
            const int n = task.nbf;
            const int oa = bra.offa, ob = bra.offb, oc = ket.offa, od = ket.offb;
            if (!(task.debug_flags & 4)) {
                // J[a,b] += sum_cd V PJ[c,d]   (kept in registers across this bra's kets)
                double pcd[NCD];
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int d = 0; d < ND; ++d) pcd[c * ND + d] = task.PJ[(size_t)(oc + c) * n + od + d];
#pragma unroll
                for (int ab = 0; ab < NAB; ++ab) {
                    double sacc = jab[ab];
#pragma unroll
                    for (int cd = 0; cd < NCD; ++cd) sacc = fma(V[ab * NCD + cd], pcd[cd], sacc);
                    jab[ab] = sacc;
                }
                // J[c,d] += sum_ab V PJ[a,b]
                double pab[NAB];
#pragma unroll
                for (int a = 0; a < NA; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b) pab[a * NB + b] = task.PJ[(size_t)(oa + a) * n + ob + b];
#pragma unroll
                for (int cd = 0; cd < NCD; ++cd) {
                    double sacc = 0.0;
#pragma unroll
                    for (int ab = 0; ab < NAB; ++ab) sacc = fma(V[ab * NCD + cd], pab[ab], sacc);
                    atomicAdd(task.J + (size_t)(oc + cd / ND) * n + od + cd % ND, task.jscale * sacc);
                }
            }
            if (ROWS && !rows_ready) {
                mbar_wait(&row_bar, row_phase);
                rows_ready = true;
            }
            for (int sp = 0; sp < ((task.debug_flags & 5) ? 0 : task.nspin); ++sp) {
                const double *P = task.PK[sp];
                double *K = task.K[sp];
                // row bases of the two bra shells: shared memory when staged, the square matrix in global otherwise
                const double *Pa = (ROWS && use_rows) ? srows : P + (size_t)oa * n;
                const double *Pb = (ROWS && use_rows) ? srows + (size_t)NA * n : P + (size_t)ob * n;
                // K[a,c] += sum_bd V P[b,d] ; K[a,d] += sum_bc V P[b,c]
                double pbd[NB * ND], pbc[NB * NC];
#pragma unroll
                for (int b = 0; b < NB; ++b) {
#pragma unroll
                    for (int d = 0; d < ND; ++d) pbd[b * ND + d] = Pb[(size_t)b * n + od + d];
#pragma unroll
                    for (int c = 0; c < NC; ++c) pbc[b * NC + c] = Pb[(size_t)b * n + oc + c];
                }
#pragma unroll
                for (int a = 0; a < NA; ++a) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        double sacc = 0.0;
#pragma unroll
                        for (int b = 0; b < NB; ++b)
#pragma unroll
                            for (int d = 0; d < ND; ++d) sacc = fma(V[((a * NB + b) * NC + c) * ND + d], pbd[b * ND + d], sacc);
                        atomicAdd(K + (size_t)(oa + a) * n + oc + c, sacc);
                    }
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        double sacc = 0.0;
#pragma unroll
                        for (int b = 0; b < NB; ++b)
#pragma unroll
                            for (int c = 0; c < NC; ++c) sacc = fma(V[((a * NB + b) * NC + c) * ND + d], pbc[b * NC + c], sacc);
                        atomicAdd(K + (size_t)(oa + a) * n + od + d, sacc);
                    }
                }
                // K[b,c] += sum_ad V P[a,d] ; K[b,d] += sum_ac V P[a,c]
                double pad[NA * ND], pac[NA * NC];
#pragma unroll
                for (int a = 0; a < NA; ++a) {
#pragma unroll
                    for (int d = 0; d < ND; ++d) pad[a * ND + d] = Pa[(size_t)a * n + od + d];
#pragma unroll
                    for (int c = 0; c < NC; ++c) pac[a * NC + c] = Pa[(size_t)a * n + oc + c];
                }
#pragma unroll
                for (int b = 0; b < NB; ++b) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        double sacc = 0.0;
#pragma unroll
                        for (int a = 0; a < NA; ++a)
#pragma unroll
                            for (int d = 0; d < ND; ++d) sacc = fma(V[((a * NB + b) * NC + c) * ND + d], pad[a * ND + d], sacc);
                        atomicAdd(K + (size_t)(ob + b) * n + oc + c, sacc);
                    }
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        double sacc = 0.0;
#pragma unroll
                        for (int a = 0; a < NA; ++a)
#pragma unroll
                            for (int c = 0; c < NC; ++c) sacc = fma(V[((a * NB + b) * NC + c) * ND + d], pac[a * NC + c], sacc);
                        atomicAdd(K + (size_t)(ob + b) * n + od + d, sacc);
                    }
                }
            }
        }
        // ---- J_ab: one reduction per bra (warp shuffle, then one red per warp)
        int bnext;
        {
            const int n = task.nbf;
#pragma unroll
            for (int ab = 0; ab < NAB; ++ab) {
                double v = jab[ab];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) jab_red[warp][ab] = v;
            }
            __syncthreads();
            bnext = s_next;   // written by thread 0 at the top of this iteration; re-written only after the barrier below
            if (tid < NAB) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < nthreads / 32; ++w) v += jab_red[w][tid];
                if (v != 0.0) atomicAdd(task.J + (size_t)(bra.offa + tid / NB) * n + bra.offb + tid % NB, task.jscale * v);
            }
        }
        // The issuing thread makes sure this bra's row copy has landed even if none of its own quartets reached the exchange
        // step (all skipped by start_shell / the value cut): the next bra's copy reuses the buffer and the barrier phase.
        if (ROWS && use_rows && tid == 0 && !rows_ready) mbar_wait(&row_bar, row_phase);
        __syncthreads();   // everyone is done with stage[s], the staged rows and jab_red before they are overwritten
        if (ROWS && use_rows) row_phase ^= 1u;
        bi = bnext;
        s ^= 1;
    }
    if (task.counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_quart += __shfl_xor_sync(0xffffffffu, n_quart, o);
            n_primq += __shfl_xor_sync(0xffffffffu, n_primq, o);
            n_cand += __shfl_xor_sync(0xffffffffu, n_cand, o);
        }
        if (lane == 0) {
            if (task.cand_counter) atomicAdd(task.cand_counter, n_cand);
            atomicAdd(task.counters, n_quart);
            atomicAdd(task.counters + 1, n_primq);
        }
    }
}

}  // namespace ub200
