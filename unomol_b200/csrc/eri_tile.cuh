// unomol_b200/csrc/eri_tile.cuh -- bra-tile / ket-stationary fused ERI + J/K kernel for the low-angular-momentum classes.
//
// Why.  The round-1 register kernel (eri_reg.cuh: CTA = one bra pair, thread = one ket) was bound by the LSU -> L2 path,
// not by FP64: every quartet cost ~20 scattered 32-byte load sectors (the ket's hot record and primitives, six density
// gathers) and 5..19 FP64 reds (profiles/r1_ncu_water154_top2_final.txt: L2 at 83 % of its throughput, FP64 pipe at 37 %;
// profiles/exp_red_rate.cu: the reds alone were 40 % of their measured ceiling).  Here the work item of a CTA is a TILE
// of up to 8 bra pairs (a, b_j) that share their first shell a, and a thread keeps ONE ket (c, d) for the whole tile:
//
//   * the ket's hot record and its primitive pairs are fetched once per tile (primitives into a per-thread,
//     conflict-free shared-memory slot: element (prim, field) of thread t at [(prim * NF2 + field) * T + t]);
//   * K[a,c], K[a,d] and J[c,d] are accumulated in REGISTERS over the bras of the tile and flushed once per ket,
//     P[a,c], P[a,d], P_J[c,d] are gathered once per ket; what remains per quartet is K[b_j,c], K[b_j,d] (reds) and
//     P[b_j,c], P[b_j,d] (gathers): (ss|ss) 5 -> 2 + 3/B reds, (ps|ss) 9 -> 2 + 7/B, (ps|ps) 19 -> 4 + 15/B;
//   * J[a,b_j] is accumulated per thread in shared memory (FP64 shared atomics are CAS loops; a private slot per
//     thread is a plain load/add/store) and reduced once per tile;
//   * the tile's ShellPair records, Schwarz ket counts and primitive pairs arrive by TMA bulk copies into a
//     double-buffered stage (cp.async.bulk + mbarrier), the Boys grid of rys_roots.cuh sits in shared memory.
//
// Same arithmetic as eri_reg.cuh / eri_generic.cuh (reference TwoElectronInts.cpp:420-509 for the integrals,
// :699-820 for the digestion); the lanes of a warp still hold Schwarz-neighbouring kets of one bra at any time, which
// the run-ordered walk of round 1 had given up.
#pragma once
#include "eri_reg.cuh"

namespace ub200 {

constexpr int TILE_THREADS = 128;
constexpr int TILE_MAXB = 8;       // bras per tile (slots of the padded tile-ordered bra list)
constexpr int TILE_KC_BYTES = 32;  // TILE_MAXB ints

#ifndef TILE_STAGE_EXP    // A/B switch: exp(-X_i) column of the two-root band in shared memory
#define TILE_STAGE_EXP 1
#endif

struct TileLayout {
    unsigned stage_bytes, off_jab, off_kp, off_boys, total;
};
// dynamic shared memory: [stage 0][stage 1][J_ab partials][ket primitive slots][Boys grid]
__host__ __device__ inline TileLayout tile_layout(int maxbp, int tile_b, int nab, int kslots, int nf2, int nboys) {
    TileLayout L;
    L.stage_bytes = (unsigned)(TILE_MAXB * sizeof(ShellPair) + TILE_KC_BYTES + TILE_MAXB * maxbp * sizeof(PrimPair));
    L.off_jab = 2u * L.stage_bytes;
    L.off_kp = L.off_jab + (unsigned)(tile_b * nab * TILE_THREADS * sizeof(double));
    L.off_boys = L.off_kp + (unsigned)(kslots * nf2 * TILE_THREADS * 16);
    L.total = L.off_boys + (unsigned)(nboys * 16);
    return L;
}
// Boys grid entries a class needs in shared memory (rys_roots.cuh): one root up to X = 35, two roots up to X = 15
// (parity mode) or 46 (exact mode); three and more roots read their polynomial tables from global memory
// rows of the exp(-X_i) column the reference-compatible two-root band reads (X <= 40 and a rounding error), kept compact (stride 1)
// behind the F_3 rows in parity mode: the band's exp(-X_i) came from global memory and was 7.8 % of the stall samples of (ps|ps)
constexpr int TILE_EXPCOL_N = 40 * RYS_FP_HINV + 2;
__host__ __device__ inline int tile_boys_rows(int nroots, int rys2_exact) {
    if (nroots == 1) return RYS_F0POLY_TAB_NPTS;
    if (nroots == 2) return (rys2_exact ? RYS_BOYS_XMAX : 15) * RYS_FP_HINV + 2;
    return 0;
}
__host__ __device__ inline int tile_boys_entries(int nroots, int rys2_exact) {   // 16-byte entries
    int n = tile_boys_rows(nroots, rys2_exact) * RYS_FP_STRIDE / 2;
#if TILE_STAGE_EXP
    if (nroots == 2 && !rys2_exact) n += TILE_EXPCOL_N / 2;
#endif
    return n;
}

// resident CTAs per SM the register allocation is capped for (ncu: the FP64 pipe waits on dependent results, "wait" is the top
// stall; more warps per scheduler hide that latency).  Measured alternatives in profiles/ (TILE_MINB_SET).
#ifndef TILE_MINB_SET
#define TILE_MINB_SET 0
#endif
#ifndef TILE_MINB_PPSS
#define TILE_MINB_PPSS 3
#endif
#ifndef TILE_MINB_PSPS
#define TILE_MINB_PSPS 1
#endif
// A/B switches of round-2 changes (1 = on): far / near split of the root evaluation (one reciprocal square root per far primitive
// quartet), fused right-hand side of the primitive screen, integer comparison of the block maximum with the storage threshold
#ifdef TILE_NO_STATS      // experiment: what the per-launch statistics counters cost (the bench needs them: quartets, model flops)
#define TILE_STAT(x)
#else
#define TILE_STAT(x) x
#endif
#ifndef TILE_FAR_SPLIT
#define TILE_FAR_SPLIT 1
#endif
#ifndef TILE_SCAN_FMA
#define TILE_SCAN_FMA 1
#endif
#ifndef TILE_SPEC_SCAN
#define TILE_SPEC_SCAN 1
#endif
#if TILE_SPEC_SCAN && !TILE_SCAN_FMA
#error "TILE_SPEC_SCAN uses the per-ket-primitive bounds of TILE_SCAN_FMA"
#endif
#ifndef TILE_PREFETCH_P
#define TILE_PREFETCH_P 1
#endif
#ifndef TILE_FAST_TRANSITION
#define TILE_FAST_TRANSITION 1
#endif
#if TILE_FAST_TRANSITION && !TILE_SCAN_FMA
#error "TILE_FAST_TRANSITION uses the per-ket-primitive bounds of TILE_SCAN_FMA"
#endif
#ifndef TILE_INT_VMAX
#define TILE_INT_VMAX 1
#endif
__host__ __device__ constexpr int tile_minb(int lab, int lcd, int nspin) {
#if TILE_MINB_SET == 1
    return lab == 0 ? 6 : (lab == 1 && lcd == 0) ? 5 : (lab == 1) ? 3 : (lcd == 0) ? 4 : 2;
#elif TILE_MINB_SET == 2
    return lab == 0 ? 8 : (lab == 1 && lcd == 0) ? 6 : (lab == 1) ? 4 : (lcd == 0) ? 5 : 3;
#elif TILE_MINB_SET == 3
    return lab == 0 ? 4 : (lab == 1 && lcd == 0) ? 4 : (lab == 1) ? 3 : (lcd == 0) ? 3 : 2;
#elif TILE_MINB_SET == 4
    return (lab == 1 && lcd == 0) ? 4 : 1;
#else
    // (ss|ss) needs 134 registers with the speculative screen and fits 128 without a spill: four resident CTAs instead of three
    // (360.5 against 370.3 ms per (H2O)154 build); the UHF instance of (ps|ss) needs 200 and fits the 168 of three resident
    // CTAs (UHF build 496 -> 472 ms); the RHF instance of (pp|ss) gains a third CTA at 168 registers
    // (40 bytes spilled, 353.5 -> 350.9 ms).  Every other cap tried costs more than the extra warps give back.
    return (lab == 0 && lcd == 0) ? 4 : (lab == 1 && lcd == 0) ? 3 : (lab == 2 && lcd == 0 && nspin == 1) ? TILE_MINB_PPSS : (lab == 1 && lcd == 1) ? TILE_MINB_PSPS : 1;
#endif
}

template <int LA, int LB, int LC, int LD, int NSPIN>
__global__ void __launch_bounds__(TILE_THREADS, tile_minb(LA + LB, LC + LD, NSPIN)) eri_tile_kernel(const ClassTask task) {
    using C = QC<LA, LB, LC, LD>;
    constexpr int NR = C::NR, GI = C::GI, GJ = C::GJ, NE = C::NE, NF = C::NF, NEF = C::NEF;
    constexpr int NA = C::NA, NB = C::NB, NC = C::NC, ND = C::ND, NAB = C::NAB, NCD = C::NCD, NINT = C::NINT;
    constexpr int NF2 = (GJ > 1) ? 5 : 3;     // double2 fields of a ket primitive this class reads
    constexpr int T = TILE_THREADS;
    extern __shared__ __align__(16) unsigned char tsm[];
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ int s_next;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int maxbp = task.tile_maxbp, kslots = task.kslots;
    const int nboys = task.stage_table ? tile_boys_entries(NR, task.rys.rys2_exact) : 0;   // 0: the table is read through L1
    const TileLayout lay = tile_layout(maxbp, task.tile_b, NAB, kslots, NF2, nboys);
    double *sjab = reinterpret_cast<double *>(tsm + lay.off_jab);       // [(j * NAB + ab) * T + tid]
    double2 *skp = reinterpret_cast<double2 *>(tsm + lay.off_kp);       // [(prim * NF2 + field) * T + tid]
    RysTables rys = task.rys;
    if (nboys > 0) {
        // one root: the Taylor rows of F_0 (boys_poly01); two roots: those of F_3 over the moment range (boys_poly03)
        double2 *sb = reinterpret_cast<double2 *>(tsm + lay.off_boys);
        const double2 *gb = reinterpret_cast<const double2 *>(NR == 1 ? task.rys.f0poly : task.rys.f3poly_glob);
        const int nrow2 = tile_boys_rows(NR, task.rys.rys2_exact) * RYS_FP_STRIDE / 2;
        for (int i = tid; i < nrow2; i += T) sb[i] = gb[i];
        if (NR == 1) rys.f0poly = reinterpret_cast<const double *>(sb);
        else rys.f3poly = reinterpret_cast<const double *>(sb);
#if TILE_STAGE_EXP
        if (NR == 2 && !task.rys.rys2_exact) {
            double *se = reinterpret_cast<double *>(sb + nrow2);
            for (int i = tid; i < TILE_EXPCOL_N; i += T) se[i] = task.rys.expcol[i * task.rys.exp_stride];
            rys.expcol = se;
            rys.exp_stride = 1;
        }
#endif
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto stage_pairs = [&](int s) { return reinterpret_cast<const ShellPair *>(tsm + (size_t)s * lay.stage_bytes); };
    auto stage_kc = [&](int s) { return reinterpret_cast<const int *>(tsm + (size_t)s * lay.stage_bytes + TILE_MAXB * sizeof(ShellPair)); };
    auto stage_prims = [&](int s) {
        return reinterpret_cast<const PrimPair *>(tsm + (size_t)s * lay.stage_bytes + TILE_MAXB * sizeof(ShellPair) + TILE_KC_BYTES);
    };
    // thread 0: TMA one tile (slot block t of the padded tile-ordered lists) into stage s
    auto issue = [&](int t, int s) {
        const ShellPair *gp = task.tbra + (size_t)t * TILE_MAXB;
        int np[TILE_MAXB], po[TILE_MAXB];
        unsigned bytes = (unsigned)(TILE_MAXB * sizeof(ShellPair) + TILE_KC_BYTES);
#pragma unroll
        for (int j = 0; j < TILE_MAXB; ++j) {
            np[j] = gp[j].nprim;
            po[j] = gp[j].prim_off;
            bytes += (unsigned)(np[j] * sizeof(PrimPair));
        }
        unsigned char *dst = tsm + (size_t)s * lay.stage_bytes;
        mbar_expect_tx(&bars[s], bytes);
        tma_bulk_g2s(dst, gp, (unsigned)(TILE_MAXB * sizeof(ShellPair)), &bars[s]);
        tma_bulk_g2s(dst + TILE_MAXB * sizeof(ShellPair), task.ket_count + (size_t)t * TILE_MAXB, TILE_KC_BYTES, &bars[s]);
        PrimPair *pd = reinterpret_cast<PrimPair *>(dst + TILE_MAXB * sizeof(ShellPair) + TILE_KC_BYTES);
#pragma unroll
        for (int j = 0; j < TILE_MAXB; ++j)
            if (np[j] > 0) tma_bulk_g2s(pd + (size_t)j * maxbp, task.prims + po[j], (unsigned)(np[j] * sizeof(PrimPair)), &bars[s]);
    };
    // Tile sequence of this CTA: positions in task.tile_order (heaviest first).  Dynamic: thread 0 claims the next position
    // from a counter shared by every CTA of every rank (work stealing; see eri_reg.cuh); static fallback: snake order.
    auto pos_of = [&](int j) { return task.nranks * j + ((j & 1) ? task.nranks - 1 - task.rank : task.rank); };
    int jb = blockIdx.x, cpos = 0, cend = 0;
    bool static_done = false;
    auto advance = [&]() -> int {   // thread 0 only; returns a tile id or -1
        int p;
        if (task.work_counter) {
            if (cpos >= cend) {
                const long long blk = claim_block(task, static_done) * task.chunk;
                cpos = (int)min(blk, (long long)INT_MAX - task.chunk);
                cend = cpos + task.chunk;
            }
            p = cpos++;
        } else {
            p = pos_of(jb);
            jb += gridDim.x;
        }
        return p < task.ntiles * task.tile_slices ? p : -1;
    };
    // a work item is (tile, ket slice): item w -> tile_order[w / tile_slices], kets slice, slice + tile_slices, ... in units of T
    const int nslice = task.tile_slices;
    unsigned phase[2] = {0u, 0u};
    int s = 0;
    if (tid == 0) {
        const int w0 = advance();
        s_next = w0;
        if (w0 >= 0) issue(task.tile_order[w0 / nslice], 0);
    }
    __syncthreads();
    int wi = s_next;
    __syncthreads();
    unsigned long long n_quart = 0, n_primq = 0, n_cand = 0;
    const double cut2 = task.prim_cut * task.prim_cut;
    const int n = task.nbf;
    // element (r, c) of the square matrices as a 32-bit offset (leading dimension <= 65535, enforced in build_plans): the address is
    // then one IMAD.WIDE.U32 instead of a 64-bit multiply-add chain, ~25 integer instructions per quartet in the digestion
#define EIDX(r, c) ((unsigned)(r) * (unsigned)n + (unsigned)(c))

    while (wi >= 0) {
        if (tid == 0) {
            const int w1 = advance();
            s_next = w1;
            if (w1 >= 0) issue(task.tile_order[w1 / nslice], s ^ 1);     // prefetch the next tile while this one is computed
        }
        const int ti = task.tile_order[wi / nslice], slice = wi % nslice;
        mbar_wait(&bars[s], phase[s]);
        phase[s] ^= 1u;
        const ShellPair *bras = stage_pairs(s);
        const int *kc = stage_kc(s);
        const PrimPair *bprims = stage_prims(s);
        int kmax = 0, nb = 0;
#pragma unroll
        for (int j = 0; j < TILE_MAXB; ++j) {
            kmax = max(kmax, kc[j]);
            if (bras[j].nprim > 0) nb = j + 1;
        }
        const int oa = bras[0].offa, sha = bras[0].sha;
        for (int e = 0; e < nb * NAB; ++e) sjab[e * T + tid] = 0.0;

        for (int ki = slice * T + tid; ki < kmax; ki += T * nslice) {
            const KetHot ket = load_streaming(task.ket_hot + ki);   // 32 B per thread, coalesced
            const int oc = ket.offa, od = ket.offb, nkp = ket.nprim;
            const bool dump = task.out != nullptr;
            // per-ket digestion state (registers): gathered once, flushed once per tile
            double pjcd[NCD], jcd[NCD], pac[NSPIN][NA * NC], pad[NSPIN][NA * ND], kac[NSPIN][NA * NC], kad[NSPIN][NA * ND];
            bool loaded = false, touched = false;
            // ket primitives: into this thread's shared-memory slots when they fit, else read from global memory per bra
            const double2 *kbase;
            int kpstride, kfstride;
            {
                const double2 *gsrc = reinterpret_cast<const double2 *>(task.prims + ket.prim_off);
                if (nkp <= kslots) {
                    for (int ik = 0; ik < nkp; ++ik)
#pragma unroll
                        for (int f = 0; f < NF2; ++f) skp[(ik * NF2 + f) * T + tid] = __ldcs(gsrc + ik * 5 + f);
                    kbase = skp + tid; kpstride = NF2 * T; kfstride = T;
                } else {
                    kbase = gsrc; kpstride = 5; kfstride = 1;
                }
            }
            double cdx = 0.0, cdy = 0.0, cdz = 0.0;
            if constexpr (LD > 0) {
                const ShellPair *kfull = task.ket + ki;
                cdx = kfull->AB[0]; cdy = kfull->AB[1]; cdz = kfull->AB[2];
            }
            // statistics of this ket in 32 bits (at most 8 x 36 x 36 per ket), added to the 64-bit totals once per ket: a 64-bit
            // increment per candidate and per survivor is two instructions in the innermost loops
            unsigned c_quart = 0, c_primq = 0, c_cand = 0;
            const double half = dump ? 1.0 : 0.5;
            const double ksym = (ket.sha == ket.shb) ? half : 1.0;

            for (int j = 0; j < nb; ++j) {
                if (ki >= kc[j]) continue;
                const ShellPair &bra = bras[j];
                if (max(max(sha, bra.shb), max(ket.sha, ket.shb)) < task.start_shell) continue;
                const PrimPair *bp = bprims + (size_t)j * maxbp;
                const int nbp = bra.nprim;
                const double bumax = bra.umax, pminb = bra.pmin;
                // The density elements this quartet's digestion gathers from rows b_j are independent of the integrals: they are requested
                // HERE, before the primitive loop, so that the L2 round trip of the scattered gathers (the long-scoreboard stalls of the
                // digestion, ncu: 9..18 % of the stall samples) runs in the shadow of the loop.  Only where that costs few registers (the UHF
                // instances lose more to the registers than they gain: 467 -> 485 ms), and not in dump mode (no density there).
                constexpr bool PREF_K = TILE_PREFETCH_P && NSPIN == 1 && NB * (NC + ND) <= 8;
                constexpr bool PREF_J = PREF_K && NAB <= 3;
                const int ob = bra.offb;
                double qbd[NSPIN][NB * ND], qbc[NSPIN][NB * NC], qab[NAB];
                if (PREF_K && !dump) {
#pragma unroll
                    for (int sp = 0; sp < NSPIN; ++sp) {
                        const double *P = task.PK[sp];
#pragma unroll
                        for (int b = 0; b < NB; ++b) {
#pragma unroll
                            for (int d = 0; d < ND; ++d) qbd[sp][b * ND + d] = P[EIDX(ob + b, od + d)];
#pragma unroll
                            for (int c = 0; c < NC; ++c) qbc[sp][b * NC + c] = P[EIDX(ob + b, oc + c)];
                        }
                    }
                }
                if (PREF_J && !dump) {
#pragma unroll
                    for (int a = 0; a < NA; ++a)
#pragma unroll
                        for (int b = 0; b < NB; ++b) qab[a * NB + b] = task.PJ[EIDX(oa + a, ob + b)];
                }
                double acc[NEF];
#pragma unroll
                for (int m = 0; m < NEF; ++m) acc[m] = 0.0;
                // Primitive quartets: the reference's sr < 1e-12 cut (TwoElectronInts.cpp:478-479) as (SR u_b u_k)^2 < cut^2 (p+q);
                // both lists sorted by u descending -> early exits; scan-then-evaluate as in eri_reg.cuh.
                int ik = 0, ib = 0;
                bool have_k = false;
                double ku = 0.0, kp_ = 0.0, kcf = 0.0, kP0 = 0.0, kP1 = 0.0, kP2 = 0.0, kip = 0.0, kA0 = 0.0, kA1 = 0.0, kA2 = 0.0;
                double tk = 0.0;
#if TILE_SCAN_FMA
                double ck = 0.0, ckm = 0.0;
#endif
#if TILE_SPEC_SCAN
                bool spec_ok = false, spec_end = false;
#endif
                for (;;) {
#if TILE_SPEC_SCAN
                    bool found = spec_ok;      // the candidate after the last survivor was tested while that one was evaluated
                    if (!found)
#else
                    bool found = false;
#endif
                    while (ik < nkp) {
#if TILE_FAST_TRANSITION
                        if (!have_k) {
                            // Next ket primitive.  Everything this step can need is requested at once -- the primitive's record AND the
                            // first bra candidate -- and the bound tests and the first candidate's screen are computed side by side before
                            // the first branch: one shared-memory round trip and one dependent chain instead of three in series (ncu
                            // source view: this general scan was 30 % of the stall samples of (ps|ss) with 19 % of its instructions).
                            const double2 up = kbase[ik * kpstride];
                            const double2 c1 = kbase[ik * kpstride + kfstride], c2 = kbase[ik * kpstride + 2 * kfstride];
                            double2 c3 = make_double2(0.0, 0.0), c4 = c3;
                            if constexpr (GJ > 1) { c3 = kbase[ik * kpstride + 3 * kfstride]; c4 = kbase[ik * kpstride + 4 * kfstride]; }
                            const double b0u = bp[0].u, b0p = bp[0].p;
                            tk = SR_TERM * up.x;
                            const double tb = tk * bumax, tb2 = tb * tb;
                            const double ckm_n = cut2 * (pminb + up.y), ck_n = cut2 * up.y;
                            const double t0 = tk * b0u, t02 = t0 * t0;
                            const bool pass0 = t02 >= fma(cut2, b0p, ck_n), end0 = t02 < ckm_n;
                            if (tb2 < cut2 * pminb) { ik = nkp; break; }
                            if (tb2 < ckm_n) { ++ik; continue; }
                            ku = up.x; kp_ = up.y; ckm = ckm_n; ck = ck_n;
                            kcf = c1.x; kP0 = c1.y; kP1 = c2.x; kP2 = c2.y;
                            if constexpr (GJ > 1) { kip = c3.x; kA0 = c3.y; kA1 = c4.x; kA2 = c4.y; }
                            TILE_STAT(++c_cand);
                            if (pass0) { have_k = true; ib = 0; found = true; break; }
                            if (end0) { ++ik; continue; }          // no bra primitive survives against this ket primitive
                            have_k = true;
                            ib = 1;
                        }
#else
                        if (!have_k) {
                            const double2 up = kbase[ik * kpstride];
                            ku = up.x; kp_ = up.y;
                            tk = SR_TERM * ku;
                            const double tb = tk * bumax;
                            if (tb * tb < cut2 * pminb) { ik = nkp; break; }
#if TILE_SCAN_FMA
                            ckm = cut2 * (pminb + kp_);
                            if (tb * tb < ckm) { ++ik; continue; }
                            ck = cut2 * kp_;
#else
                            if (tb * tb < cut2 * (pminb + kp_)) { ++ik; continue; }
#endif
                            have_k = true;
                            ib = 0;
                            const double2 c1 = kbase[ik * kpstride + kfstride], c2 = kbase[ik * kpstride + 2 * kfstride];
                            kcf = c1.x; kP0 = c1.y; kP1 = c2.x; kP2 = c2.y;
                            if constexpr (GJ > 1) {
                                const double2 c3 = kbase[ik * kpstride + 3 * kfstride], c4 = kbase[ik * kpstride + 4 * kfstride];
                                kip = c3.x; kA0 = c3.y; kA1 = c4.x; kA2 = c4.y;
                            }
                        }
#endif
                        while (ib < nbp) {
                            const double t = tk * bp[ib].u;
                            TILE_STAT(++c_cand);
#if TILE_SCAN_FMA
                            // (SR u_b u_k)^2 against cut^2 (p + q) as one fma on the candidate's p (cut^2 q and the list-end bound
                            // cut^2 (p_min + q) are per-ket-primitive values)
                            const double t2c = t * t;
                            if (t2c >= fma(cut2, bp[ib].p, ck)) { found = true; break; }
                            if (t2c < ckm) { ib = nbp; break; }
#else
                            if (t * t >= cut2 * (bp[ib].p + kp_)) { found = true; break; }
                            if (t * t < cut2 * (pminb + kp_)) { ib = nbp; break; }
#endif
                            ++ib;
                        }
                        if (found) break;
                        have_k = false;
                        ++ik;
                    }
                    if (!found) break;
#if TILE_SPEC_SCAN
                    {
                        // The survivors of one ket primitive are (all but) a prefix of the bra primitives, so the candidate after a survivor
                        // is the next survivor two times out of three.  Its test is independent of the evaluation below and is issued
                        // first: the screen's chain (LDS, 2 DMUL, DFMA, DSETP) then runs in the shadow of the evaluation's instead of in
                        // series with it (ncu source view: the scan was 34 % of the stall samples of (ps|ss)).  Reading one record past
                        // the bra's primitives is harmless (shared memory), the result is masked.
                        const int ibn = ib + 1;
                        const double tn = tk * bp[ibn].u;
                        const double t2n = tn * tn;
                        const bool inb = ibn < nbp;
                        spec_ok = inb && (t2n >= fma(cut2, bp[ibn].p, ck));
                        spec_end = !inb || (t2n < ckm);
                        TILE_STAT(c_cand += inb ? 1u : 0u);
                    }
#endif
                    {
                        const double bpv = bp[ib].p;
                        const double txp = bpv + kp_;
                        TILE_STAT(++c_primq);
                        const double bP0 = bp[ib].P[0], bP1 = bp[ib].P[1], bP2 = bp[ib].P[2];
                        const double pq0 = bP0 - kP0, pq1 = bP1 - kP1, pq2 = bP2 - kP2;
                        const double bPA0 = bp[ib].PA[0], bPA1 = bp[ib].PA[1], bPA2 = bp[ib].PA[2];
#if TILE_FAR_SPLIT
                        // X = p q |PQ|^2 / (p + q).  Most primitive quartets of a large molecule are FAR (X beyond the asymptotic limit
                        // of the quadrature), and there every factor 1 / (p + q) cancels: sr w_i = k0 sqrt(pi/4) W_i (p q |PQ|^2)^(-1/2),
                        // t_i^2 / (p + q) = R_i / (p q |PQ|^2).  The far branch is decided without a division (p q |PQ|^2 against
                        // x_far (p + q)) and needs ONE reciprocal square root; the near branch keeps 1/sqrt(p + q) and the tables.
                        const double k0 = (tk * bp[ib].u) * (bp[ib].c * kcf);
                        const double pqr = bpv * kp_ * (pq0 * pq0 + pq1 * pq1 + pq2 * pq2);
                        const double xfar = (NR == 1) ? RYS_X_ASYM1 : (rys.rys2_exact ? (double)RYS_BOYS_XMAX : 40.0);
                        const bool far = pqr > xfar * txp;
                        if constexpr (NR == 1 && GI * GJ == 1) {
                            if (far) {
                                acc[0] = fma(k0 * RYS_SQRT_PI_4, rys_rsqrt(pqr), acc[0]);
                            } else {
                                const double rtx = rys_rsqrt(txp);
                                double f0, f1;
                                boys_poly01<false>(pqr * (rtx * rtx), rys.f0poly, f0, f1);
                                acc[0] = fma(k0 * rtx, f0, acc[0]);
                            }
                        } else if constexpr (NR == 1) {
                            // (ps|ss): one root, G[1][0] = C per axis: sr*w*C = sr*(PA*w - q/(p+q)*PQ*F1)
                            double a, bq;
                            if (far) {
                                const double rs = rys_rsqrt(pqr);
                                a = k0 * RYS_SQRT_PI_4 * rs;              // sr F_0
                                bq = (0.5 * a) * kp_ * (rs * rs);         // sr F_1 q / (p + q),  F_1 = F_0 / 2X
                            } else {
                                const double rtx = rys_rsqrt(txp);
                                const double itx = rtx * rtx, sr = k0 * rtx;
                                double w, f1;
                                boys_poly01<true>(pqr * itx, rys.f0poly, w, f1);
                                a = sr * w;
                                bq = sr * f1 * kp_ * itx;
                            }
                            acc[0] = fma(a, bPA0, fma(-bq, pq0, acc[0]));
                            acc[1] = fma(a, bPA1, fma(-bq, pq1, acc[1]));
                            acc[2] = fma(a, bPA2, fma(-bq, pq2, acc[2]));
                        } else {
                            static_assert(NR == 2, "the tile kernels cover the classes with one and two roots");
                            double fr[NR], ws[NR];     // t_i^2 / (p + q) and sr w_i
                            if (far) {
                                const double rs = rys_rsqrt(pqr);
                                const double rs2 = rs * rs, ksr = k0 * RYS_SQRT_PI_4 * rs;
#pragma unroll
                                for (int ir = 0; ir < NR; ++ir) {
                                    fr[ir] = rys_herm_n<NR>(0, ir) * rs2;
                                    ws[ir] = rys_herm_n<NR>(1, ir) * ksr;
                                }
                            } else {
                                const double rtx = rys_rsqrt(txp);
                                const double itx = rtx * rtx, sr = k0 * rtx;
                                double rt[NR], wt[NR];
                                rys2_near_t2(pqr * itx, rt, wt, rys);      // rt[] = t^2
#pragma unroll
                                for (int ir = 0; ir < NR; ++ir) {
                                    fr[ir] = rt[ir] * itx;
                                    ws[ir] = wt[ir] * sr;
                                }
                            }
#pragma unroll
                            for (int ir = 0; ir < NR; ++ir) {
                                const double fff = fr[ir];
                                const double wsr = ws[ir];
#else
                        const double rtx = rys_rsqrt(txp);
                        const double itx = rtx * rtx;
                        double sr = tk * bp[ib].u * rtx;
                        sr *= bp[ib].c * kcf;
                        const double X = bpv * kp_ * itx * (pq0 * pq0 + pq1 * pq1 + pq2 * pq2);
                        if constexpr (NR == 1 && GI * GJ == 1) {
                            acc[0] = fma(sr, rys1_f0(X, rys), acc[0]);
                        } else if constexpr (NR == 1) {
                            // (ps|ss): one root, G[1][0] = C per axis: sr*w*C = sr*(PA*w - q/(p+q)*PQ*F1)
                            double w, f1;
                            rys1_f0f1(X, w, f1, rys);
                            const double a = sr * w, bq = sr * f1 * kp_ * itx;
                            acc[0] = fma(a, bPA0, fma(-bq, pq0, acc[0]));
                            acc[1] = fma(a, bPA1, fma(-bq, pq1, acc[1]));
                            acc[2] = fma(a, bPA2, fma(-bq, pq2, acc[2]));
                        } else {
                            double rt[NR], wt[NR];
                            rys_t2<NR>(X, rt, wt, rys);      // rt[] = t^2
#pragma unroll
                            for (int ir = 0; ir < NR; ++ir) {
                                const double fff = rt[ir] * itx;
                                const double wsr = wt[ir] * sr;
#endif
                                const double B00 = 0.5 * fff;
                                const double B1 = (0.5 - B00 * kp_) * bp[ib].ip;
                                const double B1p = (0.5 - B00 * bpv) * kip;
                                double g[3][GI][GJ];
#pragma unroll
                                for (int ax = 0; ax < 3; ++ax) {
                                    const double pq = ax == 0 ? pq0 : (ax == 1 ? pq1 : pq2);
                                    const double kA = ax == 0 ? kA0 : (ax == 1 ? kA1 : kA2);
                                    const double Cc = (ax == 0 ? bPA0 : (ax == 1 ? bPA1 : bPA2)) - kp_ * pq * fff;
                                    const double Cp = kA + bpv * pq * fff;
                                    const double scale = (ax == 2) ? wsr : 1.0;
                                    g[ax][0][0] = scale;
                                    if constexpr (GJ > 1) {
                                        g[ax][0][1] = Cp * scale;
#pragma unroll
                                        for (int jj = 1; jj < GJ - 1; ++jj) g[ax][0][jj + 1] = jj * B1p * g[ax][0][jj - 1] + Cp * g[ax][0][jj];
                                    }
#pragma unroll
                                    for (int i = 1; i < GI; ++i) {
                                        g[ax][i][0] = (i > 1 ? (i - 1) * B1 * g[ax][i > 1 ? i - 2 : 0][0] : 0.0) + Cc * g[ax][i - 1][0];
                                        if constexpr (GJ > 1) {
                                            g[ax][i][1] = i * B00 * g[ax][i - 1][0] + Cp * g[ax][i][0];
#pragma unroll
                                            for (int jj = 1; jj < GJ - 1; ++jj)
                                                g[ax][i][jj + 1] = jj * B1p * g[ax][i][jj - 1] + i * B00 * g[ax][i - 1][jj] + Cp * g[ax][i][jj];
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
#if TILE_SPEC_SCAN
                    // next candidate of the general scan: ib + 1 is a survivor (spec_ok), ends this ket primitive's run, or is skipped
                    ib = spec_ok ? ib + 1 : (spec_end ? nbp : ib + 2);
#else
                    ++ib;
#endif
                }
                TILE_STAT(++c_quart);
                // ---- horizontal transfer in registers: ket, then bra (reference Rys.hpp:173-192, once per contracted quartet)
                const double abx = bra.AB[0], aby = bra.AB[1], abz = bra.AB[2];
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
                // symmetry weight of the quartet in the triangular sums: 1/2 per pair of identical shells, 1/2 on the diagonal of a
                // same-list launch (the ket's own factor is per-ket state; dump mode takes the plain block)
                double sym = (sha == bra.shb) ? half * ksym : ksym;
                if (task.same_class && bra.pairid == ket.pairid) sym *= half;
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
                if (__builtin_expect(dump, 0)) {
                    // dump mode (test hook): the contracted block as computed by THIS kernel, no symmetry factor
                    double *dst = task.out + task.task_out[(size_t)ti * TILE_MAXB + j] + (size_t)ki * NINT;
#pragma unroll
                    for (int o = 0; o < NINT; ++o) dst[o] = V[o];
                    continue;
                }
                // blocks entirely at or below the reference's storage threshold never reach its G (TwoElectronInts.cpp:513,667-671)
                {
#if TILE_INT_VMAX
                    // max |V| against the threshold on the HIGH WORDS (|x| orders like its bit pattern: two integer instructions per
                    // element instead of three on the FP64 pipe).  A block is skipped when every high word is below the threshold's;
                    // a tie of the high words (max |V| within 2^-20 below the threshold) keeps the block -- its integrals are at
                    // the reference's 1e-14, far below the 1e-12 the parity tests resolve, and the exact comparison, if-converted
                    // by the compiler, cost 24 instructions on every quartet for a case that almost never happens.
                    const double thr = task.value_cut * sym;
                    unsigned hmax = 0u;
#pragma unroll
                    for (int o = 0; o < NINT; ++o) hmax = max(hmax, (unsigned)__double2hiint(V[o]) & 0x7fffffffu);
                    if (hmax < (unsigned)__double2hiint(thr)) continue;
#else
                    double vmax = 0.0;
#pragma unroll
                    for (int o = 0; o < NINT; ++o) vmax = fmax(vmax, fabs(V[o]));
                    if (vmax <= task.value_cut * sym) continue;
#endif
                }
                // ---- J/K digestion (reference TwoElectronInts.cpp:699-820, shell-block form)
                if (!loaded) {
                    loaded = true;
#pragma unroll
                    for (int c = 0; c < NC; ++c)
#pragma unroll
                        for (int d = 0; d < ND; ++d) {
                            pjcd[c * ND + d] = task.PJ[EIDX(oc + c, od + d)];
                            jcd[c * ND + d] = 0.0;
                        }
#pragma unroll
                    for (int sp = 0; sp < NSPIN; ++sp) {
                        const double *P = task.PK[sp];
#pragma unroll
                        for (int a = 0; a < NA; ++a) {
#pragma unroll
                            for (int c = 0; c < NC; ++c) { pac[sp][a * NC + c] = P[EIDX(oa + a, oc + c)]; kac[sp][a * NC + c] = 0.0; }
#pragma unroll
                            for (int d = 0; d < ND; ++d) { pad[sp][a * ND + d] = P[EIDX(oa + a, od + d)]; kad[sp][a * ND + d] = 0.0; }
                        }
                    }
                }
                touched = true;
                {
                    // J[a,b_j] += sum_cd V PJ[c,d]   (per-thread partials in shared memory, reduced once per tile)
#pragma unroll
                    for (int ab = 0; ab < NAB; ++ab) {
                        double sacc = 0.0;
#pragma unroll
                        for (int cd = 0; cd < NCD; ++cd) sacc = fma(V[ab * NCD + cd], pjcd[cd], sacc);
                        sjab[(j * NAB + ab) * T + tid] += sacc;
                    }
                    // J[c,d] += sum_ab V PJ[a,b_j]   (registers)
                    double pab[NAB];
#pragma unroll
                    for (int a = 0; a < NA; ++a)
#pragma unroll
                        for (int b = 0; b < NB; ++b) pab[a * NB + b] = PREF_J ? qab[a * NB + b] : task.PJ[EIDX(oa + a, ob + b)];
#pragma unroll
                    for (int cd = 0; cd < NCD; ++cd) {
                        double sacc = jcd[cd];
#pragma unroll
                        for (int ab = 0; ab < NAB; ++ab) sacc = fma(V[ab * NCD + cd], pab[ab], sacc);
                        jcd[cd] = sacc;
                    }
                }
#pragma unroll
                for (int sp = 0; sp < NSPIN; ++sp) {
                    const double *P = task.PK[sp];
                    double *K = task.K[sp];
                    // K[a,c] += sum_bd V P[b,d] ; K[a,d] += sum_bc V P[b,c]   (registers)
                    double pbd[NB * ND], pbc[NB * NC];
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
#pragma unroll
                        for (int d = 0; d < ND; ++d) pbd[b * ND + d] = PREF_K ? qbd[sp][b * ND + d] : P[EIDX(ob + b, od + d)];
#pragma unroll
                        for (int c = 0; c < NC; ++c) pbc[b * NC + c] = PREF_K ? qbc[sp][b * NC + c] : P[EIDX(ob + b, oc + c)];
                    }
#pragma unroll
                    for (int a = 0; a < NA; ++a) {
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            double sacc = kac[sp][a * NC + c];
#pragma unroll
                            for (int b = 0; b < NB; ++b)
#pragma unroll
                                for (int d = 0; d < ND; ++d) sacc = fma(V[((a * NB + b) * NC + c) * ND + d], pbd[b * ND + d], sacc);
                            kac[sp][a * NC + c] = sacc;
                        }
#pragma unroll
                        for (int d = 0; d < ND; ++d) {
                            double sacc = kad[sp][a * ND + d];
#pragma unroll
                            for (int b = 0; b < NB; ++b)
#pragma unroll
                                for (int c = 0; c < NC; ++c) sacc = fma(V[((a * NB + b) * NC + c) * ND + d], pbc[b * NC + c], sacc);
                            kad[sp][a * ND + d] = sacc;
                        }
                    }
                    // K[b_j,c] += sum_ad V P[a,d] ; K[b_j,d] += sum_ac V P[a,c]   (the reds that remain per quartet)
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            double sacc = 0.0;
#pragma unroll
                            for (int a = 0; a < NA; ++a)
#pragma unroll
                                for (int d = 0; d < ND; ++d) sacc = fma(V[((a * NB + b) * NC + c) * ND + d], pad[sp][a * ND + d], sacc);
                            atomicAdd(K + EIDX(ob + b, oc + c), sacc);
                        }
#pragma unroll
                        for (int d = 0; d < ND; ++d) {
                            double sacc = 0.0;
#pragma unroll
                            for (int a = 0; a < NA; ++a)
#pragma unroll
                                for (int c = 0; c < NC; ++c) sacc = fma(V[((a * NB + b) * NC + c) * ND + d], pac[sp][a * NC + c], sacc);
                            atomicAdd(K + EIDX(ob + b, od + d), sacc);
                        }
                    }
                }
            }
            n_quart += c_quart; n_primq += c_primq; n_cand += c_cand;
            if (touched) {
                // flush this ket's register accumulators: J[c,d], K[a,c], K[a,d]
#pragma unroll
                for (int cd = 0; cd < NCD; ++cd) atomicAdd(task.J + EIDX(oc + cd / ND, od + cd % ND), task.jscale * jcd[cd]);
#pragma unroll
                for (int sp = 0; sp < NSPIN; ++sp) {
                    double *K = task.K[sp];
#pragma unroll
                    for (int a = 0; a < NA; ++a) {
#pragma unroll
                        for (int c = 0; c < NC; ++c) atomicAdd(K + EIDX(oa + a, oc + c), kac[sp][a * NC + c]);
#pragma unroll
                        for (int d = 0; d < ND; ++d) atomicAdd(K + EIDX(oa + a, od + d), kad[sp][a * ND + d]);
                    }
                }
            }
        }
        // ---- J[a,b_j]: reduce the per-thread partials of this tile (one red per element)
        __syncthreads();
        const int wnext = s_next;   // written by thread 0 at the top of this iteration
        if (!task.out) {
            const int nel = nb * NAB;
            for (int e = warp; e < nel; e += T / 32) {
                double v = 0.0;
#pragma unroll
                for (int q = 0; q < T / 32; ++q) v += sjab[e * T + q * 32 + lane];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && v != 0.0) {
                    const int j = e / NAB, ab = e % NAB;
                    atomicAdd(task.J + EIDX(oa + ab / NB, bras[j].offb + ab % NB), task.jscale * v);
                }
            }
        }
        __syncthreads();   // everyone is done with stage[s] and the partials before they are overwritten
        wi = wnext;
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

#undef EIDX

}  // namespace ub200
