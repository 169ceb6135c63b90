// unomol_b200/csrc/unomol_types.h -- plain structs shared by host and device code.
#pragma once
#include <climits>
#include <cstdint>
#include "rys_roots.cuh"

namespace ub200 {

constexpr int MAXL = 4;            // per-shell angular momentum handled (s..g; reference Basis.hpp:222)
constexpr int NPAIRCLASS = 15;     // ss ps pp ds dp dd fs fp fd ff gs gp gd gf gg  (la >= lb; id = la*(la+1)/2 + lb)
constexpr int NSPDCLASS = 6;       // pair classes below this id (l <= 2) have class-templated kernels; a quartet with an f or g
                                   // shell goes to the runtime-L kernel (eri_highl.cuh)
constexpr int NBUCKET = 6;         // pair lists are further split by primitive-pair count (1, 2-3, 4-6, 7-12, 13-24, 25+)
constexpr int NBLOCK = 4;          // ... and by spatial block (slab of shell indices) so that for large N a launch's rows x columns
                                   // footprint in the square P/J/K matrices stays L2-resident
constexpr int NSUB = NBUCKET * NBLOCK;
constexpr int NGROUP = NPAIRCLASS * NSUB;   // group id = class * NSUB + bucket * NBLOCK + block; one launch per (bra group >= ket group)
constexpr double SR_TERM = 34.9868366552497250;  // 2*pi^(5/2), reference TwoElectronInts.cpp:427

// One primitive pair of a shell pair.  Replaces the per-quartet recomputation of p, P, P-A and
// exp(-ab|AB|^2/p) in the reference's inner loops (TwoElectronInts.cpp:439-460).  80 B, 16-B aligned.
struct __attribute__((aligned(16))) PrimPair {
    // hot 48 bytes (all a ket of s functions ever needs): scan fields first
    double u;        // exp(-a b |AB|^2 / p) / p
    double p;        // alpha_a + alpha_b
    double c;        // c_a c_b, doubled for off-diagonal primitive pairs of a same-shell pair (:444-450)
    double P[3];     // Gaussian product centre
    // cold 32 bytes: only when the pair carries angular momentum / more than one root
    double ip;       // 1/p
    double PA[3];    // P - A   (A = centre of the first, higher-l, shell)
};

// The 32 bytes of a shell pair that the ket side of a quartet reads per thread (one sector, coalesced across
// the lanes of a warp); the full ShellPair record is what the bra side stages through TMA.
struct __attribute__((aligned(16))) KetHot {
    int offa, offb;  // first basis function of shell a / b
    int prim_off;    // first PrimPair
    int nprim;
    int sha, shb;    // shell indices
    int pairid;      // canonical id
    int pad;
};

// One shell pair, first shell has l_a >= l_b.  96 B.
struct __attribute__((aligned(16))) ShellPair {
    double pmin;     // smallest primitive-pair exponent sum p of this pair (bound for the primitive cut)
    double umax;     // largest u of this pair's primitive pairs (they are stored sorted by u, descending)
    double spare;
    double AB[3];    // A - B
    double Q;        // Schwarz bound sqrt(max |(ab|ab)|)
    int offa, offb;  // first basis function of shell a / b
    int prim_off;    // first PrimPair
    int nprim;       // number of PrimPairs
    int sha, shb;    // shell indices (a = higher l; ties keep the larger index first)
    int pairid;      // canonical id  max(sh)*(max(sh)+1)/2 + min(sh)
    int pad;
};

// Work descriptor of one (bra class, ket class) launch.
struct ClassTask {
    const ShellPair *bra;     // bra pairs of this class, sorted by Q descending
    const ShellPair *ket;     // ket pairs
    const KetHot *ket_hot;    // the same list, hot fields only
    const PrimPair *prims;
    const int *ket_count;     // per bra: number of leading kets to visit (Schwarz prefix, triangular cap); tile kernel: per
                              // slot of the padded tile-ordered bra list (0 for padding slots)
    // bra-tile kernel (eri_tile.cuh): bras regrouped into tiles of <= tile_b pairs that share their first shell
    const ShellPair *tbra;    // tile-ordered copy of the bra list, TILE_MAXB slots per tile (padding slots have nprim = 0)
    const int *tile_order;    // tile ids of this launch, heaviest first
    int ntiles;               // entries of tile_order
    int tile_slices;          // work items per tile: the kets of a tile are dealt to this many CTAs in interleaved chunks of
                              // TILE_THREADS (lists with few tiles would not fill the GPU otherwise); >= 1
    int tile_b;               // bras per tile of this bra class (<= TILE_MAXB)
    int tile_maxbp;           // largest primitive-pair count of a bra (stride of the staged primitive slots)
    int stage_table;          // tile kernel: copy the Boys table of the class into shared memory (set by the launcher)
    int kslots;               // ket primitive pairs a thread keeps in its shared-memory slots (kets with more read global memory)
    const long long *ket_prefix;  // runtime-L kernel only: exclusive prefix sum of ket_count over the bras [nbra + 1]; its work
                              // items are single shell quartets (a (gg|gg) block is 50 625 integrals), not bras
    int nbra, nket;
    int same_class;           // bra and ket lists are the same list (triangular, diagonal gets 1/2)
    int start_shell;          // quartet kept iff max shell index >= start_shell
    int rank, nranks;         // static fallback: bras are dealt to ranks in snake order
    unsigned long long *work_counter;  // dynamic self-scheduling: next unclaimed bra of this launch (device-local, or on rank 0's
                              // GPU and IPC/NVLink-mapped into every rank: work stealing across the GPUs of the box); null = static
    // Static-plus-stealing split across the GPUs of the box: the launch's work is cut into blocks of `chunk`
    // items, heaviest first.  Blocks [0, static_blocks) are dealt to the ranks in turn (block B belongs to rank B % nranks: a
    // cost-weighted block-cyclic assignment, the blocks being cost-sorted) and a rank walks ITS blocks through a counter in its
    // own memory; the remaining blocks, the light tail, are claimed by whoever is free from the shared counter on rank 0's GPU.
    unsigned long long *local_counter; // null: every block comes from work_counter (single GPU, or option static_fraction = 0)
    long long static_blocks;           // multiple of nranks
    int chunk;                // items per block (bras claimed per atomic)
    int bra_split;            // generic kernel: warps that share the kets of one bra (work item = (bra, slice)); >= 1
    double prim_cut;          // reference's sr < 1e-12 cut
    double value_cut;         // reference's |val| > 1e-14 storage threshold (TwoElectronInts.cpp:513)
    // digestion
    int nbf, nspin;
    const double *PJ;         // square density for the Coulomb term: P itself (RHF, same array as PK[0]) or PA+PB (UHF)
    double jscale;            // 4 (RHF: G = 2J-K from half accumulators) or 2 (UHF), applied when J is flushed
    const double *PK[2];      // square, per spin, exchange
    double *J;                // square accumulators (upper/lower mixed; symmetrised afterwards)
    double *K[2];
    // dump / schwarz modes
    const int2 *task_list;    // explicit (bra index, ket index) list
    const long long *task_out;// output offset per task
    int ntask;
    double *out;              // dump: blocks; schwarz: one value per task
    unsigned long long *counters;  // [0] quartets evaluated, [1] primitive quartets surviving the cut
    unsigned long long *cand_counter;  // primitive-quartet candidates tested against the cut (all launches)
    RysTables rys;                 // Boys grid / piecewise Rys tables on this device (rys_tables.cu) + the two-root mode
    int debug_flags;               // profiling experiments only (results invalid): 1 skip exchange digestion, 2 skip root evaluation, 4 skip all digestion
};

enum Mode { MODE_DIGEST = 0, MODE_DUMP = 1, MODE_SCHWARZ = 2 };

#ifdef __CUDACC__
// Next block of work of a launch (see ClassTask::local_counter): this rank's next static block while there is one, then a
// block of the shared tail.  `static_done` is the caller's (one claiming thread's) state, initially false.  The caller
// compares the returned block with the number of blocks of the launch.
__device__ __forceinline__ long long claim_block(const ClassTask &task, bool &static_done) {
    if (task.local_counter && !static_done) {
        const long long q = (long long)atomicAdd(task.local_counter, 1ULL);
        const long long b = q * task.nranks + task.rank;
        if (b < task.static_blocks) return b;
        static_done = true;
    }
    return (task.local_counter ? task.static_blocks : 0) + (long long)atomicAdd_system(task.work_counter, 1ULL);
}
#endif

}  // namespace ub200
