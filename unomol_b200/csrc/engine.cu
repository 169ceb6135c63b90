// unomol_b200/csrc/engine.cu -- host engine + C ABI of libunomol_b200.so (see include/unomol_b200.h).
//
// Replaces the reference's TwoElectronInts (TwoElectronInts.hpp:79-110, TwoElectronInts.cpp:511-869) with an
// integral-direct GPU Fock build:
//   create / set_geometry : shell-pair + primitive-pair tables (the quantities the reference recomputes per
//                           quartet at TwoElectronInts.cpp:439-460), pairs binned by angular class,
//                           Schwarz bounds from the diagonal quartets on the GPU, pairs sorted by bound,
//                           per-bra ket prefix counts for the Q_ab*Q_cd >= tau test;
//   fock_rhf / fock_uhf   : P packed -> square, 21 class launches of the fused ERI+digestion kernel,
//                           symmetrise + pack -> G.
// No CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <array>
#include <atomic>
#include <thread>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <functional>
#include <numeric>
#include <vector>
#include "engine.h"
#include "eri_highl.cuh"

using namespace ub200;

#define CUDA_TRY(h, expr)                                                                        \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            (h)->last_error = std::string(#expr) + ": " + cudaGetErrorString(_e);                \
            fprintf(stderr, "unomol_b200: CUDA error %s\n", (h)->last_error.c_str());            \
            return UNOMOL_E_CUDA;                                                                \
        }                                                                                        \
    } while (0)

// device allocation released on every exit path (the CUDA_TRY early returns used to leak their temporaries)
template <class T>
struct DevBuf {
    T *p = nullptr;
    bool alloc(size_t n) { return cudaMalloc(&p, sizeof(T) * std::max<size_t>(n, 1)) == cudaSuccess; }
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

namespace ub200 {

// bra-class launchers (eri_class_<n>.cu)
#define DECL_BRA(a, b)                                                                                         \
    cudaError_t launch_bra_class_##a##b(int ket_class, const ClassTask &task, int mode, int grid, cudaStream_t); \
    int groups_bra_class_##a##b(int ket_class);
DECL_BRA(0, 0) DECL_BRA(1, 0) DECL_BRA(1, 1) DECL_BRA(2, 0) DECL_BRA(2, 1) DECL_BRA(2, 2)
#undef DECL_BRA

cudaError_t launch_quartet_class(int cb, int ck, const ClassTask &task, int mode, int grid, cudaStream_t s) {
    switch (cb) {
        case 0: return launch_bra_class_00(ck, task, mode, grid, s);
        case 1: return launch_bra_class_10(ck, task, mode, grid, s);
        case 2: return launch_bra_class_11(ck, task, mode, grid, s);
        case 3: return launch_bra_class_20(ck, task, mode, grid, s);
        case 4: return launch_bra_class_21(ck, task, mode, grid, s);
        case 5: return launch_bra_class_22(ck, task, mode, grid, s);
    }
    return cudaErrorInvalidValue;
}
int class_groups_per_cta(int cb, int ck) {
    switch (cb) {
        case 0: return groups_bra_class_00(ck);
        case 1: return groups_bra_class_10(ck);
        case 2: return groups_bra_class_11(ck);
        case 3: return groups_bra_class_20(ck);
        case 4: return groups_bra_class_21(ck);
        case 5: return groups_bra_class_22(ck);
    }
    return 1;
}

// Any (bra pair class, ket pair class): class-templated kernels for s/p/d, the runtime-L kernel as soon as one shell is f or g.
static cudaError_t ensure_hl_scratch(unomol_b200 *h) {
    if (h->d_hl_scratch) return cudaSuccess;
    const long long nc = (h->basis.maxl + 1) * (h->basis.maxl + 2) / 2;
    h->hl_slab = nc * nc * nc * nc;
    return cudaMalloc(&h->d_hl_scratch, sizeof(double) * (size_t)h->hl_slab * unomol_b200::HL_GRID);
}
static inline bool is_highl(int cb, int ck) { return cb >= NSPDCLASS || ck >= NSPDCLASS; }
cudaError_t launch_any_class(unomol_b200 *h, int cb, int ck, const ClassTask &task, int mode, int grid, cudaStream_t s) {
    if (!is_highl(cb, ck)) return launch_quartet_class(cb, ck, task, mode, grid, s);
    cudaError_t e = ensure_hl_scratch(h);
    if (e != cudaSuccess) return e;
    HighLArgs hl;
    pair_class_l(cb, hl.la, hl.lb);
    pair_class_l(ck, hl.lc, hl.ld);
    hl.scratch = h->d_hl_scratch;
    hl.slab = h->hl_slab;
    hl.rys = h->rys;
    hl.all_rys = h->all_rys;
    return launch_highl(task, hl, mode, std::min(grid, unomol_b200::HL_GRID), s);
}
static int any_groups_per_cta(int cb, int ck) { return is_highl(cb, ck) ? 1 : class_groups_per_cta(cb, ck); }

// algorithmic FLOPs of the reference's Rys algorithm per PRIMITIVE quartet.
double model_flops_per_primitive_quartet(int la, int lb, int lc, int ld) {
    const int La = la + lb, Lb = lc + ld, n = (La + Lb) / 2 + 1;
    auto mx = [](int a) { return a > 0 ? a : 0; };
    double R = (La < 2 && Lb < 2) ? 2.0 : 2.0 + 10.0 * mx(Lb - 1) + mx(La - 1) * (8.0 + 7.0 * mx(Lb - 1));
    auto S = [](int a, int b) { return 4.0 * (a + 1) * (b + 1) + 2.0 * (a + 1); };
    double F = 40.0 + (48.0 * n + 10.0) + 3.0 * n * R;
    auto comps = [](int l, std::vector<std::array<int, 3>> &v) {
        for (int lx = l; lx >= 0; --lx)
            for (int ly = l - lx; ly >= 0; --ly) v.push_back({lx, ly, l - lx - ly});
    };
    std::vector<std::array<int, 3>> cb, cd;
    comps(lb, cb);
    comps(ld, cd);
    const int na = (la + 1) * (la + 2) / 2, nc = (lc + 1) * (lc + 2) / 2;
    double comp = 0.0;
    for (auto &b : cb)
        for (auto &d : cd) comp += n * (S(b[0], d[0]) + S(b[1], d[1]) + S(b[2], d[2]) + 4.0) + 2.0;
    F += comp * na * nc;
    return F;
}

// ---------------------------------------------------------------- small utility kernels
__global__ void unpack_density_kernel(const double *__restrict__ PA, const double *__restrict__ PB, int n, int ld, int nspin,
                                      double *__restrict__ PJ, double *__restrict__ PK0, double *__restrict__ PK1) {
    // packed lower-triangular -> square symmetric with leading dimension ld (even, so that every row is 16-byte
    // aligned for TMA row copies).  The Coulomb scale (4 for RHF: G = 2J-K with J,K symmetrised from half
    // accumulators; 2 for UHF with PJ = PA+PB) is applied by the kernels when J is flushed (ClassTask::jscale).
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    int i = (int)(idx / n), j = (int)(idx % n);
    int a = i > j ? i : j, b = i > j ? j : i;
    size_t p = (size_t)a * (a + 1) / 2 + b;
    size_t o = (size_t)i * ld + j;
    double pa = PA[p];
    if (nspin == 1) {
        PK0[o] = pa;          // RHF: the Coulomb term reads the same array (one N^2 matrix less in L2)
    } else {
        double pb = PB[p];
        PJ[o] = pa + pb;
        PK0[o] = pa;
        PK1[o] = pb;
    }
}

__global__ void pack_fock_kernel(const double *__restrict__ J, const double *__restrict__ K0,
                                 const double *__restrict__ K1, int n, int ld, int nspin, double *__restrict__ G0,
                                 double *__restrict__ G1) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t no2 = (size_t)n * (n + 1) / 2;
    if (idx >= no2) return;
    // invert idx = i(i+1)/2 + j
    int i = (int)floor((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while ((size_t)i * (i + 1) / 2 > idx) --i;
    while ((size_t)(i + 1) * (i + 2) / 2 <= idx) ++i;
    int j = (int)(idx - (size_t)i * (i + 1) / 2);
    size_t ij = (size_t)i * ld + j, ji = (size_t)j * ld + i;
    double jj = J[ij] + J[ji];
    G0[idx] = jj - (K0[ij] + K0[ji]);
    if (nspin == 2) G1[idx] = jj - (K1[ij] + K1[ji]);
}

}  // namespace ub200

// ---------------------------------------------------------------- pair tables
static void free_pairs(unomol_b200 *h) {
    for (int c = 0; c < NGROUP; ++c) {
        if (h->cls[c].d_pairs) cudaFree(h->cls[c].d_pairs);
        if (h->cls[c].d_hot) cudaFree(h->cls[c].d_hot);
        if (h->cls[c].d_tpairs) cudaFree(h->cls[c].d_tpairs);
        h->cls[c].d_pairs = nullptr;
        h->cls[c].d_hot = nullptr;
        h->cls[c].d_tpairs = nullptr;
        h->cls[c].cap_pairs = h->cls[c].cap_tp = 0;
        h->cls[c].slot_pos.clear();
        h->cls[c].pos_slot.clear();
        h->cls[c].ntiles = 0;
        h->cls[c].pairs.clear();
        h->cls[c].n = 0;
    }
    if (h->d_prims) cudaFree(h->d_prims);
    h->d_prims = nullptr;
    h->plans.clear();          // their arrays live in d_plan_pool
    h->pairs_ready = false;
}

static int build_plans(unomol_b200 *h);

// group (= pair list) of a shell pair: angular class x primitive-count bucket x spatial block
static int group_of_pair(const unomol_b200 *h, const ShellPair &sp, int cls) {
    // lanes of a warp take different kets of one list: keep their primitive loop lengths similar (bucket);
    // spatial block = slab of shell indices of the pair's larger shell (inputs list atoms, hence shells, in spatial order)
    const int np = sp.nprim, ns = h->basis.nshell;
    const int bucket = !h->bucketed ? 0 : (np <= 1 ? 0 : np <= 3 ? 1 : np <= 6 ? 2 : np <= 12 ? 3 : np <= 24 ? 4 : 5);
    const int block = std::min(h->nblock - 1, (int)((long long)std::max(sp.sha, sp.shb) * h->nblock / ns));
    return cls * NSUB + bucket * NBLOCK + block;
}

// One shell pair (i >= j) on the host: ShellPair record + its primitive pairs with the exact prune, sorted by u descending.
// Same arithmetic as pair_device.cu.  Returns false when the pair is pruned.
static bool make_pair_host(const HostBasis &B, int i, int j, const std::vector<double> &amin, double umax, double prim_cut,
                           ShellPair &sp, std::vector<PrimPair> &keep) {
    int a = i, b = j;   // first shell = higher l (reference swaps so that l1 >= l2, TwoElectronInts.cpp:563-580)
    if (B.lv[i] < B.lv[j]) { a = j; b = i; }
    const double *A = &B.xyz[3 * B.cen[a]], *Bc = &B.xyz[3 * B.cen[b]];
    double ab2 = 0.0;
    for (int x = 0; x < 3; ++x) {
        sp.AB[x] = A[x] - Bc[x];
        ab2 += sp.AB[x] * sp.AB[x];
    }
    {
        const double pm = amin[a] + amin[b], mu = amin[a] * amin[b] / pm;
        if (SR_TERM * std::exp(-mu * ab2) / pm * umax / std::sqrt(pm) * 1.0000001 < prim_cut) return false;
    }
    const bool same = (a == b);   // the reference's pointer test al1==al2 (TwoElectronInts.cpp:444)
    keep.clear();
    for (int ia = 0; ia < B.npr[a]; ++ia) {
        const double axp = B.alpha[B.poff[a] + ia], c1 = B.coef[B.poff[a] + ia];
        const int jend = same ? ia + 1 : B.npr[b];
        for (int ib = 0; ib < jend; ++ib) {
            const double bxp = B.alpha[B.poff[b] + ib], c2 = B.coef[B.poff[b] + ib];
            PrimPair pp;
            pp.p = axp + bxp;
            pp.ip = 1.0 / pp.p;
            pp.u = std::exp(-axp * bxp * ab2 * pp.ip) * pp.ip;
            if (SR_TERM * pp.u * umax / std::sqrt(pp.p) * 1.0000001 < prim_cut) continue;
            for (int x = 0; x < 3; ++x) {
                pp.P[x] = (axp * A[x] + bxp * Bc[x]) * pp.ip;
                pp.PA[x] = pp.P[x] - A[x];
            }
            pp.c = c1 * c2 * ((same && ia != ib) ? 2.0 : 1.0);
            keep.push_back(pp);
        }
    }
    if (keep.empty()) return false;
    // primitive pairs sorted by u (descending) so the kernels can leave the primitive loops as soon as the
    // bound SR*u_bra*u_ket/sqrt(pmin) drops below the cut
    std::stable_sort(keep.begin(), keep.end(), [](const PrimPair &x, const PrimPair &y) { return x.u > y.u; });
    sp.umax = keep.front().u;
    sp.pmin = keep.front().p;
    for (auto &pp : keep) sp.pmin = std::min(sp.pmin, pp.p);
    sp.spare = 0.0;
    sp.Q = 0.0;
    sp.offa = B.off[a]; sp.offb = B.off[b];
    sp.sha = a; sp.shb = b;
    sp.pairid = i * (i + 1) / 2 + j;
    sp.pad = 0;
    sp.prim_off = 0;
    sp.nprim = (int)keep.size();
    return true;
}

// Upload one pair list (already sorted by Q descending): full records, hot mirror, tile order, position maps.  Device buffers
// only grow; the host mirrors (hot_host, tp_host) persist so that the copies need no synchronisation here.
static int finalize_list(unomol_b200 *h, int c) {
    PairClassList &L = h->cls[c];
    L.n = (int)L.pairs.size();
    L.slot_pos.clear(); L.pos_slot.clear(); L.ntiles = 0; L.maxnp = 0;
    if (!L.n) return UNOMOL_OK;
    if ((size_t)L.n > L.cap_pairs) {
        if (L.d_pairs) cudaFree(L.d_pairs);
        if (L.d_hot) cudaFree(L.d_hot);
        L.d_pairs = nullptr; L.d_hot = nullptr;
        L.cap_pairs = (size_t)L.n + (size_t)L.n / 8 + 64;
        CUDA_TRY(h, cudaMalloc(&L.d_pairs, sizeof(ShellPair) * L.cap_pairs));
        CUDA_TRY(h, cudaMalloc(&L.d_hot, sizeof(KetHot) * L.cap_pairs));
    }
    CUDA_TRY(h, cudaMemcpyAsync(L.d_pairs, L.pairs.data(), sizeof(ShellPair) * L.n, cudaMemcpyHostToDevice, h->stream));
    L.hot_host.resize(L.n);
    for (int i = 0; i < L.n; ++i) {
        const ShellPair &sp = L.pairs[i];
        L.hot_host[i] = KetHot{sp.offa, sp.offb, sp.prim_off, sp.nprim, sp.sha, sp.shb, sp.pairid, 0};
        h->pair_cls[sp.pairid] = c;
        h->pair_pos[sp.pairid] = i;
        L.maxnp = std::max(L.maxnp, sp.nprim);
    }
    CUDA_TRY(h, cudaMemcpyAsync(L.d_hot, L.hot_host.data(), sizeof(KetHot) * L.n, cudaMemcpyHostToDevice, h->stream));
    // tile order for eri_tile.cuh: by first shell, Q descending inside a shell (= list position ascending)
    if (c / NSUB < NSPDCLASS && tile_class_available(c / NSUB, 0)) {
        const int tb = tile_b_of_class(c / NSUB);
        // positions ordered by first shell, list order (= Q descending) inside a shell: counting sort, linear
        const int nsh = h->basis.nshell;
        std::vector<int> start(nsh + 1, 0), ord(L.n);
        for (int i = 0; i < L.n; ++i) ++start[L.pairs[i].sha + 1];
        for (int sh = 0; sh < nsh; ++sh) start[sh + 1] += start[sh];
        for (int i = 0; i < L.n; ++i) ord[start[L.pairs[i].sha]++] = i;
        L.pos_slot.assign(L.n, -1);
        L.slot_pos.reserve(((size_t)L.n / tb + nsh) * TILE_SLOTS);
        int fill = 0, cur_sh = -1;
        for (int o : ord) {
            if (L.pairs[o].sha != cur_sh || fill == tb) {
                L.slot_pos.resize(L.slot_pos.size() + TILE_SLOTS, -1);   // open a new tile
                fill = 0;
                cur_sh = L.pairs[o].sha;
            }
            const int slot = (int)L.slot_pos.size() - TILE_SLOTS + fill++;
            L.slot_pos[slot] = o;
            L.pos_slot[o] = slot;
        }
        L.ntiles = (int)L.slot_pos.size() / TILE_SLOTS;
        L.tp_host.resize(L.slot_pos.size());
        for (size_t sl = 0; sl < L.tp_host.size(); ++sl) {
            if (L.slot_pos[sl] >= 0) L.tp_host[sl] = L.pairs[L.slot_pos[sl]];
            else { memset(&L.tp_host[sl], 0, sizeof(ShellPair)); L.tp_host[sl].sha = L.tp_host[sl].shb = -1; L.tp_host[sl].pairid = -1; }
        }
        if (L.tp_host.size() > L.cap_tp) {
            if (L.d_tpairs) cudaFree(L.d_tpairs);
            L.d_tpairs = nullptr;
            L.cap_tp = L.tp_host.size() + L.tp_host.size() / 8 + 64 * TILE_SLOTS;
            CUDA_TRY(h, cudaMalloc(&L.d_tpairs, sizeof(ShellPair) * L.cap_tp));
        }
        CUDA_TRY(h, cudaMemcpyAsync(L.d_tpairs, L.tp_host.data(), sizeof(ShellPair) * L.tp_host.size(), cudaMemcpyHostToDevice, h->stream));
    }
    return UNOMOL_OK;
}

// Schwarz bounds Q = sqrt(max |(ab|ab)|) of the pairs of one class (all of class `cls`, any group), on the GPU (MODE_SCHWARZ)
static int schwarz_of_pairs(unomol_b200 *h, int cls, std::vector<ShellPair> &pairs) {
    const int n = (int)pairs.size();
    if (!n) return UNOMOL_OK;
    // grow-only scratch: [ShellPair n][int2 n][double n]
    const size_t need = (sizeof(ShellPair) + sizeof(int2) + sizeof(double)) * (size_t)n + 64;
    if (need > h->schwarz_scratch_cap) {
        if (h->d_schwarz_scratch) cudaFree(h->d_schwarz_scratch);
        h->d_schwarz_scratch = nullptr;
        h->schwarz_scratch_cap = need + need / 4;
        if (cudaMalloc(&h->d_schwarz_scratch, h->schwarz_scratch_cap) != cudaSuccess) { h->schwarz_scratch_cap = 0; return UNOMOL_E_NOMEM; }
    }
    struct { ShellPair *p; } d_p{reinterpret_cast<ShellPair *>(h->d_schwarz_scratch)};
    struct { int2 *p; } d_tl{reinterpret_cast<int2 *>(h->d_schwarz_scratch + sizeof(ShellPair) * (size_t)n)};
    struct { double *p; } d_q{reinterpret_cast<double *>(h->d_schwarz_scratch + (sizeof(ShellPair) + sizeof(int2)) * (size_t)n)};
    std::vector<int2> tl(n);
    for (int i = 0; i < n; ++i) tl[i] = make_int2(i, i);
    CUDA_TRY(h, cudaMemcpyAsync(d_p.p, pairs.data(), sizeof(ShellPair) * n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(d_tl.p, tl.data(), sizeof(int2) * n, cudaMemcpyHostToDevice, h->stream));
    ClassTask task{};
    task.rys = h->rys;
    task.bra = d_p.p; task.ket = d_p.p; task.prims = h->d_prims;
    task.nbra = n; task.nket = n;
    // no primitive cut here: the bound must hold for quartets whose partner pair is strong, where the
    // reference's sr<1e-12 test passes although it would fail on the weak pair's own diagonal
    task.prim_cut = 0.0;
    task.task_list = d_tl.p; task.ntask = n; task.out = d_q.p;
    const int groups = any_groups_per_cta(cls, cls);
    const int grid = std::min((n + groups - 1) / groups, h->nsm * 16);
    CUDA_TRY(h, launch_any_class(h, cls, cls, task, MODE_SCHWARZ, grid, h->stream));
    std::vector<double> q(n);
    CUDA_TRY(h, cudaMemcpyAsync(q.data(), d_q.p, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n; ++i) pairs[i].Q = q[i];
    return UNOMOL_OK;
}

// Shell-pair + primitive-pair precompute (host), Schwarz bounds (GPU), sort, upload.
static int build_pairs(unomol_b200 *h) {
    free_pairs(h);
    cudaEventRecord(h->ev2, h->stream);
    const HostBasis &B = h->basis;
    const int ns = B.nshell;
    h->pair_cls.assign((size_t)ns * (ns + 1) / 2, -1);
    h->pair_pos.assign((size_t)ns * (ns + 1) / 2, -1);
    h->h_prims.clear();
    // Exact prune.  A primitive quartet is skipped by the reference when sr = SR*u12*u34/sqrt(p+q) < prim_cut
    // (TwoElectronInts.cpp:478-479).  u34 <= umax and sqrt(p+q) > sqrt(p12), so a primitive pair with
    // SR*u12*umax/sqrt(p12) < prim_cut can never survive against any partner and is dropped; umax = 1/(2 alpha_min)
    // is attained by the most diffuse shell paired with itself.  A whole shell pair is rejected before any exp()
    // when even its most diffuse primitive pair fails that test (u, 1/p and 1/sqrt(p) are all largest there).
    // Rows of the (i >= j) pair triangle are processed by a few host threads; the result is concatenated in (i, j)
    // order so the tables do not depend on the thread count.
    long long nprim = 0, nkept = 0;
    // The McMurchie-Davidson path (l_tot > 8) has NO primitive cut in the reference (TwoElectronInts.cpp:269-418), so
    // with f/g shells in the basis nothing may be pruned up front; the Rys classes apply the cut per primitive quartet.
    h->has_highl = false;
    for (int s = 0; s < ns; ++s) h->has_highl |= (B.lv[s] > 2);
    const double prune_cut = h->has_highl ? 0.0 : h->prim_cut;
    const bool bucketed = (long long)ns * (ns + 1) / 2 >= h->bucket_min_pairs;
    // column/row blocking: keep (rows of a bra block) x (columns of a ket block) x 8 B x ~2.5 matrices within ~48 MB of L2
    int nblock = h->col_blocks;
    if (nblock <= 0) {
        // measured (profiles/exp_blocks.py): N=2002 (80 MB) is fastest unblocked, N=4004 (320 MB) 1.56x faster with 3 blocks
        const double foot = 2.5 * 8.0 * (double)B.nbf * (double)B.nbf;
        nblock = foot <= 100e6 ? 1 : (int)std::ceil(std::sqrt(foot / 48e6));
    }
    nblock = std::max(1, std::min(nblock, NBLOCK));
    h->bucketed = bucketed;
    h->nblock = nblock;
    auto group_of = [&](const ShellPair &sp, int cls) { return group_of_pair(h, sp, cls); };
    if (h->device_pairs) {
        // primitive-pair generation, exact prune and per-pair sort on the GPU (pair_device.cu); the host only groups
        std::vector<ShellPair> kept;
        std::vector<int> kcls;
        int rc = build_pair_tables_device(h, prune_cut, kept, kcls, &h->d_prims, &nprim);
        if (rc) return rc;
        nkept = (long long)kept.size();
        for (size_t i = 0; i < kept.size(); ++i) h->cls[group_of(kept[i], kcls[i])].pairs.push_back(kept[i]);
    } else {
    std::vector<double> amin(ns);
    double umax = 0.0;
    for (int s = 0; s < ns; ++s) {
        amin[s] = B.alpha[B.poff[s]];
        for (int k = 1; k < B.npr[s]; ++k) amin[s] = std::min(amin[s], B.alpha[B.poff[s] + k]);
        umax = std::max(umax, 0.5 / amin[s]);
    }
    struct OutPair { ShellPair sp; int cls; int first, count; };
    struct RowOut { std::vector<OutPair> pairs; std::vector<PrimPair> prims; };
    std::vector<RowOut> rows(ns);
    const double prim_cut = prune_cut;
    auto do_row = [&](int i) {
        RowOut &R = rows[i];
        std::vector<PrimPair> keep;
        for (int j = 0; j <= i; ++j) {
            ShellPair sp;
            if (!make_pair_host(B, i, j, amin, umax, prim_cut, sp, keep)) continue;
            R.pairs.push_back({sp, pair_class_id(B.lv[sp.sha], B.lv[sp.shb]), (int)R.prims.size(), (int)keep.size()});
            R.prims.insert(R.prims.end(), keep.begin(), keep.end());
        }
    };
    {
        const int nthr = std::max(1, std::min<int>(16, std::min<int>((int)std::thread::hardware_concurrency(), ns / 64)));
        if (nthr <= 1) {
            for (int i = 0; i < ns; ++i) do_row(i);
        } else {
            std::atomic<int> next(0);
            std::vector<std::thread> pool;
            for (int t = 0; t < nthr; ++t)
                pool.emplace_back([&]() { for (int i = next.fetch_add(1); i < ns; i = next.fetch_add(1)) do_row(ns - 1 - i); });
            for (auto &t : pool) t.join();
        }
    }
    for (int i = 0; i < ns; ++i) {
        for (auto &o : rows[i].pairs) {
            o.sp.prim_off = (int)h->h_prims.size();
            h->h_prims.insert(h->h_prims.end(), rows[i].prims.begin() + o.first, rows[i].prims.begin() + o.first + o.count);
            h->cls[group_of(o.sp, o.cls)].pairs.push_back(o.sp);
            nprim += o.count;
            ++nkept;
        }
        RowOut().pairs.swap(rows[i].pairs);
        std::vector<PrimPair>().swap(rows[i].prims);
    }
    if (nprim) {
        CUDA_TRY(h, cudaMalloc(&h->d_prims, sizeof(PrimPair) * nprim));
        CUDA_TRY(h, cudaMemcpyAsync(h->d_prims, h->h_prims.data(), sizeof(PrimPair) * nprim, cudaMemcpyHostToDevice,
                                    h->stream));
    }
    }   // host path
    h->stats.n_shell_pairs = (long long)ns * (ns + 1) / 2;
    h->stats.n_pairs_kept = nkept;
    h->stats.n_prim_pairs = nprim;
    // Schwarz bounds: diagonal quartet (ab|ab) of every kept pair, on the GPU (MODE_SCHWARZ)
    for (int c = 0; c < NGROUP; ++c) {
        PairClassList &L = h->cls[c];
        if (L.pairs.empty()) { L.n = 0; continue; }
        int rcq = schwarz_of_pairs(h, c / NSUB, L.pairs);
        if (rcq) return rcq;
        std::stable_sort(L.pairs.begin(), L.pairs.end(), [](const ShellPair &x, const ShellPair &y) { return x.Q > y.Q; });
        rcq = finalize_list(h, c);
        if (rcq) return rcq;
    }
    h->n_prims_full = nprim;
    h->d_prims_cap = nprim;
    h->inc_shells.clear();
    h->xyz_built = h->basis.xyz;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    int rc = build_plans(h);
    if (rc) return rc;
    cudaEventRecord(h->ev3, h->stream);
    cudaEventSynchronize(h->ev3);
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev2, h->ev3);
    h->stats.precompute_ms = ms;
    h->pairs_ready = true;
    return UNOMOL_OK;
}

// Geometry update with a few moved shells: every shell pair that contains one is rebuilt (host: primitive pairs + exact
// prune; GPU: Schwarz bound), merged back into the Q-sorted lists; all other pairs keep their records, their primitive pairs
// in d_prims and their bounds.  The plans (Schwarz prefixes, tile orders) are rebuilt.
static int update_pairs_incremental(unomol_b200 *h, const std::vector<int> &moved) {
    cudaEventRecord(h->ev2, h->stream);
    static const bool trace = std::getenv("UNOMOL_TIME_UPDATE") != nullptr;
    auto tnow = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = tnow(), t1 = t0, t2 = t0, t3 = t0, t4 = t0;
    const HostBasis &B = h->basis;
    const int ns = B.nshell;
    std::vector<char> mark(ns, 0);
    for (int sidx : moved) mark[sidx] = 1;
    std::vector<double> amin(ns);
    double umax = 0.0;
    for (int sh = 0; sh < ns; ++sh) {
        amin[sh] = B.alpha[B.poff[sh]];
        for (int k = 1; k < B.npr[sh]; ++k) amin[sh] = std::min(amin[sh], B.alpha[B.poff[sh] + k]);
        umax = std::max(umax, 0.5 / amin[sh]);
    }
    const double prune_cut = h->has_highl ? 0.0 : h->prim_cut;
    // rebuilt pairs, by class
    std::vector<ShellPair> dyn[NPAIRCLASS];
    std::vector<PrimPair> dprims, keep;
    for (int m : moved)
        for (int j = 0; j < ns; ++j) {
            if (mark[j] && j > m) continue;            // both moved: generated once, from the larger index
            const int hi = std::max(m, j), lo = std::min(m, j);
            ShellPair sp;
            if (!make_pair_host(B, hi, lo, amin, umax, prune_cut, sp, keep)) continue;
            sp.prim_off = (int)(h->n_prims_full + (long long)dprims.size());
            dprims.insert(dprims.end(), keep.begin(), keep.end());
            dyn[pair_class_id(B.lv[sp.sha], B.lv[sp.shb])].push_back(sp);
        }
    // reserved tail of d_prims: sized once for the largest possible set of primitive pairs of these shells
    if (h->inc_shells != moved || h->n_prims_full + (long long)dprims.size() > h->d_prims_cap) {
        long long bound = 0;
        for (int m : moved)
            for (int j = 0; j < ns; ++j) bound += (long long)B.npr[m] * B.npr[j];
        const long long cap = h->n_prims_full + bound;
        if (cap > h->d_prims_cap) {
            PrimPair *np = nullptr;
            CUDA_TRY(h, cudaMalloc(&np, sizeof(PrimPair) * std::max<long long>(cap, 1)));
            if (h->d_prims && h->n_prims_full)
                CUDA_TRY(h, cudaMemcpyAsync(np, h->d_prims, sizeof(PrimPair) * h->n_prims_full, cudaMemcpyDeviceToDevice, h->stream));
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            if (h->d_prims) cudaFree(h->d_prims);
            h->d_prims = np;
            h->d_prims_cap = cap;
        }
        h->inc_shells = moved;
    }
    if (!dprims.empty())
        CUDA_TRY(h, cudaMemcpyAsync(h->d_prims + h->n_prims_full, dprims.data(), sizeof(PrimPair) * dprims.size(), cudaMemcpyHostToDevice,
                                    h->stream));
    t1 = tnow();
    for (int cls = 0; cls < NPAIRCLASS; ++cls) {
        int rc = schwarz_of_pairs(h, cls, dyn[cls]);
        if (rc) return rc;
    }
    t2 = tnow();
    // drop the old records of the moved shells, merge the new ones in by Q
    std::vector<char> touched(NGROUP, 0);
    for (int c = 0; c < NGROUP; ++c) {
        auto &P = h->cls[c].pairs;
        const size_t before = P.size();
        for (auto &sp : P)
            if (mark[sp.sha] || mark[sp.shb]) { h->pair_cls[sp.pairid] = -1; h->pair_pos[sp.pairid] = -1; }
        P.erase(std::remove_if(P.begin(), P.end(), [&](const ShellPair &sp) { return mark[sp.sha] || mark[sp.shb]; }), P.end());
        if (P.size() != before) touched[c] = 1;
    }
    std::vector<size_t> nstatic(NGROUP);
    for (int c = 0; c < NGROUP; ++c) nstatic[c] = h->cls[c].pairs.size();
    for (int cls = 0; cls < NPAIRCLASS; ++cls)
        for (auto &sp : dyn[cls]) {
            const int c = group_of_pair(h, sp, cls);
            h->cls[c].pairs.push_back(sp);
            touched[c] = 1;
        }
    long long nkept = 0, nprim = 0;
    const auto byQ = [](const ShellPair &x, const ShellPair &y) { return x.Q > y.Q; };
    for (int c = 0; c < NGROUP; ++c) {
        if (touched[c]) {
            auto &P = h->cls[c].pairs;
            // the kept records are still sorted: sort the few new ones and merge (linear)
            std::stable_sort(P.begin() + nstatic[c], P.end(), byQ);
            std::inplace_merge(P.begin(), P.begin() + nstatic[c], P.end(), byQ);
            int rc = finalize_list(h, c);
            if (rc) return rc;
        }
        nkept += (long long)h->cls[c].pairs.size();
        for (auto &sp : h->cls[c].pairs) nprim += sp.nprim;
    }
    h->stats.n_pairs_kept = nkept;
    h->stats.n_prim_pairs = nprim;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    t3 = tnow();
    int rc = build_plans(h);
    if (rc) return rc;
    t4 = tnow();
    if (trace)
        fprintf(stderr, "unomol_b200: incremental update: pairs+prims %.3f ms, Schwarz %.3f ms, lists %.3f ms, plans %.3f ms\n", t1 - t0, t2 - t1,
                t3 - t2, t4 - t3);
    h->xyz_built = h->basis.xyz;
    cudaEventRecord(h->ev3, h->stream);
    cudaEventSynchronize(h->ev3);
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev2, h->ev3);
    h->stats.precompute_ms = ms;
    ++h->stats.n_incremental_updates;
    return UNOMOL_OK;
}

// per (bra class >= ket class): prefix counts of kets passing Q_bra*Q_ket >= tau (kets sorted descending)
static int build_plans(unomol_b200 *h) {
    // The per-plan arrays (ket counts, tile ket counts, tile orders, ket prefixes) live in ONE device pool filled by one copy:
    // a plan rebuild used to cost ~100 cudaMalloc/cudaFree/synchronize calls, most of an incremental geometry update.
    h->plans.clear();
    // pinned, grow-only staging buffer (no zero-fill, DMA straight from it)
    size_t stage_size = 0;
    bool stage_oom = false;
    struct PoolRef { size_t kc = (size_t)-1, kct = (size_t)-1, order = (size_t)-1, pre = (size_t)-1; };
    std::vector<PoolRef> refs;
    auto push = [&](const void *src, size_t bytes) {
        const size_t off = (stage_size + 255) & ~(size_t)255;
        const size_t end = off + std::max<size_t>(bytes, 16);
        if (end > h->plan_stage_cap) {
            const size_t ncap = end + end / 2 + (1 << 16);
            unsigned char *np = nullptr;
            if (cudaMallocHost(&np, ncap) != cudaSuccess) { stage_oom = true; return (size_t)0; }
            if (h->plan_stage && stage_size) memcpy(np, h->plan_stage, stage_size);
            if (h->plan_stage) cudaFreeHost(h->plan_stage);
            h->plan_stage = np;
            h->plan_stage_cap = ncap;
        }
        memcpy(h->plan_stage + off, src, bytes);
        stage_size = end;
        return off;
    };
    long long total = 0;
    std::vector<int> kc, kct, order;
    std::vector<std::pair<long long, int>> cost;
    for (int cb = 0; cb < NGROUP; ++cb)
        for (int ck = 0; ck <= cb; ++ck) {
            const PairClassList &Lb = h->cls[cb], &Lk = h->cls[ck];
            if (!Lb.n || !Lk.n) continue;
            ComboPlan plan;
            plan.cb = cb; plan.ck = ck;
            kc.assign(Lb.n, 0);
            int sweep = Lk.n;     // both lists are sorted by Q descending: the ket prefix only shrinks along the bras
            for (int i = 0; i < Lb.n; ++i) {
                int cnt;
                if (h->tau <= 0.0) cnt = Lk.n;
                else {
                    const double need = h->tau / Lb.pairs[i].Q;   // Q_ket >= need
                    while (sweep > 0 && !(Lk.pairs[sweep - 1].Q >= need)) --sweep;   // = first index with Q < need
                    cnt = sweep;
                }
                if (cb == ck) cnt = std::min(cnt, i + 1);   // canonical: ket position <= bra position
                kc[i] = cnt;
                plan.nquartets += cnt;
                if (cnt > 0) plan.nbra_eff = i + 1;
            }
            plan.nquartets_eff = plan.nquartets;
            total += plan.nquartets;
            if (plan.nbra_eff == 0) continue;
            {
                int la, lb, lc, ld;
                pair_class_l(cb / NSUB, la, lb);
                pair_class_l(ck / NSUB, lc, ld);
                plan.cost = (double)plan.nquartets * model_flops_per_primitive_quartet(la, lb, lc, ld);
            }
            int maxbp = 0;
            for (int i = 0; i < plan.nbra_eff; ++i) maxbp = std::max(maxbp, Lb.pairs[i].nprim);
            plan.highl = is_highl(cb / NSUB, ck / NSUB);
            // Thread-per-quartet kernels evaluate the Rys roots per lane: fine for the closed-form 1- and 2-root routines, but
            // 3..5 roots read 13 x 2n table coefficients per lane from global memory (rys_roots.cuh), which the generic kernel
            // amortises over the lanes of a quartet group.  Measured on SF6/TZ2P: (dp|pp) 5.3 ms in the register kernel, the
            // whole build 6.1 ms with the generic kernel for every class (profiles/exp_sf6.py).  So: register kernels for <= 2
            // roots, and for 3 roots only on long lists ((pp|pp) of the water clusters).
            int nroots;
            {
                int la, lb, lc, ld;
                pair_class_l(cb / NSUB, la, lb);
                pair_class_l(ck / NSUB, lc, ld);
                nroots = (la + lb + lc + ld) / 2 + 1;
            }
            const bool roots_ok = nroots <= 2 || (nroots == 3 && plan.nquartets >= 1000000) || h->use_reg_kernels == 2;
            plan.use_reg = !plan.highl && h->use_reg_kernels && roots_ok && reg_class_available(cb / NSUB, ck / NSUB) && maxbp <= reg_max_bra_prims();
            // (the tile kernels address the square matrices with 32-bit element offsets: leading dimension <= 65535)
            plan.use_tile = plan.use_reg && h->use_tile_kernels && tile_class_available(cb / NSUB, ck / NSUB) && Lb.ntiles > 0 &&
                            Lb.maxnp <= TILE_MAX_BRA_PRIMS && h->basis.nbf < 65535;
            if (plan.use_tile) {
                plan.maxbp = Lb.maxnp;
                // ket primitives a thread keeps in shared memory: every ket of the list when that fits the budget
                int maxkp = 0;
                const int kmaxvis = *std::max_element(kc.begin(), kc.end());
                for (int i = 0; i < kmaxvis; ++i) maxkp = std::max(maxkp, Lk.pairs[i].nprim);
                plan.kslots = std::min(maxkp, h->tile_kslots);
                while (plan.kslots > 0 && tile_smem_bytes(cb / NSUB, ck / NSUB, plan.maxbp, plan.kslots, h->rys.rys2_exact) > 100 * 1024) --plan.kslots;
                if (tile_smem_bytes(cb / NSUB, ck / NSUB, plan.maxbp, plan.kslots, h->rys.rys2_exact) > 200 * 1024) plan.use_tile = false;
            }
            kct.clear(); order.clear();
            if (plan.use_tile) {
                kct.assign(Lb.slot_pos.size(), 0);
                cost.assign(Lb.ntiles, std::pair<long long, int>(0, 0));
                for (int t = 0; t < Lb.ntiles; ++t) {
                    long long w = 0;
                    for (int j = 0; j < TILE_SLOTS; ++j) {
                        const int pos = Lb.slot_pos[(size_t)t * TILE_SLOTS + j];
                        if (pos >= 0) { kct[(size_t)t * TILE_SLOTS + j] = kc[pos]; w += kc[pos]; }
                    }
                    cost[t] = {w, t};
                }
                std::stable_sort(cost.begin(), cost.end(), [](const std::pair<long long, int> &x, const std::pair<long long, int> &y) { return x.first > y.first; });
                for (auto &ct : cost) if (ct.first > 0) order.push_back(ct.second);
                plan.ntiles = (int)order.size();
                // A work item of a CTA is (tile, ket slice).  Lists with few tiles deal the kets of a tile to several CTAs so that
                // the launch still offers ~8 CTAs per SM; a slice should keep at least two rounds of TILE_THREADS kets.
                const int kmax = *std::max_element(kc.begin(), kc.end());
                const int want = plan.ntiles > 0 ? (8 * h->nsm + plan.ntiles - 1) / plan.ntiles : 1;
                plan.tile_slices = std::max(1, std::min(want, kmax / 256));
                // small molecules (SF6/TZ2P: 17 ms with tiles, 11 ms without, 6 ms with the generic kernel): too little work per
                // launch for tiles of 8 bras to pay off
                if (plan.nquartets < 1000000 && h->use_tile_kernels != 2) plan.use_tile = false;
            }
            PoolRef ref;
            if (plan.use_tile) {
                ref.kct = push(kct.data(), sizeof(int) * kct.size());
                ref.order = push(order.data(), sizeof(int) * order.size());
            }
            ref.kc = push(kc.data(), sizeof(int) * Lb.n);
            if (plan.highl) {
                std::vector<long long> pre(plan.nbra_eff + 1, 0);
                for (int i = 0; i < plan.nbra_eff; ++i) pre[i + 1] = pre[i] + kc[i];
                ref.pre = push(pre.data(), sizeof(long long) * pre.size());
            }
            refs.push_back(ref);
            h->plans.push_back(plan);
        }
    if (stage_oom) return UNOMOL_E_NOMEM;
    if (stage_size > h->plan_pool_cap) {
        if (h->d_plan_pool) cudaFree(h->d_plan_pool);
        h->d_plan_pool = nullptr;
        h->plan_pool_cap = stage_size + stage_size / 4 + 4096;
        if (cudaMalloc(&h->d_plan_pool, h->plan_pool_cap) != cudaSuccess) { h->plan_pool_cap = 0; return UNOMOL_E_NOMEM; }
    }
    if (stage_size) CUDA_TRY(h, cudaMemcpyAsync(h->d_plan_pool, h->plan_stage, stage_size, cudaMemcpyHostToDevice, h->stream));
    for (size_t ip = 0; ip < h->plans.size(); ++ip) {
        ComboPlan &pl = h->plans[ip];
        const PoolRef &r = refs[ip];
        pl.d_ket_count = reinterpret_cast<int *>(h->d_plan_pool + r.kc);
        pl.d_kc_tile = r.kct != (size_t)-1 ? reinterpret_cast<int *>(h->d_plan_pool + r.kct) : nullptr;
        pl.d_tile_order = r.order != (size_t)-1 ? reinterpret_cast<int *>(h->d_plan_pool + r.order) : nullptr;
        pl.d_ket_prefix = r.pre != (size_t)-1 ? reinterpret_cast<long long *>(h->d_plan_pool + r.pre) : nullptr;
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const long long np = (long long)h->basis.nshell * (h->basis.nshell + 1) / 2;
    h->stats.n_quartets_total = np * (np + 1) / 2;
    (void)total;
    if (h->plans.size() + 1 > h->counters_cap) {
        if (h->d_counters) cudaFree(h->d_counters);
        h->d_counters = nullptr;
        h->counters_cap = h->plans.size() + 17;
        if (cudaMalloc(&h->d_counters, sizeof(unsigned long long) * 2 * h->counters_cap) != cudaSuccess) { h->counters_cap = 0; return UNOMOL_E_NOMEM; }
    }
    return UNOMOL_OK;
}

// ---------------------------------------------------------------- Fock build
static int ensure_buffers(unomol_b200 *h) {
    if (h->d_PJ) return UNOMOL_OK;
    const size_t n = h->basis.nbf, ld = (n + 1) & ~(size_t)1, nn = ld * ld, no2 = n * (n + 1) / 2;
    for (int s = 0; s < 2; ++s) {
        CUDA_TRY(h, cudaMalloc(&h->d_Ppacked[s], sizeof(double) * no2));
        CUDA_TRY(h, cudaMalloc(&h->d_Gpacked[s], sizeof(double) * no2));
        CUDA_TRY(h, cudaMalloc(&h->d_PK[s], sizeof(double) * nn));
        CUDA_TRY(h, cudaMalloc(&h->d_K[s], sizeof(double) * nn));
    }
    CUDA_TRY(h, cudaMalloc(&h->d_PJ, sizeof(double) * nn));
    CUDA_TRY(h, cudaMalloc(&h->d_J, sizeof(double) * nn));
    CUDA_TRY(h, cudaMallocHost(&h->h_pinned, sizeof(double) * no2 * 4));
    return UNOMOL_OK;
}

typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);

// device-resident core: packed dP[nspin] -> packed dG[nspin] (overwritten), on h->stream
static int fock_device(unomol_b200 *h, int nspin, const double *dPA, const double *dPB, double *dGA, double *dGB) {
    if (!h->pairs_ready) {
        int rc = build_pairs(h);
        if (rc) return rc;
    }
    int rc = ensure_buffers(h);
    if (rc) return rc;
    const int n = h->basis.nbf, ld = (n + 1) & ~1;
    const size_t nn = (size_t)ld * ld, no2 = (size_t)n * (n + 1) / 2;
    cudaStream_t st = h->stream;
    cudaEventRecord(h->ev0, st);
    unpack_density_kernel<<<(unsigned)(((size_t)n * n + 255) / 256), 256, 0, st>>>(dPA, dPB, n, ld, nspin, h->d_PJ, h->d_PK[0], h->d_PK[1]);
    CUDA_TRY(h, cudaMemsetAsync(h->d_J, 0, sizeof(double) * nn, st));
    for (int s = 0; s < nspin; ++s) CUDA_TRY(h, cudaMemsetAsync(h->d_K[s], 0, sizeof(double) * nn, st));
    CUDA_TRY(h, cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * 2 * (h->plans.size() + 1), st));
    cudaEventRecord(h->ev2, st);
    int nlaunch = 2;
    int n_tile = 0, n_reg = 0, n_rows = 0, n_gen = 0, n_hl = 0;
    // work counters of this build (see engine.h)
    unsigned long long *work = nullptr;
    if (h->d_work_shared && h->work_owner) {
        // The owner clears the set this build does NOT use, in EVERY build (stealing enabled or not): build k uses set k & 1, so a
        // set is always clean again before its next use, also across work_stealing being switched off and on (build_count is
        // never reset by that option: all ranks keep counting the same builds).  The collective that ends a build orders the
        // reset before the next build of the other ranks.
        const int par = (int)(h->build_count & 1);
        CUDA_TRY(h, cudaMemsetAsync(h->d_work_shared + (size_t)(par ^ 1) * unomol_b200::MAXPLAN, 0,
                                    sizeof(unsigned long long) * unomol_b200::MAXPLAN, st));
    }
    if (h->d_work_shared && h->steal_enabled) {
        const int par = (int)(h->build_count & 1);
        work = h->d_work_shared + (size_t)par * unomol_b200::MAXPLAN;
    } else if (h->nranks == 1) {
        if (!h->d_work_local) CUDA_TRY(h, cudaMalloc(&h->d_work_local, sizeof(unsigned long long) * unomol_b200::MAXPLAN));
        CUDA_TRY(h, cudaMemsetAsync(h->d_work_local, 0, sizeof(unsigned long long) * std::max<size_t>(1, h->plans.size()), st));
        work = h->d_work_local;
    }
    // static-plus-stealing split (ClassTask::local_counter): this rank's own counters for the statically dealt blocks
    unsigned long long *local = nullptr;
    if (work && h->nranks > 1 && h->static_fraction > 0.0) {
        if (!h->d_work_local) CUDA_TRY(h, cudaMalloc(&h->d_work_local, sizeof(unsigned long long) * unomol_b200::MAXPLAN));
        CUDA_TRY(h, cudaMemsetAsync(h->d_work_local, 0, sizeof(unsigned long long) * std::max<size_t>(1, h->plans.size()), st));
        local = h->d_work_local;
    }
    // blocks of a launch that are dealt statically: the given share of its blocks, a whole number per rank
    auto static_blocks = [&](long long items, int chunk) -> long long {
        const long long nblocks = (items + chunk - 1) / chunk;
        return (long long)(h->static_fraction * (double)nblocks / h->nranks) * h->nranks;
    };
    ++h->build_count;
    // The class launches are independent (they only meet in the FP64 reds on J/K): spread them over a few streams,
    // largest first, so that the small launches of small molecules (SF6: 24 launches of a few hundred CTAs)
    // overlap instead of each leaving most SMs idle.
    if (!h->aux[0]) {
        for (int a = 0; a < unomol_b200::NAUX; ++a) {
            CUDA_TRY(h, cudaStreamCreateWithFlags(&h->aux[a], cudaStreamNonBlocking));
            CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_join[a], cudaEventDisableTiming));
        }
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_fork, st));
    for (int a = 0; a < unomol_b200::NAUX; ++a) CUDA_TRY(h, cudaStreamWaitEvent(h->aux[a], h->ev_fork, 0));
    std::vector<size_t> order(h->plans.size());
    std::iota(order.begin(), order.end(), (size_t)0);
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return h->plans[x].cost > h->plans[y].cost; });
    static const bool debug_plans = getenv("UNOMOL_DEBUG_PLANS") != nullptr;   // launch list of the first build on stderr
    for (size_t io = 0; io < order.size(); ++io) {
        const size_t ip = order[io];
        const ComboPlan &pl = h->plans[ip];
        if (debug_plans && h->build_count == 1)
            fprintf(stderr, "launch %3zu: class (%d|%d) bucket (%d|%d) block (%d|%d) %s quartets %lld nbra %d nket %d tiles %d slices %d kslots %d maxbp %d ket maxnp %d\n",
                    io, pl.cb / NSUB, pl.ck / NSUB, pl.cb % NSUB / NBLOCK, pl.ck % NSUB / NBLOCK, pl.cb % NBLOCK, pl.ck % NBLOCK,
                    pl.highl ? "highl" : pl.use_tile ? "tile" : pl.use_reg ? "reg" : "generic", pl.nquartets, pl.nbra_eff, h->cls[pl.ck].n,
                    pl.ntiles, pl.tile_slices, pl.kslots, pl.maxbp, h->cls[pl.ck].maxnp);
        // the runtime-L launches share one scratch area: keep them on one stream
        cudaStream_t st = pl.highl ? h->aux[0] : h->aux[io % unomol_b200::NAUX];
        ClassTask task{};
        task.rys = h->rys;
        task.bra = h->cls[pl.cb].d_pairs; task.ket = h->cls[pl.ck].d_pairs; task.prims = h->d_prims;
        task.ket_hot = h->cls[pl.ck].d_hot;
        task.ket_count = pl.d_ket_count;
        task.ket_prefix = pl.d_ket_prefix;
        task.nbra = pl.nbra_eff; task.nket = h->cls[pl.ck].n;
        task.same_class = (pl.cb == pl.ck);
        task.start_shell = h->start_shell;
        task.rank = h->rank; task.nranks = h->nranks;
        task.prim_cut = h->prim_cut;
        task.value_cut = h->value_cut;
        task.nbf = ld; task.nspin = nspin;   // kernels use nbf only as the leading dimension of the square matrices
        task.PJ = (nspin == 1) ? h->d_PK[0] : h->d_PJ; task.PK[0] = h->d_PK[0]; task.PK[1] = h->d_PK[1];
        task.jscale = (nspin == 1) ? 4.0 : 2.0;
        task.J = h->d_J; task.K[0] = h->d_K[0]; task.K[1] = h->d_K[1];
        task.counters = h->d_counters + 2 * ip;
        task.debug_flags = h->debug_flags;
        task.work_counter = work ? work + ip : nullptr;
        task.local_counter = local ? local + ip : nullptr;
        task.cand_counter = h->d_counters + 2 * h->plans.size();
        const int nmine = work ? pl.nbra_eff : (pl.nbra_eff + h->nranks - 1) / h->nranks;
        task.chunk = std::max(1, std::min(8, pl.nbra_eff / (h->nsm * 16 * 8 * h->nranks)));
        if (pl.highl) {
            // work items of the runtime-L kernel are single quartets
            const long long nq = work ? pl.nquartets_eff : (pl.nquartets_eff + h->nranks - 1) / h->nranks;
            task.chunk = 1;
            task.static_blocks = static_blocks(pl.nquartets_eff, 1);
            CUDA_TRY(h, launch_any_class(h, pl.cb / NSUB, pl.ck / NSUB, task, MODE_DIGEST, (int)std::min<long long>(nq, 1 << 20), st));
            ++n_hl;
        } else if (pl.use_tile) {
            ++n_tile;
            task.tbra = h->cls[pl.cb].d_tpairs;
            task.ket_count = pl.d_kc_tile;
            task.tile_order = pl.d_tile_order;
            task.ntiles = pl.ntiles;
            task.tile_b = tile_b_of_class(pl.cb / NSUB);
            task.tile_maxbp = pl.maxbp;
            task.kslots = pl.kslots;
            task.chunk = 1;
            task.tile_slices = pl.tile_slices;
            const int items = pl.ntiles * pl.tile_slices;
            task.static_blocks = static_blocks(items, 1);
            const int tmine = work ? items : (items + h->nranks - 1) / h->nranks;
            CUDA_TRY(h, launch_tile_class(pl.cb / NSUB, pl.ck / NSUB, task, std::min(tmine, h->nsm * 8), st));
        } else if (pl.use_reg) {
            task.static_blocks = static_blocks(pl.nbra_eff, task.chunk);
            CUDA_TRY(h, launch_reg_class(pl.cb / NSUB, pl.ck / NSUB, task, std::min(nmine, h->nsm * 16), st, h->stage_rows != 0));
            ++n_reg;
            if (h->stage_rows && nspin == 1 && reg_rows_fit(pl.cb / NSUB, ld)) ++n_rows;
        } else {
            // generic kernel: a warp takes (bra, slice) items; split the kets of a bra over several warps when the list has
            // fewer bras than the GPU keeps warps busy (small molecules), every rank choosing the same split
            const int warps = h->nsm * 16 * 2;
            int split = pl.nbra_eff >= warps ? 1 : std::min(64, (warps + pl.nbra_eff - 1) / std::max(1, pl.nbra_eff));
            if (!h->bra_split_enabled) split = 1;
            task.bra_split = split;
            const long long items = (long long)(work ? pl.nbra_eff : nmine) * split;
            task.chunk = 1;      // the generic kernel's warps claim single (bra, slice) items
            task.static_blocks = static_blocks((long long)pl.nbra_eff * split, 1);
            const int ctas = (int)std::min<long long>((items + 3) / 4, h->nsm * 32);   // 4 warps per CTA
            CUDA_TRY(h, launch_quartet_class(pl.cb / NSUB, pl.ck / NSUB, task, MODE_DIGEST, std::max(ctas, 1), st));
            ++n_gen;
        }
        ++nlaunch;
    }
    for (int a = 0; a < unomol_b200::NAUX; ++a) {
        CUDA_TRY(h, cudaEventRecord(h->ev_join[a], h->aux[a]));
        CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_join[a], 0));
    }
    cudaEventRecord(h->ev3, st);
    pack_fock_kernel<<<(unsigned)((no2 + 255) / 256), 256, 0, st>>>(h->d_J, h->d_K[0], h->d_K[1], n, ld, nspin, dGA, dGB);
    ++nlaunch;
    CUDA_TRY(h, cudaGetLastError());
    if (h->nccl_comm && h->nranks > 1) {
        static std::atomic<nccl_allreduce_fn> fn_cache{nullptr};   // one process may drive several GPUs from several threads
        nccl_allreduce_fn fn = fn_cache.load();
        if (!fn) { fn = (nccl_allreduce_fn)dlsym(RTLD_DEFAULT, "ncclAllReduce"); fn_cache.store(fn); }
        if (!fn) { h->last_error = "ncclAllReduce not found in the process (load libnccl first)"; return UNOMOL_E_NCCL; }
        // ncclDouble = 8, ncclSum = 0 (nccl.h)
        if (fn(dGA, dGA, no2, 8, 0, h->nccl_comm, st) != 0) return UNOMOL_E_NCCL;
        if (nspin == 2 && fn(dGB, dGB, no2, 8, 0, h->nccl_comm, st) != 0) return UNOMOL_E_NCCL;
    }
    cudaEventRecord(h->ev1, st);
    h->stats.n_launches = nlaunch;
    h->stats.n_tile_launches = n_tile; h->stats.n_reg_launches = n_reg; h->stats.n_rows_launches = n_rows;
    h->stats.n_generic_launches = n_gen; h->stats.n_highl_launches = n_hl;
    return UNOMOL_OK;
}

static int finish_stats(unomol_b200 *h) {
    CUDA_TRY(h, cudaEventSynchronize(h->ev1));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->stats.last_fock_ms = ms;
    cudaEventElapsedTime(&ms, h->ev2, h->ev3);
    h->stats.last_eri_kernel_ms = ms;
    std::vector<unsigned long long> c(2 * (h->plans.size() + 1));
    CUDA_TRY(h, cudaMemcpy(c.data(), h->d_counters, sizeof(unsigned long long) * c.size(), cudaMemcpyDeviceToHost));
    long long nq = 0;
    double fl = 0.0;
    for (size_t ip = 0; ip < h->plans.size(); ++ip) {
        int la, lb, lc, ld;
        pair_class_l(h->plans[ip].cb / NSUB, la, lb);
        pair_class_l(h->plans[ip].ck / NSUB, lc, ld);
        nq += (long long)c[2 * ip];
        fl += (double)c[2 * ip + 1] * model_flops_per_primitive_quartet(la, lb, lc, ld);
    }
    h->stats.n_quartets = nq;
    h->stats.model_flops = fl;
    h->stats.n_prim_candidates = (long long)c[2 * h->plans.size()];
    h->stats.n_prim_quartets = 0;
    for (size_t ip = 0; ip < h->plans.size(); ++ip) h->stats.n_prim_quartets += (long long)c[2 * ip + 1];
    return UNOMOL_OK;
}

// ---------------------------------------------------------------- C ABI
extern "C" {

const char *unomol_b200_version(void) { return "unomol_b200 0.1 (sm_100a)"; }

const char *unomol_b200_strerror(int code) {
    switch (code) {
        case UNOMOL_OK: return "ok";
        case UNOMOL_E_ARG: return "bad argument";
        case UNOMOL_E_CUDA: return "CUDA failure or no CUDA device (there is no CPU fallback)";
        case UNOMOL_E_UNSUPPORTED: return "angular momentum above g (l > 4) is not supported (as in the reference, Basis.hpp:222)";
        case UNOMOL_E_NOMEM: return "out of memory";
        case UNOMOL_E_STATE: return "call order";
        case UNOMOL_E_NCCL: return "NCCL failure";
    }
    return "unknown error";
}

int unomol_b200_create(const unomol_basis_desc *b, int start_shell, int device, int rank, int nranks,
                       unomol_b200_t **out) {
    if (!b || !out || b->nshell <= 0 || nranks < 1 || rank < 0 || rank >= nranks) return UNOMOL_E_ARG;
    for (int s = 0; s < b->nshell; ++s)
        if (b->lv[s] > MAXL) return UNOMOL_E_UNSUPPORTED;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev) return UNOMOL_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return UNOMOL_E_CUDA;
    unomol_b200 *h = new unomol_b200;
    h->device = device; h->rank = rank; h->nranks = nranks; h->start_shell = start_shell;
    HostBasis &B = h->basis;
    B.nshell = b->nshell; B.nbf = b->nbf; B.ncen = b->ncen; B.maxl = b->maxl;
    B.npr.assign(b->npr, b->npr + b->nshell);
    B.lv.assign(b->lv, b->lv + b->nshell);
    for (int s = 0; s < b->nshell; ++s) B.maxl = std::max(B.maxl, B.lv[s]);   // sizes the runtime-L scratch: do not trust the header field
    B.cen.assign(b->cen, b->cen + b->nshell);
    B.off.assign(b->off, b->off + b->nshell);
    B.poff.assign(b->poff, b->poff + b->nshell);
    int nprim = 0;
    for (int s = 0; s < b->nshell; ++s) nprim = std::max(nprim, b->poff[s] + b->npr[s]);
    B.alpha.assign(b->alpha, b->alpha + nprim);
    B.coef.assign(b->coef, b->coef + nprim);
    B.xyz.assign(b->xyz, b->xyz + 3 * b->ncen);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return UNOMOL_E_CUDA; }
    cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1); cudaEventCreate(&h->ev2); cudaEventCreate(&h->ev3);
    if (rys_device_tables(&h->rys) != cudaSuccess) { unomol_b200_destroy(h); return UNOMOL_E_CUDA; }
    if (cudaDeviceGetAttribute(&h->nsm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || h->nsm <= 0) h->nsm = 148;
    h->stats.nbf = B.nbf; h->stats.nshell = B.nshell; h->stats.rank = rank; h->stats.nranks = nranks;
    int rc = build_pairs(h);
    if (rc) { unomol_b200_destroy(h); return rc; }
    *out = h;
    return UNOMOL_OK;
}

void unomol_b200_destroy(unomol_b200_t *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    free_pairs(h);
    for (int s = 0; s < 2; ++s) {
        cudaFree(h->d_Ppacked[s]); cudaFree(h->d_Gpacked[s]); cudaFree(h->d_PK[s]); cudaFree(h->d_K[s]);
    }
    cudaFree(h->d_PJ); cudaFree(h->d_J); cudaFree(h->d_counters); cudaFree(h->d_hl_scratch);
    cudaFree(h->d_plan_pool); cudaFree(h->d_schwarz_scratch);
    if (h->plan_stage) cudaFreeHost(h->plan_stage);
    cudaFree(h->d_work_local);
    if (h->d_work_shared && !h->work_borrowed) { if (h->work_owner) cudaFree(h->d_work_shared); else cudaIpcCloseMemHandle(h->d_work_shared); }
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    unomol_scf_free(h);
    if (h->ev0) { cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaEventDestroy(h->ev2); cudaEventDestroy(h->ev3); }
    for (int a = 0; a < unomol_b200::NAUX; ++a) {
        if (h->aux[a]) cudaStreamDestroy(h->aux[a]);
        if (h->ev_join[a]) cudaEventDestroy(h->ev_join[a]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int unomol_b200_set_option(unomol_b200_t *h, const char *name, double value) {
    if (!h || !name) return UNOMOL_E_ARG;
    cudaSetDevice(h->device);
    if (!strcmp(name, "schwarz_tau")) {
        h->tau = value;
        if (h->pairs_ready) return build_plans(h);
        return UNOMOL_OK;
    }
    if (!strcmp(name, "prim_cut")) { h->prim_cut = value; h->pairs_ready = false; return UNOMOL_OK; }
    if (!strcmp(name, "bucket_min_pairs")) { h->bucket_min_pairs = (int)value; h->pairs_ready = false; return UNOMOL_OK; }
    if (!strcmp(name, "work_stealing")) { h->steal_enabled = value != 0.0; return UNOMOL_OK; }
    // share of a launch's blocks that is dealt to the ranks statically (block-cyclic over the cost-sorted blocks); the rest is
    // stolen from the shared counter.  0 = every block from the shared counter (round-2 behaviour up to this option).
    if (!strcmp(name, "static_fraction")) { h->static_fraction = std::min(1.0, std::max(0.0, value)); return UNOMOL_OK; }
    if (!strcmp(name, "device_pairs")) { h->device_pairs = (int)value; h->pairs_ready = false; return UNOMOL_OK; }
    if (!strcmp(name, "col_blocks")) { h->col_blocks = (int)value; h->pairs_ready = false; return UNOMOL_OK; }
    if (!strcmp(name, "stage_rows")) { h->stage_rows = (int)value; return UNOMOL_OK; }
    if (!strcmp(name, "bra_split")) { h->bra_split_enabled = value != 0.0; return UNOMOL_OK; }
    if (!strcmp(name, "value_cut")) { h->value_cut = value; return UNOMOL_OK; }
    if (!strcmp(name, "debug_flags")) { h->debug_flags = (int)value; return UNOMOL_OK; }
    // two-root quadrature: 0 = reproduce the reference's behaviour for 15 < X <= 40 (parity, default), 1 = exact.  The
    // Schwarz bounds depend on it, so the pair tables are rebuilt.
    if (!strcmp(name, "rys2_exact")) { h->rys.rys2_exact = value != 0.0; h->pairs_ready = false; return UNOMOL_OK; }
    if (!strcmp(name, "tile_kernels")) {
        h->use_tile_kernels = (int)value;    // 0 = off, 1 = lists with enough tiles to fill the GPU, 2 = always (tests)
        if (h->pairs_ready) return build_plans(h);
        return UNOMOL_OK;
    }
    if (!strcmp(name, "dump_kernel")) { h->dump_kernel = (int)value; return UNOMOL_OK; }
    if (!strcmp(name, "tile_kslots")) { h->tile_kslots = std::max(0, std::min(6, (int)value)); h->pairs_ready = false; return UNOMOL_OK; }
    // 1: quartets with l_tot > 8 also go through the Rys quadrature (6..9 roots), which is what the reference's MPI build asks of
    // its Rys::rootN (Rys.cpp:231-312); 0 (default): the serial reference's rule, McMurchie-Davidson above l_tot = 8.  Schwarz
    // bounds of f/g pairs come from the same kernel, so the pair tables are rebuilt.
    if (!strcmp(name, "all_rys")) { h->all_rys = value != 0.0; h->pairs_ready = false; return UNOMOL_OK; }
    if (!strcmp(name, "incremental_geometry")) { h->incremental = (int)value; return UNOMOL_OK; }
    if (!strcmp(name, "reg_kernels")) {
        h->use_reg_kernels = (int)value;     // 0 = generic kernel only, 1 = by class and list length, 2 = every available class
        if (h->pairs_ready) return build_plans(h);
        return UNOMOL_OK;
    }
    return UNOMOL_E_ARG;
}

int unomol_b200_set_geometry(unomol_b200_t *h, const double *xyz) {
    if (!h || !xyz) return UNOMOL_E_ARG;
    cudaSetDevice(h->device);
    h->basis.xyz.assign(xyz, xyz + 3 * h->basis.ncen);
    // which centres moved since the tables were built?  Few (the polarisation-potential scan moves one, reference
    // RHF.hpp:351-354): rebuild only the shell pairs that contain one of their shells.
    if (h->pairs_ready && h->incremental && h->xyz_built.size() == h->basis.xyz.size()) {
        std::vector<char> cmoved(h->basis.ncen, 0);
        int nmoved = 0;
        for (int c = 0; c < h->basis.ncen; ++c)
            for (int x = 0; x < 3; ++x)
                if (h->basis.xyz[3 * c + x] != h->xyz_built[3 * c + x]) { if (!cmoved[c]) ++nmoved; cmoved[c] = 1; }
        if (nmoved == 0) return UNOMOL_OK;
        std::vector<int> moved;
        for (int sh = 0; sh < h->basis.nshell; ++sh)
            if (cmoved[h->basis.cen[sh]]) moved.push_back(sh);
        // the reserved tail of d_prims belongs to ONE set of moved shells; a different set needs a full rebuild first
        const bool same_set = h->inc_shells.empty() || h->inc_shells == moved;
        if (same_set && (long long)moved.size() * 4 <= h->basis.nshell) return update_pairs_incremental(h, moved);
    }
    return build_pairs(h);
}

int unomol_b200_fock_rhf_device(unomol_b200_t *h, const double *dP, double *dG, int async) {
    if (!h || !dP || !dG) return UNOMOL_E_ARG;
    cudaSetDevice(h->device);
    int rc = fock_device(h, 1, dP, nullptr, dG, nullptr);
    if (rc || async) return rc;
    return finish_stats(h);
}

int unomol_b200_fock_uhf_device(unomol_b200_t *h, const double *dPA, const double *dPB, double *dGA, double *dGB,
                                int async) {
    if (!h || !dPA || !dPB || !dGA || !dGB) return UNOMOL_E_ARG;
    cudaSetDevice(h->device);
    int rc = fock_device(h, 2, dPA, dPB, dGA, dGB);
    if (rc || async) return rc;
    return finish_stats(h);
}

// Host passes over the packed matrices of the host-pointer entry (staging copy of P, G += result): at 2002 functions they are
// 16 MB each and, single-threaded, were ~5 ms of the 8 ms the host-pointer call costs over the device-resident one.  Large
// matrices are cut into one range per worker thread; small ones (every molecule of the reference's test set) stay inline.
static void host_ranges(size_t n, const std::function<void(size_t, size_t)> &fn) {
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t nt = n < (size_t)1 << 19 ? 1 : std::min<size_t>(8, hw ? hw : 1);
    if (nt <= 1) { fn((size_t)0, n); return; }
    std::vector<std::thread> th;
    const size_t chunk = (n + nt - 1) / nt;
    for (size_t t = 1; t < nt; ++t) th.emplace_back([=, &fn] { fn(std::min(n, t * chunk), std::min(n, (t + 1) * chunk)); });
    fn((size_t)0, std::min(n, chunk));
    for (auto &x : th) x.join();
}

static int fock_host(unomol_b200 *h, int nspin, const double *PA, const double *PB, double *GA, double *GB) {
    cudaSetDevice(h->device);
    int rc = ensure_buffers(h);
    if (rc) return rc;
    const size_t n = h->basis.nbf, no2 = n * (n + 1) / 2;
    const double *Ph[2] = {PA, PB};
    double *Gh[2] = {GA, GB};
    for (int s = 0; s < nspin; ++s) {
        double *stage = h->h_pinned + s * no2;
        const double *srcp = Ph[s];
        host_ranges(no2, [=](size_t a, size_t b) { memcpy(stage + a, srcp + a, sizeof(double) * (b - a)); });
        CUDA_TRY(h, cudaMemcpyAsync(h->d_Ppacked[s], h->h_pinned + s * no2, sizeof(double) * no2, cudaMemcpyHostToDevice,
                                    h->stream));
    }
    rc = fock_device(h, nspin, h->d_Ppacked[0], h->d_Ppacked[1], h->d_Gpacked[0], h->d_Gpacked[1]);
    if (rc) return rc;
    for (int s = 0; s < nspin; ++s)
        CUDA_TRY(h, cudaMemcpyAsync(h->h_pinned + (2 + s) * no2, h->d_Gpacked[s], sizeof(double) * no2,
                                    cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int s = 0; s < nspin; ++s) {
        const double *src = h->h_pinned + (2 + s) * no2;
        double *dst = Gh[s];
        host_ranges(no2, [=](size_t a, size_t b) { for (size_t i = a; i < b; ++i) dst[i] += src[i]; });   // the reference accumulates into G
    }
    return finish_stats(h);
}

int unomol_b200_fock_rhf(unomol_b200_t *h, const double *P, double *G) {
    if (!h || !P || !G) return UNOMOL_E_ARG;
    return fock_host(h, 1, P, nullptr, G, nullptr);
}

int unomol_b200_fock_uhf(unomol_b200_t *h, const double *PA, const double *PB, double *GA, double *GB) {
    if (!h || !PA || !PB || !GA || !GB) return UNOMOL_E_ARG;
    return fock_host(h, 2, PA, PB, GA, GB);
}

int unomol_b200_stats(unomol_b200_t *h, unomol_b200_stats_t *out) {
    if (!h || !out) return UNOMOL_E_ARG;
    *out = h->stats;
    return UNOMOL_OK;
}

int unomol_b200_steal_export(unomol_b200_t *h, void *handle64) {
    if (!h || !handle64) return UNOMOL_E_ARG;
    cudaSetDevice(h->device);
    if (!h->d_work_shared) {
        CUDA_TRY(h, cudaMalloc(&h->d_work_shared, sizeof(unsigned long long) * 2 * unomol_b200::MAXPLAN));
        CUDA_TRY(h, cudaMemset(h->d_work_shared, 0, sizeof(unsigned long long) * 2 * unomol_b200::MAXPLAN));
        h->work_owner = true;
    }
    else if (h->work_owner) {
        // re-export: start from clean counters (no build may be in flight when handles are exchanged)
        CUDA_TRY(h, cudaDeviceSynchronize());
        CUDA_TRY(h, cudaMemset(h->d_work_shared, 0, sizeof(unsigned long long) * 2 * unomol_b200::MAXPLAN));
    }
    cudaIpcMemHandle_t ipc;
    CUDA_TRY(h, cudaIpcGetMemHandle(&ipc, h->d_work_shared));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    memcpy(handle64, &ipc, 64);
    h->build_count = 0;
    return UNOMOL_OK;
}

int unomol_b200_steal_import(unomol_b200_t *h, const void *handle64) {
    if (!h || !handle64) return UNOMOL_E_ARG;
    cudaSetDevice(h->device);
    if (h->d_work_shared) return UNOMOL_E_STATE;
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, handle64, 64);
    void *p = nullptr;
    CUDA_TRY(h, cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->d_work_shared = (unsigned long long *)p;
    h->work_owner = false;
    h->work_imported = true;
    h->build_count = 0;
    return UNOMOL_OK;
}

int unomol_b200_steal_share(unomol_b200_t *owner, unomol_b200_t *peer) {
    if (!owner || !peer || owner == peer) return UNOMOL_E_ARG;
    if (peer->d_work_shared) return UNOMOL_E_STATE;
    cudaSetDevice(owner->device);
    if (!owner->d_work_shared) {
        CUDA_TRY(owner, cudaMalloc(&owner->d_work_shared, sizeof(unsigned long long) * 2 * unomol_b200::MAXPLAN));
        CUDA_TRY(owner, cudaMemset(owner->d_work_shared, 0, sizeof(unsigned long long) * 2 * unomol_b200::MAXPLAN));
        owner->work_owner = true;
        owner->build_count = 0;
    }
    cudaSetDevice(peer->device);
    if (peer->device != owner->device) {
        int can = 0;
        CUDA_TRY(peer, cudaDeviceCanAccessPeer(&can, peer->device, owner->device));
        if (!can) { peer->last_error = "no peer access between the two devices"; return UNOMOL_E_CUDA; }
        cudaError_t e = cudaDeviceEnablePeerAccess(owner->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return UNOMOL_E_CUDA; }
        cudaGetLastError();
    }
    peer->d_work_shared = owner->d_work_shared;
    peer->work_owner = false;
    peer->work_borrowed = true;
    peer->build_count = owner->build_count;
    return UNOMOL_OK;
}

int unomol_b200_attach_nccl(unomol_b200_t *h, void *comm) {
    if (!h) return UNOMOL_E_ARG;
    h->nccl_comm = comm;
    return UNOMOL_OK;
}

int unomol_b200_device_buffers(unomol_b200_t *h, void **stream, double **dP, double **dG) {
    if (!h) return UNOMOL_E_ARG;
    cudaSetDevice(h->device);
    int rc = ensure_buffers(h);
    if (rc) return rc;
    if (stream) *stream = (void *)h->stream;
    if (dP) { dP[0] = h->d_Ppacked[0]; dP[1] = h->d_Ppacked[1]; }
    if (dG) { dG[0] = h->d_Gpacked[0]; dG[1] = h->d_Gpacked[1]; }
    return UNOMOL_OK;
}

int unomol_b200_schwarz(unomol_b200_t *h, double *Q) {
    if (!h || !Q) return UNOMOL_E_ARG;
    if (!h->pairs_ready) { int rc = build_pairs(h); if (rc) return rc; }
    const size_t np = (size_t)h->basis.nshell * (h->basis.nshell + 1) / 2;
    for (size_t i = 0; i < np; ++i) Q[i] = 0.0;
    for (int c = 0; c < NGROUP; ++c)
        for (auto &sp : h->cls[c].pairs) Q[sp.pairid] = sp.Q;
    return UNOMOL_OK;
}

// ---- test hooks -------------------------------------------------------------------------------------
// locate the computed orientation of shell pair (i,j): class, position, and whether (i,j) is (a,b) or (b,a)
static bool locate_pair(unomol_b200 *h, int i, int j, int &cls, int &pos, bool &swapped) {
    int hi = std::max(i, j), lo = std::min(i, j);
    int id = hi * (hi + 1) / 2 + lo;
    cls = h->pair_cls[id];
    pos = h->pair_pos[id];
    if (pos < 0) return false;
    const ShellPair &sp = h->cls[cls].pairs[pos];
    swapped = (sp.sha != i);   // stored first shell differs from the caller's first shell
    if (i == j) swapped = false;
    return true;
}

// the plan a Fock build uses for (bra group, ket group), or null
static const ComboPlan *find_plan(unomol_b200 *h, int cb, int ck) {
    for (auto &p : h->plans) if (p.cb == cb && p.ck == ck) return &p;
    return nullptr;
}

// One quartet (bra position pb of group cb, ket position pk of group ck) through the kernel a Fock build uses for the
// class: the bra-tile kernel or the register kernel in their dump mode (option "dump_kernel" = 1).  Returns 1 when
// that kernel ran, 0 when the class belongs to the generic kernel, < 0 on error.
static int dump_quartet_hot(unomol_b200 *h, int cb, int ck, int pb, int pk, double *d_out) {
    const ComboPlan *pl = find_plan(h, cb, ck);
    if (!pl || (!pl->use_tile && !pl->use_reg)) return 0;
    const PairClassList &Lb = h->cls[cb], &Lk = h->cls[ck];
    unsigned char hostbuf[128];
    memset(hostbuf, 0, sizeof(hostbuf));           // [0,32) ket counts of the tile's slots, [32,36) tile id 0, [64,128) offsets
    unsigned char *d_buf = nullptr;
    if (cudaMalloc(&d_buf, sizeof(hostbuf)) != cudaSuccess) return -UNOMOL_E_NOMEM;
    ClassTask task{};
    task.rys = h->rys;
    task.prims = h->d_prims;
    task.ket_hot = Lk.d_hot + pk;
    task.ket = Lk.d_pairs + pk;
    task.nket = 1;
    task.nranks = 1;
    task.prim_cut = h->prim_cut;
    task.value_cut = h->value_cut;
    task.task_out = reinterpret_cast<const long long *>(d_buf + 64);
    task.out = d_out;
    cudaError_t e;
    if (pl->use_tile) {
        const int slot = Lb.pos_slot[pb];
        reinterpret_cast<int *>(hostbuf)[slot % TILE_SLOTS] = 1;
        task.tbra = Lb.d_tpairs + (size_t)(slot / TILE_SLOTS) * TILE_SLOTS;
        task.ket_count = reinterpret_cast<const int *>(d_buf);
        task.tile_order = reinterpret_cast<const int *>(d_buf + 32);
        task.ntiles = 1;
        task.tile_b = tile_b_of_class(cb / NSUB);
        task.tile_maxbp = Lb.maxnp;
        task.kslots = std::min(6, pl->kslots);
        task.tile_slices = 1;
        task.chunk = 1;
        task.nspin = 1;
        cudaMemcpyAsync(d_buf, hostbuf, sizeof(hostbuf), cudaMemcpyHostToDevice, h->stream);
        e = launch_tile_class(cb / NSUB, ck / NSUB, task, 1, h->stream);
    } else {
        reinterpret_cast<int *>(hostbuf)[0] = 1;
        task.bra = Lb.d_pairs + pb;
        task.nbra = 1;
        task.ket_count = reinterpret_cast<const int *>(d_buf);
        task.chunk = 1;
        task.nspin = 1;
        cudaMemcpyAsync(d_buf, hostbuf, sizeof(hostbuf), cudaMemcpyHostToDevice, h->stream);
        e = launch_reg_class(cb / NSUB, ck / NSUB, task, 1, h->stream, false);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_buf);
    if (e != cudaSuccess) {
        h->last_error = std::string("dump_quartet_hot: ") + cudaGetErrorString(e);
        return -UNOMOL_E_CUDA;
    }
    return 1;
}

int unomol_b200_eri_quartet(unomol_b200_t *h, int ish, int jsh, int ksh, int lsh, double *out) {
    if (!h || !out) return UNOMOL_E_ARG;
    const HostBasis &B = h->basis;
    const int ns = B.nshell;
    if (ish < 0 || jsh < 0 || ksh < 0 || lsh < 0 || ish >= ns || jsh >= ns || ksh >= ns || lsh >= ns) return UNOMOL_E_ARG;
    cudaSetDevice(h->device);
    if (!h->pairs_ready) { int rc = build_pairs(h); if (rc) return rc; }
    auto nc = [&](int s) { return (B.lv[s] + 1) * (B.lv[s] + 2) / 2; };
    const int n1 = nc(ish), n2 = nc(jsh), n3 = nc(ksh), n4 = nc(lsh);
    const int ntot = n1 * n2 * n3 * n4;
    for (int i = 0; i < ntot; ++i) out[i] = 0.0;
    int c1, p1, c2, p2;
    bool sw1, sw2;
    if (!locate_pair(h, ish, jsh, c1, p1, sw1) || !locate_pair(h, ksh, lsh, c2, p2, sw2)) return UNOMOL_OK;  // exactly zero
    const bool braket_swapped = c1 < c2;   // kernels exist for bra class >= ket class
    const int cb = braket_swapped ? c2 : c1, ck = braket_swapped ? c1 : c2;
    const int pb = braket_swapped ? p2 : p1, pk = braket_swapped ? p1 : p2;
    int2 tl = make_int2(pb, pk);
    long long off0 = 0;
    DevBuf<int2> d_tl;
    DevBuf<long long> d_off;
    DevBuf<double> d_out;
    if (!d_tl.alloc(1) || !d_off.alloc(1) || !d_out.alloc(ntot)) return UNOMOL_E_NOMEM;
    CUDA_TRY(h, cudaMemsetAsync(d_out.p, 0, sizeof(double) * ntot, h->stream));
    int ran_hot = 0;
    if (h->dump_kernel == 1) {
        ran_hot = dump_quartet_hot(h, cb, ck, pb, pk, d_out.p);
        if (ran_hot < 0) return -ran_hot;
    }
    h->stats.last_dump_kernel = ran_hot;   // lets the tests assert which kernel produced the block
    if (!ran_hot) {
        CUDA_TRY(h, cudaMemcpyAsync(d_tl.p, &tl, sizeof(int2), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(d_off.p, &off0, sizeof(long long), cudaMemcpyHostToDevice, h->stream));
        ClassTask task{};
        task.rys = h->rys;
        task.bra = h->cls[cb].d_pairs; task.ket = h->cls[ck].d_pairs; task.prims = h->d_prims;
        task.nbra = h->cls[cb].n; task.nket = h->cls[ck].n;
        task.prim_cut = h->prim_cut;
        task.task_list = d_tl.p; task.task_out = d_off.p; task.ntask = 1; task.out = d_out.p;
        CUDA_TRY(h, launch_any_class(h, cb / NSUB, ck / NSUB, task, MODE_DUMP, 1, h->stream));
    }
    std::vector<double> blk(ntot);
    CUDA_TRY(h, cudaMemcpyAsync(blk.data(), d_out.p, sizeof(double) * ntot, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    // block layout [a][b][c][d] in the kernel's orientation -> caller's [i][j][k][l]
    // kernel bra pair = (A,B) shells, ket pair = (C,D)
    const ShellPair &SB = h->cls[cb].pairs[pb], &SK = h->cls[ck].pairs[pk];
    const int NA = nc(SB.sha), NB = nc(SB.shb), NC = nc(SK.sha), ND = nc(SK.shb);
    (void)NA;
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j)
            for (int k = 0; k < n3; ++k)
                for (int l = 0; l < n4; ++l) {
                    // components in the pair orientation of the caller's bra (ish,jsh) and ket (ksh,lsh)
                    int b1 = sw1 ? j : i, b2 = sw1 ? i : j;   // (first,second) of pair (ish,jsh) as stored
                    int k1 = sw2 ? l : k, k2 = sw2 ? k : l;
                    int a, b, c, d;
                    if (!braket_swapped) { a = b1; b = b2; c = k1; d = k2; }
                    else { a = k1; b = k2; c = b1; d = b2; }
                    out[((i * n2 + j) * n3 + k) * n4 + l] = blk[((a * NB + b) * NC + c) * ND + d];
                }
    return UNOMOL_OK;
}

int unomol_b200_dump_eris(unomol_b200_t *h, double thresh, unomol_twoint *buf, size_t cap, size_t *nout) {
    if (!h || !nout) return UNOMOL_E_ARG;
    cudaSetDevice(h->device);
    if (!h->pairs_ready) { int rc = build_pairs(h); if (rc) return rc; }
    const HostBasis &B = h->basis;
    auto nc = [&](int s) { return (B.lv[s] + 1) * (B.lv[s] + 2) / 2; };
    // every (bra,ket) combination of kept pairs, canonical orientation, dense offsets per combo
    struct Combo { int cb, ck; long long base; int nint; };
    std::vector<Combo> combos;
    long long total = 0;
    for (int cb = 0; cb < NGROUP; ++cb)
        for (int ck = 0; ck <= cb; ++ck) {
            const int nb = h->cls[cb].n, nk = h->cls[ck].n;
            if (!nb || !nk) continue;
            int la, lb, lc, ld;
            pair_class_l(cb / NSUB, la, lb); pair_class_l(ck / NSUB, lc, ld);
            const int nint = ((la + 1) * (la + 2) / 2) * ((lb + 1) * (lb + 2) / 2) * ((lc + 1) * (lc + 2) / 2) * ((ld + 1) * (ld + 2) / 2);
            const long long ntask = (cb == ck) ? (long long)nb * (nb + 1) / 2 : (long long)nb * nk;
            combos.push_back({cb, ck, total, nint});
            total += ntask * nint;
        }
    if (total > (1LL << 28)) return UNOMOL_E_NOMEM;   // 2 GiB of doubles: this hook is for small systems
    double *d_out = nullptr;
    CUDA_TRY(h, cudaMalloc(&d_out, sizeof(double) * std::max<long long>(total, 1)));
    if (cudaMemsetAsync(d_out, 0, sizeof(double) * std::max<long long>(total, 1), h->stream) != cudaSuccess) { cudaFree(d_out); return UNOMOL_E_CUDA; }
    for (auto &cmb : combos) {
        const int nb = h->cls[cmb.cb].n, nk = h->cls[cmb.ck].n;
        std::vector<int2> tl;
        std::vector<long long> off;
        for (int i = 0; i < nb; ++i) {
            const int jmax = (cmb.cb == cmb.ck) ? i + 1 : nk;
            for (int j = 0; j < jmax; ++j) {
                tl.push_back(make_int2(i, j));
                const long long t = (cmb.cb == cmb.ck) ? (long long)i * (i + 1) / 2 + j : (long long)i * nk + j;
                off.push_back(cmb.base + t * cmb.nint);
            }
        }
        const ComboPlan *pl = h->dump_kernel == 1 ? find_plan(h, cmb.cb, cmb.ck) : nullptr;
        if (pl && (pl->use_tile || pl->use_reg)) {
            // the kernel a Fock build uses for this class, in its dump mode: every (bra, ket) pair of the combination
            const PairClassList &Lb = h->cls[cmb.cb];
            const bool same = cmb.cb == cmb.ck;
            const size_t nslot = pl->use_tile ? Lb.slot_pos.size() : (size_t)nb;
            std::vector<int> kcnt(nslot, 0);
            std::vector<long long> offs(nslot, 0);
            for (size_t sl = 0; sl < nslot; ++sl) {
                const int pos = pl->use_tile ? Lb.slot_pos[sl] : (int)sl;
                if (pos < 0) continue;
                kcnt[sl] = same ? pos + 1 : nk;
                offs[sl] = cmb.base + (same ? (long long)pos * (pos + 1) / 2 : (long long)pos * nk) * cmb.nint;
            }
            std::vector<int> order(pl->use_tile ? Lb.ntiles : 0);
            std::iota(order.begin(), order.end(), 0);
            DevBuf<int> d_kc, d_order;
            DevBuf<long long> d_offs;
            if (!d_kc.alloc(nslot) || !d_offs.alloc(nslot) || !d_order.alloc(order.size())) { cudaFree(d_out); return UNOMOL_E_NOMEM; }
            cudaMemcpyAsync(d_kc.p, kcnt.data(), sizeof(int) * nslot, cudaMemcpyHostToDevice, h->stream);
            cudaMemcpyAsync(d_offs.p, offs.data(), sizeof(long long) * nslot, cudaMemcpyHostToDevice, h->stream);
            if (!order.empty()) cudaMemcpyAsync(d_order.p, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice, h->stream);
            ClassTask task{};
            task.rys = h->rys;
            task.bra = Lb.d_pairs; task.ket = h->cls[cmb.ck].d_pairs; task.ket_hot = h->cls[cmb.ck].d_hot; task.prims = h->d_prims;
            task.nbra = nb; task.nket = nk; task.prim_cut = h->prim_cut; task.value_cut = h->value_cut;
            task.nranks = 1; task.chunk = 1; task.nspin = 1;
            task.ket_count = d_kc.p; task.task_out = d_offs.p; task.out = d_out;
            cudaError_t e;
            if (pl->use_tile) {
                task.tbra = Lb.d_tpairs; task.tile_order = d_order.p; task.ntiles = Lb.ntiles;
                task.tile_b = tile_b_of_class(cmb.cb / NSUB); task.tile_maxbp = Lb.maxnp; task.kslots = pl->kslots;
                task.tile_slices = 1;
                e = launch_tile_class(cmb.cb / NSUB, cmb.ck / NSUB, task, std::min(Lb.ntiles, h->nsm * 8), h->stream);
            } else {
                e = launch_reg_class(cmb.cb / NSUB, cmb.ck / NSUB, task, std::min(nb, h->nsm * 16), h->stream, false);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) { cudaFree(d_out); h->last_error = cudaGetErrorString(e); return UNOMOL_E_CUDA; }
            continue;
        }
        DevBuf<int2> d_tl;
        DevBuf<long long> d_off;
        if (!d_tl.alloc(tl.size()) || !d_off.alloc(off.size())) { cudaFree(d_out); return UNOMOL_E_NOMEM; }
        cudaMemcpyAsync(d_tl.p, tl.data(), sizeof(int2) * tl.size(), cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyAsync(d_off.p, off.data(), sizeof(long long) * off.size(), cudaMemcpyHostToDevice, h->stream);
        ClassTask task{};
        task.rys = h->rys;
        task.bra = h->cls[cmb.cb].d_pairs; task.ket = h->cls[cmb.ck].d_pairs; task.prims = h->d_prims;
        task.nbra = nb; task.nket = nk; task.prim_cut = h->prim_cut;
        task.task_list = d_tl.p; task.task_out = d_off.p; task.ntask = (int)tl.size(); task.out = d_out;
        const int groups = any_groups_per_cta(cmb.cb / NSUB, cmb.ck / NSUB);
        const int grid = std::min(((int)tl.size() + groups - 1) / groups, h->nsm * 16);
        {
            cudaError_t e = launch_any_class(h, cmb.cb / NSUB, cmb.ck / NSUB, task, MODE_DUMP, grid, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) { cudaFree(d_out); h->last_error = cudaGetErrorString(e); return UNOMOL_E_CUDA; }
        }
    }
    std::vector<double> all((size_t)std::max<long long>(total, 1));
    CUDA_TRY(h, cudaMemcpy(all.data(), d_out, sizeof(double) * all.size(), cudaMemcpyDeviceToHost));
    cudaFree(d_out);
    auto combo_of = [&](int cb, int ck) -> const Combo * {
        for (auto &c : combos) if (c.cb == cb && c.ck == ck) return &c;
        return nullptr;
    };
    // the reference's loop order and canonical function filter (TwoElectronInts.cpp:541-653)
    size_t cnt = 0;
    const int ns = B.nshell;
    for (int ish = h->start_shell; ish < ns; ++ish)
        for (int jsh = 0; jsh <= ish; ++jsh) {
            int c1, p1; bool sw1;
            const bool have1 = locate_pair(h, ish, jsh, c1, p1, sw1);
            for (int ksh = 0; ksh <= ish; ++ksh)
                for (int lsh = 0; lsh <= ksh; ++lsh) {
                    int c2, p2; bool sw2;
                    const bool have2 = locate_pair(h, ksh, lsh, c2, p2, sw2);
                    const double *blk = nullptr;
                    bool bks = false;
                    int NB = 1, NC = 1, ND = 1;
                    if (have1 && have2) {
                        bks = (c1 < c2) || (c1 == c2 && p1 < p2);
                        const int cb = bks ? c2 : c1, ck = bks ? c1 : c2, pb = bks ? p2 : p1, pk = bks ? p1 : p2;
                        const Combo *cmb = combo_of(cb, ck);
                        const long long t = (cb == ck) ? (long long)pb * (pb + 1) / 2 + pk : (long long)pb * h->cls[ck].n + pk;
                        blk = all.data() + cmb->base + t * cmb->nint;
                        const ShellPair &SB = h->cls[cb].pairs[pb], &SK = h->cls[ck].pairs[pk];
                        NB = nc(SB.shb); NC = nc(SK.sha); ND = nc(SK.shb);
                    }
                    for (int ils = 0; ils < nc(ish); ++ils) {
                        const int ir = B.off[ish] + ils;
                        for (int jls = 0; jls < nc(jsh); ++jls) {
                            const int jr = B.off[jsh] + jls;
                            if (jr > ir) break;
                            for (int kls = 0; kls < nc(ksh); ++kls) {
                                const int kr = B.off[ksh] + kls;
                                if (kr > ir) break;
                                for (int lls = 0; lls < nc(lsh); ++lls) {
                                    const int lr = B.off[lsh] + lls;
                                    if (lr > kr || (ir == kr && lr > jr)) break;
                                    double v = 0.0;
                                    if (blk) {
                                        int b1 = sw1 ? jls : ils, b2 = sw1 ? ils : jls;
                                        int k1 = sw2 ? lls : kls, k2 = sw2 ? kls : lls;
                                        int a, b, c, d;
                                        if (!bks) { a = b1; b = b2; c = k1; d = k2; }
                                        else { a = k1; b = k2; c = b1; d = b2; }
                                        v = blk[((a * NB + b) * NC + c) * ND + d];
                                    }
                                    if (std::fabs(v) > thresh) {
                                        if (buf && cnt < cap) { buf[cnt].val = v; buf[cnt].i = ir; buf[cnt].j = jr; buf[cnt].k = kr; buf[cnt].l = lr; }
                                        ++cnt;
                                    }
                                }
                            }
                        }
                    }
                }
        }
    *nout = cnt;
    return UNOMOL_OK;
}

int unomol_b200_sample_quartets(unomol_b200_t *h, long long nsample, unsigned long long seed, int *shells,
                                long long *ntotal) {
    if (!h || nsample < 0 || (nsample > 0 && !shells)) return UNOMOL_E_ARG;
    cudaSetDevice(h->device);
    if (!h->pairs_ready) { int rc = build_pairs(h); if (rc) return rc; }
    // cumulative quartet counts: per plan, per bra (recomputed on the host exactly like build_plans)
    struct Row { int plan, bra; long long cum; };
    std::vector<Row> rows;
    long long total = 0;
    std::vector<std::vector<int>> kcs(h->plans.size());
    for (size_t ip = 0; ip < h->plans.size(); ++ip) {
        const ComboPlan &pl = h->plans[ip];
        const PairClassList &Lb = h->cls[pl.cb], &Lk = h->cls[pl.ck];
        for (int i = 0; i < pl.nbra_eff; ++i) {
            int cnt;
            if (h->tau <= 0.0) cnt = Lk.n;
            else {
                const double need = h->tau / Lb.pairs[i].Q;
                int lo = 0, hi = Lk.n;
                while (lo < hi) { int mid = (lo + hi) / 2; if (Lk.pairs[mid].Q >= need) lo = mid + 1; else hi = mid; }
                cnt = lo;
            }
            if (pl.cb == pl.ck) cnt = std::min(cnt, i + 1);
            if (cnt <= 0) continue;
            total += cnt;
            rows.push_back({(int)ip, i, total});
        }
    }
    if (ntotal) *ntotal = total;
    if (total == 0) return UNOMOL_OK;
    unsigned long long st = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    auto next = [&]() { st ^= st >> 12; st ^= st << 25; st ^= st >> 27; return st * 2685821657736338717ULL; };
    for (long long q = 0; q < nsample; ++q) {
        const long long r = (long long)(next() % (unsigned long long)total);
        size_t lo = 0, hi = rows.size();
        while (lo < hi) { size_t mid = (lo + hi) / 2; if (rows[mid].cum > r) hi = mid; else lo = mid + 1; }
        const Row &row = rows[lo];
        const long long before = lo ? rows[lo - 1].cum : 0;
        const int ki = (int)(r - before);
        const ComboPlan &pl = h->plans[row.plan];
        const ShellPair &B = h->cls[pl.cb].pairs[row.bra], &K = h->cls[pl.ck].pairs[ki];
        shells[4 * q + 0] = B.sha; shells[4 * q + 1] = B.shb; shells[4 * q + 2] = K.sha; shells[4 * q + 3] = K.shb;
    }
    return UNOMOL_OK;
}

namespace ub200 {
__global__ void dfma_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace ub200

double unomol_b200_model_flops(int la, int lb, int lc, int ld) { return model_flops_per_primitive_quartet(la, lb, lc, ld); }

int unomol_b200_fp64_peak(int device, double *tflops) {
    if (!tflops) return UNOMOL_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return UNOMOL_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return UNOMOL_E_CUDA;
    const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 1 << 15;
    double *d = nullptr;
    if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return UNOMOL_E_NOMEM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return UNOMOL_E_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return UNOMOL_OK;
}

}  // extern "C"
