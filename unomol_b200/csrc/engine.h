// unomol_b200/csrc/engine.h -- host-side engine behind the C ABI (include/unomol_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "../../include/unomol_b200.h"
#include "unomol_types.h"

struct unomol_b200;

namespace ub200 {

// launchers, one per quartet class, defined in eri_class_*.cu
cudaError_t launch_quartet_class(int bra_class, int ket_class, const ClassTask &task, int mode, int grid,
                                 cudaStream_t stream);
int class_groups_per_cta(int bra_class, int ket_class);
// register-resident kernels for the small classes (eri_reg_classes.cu)
bool reg_class_available(int bra_class, int ket_class);
int reg_max_bra_prims();
bool reg_rows_fit(int bra_class, int ld);   // the bra's rows of P fit the register kernels' shared-memory stage
// device-side shell-pair / primitive-pair tables (pair_device.cu)
int build_pair_tables_device(unomol_b200 *h, double prune_cut, std::vector<ShellPair> &kept, std::vector<int> &cls, PrimPair **d_prims_out,
                             long long *nprim_out);
cudaError_t launch_reg_class(int bra_class, int ket_class, const ClassTask &task, int grid, cudaStream_t stream, bool allow_rows);
// bra-tile / ket-stationary kernels for the s/p classes (eri_tile_classes.cu)
bool tile_class_available(int bra_class, int ket_class);
int tile_b_of_class(int bra_class);
size_t tile_smem_bytes(int bra_class, int ket_class, int maxbp, int kslots, int rys2_exact);
cudaError_t launch_tile_class(int bra_class, int ket_class, const ClassTask &task, int grid, cudaStream_t stream);
constexpr int TILE_SLOTS = 8;        // = TILE_MAXB of eri_tile.cuh: slots per tile in the padded tile-ordered lists
constexpr int TILE_MAX_BRA_PRIMS = 36;
// SURVEY.md 8(d) flop model per primitive quartet of class (la lb | lc ld)
double model_flops_per_primitive_quartet(int la, int lb, int lc, int ld);

inline int pair_class_id(int la, int lb) { return la * (la + 1) / 2 + lb; }
inline void pair_class_l(int cls, int &la, int &lb) {
    la = 0;
    while ((la + 1) * (la + 2) / 2 <= cls) ++la;
    lb = cls - la * (la + 1) / 2;
}
// device addresses of the generated Rys / Boys tables on the current device (rys_tables.cu)
cudaError_t rys_device_tables(RysTables *out);
// runtime-L kernel for quartets with f/g shells (eri_highl.cu)
struct HighLArgs;
cudaError_t launch_highl(const ClassTask &task, const HighLArgs &hl, int mode, int grid, cudaStream_t stream);

struct HostBasis {
    int nshell = 0, nbf = 0, ncen = 0, maxl = 0;
    std::vector<int> npr, lv, cen, off, poff;
    std::vector<double> alpha, coef, xyz;
};

struct PairClassList {
    std::vector<ShellPair> pairs;   // sorted by Q descending after Schwarz
    ShellPair *d_pairs = nullptr;
    KetHot *d_hot = nullptr;        // hot-field mirror of d_pairs
    int n = 0;
    // tile order (eri_tile.cuh): pairs regrouped by first shell, Q descending inside a shell, cut into tiles of
    // tile_b_of_class() pairs; every tile owns TILE_SLOTS slots of the padded copy d_tpairs
    ShellPair *d_tpairs = nullptr;
    std::vector<int> slot_pos;      // [ntiles * TILE_SLOTS]: position in `pairs` or -1 (padding)
    std::vector<int> pos_slot;      // [n]: slot of a pair
    int ntiles = 0;
    int maxnp = 0;                  // largest primitive-pair count in the list
    size_t cap_pairs = 0, cap_tp = 0;       // capacities of d_pairs / d_hot and of d_tpairs (grow-only)
    std::vector<KetHot> hot_host;           // host mirrors of d_hot / d_tpairs: sources of asynchronous uploads
    std::vector<ShellPair> tp_host;
};

struct ComboPlan {                  // one (bra class, ket class) launch
    int cb = 0, ck = 0;
    int nbra_eff = 0;               // leading bras that have at least one ket
    long long nquartets = 0;        // sum of ket_count (all ranks, before start_shell filter)
    long long nquartets_eff = 0;
    int *d_ket_count = nullptr;
    long long *d_ket_prefix = nullptr;   // runtime-L plans: exclusive prefix sum of the ket counts [nbra_eff + 1]
    double cost = 0.0;              // quartets x model flops: launch order (largest first)
    bool use_reg = false;           // register-resident kernel (small class, bra contraction fits the stage)
    bool use_tile = false;          // bra-tile / ket-stationary kernel (s/p classes)
    int *d_kc_tile = nullptr;       // ket counts per slot of the bra list's tile order
    int *d_tile_order = nullptr;    // tiles with work, heaviest first
    int ntiles = 0, tile_slices = 1;
    int kslots = 0, maxbp = 0;
    bool highl = false;             // contains an f or g shell: runtime-L kernel
};

}  // namespace ub200

struct unomol_b200 {
    int device = 0, rank = 0, nranks = 1, start_shell = 0;
    int nsm = 148;                  // multiprocessors of the device (grid sizing)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    static constexpr int NAUX = 4;            // class launches of one build are spread over these streams
    cudaStream_t aux[NAUX] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[NAUX] = {nullptr, nullptr, nullptr, nullptr};
    ub200::HostBasis basis;
    ub200::RysTables rys{};         // device table pointers + option "rys2_exact"
    double tau = 1e-12, prim_cut = 1e-12, value_cut = 1e-14;
    int use_reg_kernels = 1;
    int tile_kslots = 6;            // option "tile_kslots": ket primitives a thread of the tile kernels keeps in shared memory (<= 6)
    int use_tile_kernels = 1;       // option "tile_kernels": 0 falls back to the one-bra-per-CTA register kernels
    int all_rys = 0;                // option "all_rys": Rys quadrature with 6..9 roots for l_tot > 8 instead of McMurchie-Davidson
    int dump_kernel = 0;            // option "dump_kernel": 1 = eri_quartet / dump_eris run the kernel a Fock build uses for the class
    int device_pairs = 1;           // option "device_pairs": build the pair tables on the GPU (0 = threaded host path)
    int col_blocks = 0;             // option "col_blocks": spatial blocks per pair list (0 = choose from N so a launch fits L2)
    bool bra_split_enabled = true;  // option "bra_split": generic kernel deals the kets of one bra to several warps for short bra lists
    int stage_rows = 1;             // option "stage_rows": stage the bra's rows of P in shared memory (TMA) when they fit
    int debug_flags = 0;            // option "debug_flags" (profiling experiments; see ClassTask)
    int bucket_min_pairs = 20000;   // primitive-count bucketing only pays off for large pair lists   // option "reg_kernels": 0 forces the generic kernel for every class
    bool pairs_ready = false;
    // pair data
    ub200::PairClassList cls[ub200::NGROUP];   // indexed by group id (class * NSUB + bucket * NBLOCK + spatial block)
    std::vector<ub200::PrimPair> h_prims;
    ub200::PrimPair *d_prims = nullptr;
    std::vector<int> pair_cls, pair_pos;      // per canonical shell pair id: class and position (-1 = pruned)
    // incremental geometry updates (set_geometry with few moved centres: the polarisation-potential scan of the reference,
    // RHF.hpp:292-388, moves ONE centre per grid point): pairs without a moved shell keep their records, primitive pairs
    // and Schwarz bounds; the moved shells' pairs are rebuilt on the host into a reserved tail of d_prims
    bool bucketed = false;
    int nblock = 1;
    int incremental = 1;                      // option "incremental_geometry"
    long long n_prims_full = 0, d_prims_cap = 0;   // primitive pairs written by the last full build / capacity of d_prims
    std::vector<int> inc_shells;              // moved shells the reserved tail was sized for
    std::vector<double> xyz_built;            // geometry of the current tables
    std::vector<ub200::ComboPlan> plans;
    // density / Fock work buffers (device)
    double *d_Ppacked[2] = {nullptr, nullptr}, *d_Gpacked[2] = {nullptr, nullptr};
    double *d_PJ = nullptr, *d_PK[2] = {nullptr, nullptr}, *d_J = nullptr, *d_K[2] = {nullptr, nullptr};
    double *h_pinned = nullptr;               // staging for host<->device copies (4 * no2 doubles)
    unsigned long long *d_counters = nullptr; // 2 per combo
    size_t counters_cap = 0;
    unsigned char *d_plan_pool = nullptr;     // every plan's ket counts / tile orders / prefixes, one upload per build_plans
    size_t plan_pool_cap = 0;
    unsigned char *plan_stage = nullptr;      // pinned staging of the pool
    size_t plan_stage_cap = 0;
    unsigned char *d_schwarz_scratch = nullptr;
    size_t schwarz_scratch_cap = 0;
    // runtime-L kernel (f/g shells): per-CTA slabs for the Cartesian block of a quartet; its launches share one stream
    double *d_hl_scratch = nullptr;
    long long hl_slab = 0;
    static constexpr int HL_GRID = 148 * 2;
    bool has_highl = false;
    // dynamic bra scheduling: one work counter per launch.  Local (single rank) or shared: allocated on rank 0's GPU,
    // exported as a CUDA IPC handle and mapped by every other rank (system-scope atomics over NVLink); two sets,
    // alternating per build, so that the owner can reset one while the other is in use.
    static constexpr int MAXPLAN = ub200::NGROUP * (ub200::NGROUP + 1) / 2;
    unsigned long long *d_work_local = nullptr, *d_work_shared = nullptr;
    bool work_owner = false, work_imported = false, steal_enabled = true;   // option "work_stealing"
    double static_fraction = 0.0;             // option "static_fraction": share of a launch's work blocks dealt statically (N > 1).
                                              // Default 0 = pure stealing, by measurement: 8 GPUs, (H2O)154: 46.5 ms against 63.4 ms at 0.5
                                              // (the heavy first half of the cost-sorted blocks is most of the work; the stolen tail
                                              // cannot level what the static deal leaves uneven)
    bool work_borrowed = false;               // d_work_shared is another handle's allocation in this process (peer access)
    long long build_count = 0;
    unomol_b200_stats_t stats{};
    // SCF algebra
    void *cusolver = nullptr, *cublas = nullptr;
    double *d_X = nullptr, *d_F = nullptr, *d_W = nullptr, *d_T = nullptr, *d_evals = nullptr, *d_work = nullptr;
    int *d_info = nullptr;
    int lwork = 0;
    // device-resident RHF iteration: packed core Hamiltonian, previous density, two reduction scalars
    double *d_scfH = nullptr, *d_scfPold = nullptr, *d_scfRed = nullptr;
    // ... and UHF: previous beta density, beta orbital energies, {E, |dPA|^2 (slot 1), -, |dPB|^2 (slot 3)}
    double *d_scfPoldB = nullptr, *d_evalsB = nullptr, *d_scfRed4 = nullptr;
    // NCCL
    void *nccl_comm = nullptr;
    std::string last_error;
};

void unomol_scf_free(unomol_b200 *h);   // scf_device.cu
