/* include/unomol_b200.h -- C ABI of libunomol_b200.so: the B200-native, integral-direct replacement for the
 * reference's two-electron engine (class unomol::TwoElectronInts, reference TwoElectronInts.hpp:79-110).
 *
 * The reference has no FFI layer; the seam is that C++ class.  Each entry point below names the reference
 * interface it replaces.  Plain pointers and sizes only; no C++/torch types.  All matrices that cross this
 * boundary use the reference's conventions: P and G are packed lower-triangular double[nbf*(nbf+1)/2] with
 * idx = i*(i+1)/2 + j (i >= j) (reference TwoElectronInts.cpp:705-722), caller-owned, and G is ACCUMULATED
 * into (the caller zeroes it: reference RHF.hpp:89).  Every function returns 0 on success or a negative
 * UNOMOL_E_* code; unomol_b200_strerror() describes it.  There is no CPU fallback: without a CUDA device the
 * compute entry points return UNOMOL_E_CUDA.
 *
 * Threading: one host thread per handle (the reference is single-threaded and not re-entrant,
 * TwoElectronInts.cpp:529-534).  One handle drives one GPU; multi-GPU = one handle per rank (rank, nranks),
 * each computing a partial G over its share of the screened quartets; partial G's are summed either by the
 * caller (torch.distributed / NCCL all_reduce on the device buffer) or inside the library once an NCCL
 * communicator has been attached with unomol_b200_attach_nccl().
 */
#ifndef UNOMOL_B200_H
#define UNOMOL_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define UNOMOL_OK 0
#define UNOMOL_E_ARG (-1)      /* bad argument (null pointer, index out of range, ...) */
#define UNOMOL_E_CUDA (-2)     /* CUDA runtime / cuSOLVER / cuBLAS failure, or no device */
#define UNOMOL_E_UNSUPPORTED (-3) /* angular momentum above g (l > 4 per shell; reference Basis.hpp:222) */
#define UNOMOL_E_NOMEM (-4)
#define UNOMOL_E_STATE (-5)    /* call order (e.g. fock before create) */
#define UNOMOL_E_NCCL (-6)

typedef struct unomol_b200 unomol_b200_t;

/* Flattened copy of the reference's Basis / Shell / Center (reference Basis.hpp:17-256).
 * coef[] must already be through Shell::normalize (Basis.hpp:56-75); off[] = Basis::offset(shell)
 * (Basis.hpp:226-232); Cartesian components in AuxFunctions order (AuxFunctions.hpp:35-41). */
typedef struct unomol_basis_desc {
    int nshell, nbf, ncen, maxl;
    const int *npr;      /* [nshell] primitives per shell */
    const int *lv;       /* [nshell] angular momentum */
    const int *cen;      /* [nshell] centre index */
    const int *off;      /* [nshell] first basis-function index */
    const int *poff;     /* [nshell] offset of the shell's primitives in alpha/coef */
    const double *alpha; /* exponents */
    const double *coef;  /* normalised contraction coefficients */
    const double *xyz;   /* [ncen*3] bohr */
} unomol_basis_desc;

/* One stored integral, layout-compatible with the reference's TwoInts (TwoElectronInts.hpp:20-23). */
typedef struct unomol_twoint {
    double val;
    int i, j, k, l;
} unomol_twoint;

typedef struct unomol_b200_stats_t {
    long long n_shell_pairs;      /* i>=j shell pairs in range */
    long long n_pairs_kept;       /* after the exact primitive-threshold prune + Schwarz prune */
    long long n_prim_pairs;       /* primitive pairs stored */
    long long n_quartets;         /* contracted shell quartets evaluated in the last Fock build (this rank) */
    long long n_quartets_total;   /* canonical shell quartets before screening (all ranks) */
    double model_flops;           /* SURVEY.md 8(d) flop model summed over the quartets of the last build, counting
                                     every primitive quartet of each evaluated shell quartet */
    double last_fock_ms;          /* device time of the last Fock build (CUDA events on the build stream) */
    double last_eri_kernel_ms;    /* of which: fused ERI+digestion kernels */
    double precompute_ms;         /* pair data + Schwarz bounds of the last create/set_geometry */
    int n_launches;               /* kernels launched by the last Fock build */
    int nbf, nshell, rank, nranks;
    long long n_prim_quartets;    /* primitive quartets that passed the reference's sr cut in the last build */
    long long n_prim_candidates;  /* primitive quartets tested against the cut (register kernels only) */
    int n_tile_launches;          /* of n_launches: bra-tile / ket-stationary kernels (eri_tile.cuh) */
    int n_reg_launches;           /* one-bra-per-CTA register kernels (eri_reg.cuh); n_rows_launches of them staged rows of P */
    int n_rows_launches;
    int n_generic_launches;       /* shared-memory class kernels (eri_generic.cuh) */
    int n_highl_launches;         /* runtime-L kernel (f/g shells) */
    int last_dump_kernel;         /* eri_quartet: 0 = generic kernel produced the block, 1 = the Fock build's own kernel */
    int n_incremental_updates;    /* set_geometry calls served by the incremental path (few moved centres) since create */
    double onee_ms;               /* device time of the last unomol_b200_one_electron kernel */
} unomol_b200_stats_t;

/* Replaces the TwoElectronInts constructor (TwoElectronInts.hpp:83-88) + calculate() set-up
 * (TwoElectronInts.cpp:511-697): instead of computing and caching every integral it builds shell-pair data,
 * Schwarz bounds and class-sorted pair lists on the GPU.  start_shell has the reference's meaning
 * (quartets whose largest shell index >= start_shell, TwoElectronInts.cpp:541).  device = CUDA ordinal,
 * rank/nranks = this handle's share of the quartets (0/1 for a single GPU). */
int unomol_b200_create(const unomol_basis_desc *basis, int start_shell, int device, int rank, int nranks,
                       unomol_b200_t **out);
void unomol_b200_destroy(unomol_b200_t *h);

/* Options (before the next set_geometry/fock call).  Names:
 *   "schwarz_tau"   quartet kept iff Q_ab*Q_cd >= tau (default 1e-12; 0 keeps every quartet)
 *   "prim_cut"      the reference's primitive-quartet cut sr < cut (TwoElectronInts.cpp:479; default 1e-12)
 *   "value_cut"     shell-quartet blocks whose largest |integral| is <= value_cut are not digested (default 1e-14, the
 *                   reference's storage threshold, TwoElectronInts.cpp:513,667-671; 0 digests everything)
 * plus tuning/experiment switches documented in unomol_b200/csrc/engine.h (reg_kernels, stage_rows, col_blocks,
 * bucket_min_pairs, device_pairs, work_stealing, debug_flags).  Unknown names return UNOMOL_E_ARG. */
int unomol_b200_set_option(unomol_b200_t *h, const char *name, double value);

/* Replaces recalculate() (TwoElectronInts.hpp:96-98) after the caller moved centres
 * (Basis::SetCenterPosition, Basis.hpp:347-349): xyz = ncen*3 doubles. */
int unomol_b200_set_geometry(unomol_b200_t *h, const double *xyz);

/* Replaces formGmatrix(Pmat,Gmat) (TwoElectronInts.cpp:822-844 + kernel :699-747): G += 2J[P] - K[P],
 * P = C_occ C_occ^T (no factor 2).  Host pointers; H2D/D2H copies happen inside.  With nranks>1 and no NCCL
 * communicator attached the result is this rank's PARTIAL G. */
int unomol_b200_fock_rhf(unomol_b200_t *h, const double *P, double *G);
/* Replaces formGmatrix(PA,PB,GA,GB) (TwoElectronInts.cpp:846-869 + kernel2 :749-820):
 * G^s += J[PA+PB] - K[P^s]. */
int unomol_b200_fock_uhf(unomol_b200_t *h, const double *PA, const double *PB, double *GA, double *GB);

/* Device-resident variants: dP / dG are DEVICE pointers to packed matrices (G is overwritten with the
 * partial 2J-K, not accumulated), enqueued on the library stream and synchronised before return unless
 * async != 0.  Used by the on-device SCF step and by bench.py's kernel-only timing. */
int unomol_b200_fock_rhf_device(unomol_b200_t *h, const double *dP, double *dG, int async);
int unomol_b200_fock_uhf_device(unomol_b200_t *h, const double *dPA, const double *dPB, double *dGA, double *dGB,
                                int async);

/* Test hook with no reference counterpart in the class (the reference's per-quartet routine is
 * calc_two_electron_ints_rys, TwoElectronInts.cpp:420-509): all Cartesian components of the ordered shell
 * quartet (ish jsh|ksh lsh), out[((a*nb+b)*nc+c)*nd+d]. */
int unomol_b200_eri_quartet(unomol_b200_t *h, int ish, int jsh, int ksh, int lsh, double *out);

/* The reference's stored-integral list (cache of TwoInts, TwoElectronInts.cpp:623-671): unique function
 * quartets i>=j, i>=k, k>=l, (i==k => j>=l), |val|>thresh, in the reference's loop order.  Small systems only
 * (O(nbf^4) device memory).  *n receives the number of records; at most cap are written. */
int unomol_b200_dump_eris(unomol_b200_t *h, double thresh, unomol_twoint *buf, size_t cap, size_t *n);

/* Schwarz bounds Q_ij = sqrt(max|(ij|ij)|) for every shell pair i>=j, packed lower-triangular [nshell*(nshell+1)/2]. */
int unomol_b200_schwarz(unomol_b200_t *h, double *Q);

int unomol_b200_stats(unomol_b200_t *h, unomol_b200_stats_t *out);

/* Attach an initialised NCCL communicator (ncclComm_t passed as void*): fock_* then all-reduces the partial G
 * over it (ncclAllReduce, ncclDouble, ncclSum) before returning.  Replaces MPI_Reduce(Gbuf->Gmat)
 * (reference RHF_MPI.hpp:108).  The symbol is resolved from the already-loaded libnccl at run time. */
int unomol_b200_attach_nccl(unomol_b200_t *h, void *nccl_comm);

/* Work stealing across the GPUs of one box (north star: "static-plus-work-stealing split of screened quartets").
 * Rank 0 calls steal_export: it allocates the per-launch work counters on its GPU and returns a 64-byte CUDA IPC
 * handle; the caller ships the 64 bytes to the other ranks (torch.distributed broadcast, MPI_Bcast ...), each of
 * which calls steal_import.  From then on every CTA of every rank claims bras of the Schwarz-sorted lists from the
 * same counters with system-scope atomics over NVLink, heaviest first, so no GPU idles while another still has
 * work.  CONTRACT: every rank performs every Fock build, and a collective over all ranks (the all-reduce of G)
 * separates consecutive builds.  Without these calls a multi-rank handle uses the static snake-order split. */
int unomol_b200_steal_export(unomol_b200_t *h, void *handle64);
int unomol_b200_steal_import(unomol_b200_t *h, const void *handle64);
/* The same for handles that live in ONE process (one host thread per GPU, as unomol_b200/host/TwoElectronInts.hpp does
 * with UNOMOL_GPUS=N): CUDA IPC cannot re-open a handle inside its own process, so `peer` maps `owner`'s work counters
 * through CUDA peer access instead.  Same contract: every handle performs every build; the builds of all handles are
 * joined before the next one starts.  Returns UNOMOL_E_CUDA when the two devices cannot access each other. */
int unomol_b200_steal_share(unomol_b200_t *owner, unomol_b200_t *peer);

/* Device pointers of the square density / partial-G work buffers and the stream the library launches on
 * (for callers that keep the SCF on the device or time with their own events). */
int unomol_b200_device_buffers(unomol_b200_t *h, void **stream, double **dP_packed, double **dG_packed);

/* ---- device-resident SCF algebra (replaces SymmPack::rsp / sp_trans and formCmatrix / formPmatrix,
 * reference SymmPack.cpp:272-348, RHF.hpp:178-203, on cuSOLVER/cuBLAS; no CPU fallback) -------------------
 * All host-pointer matrices are packed lower-triangular unless stated. */
/* X = U s^-1/2 from the overlap (formXmatrix, RHF.hpp:214-233); kept on the device in the handle. */
int unomol_b200_scf_set_overlap(unomol_b200_t *h, const double *S);
/* F = H + G (packed, host) -> F' = X^T F X -> eigen-decomposition -> C = X W -> P = C_occ C_occ^T.
 * evals[nbf]; C (optional, may be NULL) row-major nbf*nbf with eigenvectors in columns (RHF.hpp:178-190);
 * P packed. */
int unomol_b200_scf_diag(unomol_b200_t *h, const double *F, int nocc, double *evals, double *C, double *P);

/* Device-resident RHF iteration = RestrictedHartreeFock::scf_converger + update (reference RHF.hpp:536-569, 87-112) with the
 * density, G, F and the core Hamiltonian kept on the GPU: per iteration only two scalars cross the bus.
 *   scf_load           after scf_set_overlap: packed core Hamiltonian H and starting density P (host pointers).
 *   scf_iterate_rhf    damp != 0 first replaces P by (P + P_previous)/2 (scf_converger's mixing branch; the caller
 *                      decides from the sign of the last energy change like the reference), then G = 2J-K[P],
 *                      *e_elec = 2 tr(PH) + tr(PG), F = H + G, F' = X^T F X, eigen-decomposition, C = X W,
 *                      P <- C_occ C_occ^T, *pdiff = ||P - P_previous||_F / nbf (SymmPackDiffNorm).
 *   ..._begin/_finish  the same in two halves for multi-rank callers: _begin enqueues the mixing and this rank's partial
 *                      Fock build into the packed device G (unomol_b200_device_buffers), the caller all-reduces that
 *                      buffer on the library stream (or attaches NCCL and lets the library do it), _finish does the rest.
 *   scf_fetch          current P (packed), orbital energies, C (row-major, eigenvectors in columns); any may be NULL. */
int unomol_b200_scf_load(unomol_b200_t *h, const double *H, const double *P);
int unomol_b200_scf_iterate_rhf(unomol_b200_t *h, int nocc, int damp, double *e_elec, double *pdiff);
int unomol_b200_scf_iterate_rhf_begin(unomol_b200_t *h, int damp);
int unomol_b200_scf_iterate_rhf_finish(unomol_b200_t *h, int nocc, double *e_elec, double *pdiff);
int unomol_b200_scf_fetch(unomol_b200_t *h, double *P, double *evals, double *C);

/* Device-resident UHF iteration = UnRestrictedHartreeFock::scf_converger + update (reference UHF.hpp:690-740, 101-134):
 * both spin densities, both G, F and H stay on the GPU.  damp != 0 first mixes BOTH densities with their predecessors;
 * then G_alpha, G_beta = J[PA+PB] - K[P_sigma], *e_elec = tr(PA H) + tr(PB H) + (tr(PA GA) + tr(PB GB))/2, and per spin
 * F = H + G_sigma -> X^T F X -> eigen-decomposition -> P_sigma; *pdiff = |dPA|/nbf + |dPB|/nbf (SymmPackDiffNorm, summed). */
int unomol_b200_scf_load_uhf(unomol_b200_t *h, const double *H, const double *PA, const double *PB);
int unomol_b200_scf_iterate_uhf(unomol_b200_t *h, int nocc_a, int nocc_b, int damp, double *e_elec, double *pdiff);
int unomol_b200_scf_fetch_uhf(unomol_b200_t *h, double *PA, double *PB, double *evals_a, double *evals_b);

/* One-electron matrices on the device: replaces OneElectronInts(bas, S, T, H) (reference OneElectronInts.cpp:127-202) and,
 * with M != NULL, MomentInts (Moments.cpp:94-187).  charge[ncen] = nuclear charges (a centre with charge 0, e.g. the skip
 * centre of the polarisation scan, attracts nothing).  Outputs are packed lower-triangular host arrays [no2]: overlap S,
 * kinetic energy T, core Hamiltonian H = T + V; M = 9 consecutive matrices dx dy dz qxx qxy qxz qyy qyz qzz (the reference's
 * MomInts order) about the origin, or NULL.  Geometry = the handle's current one (create / set_geometry). */
int unomol_b200_one_electron(unomol_b200_t *h, const double *charge, double *S, double *T, double *H, double *M);
/* ... plus the positron charge model of the polarisation scan folded into H: H -= GDPMInts (reference GDPMInts.cpp:86-145,
 * called at RHF.hpp:317,356) for the model centred on centre dpm_center (pass that centre with charge 0: the reference skips
 * it in the nuclear attraction, OneElectronInts.cpp:43).  dpm_center = -1: same as unomol_b200_one_electron. */
int unomol_b200_one_electron_dpm(unomol_b200_t *h, const double *charge, int dpm_center, double *S, double *T, double *H, double *M);

/* Bench support (no reference counterpart).
 * sample_quartets: draws nsample shell quartets uniformly from the screened canonical quartet list the Fock
 * build evaluates (all ranks), deterministic in seed; shells[4*q..] = (ish,jsh,ksh,lsh).  *ntotal receives the
 * size of the list.  Used to time the reference on "the same screened quartet list" (SURVEY.md 8(d)).
 * fp64_peak: measured DFMA throughput of the device in TFLOP/s (dependent-chain-free FMA loop on every SM). */
int unomol_b200_sample_quartets(unomol_b200_t *h, long long nsample, unsigned long long seed, int *shells,
                                long long *ntotal);
int unomol_b200_fp64_peak(int device, double *tflops);
/* algorithmic FLOPs of the reference's Rys algorithm per PRIMITIVE quartet of class (la lb|lc ld);
 * this is the per-unit figure behind stats.model_flops and bench.py's roofline.achieved.  Host only. */
double unomol_b200_model_flops(int la, int lb, int lc, int ld);

const char *unomol_b200_strerror(int code);
const char *unomol_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
