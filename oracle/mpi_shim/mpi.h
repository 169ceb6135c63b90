/* oracle/mpi_shim/mpi.h -- a single-node stand-in for the seven MPI entry points the reference's MPI build uses
 * (MPI_Init, MPI_Finalize, MPI_Comm_rank, MPI_Comm_size, MPI_Barrier, MPI_Bcast, MPI_Reduce; reference
 * UnomolMPI.cc:12-35, TwoElectronIntsMPI.cpp:274-275,350-354, RHF_MPI.hpp:34-35,105-109, UHF_MPI.hpp:113-123).
 *
 * TEST / BENCH INFRASTRUCTURE ONLY.  The image has no MPI (no mpi.h, mpicc, mpirun); SURVEY.md 8(c) proposes exactly this:
 * compile the UNMODIFIED reference MPI sources against a fork + shared-memory shim so that the reference's own work
 * distribution (round-robin over the lsh loop, MPI_Reduce of the partial G) can be timed on the host cores of the GPU box.
 * Ranks are processes forked inside MPI_Init (UNOMOL_MPI_SHIM_NP, default 1); collectives go through one anonymous shared
* mapping with a process-shared pthread barrier; ranks > 0 write stdout / stderr to mpi_rankNNNN.out / .err in the working directory.  Header-only; every function is static so that several reference
 * translation units can include it.  Nothing in the product path includes or links this file. */
#ifndef UNOMOL_MPI_SHIM_H
#define UNOMOL_MPI_SHIM_H
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 8
#define MPI_INT 4
#define MPI_SUM 0
#define MPI_SUCCESS 0

#define UNOMOL_SHIM_BUF_BYTES (64u << 20)     /* per collective: a packed G of 4000 functions is 64 MB */

struct unomol_shim_shared {
    pthread_barrier_t barrier;
    int nproc;
    int pad;
    unsigned char buf[1];                      /* UNOMOL_SHIM_BUF_BYTES of scratch, then nproc slots for reductions */
};

/* one definition per process image: weak so that every including translation unit shares it */
__attribute__((weak)) struct unomol_shim_shared *unomol_shim = 0;
__attribute__((weak)) int unomol_shim_rank = 0;
__attribute__((weak)) int unomol_shim_nproc = 1;
__attribute__((weak)) size_t unomol_shim_slot_bytes = 0;

static inline int MPI_Init(int *argc, char ***argv) {
    (void)argc; (void)argv;
    const char *e = getenv("UNOMOL_MPI_SHIM_NP");
    int np = e ? atoi(e) : 1;
    if (np < 1) np = 1;
    const char *sb = getenv("UNOMOL_MPI_SHIM_SLOT_MB");
    unomol_shim_slot_bytes = (size_t)(sb ? atoi(sb) : 16) << 20;     /* per-rank reduction slot */
    size_t bytes = sizeof(struct unomol_shim_shared) + UNOMOL_SHIM_BUF_BYTES + (size_t)np * unomol_shim_slot_bytes;
    void *p = mmap(0, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) { perror("mpi shim: mmap"); exit(2); }
    unomol_shim = (struct unomol_shim_shared *)p;
    unomol_shim->nproc = np;
    pthread_barrierattr_t a;
    pthread_barrierattr_init(&a);
    pthread_barrierattr_setpshared(&a, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&unomol_shim->barrier, &a, (unsigned)np);
    unomol_shim_nproc = np;
    unomol_shim_rank = 0;
    fflush(0);
    for (int r = 1; r < np; ++r) {
        pid_t pid = fork();
        if (pid < 0) { perror("mpi shim: fork"); exit(2); }
        if (pid == 0) {
            unomol_shim_rank = r;
            /* every forked rank writes its stdout / stderr to its own file (the reference prints from all ranks through
             * unbuffered streams; sharing rank 0's descriptors interleaves the pieces of a line) */
            char name[64];
            snprintf(name, sizeof(name), "mpi_rank%04d.out", r);
            if (!freopen(name, "w", stdout)) perror("mpi shim: stdout");
            snprintf(name, sizeof(name), "mpi_rank%04d.err", r);
            if (!freopen(name, "w", stderr)) perror("mpi shim: stderr");
            break;
        }
    }
    return MPI_SUCCESS;
}
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = unomol_shim_rank; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = unomol_shim_nproc; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) {
    (void)c;
    if (unomol_shim_nproc > 1) pthread_barrier_wait(&unomol_shim->barrier);
    return MPI_SUCCESS;
}
static inline size_t unomol_shim_size(MPI_Datatype t) { return t == MPI_DOUBLE ? sizeof(double) : sizeof(int); }
static inline int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm c) {
    if (unomol_shim_nproc == 1) return MPI_SUCCESS;
    size_t bytes = (size_t)count * unomol_shim_size(t);
    if (bytes > UNOMOL_SHIM_BUF_BYTES) { fprintf(stderr, "mpi shim: broadcast of %zu bytes exceeds the buffer\n", bytes); exit(2); }
    MPI_Barrier(c);                                             /* the buffer is free */
    if (unomol_shim_rank == root) memcpy(unomol_shim->buf, buf, bytes);
    MPI_Barrier(c);
    if (unomol_shim_rank != root) memcpy(buf, unomol_shim->buf, bytes);
    MPI_Barrier(c);
    return MPI_SUCCESS;
}
static inline int MPI_Reduce(const void *send, void *recv, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
    (void)op;
    size_t bytes = (size_t)count * unomol_shim_size(t);
    if (unomol_shim_nproc == 1) { if (recv != send) memcpy(recv, send, bytes); return MPI_SUCCESS; }
    if (bytes > unomol_shim_slot_bytes) { fprintf(stderr, "mpi shim: reduction of %zu bytes exceeds the slot (UNOMOL_MPI_SHIM_SLOT_MB)\n", bytes); exit(2); }
    unsigned char *slots = unomol_shim->buf + UNOMOL_SHIM_BUF_BYTES;
    MPI_Barrier(c);
    memcpy(slots + (size_t)unomol_shim_rank * unomol_shim_slot_bytes, send, bytes);
    MPI_Barrier(c);
    if (unomol_shim_rank == root) {
        /* rank order, like a linear reduction tree: deterministic */
        if (t == MPI_DOUBLE) {
            double *out = (double *)recv;
            for (int i = 0; i < count; ++i) out[i] = 0.0;
            for (int r = 0; r < unomol_shim_nproc; ++r) {
                const double *in = (const double *)(slots + (size_t)r * unomol_shim_slot_bytes);
                for (int i = 0; i < count; ++i) out[i] += in[i];
            }
        } else {
            int *out = (int *)recv;
            for (int i = 0; i < count; ++i) out[i] = 0;
            for (int r = 0; r < unomol_shim_nproc; ++r) {
                const int *in = (const int *)(slots + (size_t)r * unomol_shim_slot_bytes);
                for (int i = 0; i < count; ++i) out[i] += in[i];
            }
        }
    }
    MPI_Barrier(c);
    return MPI_SUCCESS;
}
static inline int MPI_Finalize(void) {
    fflush(0);
    if (unomol_shim_nproc > 1) {
        MPI_Barrier(MPI_COMM_WORLD);
        if (unomol_shim_rank != 0) _exit(0);                    /* forked ranks end here; rank 0 returns to main() */
        for (int r = 1; r < unomol_shim_nproc; ++r) wait(0);
    }
    return MPI_SUCCESS;
}
#endif
