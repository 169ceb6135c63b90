// unomol_b200/csrc/eri_tile_classes.cu -- instantiations + dispatcher of the bra-tile / ket-stationary kernels (eri_tile.cuh).
#include "eri_tile.cuh"
#include <atomic>
#include "engine.h"

namespace ub200 {

template <int LA, int LB, int LC, int LD, int NSPIN>
static cudaError_t launch_tile_inst(ClassTask task, int grid, size_t smem_staged, size_t smem_plain, cudaStream_t stream) {
    static std::atomic<size_t> attr_smem_dev[64];   // per device: largest dynamic shared memory size enabled so far
    int dev = 0;
    cudaGetDevice(&dev);
    std::atomic<size_t> &have = attr_smem_dev[dev & 63];
    if (smem_staged > have) {
        cudaError_t e = cudaFuncSetAttribute(eri_tile_kernel<LA, LB, LC, LD, NSPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_staged);
        if (e != cudaSuccess) return e;
        have = smem_staged;
    }
    // The Boys table is staged in shared memory unless that costs a resident CTA (these kernels are bound by the latency of
    // dependent FP64 instructions, i.e. by warps per scheduler: (ps|ss) with six ket-primitive slots dropped from three CTAs to
    // two when the table grew by 4.6 KB, 16 -> 21 ms); then it is read through L1 instead.
    size_t smem = smem_staged;
    task.stage_table = 1;
    if (smem_plain < smem_staged) {
        int occ_staged = 0, occ_plain = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_staged, eri_tile_kernel<LA, LB, LC, LD, NSPIN>, TILE_THREADS, smem_staged);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_plain, eri_tile_kernel<LA, LB, LC, LD, NSPIN>, TILE_THREADS, smem_plain);
        if (occ_plain > occ_staged) { smem = smem_plain; task.stage_table = 0; }
    }
    eri_tile_kernel<LA, LB, LC, LD, NSPIN><<<grid, TILE_THREADS, smem, stream>>>(task);
    return cudaGetLastError();
}

template <int LA, int LB, int LC, int LD>
static cudaError_t launch_tile(const ClassTask &task, int grid, cudaStream_t stream) {
    using C = QC<LA, LB, LC, LD>;
    if (grid <= 0) return cudaSuccess;
    const int nf2 = C::GJ > 1 ? 5 : 3;
    const TileLayout lay = tile_layout(task.tile_maxbp, task.tile_b, C::NAB, task.kslots, nf2, tile_boys_entries(C::NR, task.rys.rys2_exact));
    const TileLayout plain = tile_layout(task.tile_maxbp, task.tile_b, C::NAB, task.kslots, nf2, 0);
    if (task.nspin == 2) return launch_tile_inst<LA, LB, LC, LD, 2>(task, grid, lay.total, plain.total, stream);
    return launch_tile_inst<LA, LB, LC, LD, 1>(task, grid, lay.total, plain.total, stream);
}

bool tile_class_available(int cb, int ck) {
    switch (cb * 8 + ck) {
        case 0 * 8 + 0: case 1 * 8 + 0: case 1 * 8 + 1: case 2 * 8 + 0: case 2 * 8 + 1:
            return true;   // (pp|pp) stays with eri_reg.cuh: its tile version spills 840 bytes per thread
    }
    return false;
}

// bras per tile for a bra class: the per-thread J_ab partials cost tile_b * NAB * 8 bytes of shared memory per thread
int tile_b_of_class(int cb) { return cb <= 1 ? 8 : 4; }

size_t tile_smem_bytes(int cb, int ck, int maxbp, int kslots, int rys2_exact) {
    int la, lb, lc, ld;
    pair_class_l(cb, la, lb);
    pair_class_l(ck, lc, ld);
    const int nab = ((la + 1) * (la + 2) / 2) * ((lb + 1) * (lb + 2) / 2);
    const int nr = (la + lb + lc + ld) / 2 + 1;
    return tile_layout(maxbp, tile_b_of_class(cb), nab, kslots, (lc + ld) > 0 ? 5 : 3, tile_boys_entries(nr, rys2_exact)).total;
}

cudaError_t launch_tile_class(int cb, int ck, const ClassTask &task, int grid, cudaStream_t stream) {
    switch (cb * 8 + ck) {
        case 0 * 8 + 0: return launch_tile<0, 0, 0, 0>(task, grid, stream);
        case 1 * 8 + 0: return launch_tile<1, 0, 0, 0>(task, grid, stream);
        case 1 * 8 + 1: return launch_tile<1, 0, 1, 0>(task, grid, stream);
        case 2 * 8 + 0: return launch_tile<1, 1, 0, 0>(task, grid, stream);
        case 2 * 8 + 1: return launch_tile<1, 1, 1, 0>(task, grid, stream);
    }
    return cudaErrorNotSupported;
}

}  // namespace ub200
