// unomol_b200/csrc/eri_tile_classes.cu -- instantiations + dispatcher of the bra-tile / ket-stationary kernels (eri_tile.cuh).
#include "eri_tile.cuh"
#include <atomic>
#include "engine.h"

namespace ub200 {

template <int LA, int LB, int LC, int LD, int NSPIN>
static cudaError_t launch_tile_inst(const ClassTask &task, int grid, size_t smem, cudaStream_t stream) {
    static std::atomic<size_t> attr_smem_dev[64];   // per device: largest dynamic shared memory size enabled so far
    int dev = 0;
    cudaGetDevice(&dev);
    std::atomic<size_t> &have = attr_smem_dev[dev & 63];
    if (smem > have) {
        cudaError_t e = cudaFuncSetAttribute(eri_tile_kernel<LA, LB, LC, LD, NSPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        have = smem;
    }
    eri_tile_kernel<LA, LB, LC, LD, NSPIN><<<grid, TILE_THREADS, smem, stream>>>(task);
    return cudaGetLastError();
}

template <int LA, int LB, int LC, int LD>
static cudaError_t launch_tile(const ClassTask &task, int grid, cudaStream_t stream) {
    using C = QC<LA, LB, LC, LD>;
    if (grid <= 0) return cudaSuccess;
    const TileLayout lay = tile_layout(task.tile_maxbp, task.tile_b, C::NAB, task.kslots, C::GJ > 1 ? 5 : 3,
                                       tile_boys_entries(C::NR, task.rys.rys2_exact));
    if (task.nspin == 2) return launch_tile_inst<LA, LB, LC, LD, 2>(task, grid, lay.total, stream);
    return launch_tile_inst<LA, LB, LC, LD, 1>(task, grid, lay.total, stream);
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
