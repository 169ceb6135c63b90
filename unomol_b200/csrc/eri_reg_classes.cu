// unomol_b200/csrc/eri_reg_classes.cu -- instantiations + dispatcher of the register-resident class kernels
// (eri_reg.cuh) for the quartet classes with at most 32 contracted [e0|f0] intermediates.
#include "eri_reg.cuh"
#include <atomic>
#include "engine.h"

namespace ub200 {

template <int LA, int LB, int LC, int LD>
static cudaError_t launch_reg(const ClassTask &task, int grid, cudaStream_t stream, bool allow_rows) {
    static_assert(QC<LA, LB, LC, LD>::NEF <= 144, "register kernel is for small classes");
    using C = QC<LA, LB, LC, LD>;
    if (grid <= 0) return cudaSuccess;
    // stage the bra's rows of P in shared memory when they fit (RHF only): see eri_reg.cuh
    const size_t row_bytes = sizeof(double) * (size_t)(C::NA + C::NB) * task.nbf;
    if (allow_rows && task.nspin == 1 && row_bytes <= REG_ROWS_MAX_BYTES) {
        static std::atomic<bool> attr_done_dev[64];   // per device: one process may drive several GPUs
        int attr_dev = 0;
        cudaGetDevice(&attr_dev);
        std::atomic<bool> &attr_done = attr_done_dev[attr_dev & 63];
        if (!attr_done) {
            cudaFuncSetAttribute(eri_reg_kernel<LA, LB, LC, LD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)REG_ROWS_MAX_BYTES);
            attr_done = true;
        }
        eri_reg_kernel<LA, LB, LC, LD, true><<<grid, REG_THREADS_ROWS, row_bytes, stream>>>(task);
    } else {
        eri_reg_kernel<LA, LB, LC, LD, false><<<grid, REG_THREADS, 0, stream>>>(task);
    }
    return cudaGetLastError();
}

bool reg_class_available(int cb, int ck) {
    switch (cb * 8 + ck) {
        case 0 * 8 + 0: case 1 * 8 + 0: case 1 * 8 + 1: case 2 * 8 + 0: case 2 * 8 + 1:
        case 3 * 8 + 0: case 3 * 8 + 1: case 4 * 8 + 0: case 5 * 8 + 0:
        case 2 * 8 + 2: case 3 * 8 + 2: case 3 * 8 + 3: case 4 * 8 + 1:
        case 4 * 8 + 3: case 5 * 8 + 1: case 4 * 8 + 2:
            return true;
    }
    return false;
}

int reg_max_bra_prims() { return REG_MAX_BRA_PRIMS; }

bool reg_rows_fit(int cb, int ld) {
    int la, lb;
    pair_class_l(cb, la, lb);
    return sizeof(double) * (size_t)((la + 1) * (la + 2) / 2 + (lb + 1) * (lb + 2) / 2) * ld <= REG_ROWS_MAX_BYTES;
}

cudaError_t launch_reg_class(int cb, int ck, const ClassTask &task, int grid, cudaStream_t stream, bool allow_rows) {
    switch (cb * 8 + ck) {
        case 0 * 8 + 0: return launch_reg<0, 0, 0, 0>(task, grid, stream, allow_rows);
        case 1 * 8 + 0: return launch_reg<1, 0, 0, 0>(task, grid, stream, allow_rows);
        case 1 * 8 + 1: return launch_reg<1, 0, 1, 0>(task, grid, stream, allow_rows);
        case 2 * 8 + 0: return launch_reg<1, 1, 0, 0>(task, grid, stream, allow_rows);
        case 2 * 8 + 1: return launch_reg<1, 1, 1, 0>(task, grid, stream, allow_rows);
        case 3 * 8 + 0: return launch_reg<2, 0, 0, 0>(task, grid, stream, allow_rows);
        case 3 * 8 + 1: return launch_reg<2, 0, 1, 0>(task, grid, stream, allow_rows);
        case 4 * 8 + 0: return launch_reg<2, 1, 0, 0>(task, grid, stream, allow_rows);
        case 5 * 8 + 0: return launch_reg<2, 2, 0, 0>(task, grid, stream, allow_rows);
        case 2 * 8 + 2: return launch_reg<1, 1, 1, 1>(task, grid, stream, allow_rows);
        case 3 * 8 + 2: return launch_reg<2, 0, 1, 1>(task, grid, stream, allow_rows);
        case 3 * 8 + 3: return launch_reg<2, 0, 2, 0>(task, grid, stream, allow_rows);
        case 4 * 8 + 1: return launch_reg<2, 1, 1, 0>(task, grid, stream, allow_rows);
        case 4 * 8 + 3: return launch_reg<2, 1, 2, 0>(task, grid, stream, allow_rows);
        case 5 * 8 + 1: return launch_reg<2, 2, 1, 0>(task, grid, stream, allow_rows);
        case 4 * 8 + 2: return launch_reg<2, 1, 1, 1>(task, grid, stream, allow_rows);
    }
    return cudaErrorNotSupported;
}

}  // namespace ub200
