// unomol_b200/csrc/eri_class_inst.cuh -- explicit instantiation of the class kernels for one bra class.
// Each eri_class_<n>.cu defines UNOMOL_BRA_LA / UNOMOL_BRA_LB and includes this file, so the 21 quartet
// classes compile in parallel translation units.
#include "eri_generic.cuh"
#include "engine.h"

namespace ub200 {

#define BLA UNOMOL_BRA_LA
#define BLB UNOMOL_BRA_LB
constexpr int kBraClass = BLA * (BLA + 1) / 2 + BLB;

UNOMOL_INSTANTIATE_CLASS(BLA, BLB, 0, 0)
#if (BLA * (BLA + 1) / 2 + BLB) >= 1
UNOMOL_INSTANTIATE_CLASS(BLA, BLB, 1, 0)
#endif
#if (BLA * (BLA + 1) / 2 + BLB) >= 2
UNOMOL_INSTANTIATE_CLASS(BLA, BLB, 1, 1)
#endif
#if (BLA * (BLA + 1) / 2 + BLB) >= 3
UNOMOL_INSTANTIATE_CLASS(BLA, BLB, 2, 0)
#endif
#if (BLA * (BLA + 1) / 2 + BLB) >= 4
UNOMOL_INSTANTIATE_CLASS(BLA, BLB, 2, 1)
#endif
#if (BLA * (BLA + 1) / 2 + BLB) >= 5
UNOMOL_INSTANTIATE_CLASS(BLA, BLB, 2, 2)
#endif

#define UNOMOL_BRA_FN2(a, b) launch_bra_class_##a##b
#define UNOMOL_BRA_FN(a, b) UNOMOL_BRA_FN2(a, b)
#define UNOMOL_GRP_FN2(a, b) groups_bra_class_##a##b
#define UNOMOL_GRP_FN(a, b) UNOMOL_GRP_FN2(a, b)

cudaError_t UNOMOL_BRA_FN(UNOMOL_BRA_LA, UNOMOL_BRA_LB)(int ket_class, const ClassTask &task, int mode, int grid,
                                                        cudaStream_t stream) {
    switch (ket_class) {
        case 0: return launch_class<BLA, BLB, 0, 0>(task, mode, grid, stream);
#if (BLA * (BLA + 1) / 2 + BLB) >= 1
        case 1: return launch_class<BLA, BLB, 1, 0>(task, mode, grid, stream);
#endif
#if (BLA * (BLA + 1) / 2 + BLB) >= 2
        case 2: return launch_class<BLA, BLB, 1, 1>(task, mode, grid, stream);
#endif
#if (BLA * (BLA + 1) / 2 + BLB) >= 3
        case 3: return launch_class<BLA, BLB, 2, 0>(task, mode, grid, stream);
#endif
#if (BLA * (BLA + 1) / 2 + BLB) >= 4
        case 4: return launch_class<BLA, BLB, 2, 1>(task, mode, grid, stream);
#endif
#if (BLA * (BLA + 1) / 2 + BLB) >= 5
        case 5: return launch_class<BLA, BLB, 2, 2>(task, mode, grid, stream);
#endif
        default: return cudaErrorInvalidValue;
    }
}

int UNOMOL_GRP_FN(UNOMOL_BRA_LA, UNOMOL_BRA_LB)(int ket_class) {
    switch (ket_class) {
        case 0: return QC<BLA, BLB, 0, 0>::GROUPS;
#if (BLA * (BLA + 1) / 2 + BLB) >= 1
        case 1: return QC<BLA, BLB, 1, 0>::GROUPS;
#endif
#if (BLA * (BLA + 1) / 2 + BLB) >= 2
        case 2: return QC<BLA, BLB, 1, 1>::GROUPS;
#endif
#if (BLA * (BLA + 1) / 2 + BLB) >= 3
        case 3: return QC<BLA, BLB, 2, 0>::GROUPS;
#endif
#if (BLA * (BLA + 1) / 2 + BLB) >= 4
        case 4: return QC<BLA, BLB, 2, 1>::GROUPS;
#endif
#if (BLA * (BLA + 1) / 2 + BLB) >= 5
        case 5: return QC<BLA, BLB, 2, 2>::GROUPS;
#endif
        default: return 1;
    }
}

}  // namespace ub200
