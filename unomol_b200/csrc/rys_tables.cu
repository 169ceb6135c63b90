// unomol_b200/csrc/rys_tables.cu -- device copies of the generated Rys / Boys tables (rys_tables.inc) and the
// per-device pointer block the kernels receive through ClassTask::rys.
#include <cuda_runtime.h>
#include "rys_roots.cuh"

namespace ub200 {

#define RYS_TABLE(name, n) __device__ __align__(16) const double name[n]
#include "rys_tables.inc"
#include "rys_tables_hi.inc"
#include "rys_tables_poly.inc"
#undef RYS_TABLE

// Device addresses of the tables on the CURRENT device (symbols exist once per device).
cudaError_t rys_device_tables(RysTables *out) {
    void *p = nullptr;
    cudaError_t e;
    if ((e = cudaGetSymbolAddress(&p, rys_boys_tab)) != cudaSuccess) return e;
    out->boys = (const double *)p;
    if ((e = cudaGetSymbolAddress(&p, rys_boys1_tab)) != cudaSuccess) return e;
    out->boys1 = (const double *)p;
    if ((e = cudaGetSymbolAddress(&p, rys_boys0_tab)) != cudaSuccess) return e;
    out->boys0 = (const double *)p;
    if ((e = cudaGetSymbolAddress(&p, rys_f0poly_tab)) != cudaSuccess) return e;
    out->f0poly = (const double *)p;
    if ((e = cudaGetSymbolAddress(&p, rys_f3poly_tab)) != cudaSuccess) return e;
    out->f3poly = out->f3poly_glob = (const double *)p;
    out->expcol = out->f3poly_glob + RYS_FP_DEG + 1;
    out->exp_stride = RYS_FP_STRIDE;
    if ((e = cudaGetSymbolAddress(&p, rys_piece3_tab)) != cudaSuccess) return e;
    out->piece[0] = (const double *)p;
    if ((e = cudaGetSymbolAddress(&p, rys_piece4_tab)) != cudaSuccess) return e;
    out->piece[1] = (const double *)p;
    if ((e = cudaGetSymbolAddress(&p, rys_piece5_tab)) != cudaSuccess) return e;
    out->piece[2] = (const double *)p;
    if ((e = cudaGetSymbolAddress(&p, rys_piece6_tab)) != cudaSuccess) return e;
    out->piece_hi[0] = (const double *)p;
    if ((e = cudaGetSymbolAddress(&p, rys_piece7_tab)) != cudaSuccess) return e;
    out->piece_hi[1] = (const double *)p;
    if ((e = cudaGetSymbolAddress(&p, rys_piece8_tab)) != cudaSuccess) return e;
    out->piece_hi[2] = (const double *)p;
    if ((e = cudaGetSymbolAddress(&p, rys_piece9_tab)) != cudaSuccess) return e;
    out->piece_hi[3] = (const double *)p;
    out->rys2_exact = 0;
    return cudaSuccess;
}

}  // namespace ub200
