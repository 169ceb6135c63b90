// unomol_b200/csrc/scf_device.cu -- device-resident SCF linear algebra on cuSOLVER / cuBLAS (library calls).
//
// Replaces the reference's packed EISPACK-style eigensolver and transforms on the host:
//   formXmatrix (RHF.hpp:214-233)       -> unomol_b200_scf_set_overlap : S = U s U^T, X = U s^-1/2
//   sp_trans   (SymmPack.cpp:294-348)   -> F' = X^T F X                (2 x cublasDgemm)
//   rsp        (SymmPack.cpp:272-288)   -> cusolverDnDsyevd
//   formCmatrix (RHF.hpp:178-190)       -> C = X W                     (cublasDgemm)
//   formPmatrix (RHF.hpp:192-203)       -> P = C_occ C_occ^T           (cublasDgemm, k = nocc)
// Matrices live column-major on the device; symmetric ones are layout-agnostic.  No CPU fallback.
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <cstdio>
#include <vector>
#include "engine.h"

using namespace ub200;

namespace {

__global__ void unpack_symm_kernel(const double *__restrict__ packed, int n, double *__restrict__ full) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    int i = (int)(idx / n), j = (int)(idx % n);
    int a = i > j ? i : j, b = i > j ? j : i;
    full[idx] = packed[(size_t)a * (a + 1) / 2 + b];
}

__global__ void pack_symm_kernel(const double *__restrict__ full, int n, double *__restrict__ packed) {
    // one thread per (i,j), i >= j
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    int i = (int)(idx / n), j = (int)(idx % n);
    if (j > i) return;
    packed[(size_t)i * (i + 1) / 2 + j] = full[(size_t)j * n + i];
}

// X(:,j) = U(:,j) / sqrt(s_j)   (column-major); flags non-positive eigenvalues like RHF.hpp:219-225
__global__ void scale_columns_kernel(double *__restrict__ U, const double *__restrict__ s, int n, int *__restrict__ bad) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    int j = (int)(idx / n);
    double f = s[j];
    if (fabs(f) < 1.e-7 && f < 0.0) f = -f;
    if (!(f > 0.0)) { atomicExch(bad, 1); return; }
    U[idx] *= rsqrt(f);
}

// ---- device-resident RHF iteration (packed matrices, idx = i(i+1)/2 + j) ---------------------------------------
__device__ __forceinline__ int packed_row(size_t idx) {
    int i = (int)floor((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while ((size_t)i * (i + 1) / 2 > idx) --i;
    while ((size_t)(i + 1) * (i + 2) / 2 <= idx) ++i;
    return i;
}

__device__ __forceinline__ void block_sum_to(double v, double *dst) {
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(dst, v);
    }
}

// P <- (P + Pold)/2   (scf_converger, reference RHF.hpp:564-567)
__global__ void scf_damp_kernel(double *__restrict__ P, const double *__restrict__ Pold, size_t no2) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < no2) P[idx] = (P[idx] + Pold[idx]) * 0.5;
}

// E = 2 tr(P H) + tr(P G) with full-matrix traces (TraceSymmPackProduct, SymmPack.cpp:7-18; RHF.hpp:94-96);
// F = H + G unpacked to the square work matrix; Pold <- P
__global__ void scf_energy_fock_kernel(const double *__restrict__ P, const double *__restrict__ H, const double *__restrict__ G,
                                       double *__restrict__ Pold, double *__restrict__ Ffull, int n, double *__restrict__ red) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t no2 = (size_t)n * (n + 1) / 2;
    double e = 0.0;
    if (idx < no2) {
        const int i = packed_row(idx), j = (int)(idx - (size_t)i * (i + 1) / 2);
        const double p = P[idx], hh = H[idx], g = G[idx], f = hh + g;
        e = (i == j ? 1.0 : 2.0) * p * (2.0 * hh + g);
        Ffull[(size_t)i * n + j] = f;
        Ffull[(size_t)j * n + i] = f;
        Pold[idx] = p;
    }
    block_sum_to(e, red);
}

// packs the new density and accumulates ||P - Pold||_F^2 with the full-matrix weights (SymmPackDiffNorm, SymmPack.cpp:20-36)
__global__ void scf_pack_diff_kernel(const double *__restrict__ Pfull, const double *__restrict__ Pold, double *__restrict__ P, int n,
                                     double *__restrict__ red) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t no2 = (size_t)n * (n + 1) / 2;
    double d2 = 0.0;
    if (idx < no2) {
        const int i = packed_row(idx), j = (int)(idx - (size_t)i * (i + 1) / 2);
        const double p = Pfull[(size_t)j * n + i];
        const double d = p - Pold[idx];
        P[idx] = p;
        d2 = (i == j ? 1.0 : 2.0) * d * d;
    }
    block_sum_to(d2, red + 1);
}

// ---- device-resident UHF iteration ------------------------------------------------------------------------------
// E = tr(PA H) + tr(PB H) + (tr(PA GA) + tr(PB GB)) / 2 with full-matrix traces (reference UHF.hpp:111-114)
__global__ void scf_energy_uhf_kernel(const double *__restrict__ PA, const double *__restrict__ PB, const double *__restrict__ H,
                                      const double *__restrict__ GA, const double *__restrict__ GB, int n, double *__restrict__ red) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t no2 = (size_t)n * (n + 1) / 2;
    double e = 0.0;
    if (idx < no2) {
        const int i = packed_row(idx), j = (int)(idx - (size_t)i * (i + 1) / 2);
        const double pa = PA[idx], pb = PB[idx], hh = H[idx];
        e = (i == j ? 1.0 : 2.0) * ((pa + pb) * hh + 0.5 * (pa * GA[idx] + pb * GB[idx]));
    }
    block_sum_to(e, red);
}

// F = H + G unpacked to the square work matrix; Pold <- P
__global__ void scf_fock_unpack_kernel(const double *__restrict__ P, const double *__restrict__ H, const double *__restrict__ G,
                                       double *__restrict__ Pold, double *__restrict__ Ffull, int n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t no2 = (size_t)n * (n + 1) / 2;
    if (idx >= no2) return;
    const int i = packed_row(idx), j = (int)(idx - (size_t)i * (i + 1) / 2);
    const double f = H[idx] + G[idx];
    Ffull[(size_t)i * n + j] = f;
    Ffull[(size_t)j * n + i] = f;
    Pold[idx] = P[idx];
}

int scf_ensure(unomol_b200 *h) {
    if (h->cusolver) return UNOMOL_OK;
    const int n = h->basis.nbf;
    cusolverDnHandle_t cs;
    cublasHandle_t cb;
    if (cusolverDnCreate(&cs) != CUSOLVER_STATUS_SUCCESS) return UNOMOL_E_CUDA;
    if (cublasCreate(&cb) != CUBLAS_STATUS_SUCCESS) return UNOMOL_E_CUDA;
    cusolverDnSetStream(cs, h->stream);
    cublasSetStream(cb, h->stream);
    // the handles are published only after EVERY allocation succeeded: a later call must not find h->cusolver set next to
    // null work buffers.  On failure everything allocated so far is released.
    const size_t nn = (size_t)n * n;
    int lwork = 0;
    bool ok = cudaMalloc(&h->d_X, sizeof(double) * nn) == cudaSuccess && cudaMalloc(&h->d_F, sizeof(double) * nn) == cudaSuccess &&
              cudaMalloc(&h->d_W, sizeof(double) * nn) == cudaSuccess && cudaMalloc(&h->d_T, sizeof(double) * nn) == cudaSuccess &&
              cudaMalloc(&h->d_evals, sizeof(double) * n) == cudaSuccess && cudaMalloc(&h->d_info, sizeof(int) * 2) == cudaSuccess;
    int rc = ok ? UNOMOL_OK : UNOMOL_E_NOMEM;
    if (ok && cusolverDnDsyevd_bufferSize(cs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, h->d_W, n, h->d_evals, &lwork) !=
                  CUSOLVER_STATUS_SUCCESS) {
        ok = false;
        rc = UNOMOL_E_CUDA;
    }
    if (ok && cudaMalloc(&h->d_work, sizeof(double) * (lwork > 0 ? lwork : 1)) != cudaSuccess) {
        ok = false;
        rc = UNOMOL_E_NOMEM;
    }
    if (!ok) {
        cudaGetLastError();
        cudaFree(h->d_X); cudaFree(h->d_F); cudaFree(h->d_W); cudaFree(h->d_T); cudaFree(h->d_evals); cudaFree(h->d_info); cudaFree(h->d_work);
        h->d_X = h->d_F = h->d_W = h->d_T = h->d_evals = h->d_work = nullptr;
        h->d_info = nullptr;
        cusolverDnDestroy(cs);
        cublasDestroy(cb);
        return rc;
    }
    h->lwork = lwork;
    h->cusolver = cs;
    h->cublas = cb;
    return UNOMOL_OK;
}

}  // namespace

void unomol_scf_free(unomol_b200 *h) {
    if (h->cusolver) cusolverDnDestroy((cusolverDnHandle_t)h->cusolver);
    if (h->cublas) cublasDestroy((cublasHandle_t)h->cublas);
    h->cusolver = h->cublas = nullptr;
    cudaFree(h->d_X); cudaFree(h->d_F); cudaFree(h->d_W); cudaFree(h->d_T);
    cudaFree(h->d_evals); cudaFree(h->d_work); cudaFree(h->d_info);
    cudaFree(h->d_scfH); cudaFree(h->d_scfPold); cudaFree(h->d_scfRed);
    cudaFree(h->d_scfPoldB); cudaFree(h->d_evalsB); cudaFree(h->d_scfRed4);
    h->d_scfH = h->d_scfPold = h->d_scfRed = nullptr;
    h->d_scfPoldB = h->d_evalsB = h->d_scfRed4 = nullptr;
    h->d_X = h->d_F = h->d_W = h->d_T = h->d_evals = h->d_work = nullptr;
    h->d_info = nullptr;
}

extern "C" {

int unomol_b200_scf_set_overlap(unomol_b200_t *h, const double *S) {
    if (!h || !S) return UNOMOL_E_ARG;
    cudaSetDevice(h->device);
    int rc = scf_ensure(h);
    if (rc) return rc;
    const int n = h->basis.nbf;
    const size_t nn = (size_t)n * n, no2 = (size_t)n * (n + 1) / 2;
    double *d_packed = nullptr;
    if (cudaMalloc(&d_packed, sizeof(double) * no2) != cudaSuccess) return UNOMOL_E_NOMEM;
    cudaMemcpyAsync(d_packed, S, sizeof(double) * no2, cudaMemcpyHostToDevice, h->stream);
    unpack_symm_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, h->stream>>>(d_packed, n, h->d_X);
    cudaMemsetAsync(h->d_info, 0, sizeof(int) * 2, h->stream);
    if (cusolverDnDsyevd((cusolverDnHandle_t)h->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, h->d_X, n,
                         h->d_evals, h->d_work, h->lwork, h->d_info) != CUSOLVER_STATUS_SUCCESS) {
        cudaFree(d_packed);
        return UNOMOL_E_CUDA;
    }
    scale_columns_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, h->stream>>>(h->d_X, h->d_evals, n, h->d_info + 1);
    int info[2] = {0, 0};
    cudaMemcpyAsync(info, h->d_info, sizeof(int) * 2, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    cudaFree(d_packed);
    if (e != cudaSuccess || info[0] != 0) return UNOMOL_E_CUDA;
    if (info[1] != 0) {
        h->last_error = "Zero or negative eigenvalue in overlap matrix";
        return UNOMOL_E_ARG;
    }
    return UNOMOL_OK;
}

int unomol_b200_scf_diag(unomol_b200_t *h, const double *F, int nocc, double *evals, double *C, double *P) {
    if (!h || !F || !evals || !P || nocc < 0 || nocc > h->basis.nbf) return UNOMOL_E_ARG;
    if (!h->cusolver || !h->d_X) return UNOMOL_E_STATE;
    cudaSetDevice(h->device);
    const int n = h->basis.nbf;
    const size_t nn = (size_t)n * n, no2 = (size_t)n * (n + 1) / 2;
    cublasHandle_t cb = (cublasHandle_t)h->cublas;
    double *d_packed = nullptr;
    if (cudaMalloc(&d_packed, sizeof(double) * no2) != cudaSuccess) return UNOMOL_E_NOMEM;
    cudaMemcpyAsync(d_packed, F, sizeof(double) * no2, cudaMemcpyHostToDevice, h->stream);
    unpack_symm_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, h->stream>>>(d_packed, n, h->d_F);
    const double one = 1.0, zero = 0.0;
    // T = F X ; W = X^T T
    cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, h->d_F, n, h->d_X, n, &zero, h->d_T, n);
    cublasDgemm(cb, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, h->d_X, n, h->d_T, n, &zero, h->d_W, n);
    cudaMemsetAsync(h->d_info, 0, sizeof(int) * 2, h->stream);
    if (cusolverDnDsyevd((cusolverDnHandle_t)h->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, h->d_W, n,
                         h->d_evals, h->d_work, h->lwork, h->d_info) != CUSOLVER_STATUS_SUCCESS) {
        cudaFree(d_packed);
        return UNOMOL_E_CUDA;
    }
    // C = X W  (into d_T), P = C_occ C_occ^T (into d_F)
    cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, h->d_X, n, h->d_W, n, &zero, h->d_T, n);
    if (nocc > 0)
        cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, nocc, &one, h->d_T, n, h->d_T, n, &zero, h->d_F, n);
    else
        cudaMemsetAsync(h->d_F, 0, sizeof(double) * nn, h->stream);
    pack_symm_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, h->stream>>>(h->d_F, n, d_packed);
    cudaMemcpyAsync(P, d_packed, sizeof(double) * no2, cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(evals, h->d_evals, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
    std::vector<double> ccm;
    if (C) {
        ccm.resize(nn);
        cudaMemcpyAsync(ccm.data(), h->d_T, sizeof(double) * nn, cudaMemcpyDeviceToHost, h->stream);
    }
    int info = 0;
    cudaMemcpyAsync(&info, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    cudaFree(d_packed);
    if (e != cudaSuccess || info != 0) return UNOMOL_E_CUDA;
    if (C)
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < n; ++k) C[(size_t)i * n + k] = ccm[(size_t)k * n + i];   // row-major, eigenvectors in columns
    return UNOMOL_OK;
}


// ---- device-resident RHF iteration: the density, G, F and the core Hamiltonian stay on the GPU ---------------
int unomol_b200_scf_load(unomol_b200_t *h, const double *H, const double *P) {
    if (!h || !H || !P) return UNOMOL_E_ARG;
    if (!h->cusolver || !h->d_X) return UNOMOL_E_STATE;
    cudaSetDevice(h->device);
    double *dP[2], *dG[2];
    int rc = unomol_b200_device_buffers(h, nullptr, dP, dG);
    if (rc) return rc;
    const size_t n = h->basis.nbf, no2 = n * (n + 1) / 2;
    if (!h->d_scfH) {
        if (cudaMalloc(&h->d_scfH, sizeof(double) * no2) != cudaSuccess) return UNOMOL_E_NOMEM;
        if (cudaMalloc(&h->d_scfPold, sizeof(double) * no2) != cudaSuccess) return UNOMOL_E_NOMEM;
        if (cudaMalloc(&h->d_scfRed, sizeof(double) * 2) != cudaSuccess) return UNOMOL_E_NOMEM;
    }
    cudaMemcpyAsync(h->d_scfH, H, sizeof(double) * no2, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(dP[0], P, sizeof(double) * no2, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(h->d_scfPold, P, sizeof(double) * no2, cudaMemcpyHostToDevice, h->stream);
    return cudaStreamSynchronize(h->stream) == cudaSuccess ? UNOMOL_OK : UNOMOL_E_CUDA;
}

int unomol_b200_scf_iterate_rhf_begin(unomol_b200_t *h, int damp) {
    if (!h) return UNOMOL_E_ARG;
    if (!h->d_scfH) return UNOMOL_E_STATE;
    cudaSetDevice(h->device);
    double *dP[2], *dG[2];
    int rc = unomol_b200_device_buffers(h, nullptr, dP, dG);
    if (rc) return rc;
    const size_t n = h->basis.nbf, no2 = n * (n + 1) / 2;
    if (damp) scf_damp_kernel<<<(unsigned)((no2 + 255) / 256), 256, 0, h->stream>>>(dP[0], h->d_scfPold, no2);
    return unomol_b200_fock_rhf_device(h, dP[0], dG[0], /*async=*/1);
}

int unomol_b200_scf_iterate_rhf_finish(unomol_b200_t *h, int nocc, double *e_elec, double *pdiff) {
    if (!h || !e_elec || !pdiff || nocc < 0 || nocc > h->basis.nbf) return UNOMOL_E_ARG;
    if (!h->d_scfH) return UNOMOL_E_STATE;
    cudaSetDevice(h->device);
    double *dP[2], *dG[2];
    int rc = unomol_b200_device_buffers(h, nullptr, dP, dG);
    if (rc) return rc;
    const int n = h->basis.nbf;
    const size_t nn = (size_t)n * n, no2 = (size_t)n * (n + 1) / 2;
    cublasHandle_t cb = (cublasHandle_t)h->cublas;
    const unsigned blocks = (unsigned)((no2 + 255) / 256);
    cudaMemsetAsync(h->d_scfRed, 0, sizeof(double) * 2, h->stream);
    scf_energy_fock_kernel<<<blocks, 256, 0, h->stream>>>(dP[0], h->d_scfH, dG[0], h->d_scfPold, h->d_F, n, h->d_scfRed);
    const double one = 1.0, zero = 0.0;
    cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, h->d_F, n, h->d_X, n, &zero, h->d_T, n);
    cublasDgemm(cb, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, h->d_X, n, h->d_T, n, &zero, h->d_W, n);
    cudaMemsetAsync(h->d_info, 0, sizeof(int) * 2, h->stream);
    if (cusolverDnDsyevd((cusolverDnHandle_t)h->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, h->d_W, n,
                         h->d_evals, h->d_work, h->lwork, h->d_info) != CUSOLVER_STATUS_SUCCESS)
        return UNOMOL_E_CUDA;
    cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, h->d_X, n, h->d_W, n, &zero, h->d_T, n);   // C = X W
    if (nocc > 0)
        cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, nocc, &one, h->d_T, n, h->d_T, n, &zero, h->d_F, n);
    else
        cudaMemsetAsync(h->d_F, 0, sizeof(double) * nn, h->stream);
    scf_pack_diff_kernel<<<blocks, 256, 0, h->stream>>>(h->d_F, h->d_scfPold, dP[0], n, h->d_scfRed);
    double red[2] = {0.0, 0.0};
    int info = 0;
    cudaMemcpyAsync(red, h->d_scfRed, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(&info, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess || info != 0) return UNOMOL_E_CUDA;
    *e_elec = red[0];
    *pdiff = sqrt(red[1]) / n;
    return UNOMOL_OK;
}

int unomol_b200_scf_iterate_rhf(unomol_b200_t *h, int nocc, int damp, double *e_elec, double *pdiff) {
    int rc = unomol_b200_scf_iterate_rhf_begin(h, damp);
    if (rc) return rc;
    return unomol_b200_scf_iterate_rhf_finish(h, nocc, e_elec, pdiff);
}

// ---- device-resident UHF iteration = UnRestrictedHartreeFock::scf_converger + update (reference UHF.hpp:690-740, 101-134) ----
int unomol_b200_scf_load_uhf(unomol_b200_t *h, const double *H, const double *PA, const double *PB) {
    if (!h || !H || !PA || !PB) return UNOMOL_E_ARG;
    int rc = unomol_b200_scf_load(h, H, PA);
    if (rc) return rc;
    double *dP[2], *dG[2];
    rc = unomol_b200_device_buffers(h, nullptr, dP, dG);
    if (rc) return rc;
    const size_t n = h->basis.nbf, no2 = n * (n + 1) / 2;
    if (!h->d_scfPoldB) {
        if (cudaMalloc(&h->d_scfPoldB, sizeof(double) * no2) != cudaSuccess) return UNOMOL_E_NOMEM;
        if (cudaMalloc(&h->d_evalsB, sizeof(double) * n) != cudaSuccess) return UNOMOL_E_NOMEM;
        if (cudaMalloc(&h->d_scfRed4, sizeof(double) * 4) != cudaSuccess) return UNOMOL_E_NOMEM;
    }
    cudaMemcpyAsync(dP[1], PB, sizeof(double) * no2, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(h->d_scfPoldB, PB, sizeof(double) * no2, cudaMemcpyHostToDevice, h->stream);
    return cudaStreamSynchronize(h->stream) == cudaSuccess ? UNOMOL_OK : UNOMOL_E_CUDA;
}

int unomol_b200_scf_iterate_uhf(unomol_b200_t *h, int nocc_a, int nocc_b, int damp, double *e_elec, double *pdiff) {
    if (!h || !e_elec || !pdiff || nocc_a < 0 || nocc_b < 0 || nocc_a > h->basis.nbf || nocc_b > h->basis.nbf) return UNOMOL_E_ARG;
    if (!h->d_scfH || !h->d_scfPoldB) return UNOMOL_E_STATE;
    cudaSetDevice(h->device);
    double *dP[2], *dG[2];
    int rc = unomol_b200_device_buffers(h, nullptr, dP, dG);
    if (rc) return rc;
    const int n = h->basis.nbf;
    const size_t nn = (size_t)n * n, no2 = (size_t)n * (n + 1) / 2;
    const unsigned blocks = (unsigned)((no2 + 255) / 256);
    cublasHandle_t cb = (cublasHandle_t)h->cublas;
    double *Pold[2] = {h->d_scfPold, h->d_scfPoldB};
    double *ev[2] = {h->d_evals, h->d_evalsB};
    const int nocc[2] = {nocc_a, nocc_b};
    if (damp) {
        scf_damp_kernel<<<blocks, 256, 0, h->stream>>>(dP[0], Pold[0], no2);
        scf_damp_kernel<<<blocks, 256, 0, h->stream>>>(dP[1], Pold[1], no2);
    }
    rc = unomol_b200_fock_uhf_device(h, dP[0], dP[1], dG[0], dG[1], /*async=*/1);
    if (rc) return rc;
    cudaMemsetAsync(h->d_scfRed4, 0, sizeof(double) * 4, h->stream);
    scf_energy_uhf_kernel<<<blocks, 256, 0, h->stream>>>(dP[0], dP[1], h->d_scfH, dG[0], dG[1], n, h->d_scfRed4);
    const double one = 1.0, zero = 0.0;
    for (int sp = 0; sp < 2; ++sp) {
        scf_fock_unpack_kernel<<<blocks, 256, 0, h->stream>>>(dP[sp], h->d_scfH, dG[sp], Pold[sp], h->d_F, n);
        cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, h->d_F, n, h->d_X, n, &zero, h->d_T, n);
        cublasDgemm(cb, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, h->d_X, n, h->d_T, n, &zero, h->d_W, n);
        cudaMemsetAsync(h->d_info, 0, sizeof(int) * 2, h->stream);
        if (cusolverDnDsyevd((cusolverDnHandle_t)h->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, h->d_W, n, ev[sp],
                             h->d_work, h->lwork, h->d_info) != CUSOLVER_STATUS_SUCCESS)
            return UNOMOL_E_CUDA;
        int info = 0;
        cudaMemcpyAsync(&info, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, h->d_X, n, h->d_W, n, &zero, h->d_T, n);   // C = X W
        if (nocc[sp] > 0)
            cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, nocc[sp], &one, h->d_T, n, h->d_T, n, &zero, h->d_F, n);
        else
            cudaMemsetAsync(h->d_F, 0, sizeof(double) * nn, h->stream);
        scf_pack_diff_kernel<<<blocks, 256, 0, h->stream>>>(h->d_F, Pold[sp], dP[sp], n, h->d_scfRed4 + 2 * sp);
        if (cudaStreamSynchronize(h->stream) != cudaSuccess || info != 0) return UNOMOL_E_CUDA;
    }
    double red[4] = {0.0, 0.0, 0.0, 0.0};
    cudaMemcpyAsync(red, h->d_scfRed4, sizeof(double) * 4, cudaMemcpyDeviceToHost, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return UNOMOL_E_CUDA;
    *e_elec = red[0];
    *pdiff = sqrt(red[1]) / n + sqrt(red[3]) / n;      // SymmPackDiffNorm per spin, summed (UHF.hpp:124,132)
    return UNOMOL_OK;
}

int unomol_b200_scf_fetch_uhf(unomol_b200_t *h, double *PA, double *PB, double *evals_a, double *evals_b) {
    if (!h) return UNOMOL_E_ARG;
    if (!h->d_scfH || !h->d_scfPoldB) return UNOMOL_E_STATE;
    cudaSetDevice(h->device);
    double *dP[2], *dG[2];
    int rc = unomol_b200_device_buffers(h, nullptr, dP, dG);
    if (rc) return rc;
    const size_t n = h->basis.nbf, no2 = n * (n + 1) / 2;
    if (PA) cudaMemcpyAsync(PA, dP[0], sizeof(double) * no2, cudaMemcpyDeviceToHost, h->stream);
    if (PB) cudaMemcpyAsync(PB, dP[1], sizeof(double) * no2, cudaMemcpyDeviceToHost, h->stream);
    if (evals_a) cudaMemcpyAsync(evals_a, h->d_evals, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
    if (evals_b) cudaMemcpyAsync(evals_b, h->d_evalsB, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
    return cudaStreamSynchronize(h->stream) == cudaSuccess ? UNOMOL_OK : UNOMOL_E_CUDA;
}

int unomol_b200_scf_fetch(unomol_b200_t *h, double *P, double *evals, double *C) {
    if (!h) return UNOMOL_E_ARG;
    if (!h->d_scfH) return UNOMOL_E_STATE;
    cudaSetDevice(h->device);
    double *dP[2], *dG[2];
    int rc = unomol_b200_device_buffers(h, nullptr, dP, dG);
    if (rc) return rc;
    const int n = h->basis.nbf;
    const size_t nn = (size_t)n * n, no2 = (size_t)n * (n + 1) / 2;
    if (P) cudaMemcpyAsync(P, dP[0], sizeof(double) * no2, cudaMemcpyDeviceToHost, h->stream);
    if (evals) cudaMemcpyAsync(evals, h->d_evals, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
    std::vector<double> ccm;
    if (C) {
        ccm.resize(nn);
        cudaMemcpyAsync(ccm.data(), h->d_T, sizeof(double) * nn, cudaMemcpyDeviceToHost, h->stream);
    }
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return UNOMOL_E_CUDA;
    if (C)
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < n; ++k) C[(size_t)i * n + k] = ccm[(size_t)k * n + i];   // row-major, eigenvectors in columns
    return UNOMOL_OK;
}

}  // extern "C"
