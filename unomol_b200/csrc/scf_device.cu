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

int scf_ensure(unomol_b200 *h) {
    if (h->cusolver) return UNOMOL_OK;
    const int n = h->basis.nbf;
    cusolverDnHandle_t cs;
    cublasHandle_t cb;
    if (cusolverDnCreate(&cs) != CUSOLVER_STATUS_SUCCESS) return UNOMOL_E_CUDA;
    if (cublasCreate(&cb) != CUBLAS_STATUS_SUCCESS) return UNOMOL_E_CUDA;
    cusolverDnSetStream(cs, h->stream);
    cublasSetStream(cb, h->stream);
    h->cusolver = cs;
    h->cublas = cb;
    const size_t nn = (size_t)n * n;
    if (cudaMalloc(&h->d_X, sizeof(double) * nn) != cudaSuccess) return UNOMOL_E_NOMEM;
    if (cudaMalloc(&h->d_F, sizeof(double) * nn) != cudaSuccess) return UNOMOL_E_NOMEM;
    if (cudaMalloc(&h->d_W, sizeof(double) * nn) != cudaSuccess) return UNOMOL_E_NOMEM;
    if (cudaMalloc(&h->d_T, sizeof(double) * nn) != cudaSuccess) return UNOMOL_E_NOMEM;
    if (cudaMalloc(&h->d_evals, sizeof(double) * n) != cudaSuccess) return UNOMOL_E_NOMEM;
    if (cudaMalloc(&h->d_info, sizeof(int) * 2) != cudaSuccess) return UNOMOL_E_NOMEM;
    int lwork = 0;
    if (cusolverDnDsyevd_bufferSize(cs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, h->d_W, n, h->d_evals,
                                    &lwork) != CUSOLVER_STATUS_SUCCESS)
        return UNOMOL_E_CUDA;
    h->lwork = lwork;
    if (cudaMalloc(&h->d_work, sizeof(double) * (lwork > 0 ? lwork : 1)) != cudaSuccess) return UNOMOL_E_NOMEM;
    return UNOMOL_OK;
}

}  // namespace

void unomol_scf_free(unomol_b200 *h) {
    if (h->cusolver) cusolverDnDestroy((cusolverDnHandle_t)h->cusolver);
    if (h->cublas) cublasDestroy((cublasHandle_t)h->cublas);
    h->cusolver = h->cublas = nullptr;
    cudaFree(h->d_X); cudaFree(h->d_F); cudaFree(h->d_W); cudaFree(h->d_T);
    cudaFree(h->d_evals); cudaFree(h->d_work); cudaFree(h->d_info);
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

}  // extern "C"
