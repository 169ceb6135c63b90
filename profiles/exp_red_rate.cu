// profiles/exp_red_rate.cu -- microbenchmark: FP64 red.global.add throughput into an L2-resident N x N matrix with the
// access pattern of the exchange digestion (a CTA owns two rows; lanes hit scattered columns), against plain scattered
// 8-byte loads.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_red_rate exp_red_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// mode 0: reds, scattered columns of 2 CTA-owned rows; 1: reds fully random; 2: loads scattered; 3: reds, same column per warp (conflicts)
// 4: reds, consecutive columns per warp (coalesced sectors)
template <int MODE>
__global__ void k(double *M, int n, int iters, double *sink) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t row0 = (size_t)(hash(blockIdx.x) % n) * n, row1 = (size_t)(hash(blockIdx.x + 7919u) % n) * n;
    double acc = 0.0;
    for (int i = 0; i < iters; ++i) {
        const unsigned r = hash(tid * 2654435761u + i);
        unsigned col = r % n;
        if (MODE == 3) col = hash((tid >> 5) * 31u + i) % n;
        if (MODE == 4) col = (hash((tid >> 5) * 31u + i) % (n - 32)) + (threadIdx.x & 31);
        if (MODE == 0 || MODE == 3 || MODE == 4) {
            atomicAdd(M + row0 + col, 1.0);
            atomicAdd(M + row1 + (col * 7u) % n, 1.0);
        } else if (MODE == 1) {
            atomicAdd(M + (size_t)(r % n) * n + (hash(r) % n), 1.0);
            atomicAdd(M + (size_t)(hash(r + 1) % n) * n + (hash(r + 2) % n), 1.0);
        } else {
            acc += M[row0 + col] + M[row1 + (col * 7u) % n];
        }
    }
    if (acc == 12345.678) sink[0] = acc;
}

template <int MODE>
void run(const char *name, double *M, int n, double *sink) {
    const int grid = 148 * 16, block = 128, iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, block>>>(M, n, 100, sink);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(M, n, iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.2f ms  %.3e ops/s\n", name, ms, 2.0 * grid * block * (double)iters / (ms * 1e-3));
}

int main() {
    const int n = 2002;
    double *M, *sink;
    cudaMalloc(&M, sizeof(double) * n * n); cudaMemset(M, 0, sizeof(double) * n * n); cudaMalloc(&sink, 8);
    run<0>("red f64, 2 CTA rows, scattered columns", M, n, sink);
    run<1>("red f64, fully random", M, n, sink);
    run<2>("ld f64, 2 CTA rows, scattered columns", M, n, sink);
    run<3>("red f64, same column across the warp", M, n, sink);
    run<4>("red f64, consecutive columns across warp", M, n, sink);
    return 0;
}
