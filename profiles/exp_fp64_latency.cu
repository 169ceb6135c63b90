// profiles/exp_fp64_latency.cu -- dependent-issue latency and per-SMSP throughput of DFMA on sm_100a.
//   chains = independent DFMA chains per thread, warps = warps per CTA (one CTA per SM): cycles per DFMA warp-instruction
//   per SMSP.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_fp64_latency exp_fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double *out, long long *cyc, int iters, double a, double b) {
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x + c;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
void run(int warps) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * 148 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<CH><<<148, warps * 32>>>(out, cyc, iters, 0.999999, 1e-9);
    k<CH><<<148, warps * 32>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // warps per SMSP = warps/4 (at least 1); DFMA warp-instructions per SMSP = iters*CH*max(warps/4,1)
    double per_smsp = (double)iters * CH * (warps >= 4 ? warps / 4.0 : 1.0);
    printf("chains %d warps/CTA %2d: %.2f cycles per dependent step, %.2f cycles per DFMA warp-inst per SMSP\n", CH, warps,
           (double)h / iters, (double)h / per_smsp);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
