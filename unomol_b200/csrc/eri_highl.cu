// unomol_b200/csrc/eri_highl.cu -- launcher of the runtime-L kernel for quartets with f/g shells (see eri_highl.cuh).
#include <cuda_runtime.h>
#include "eri_highl.cuh"
#include <atomic>
#include "engine.h"

namespace ub200 {

template <int MODE>
__global__ void __launch_bounds__(HL_THREADS) eri_highl_kernel(const ClassTask task, const HighLArgs hl) {
    extern __shared__ __align__(16) unsigned char hl_smem_raw[];
    double *sm = reinterpret_cast<double *>(hl_smem_raw);
    __shared__ long long s_q;
    hl_init_tables(hl, sm);
    double *V = hl.scratch + (size_t)blockIdx.x * hl.slab;
    const int NAB = hl_ncart(hl.la) * hl_ncart(hl.lb), NCD = hl_ncart(hl.lc) * hl_ncart(hl.ld), NINT = NAB * NCD;
    double *red = sm + HL_OFF_RED;
    unsigned long long n_quart = 0, n_primq = 0;
    long long seq = blockIdx.x;
    bool static_done = false;      // thread 0's state of claim_block
    for (int outer = blockIdx.x;; outer += gridDim.x) {
        int bi = 0, kfirst = 0, kcount = 0;
        if (MODE == MODE_DIGEST) {
            // quartets one at a time: from the launch's work counter (dynamic, heaviest bras first; shared across ranks
            // when the counter is IPC-mapped peer memory), else static snake order over ranks -- as eri_generic.cuh,
            // but per quartet: a block can be 50 625 integrals and a launch may have fewer bras than the GPU has SMs
            const long long total = task.ket_prefix[task.nbra];
            long long q;
            if (task.work_counter) {
                __syncthreads();
                if (threadIdx.x == 0) s_q = claim_block(task, static_done);
                __syncthreads();
                q = s_q;
            } else {
                q = (long long)task.nranks * seq + ((seq & 1) ? task.nranks - 1 - task.rank : task.rank);
                seq += gridDim.x;
            }
            if (q >= total) break;
            int lo = 0, hi = task.nbra;      // largest bi with ket_prefix[bi] <= q
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (task.ket_prefix[mid] <= q) lo = mid; else hi = mid;
            }
            bi = lo;
            kfirst = (int)(q - task.ket_prefix[bi]);
            kcount = kfirst + 1;
        } else {
            if (outer >= task.ntask) break;
            bi = task.task_list[outer].x;
            kfirst = task.task_list[outer].y;
            kcount = kfirst + 1;
        }
        const ShellPair bra = task.bra[bi];
        const bool one12 = bra.AB[0] == 0.0 && bra.AB[1] == 0.0 && bra.AB[2] == 0.0;
        for (int ki = kfirst; ki < kcount; ++ki) {
            const ShellPair ket = task.ket[ki];
            if (MODE == MODE_DIGEST) {
                const int imax = max(max(bra.sha, bra.shb), max(ket.sha, ket.shb));
                if (imax < task.start_shell) continue;
            }
            const bool one34 = ket.AB[0] == 0.0 && ket.AB[1] == 0.0 && ket.AB[2] == 0.0;
            n_primq += hl_quartet_block(hl, bra, ket, task.prims, task.prim_cut, one12, one34, sm, V);
            ++n_quart;
            if (MODE == MODE_DUMP) {
                double *dst = task.out + task.task_out[outer];
                for (int o = threadIdx.x; o < NINT; o += blockDim.x) dst[o] = V[o];
            } else if (MODE == MODE_SCHWARZ) {
                double mx = 0.0;
                for (int ab = threadIdx.x; ab < NAB; ab += blockDim.x) mx = fmax(mx, fabs(V[(size_t)ab * NCD + ab]));
                mx = hl_block_max(mx, red);
                if (threadIdx.x == 0) task.out[outer] = sqrt(mx);
            } else {
                double sym = 1.0;
                if (bra.sha == bra.shb) sym *= 0.5;
                if (ket.sha == ket.shb) sym *= 0.5;
                if (task.same_class && bra.pairid == ket.pairid) sym *= 0.5;
                // blocks entirely below the reference's storage threshold |val| <= 1e-14 never reach its G
                double mx = 0.0;
                for (int o = threadIdx.x; o < NINT; o += blockDim.x) mx = fmax(mx, fabs(V[o]));
                mx = hl_block_max(mx, red);
                if (mx > task.value_cut) hl_digest(hl, task, bra, ket, V, sym);
            }
            __syncthreads();   // V is rewritten by the next quartet
        }
    }
    if (MODE == MODE_DIGEST && task.counters && threadIdx.x == 0) {
        atomicAdd(task.counters, n_quart);
        atomicAdd(task.counters + 1, n_primq);
    }
}

cudaError_t launch_highl(const ClassTask &task, const HighLArgs &hl, int mode, int grid, cudaStream_t stream) {
    static std::atomic<bool> attr_done_dev[64];   // per device: one process may drive several GPUs
    int attr_dev = 0;
    cudaGetDevice(&attr_dev);
    std::atomic<bool> &attr_done = attr_done_dev[attr_dev & 63];
    if (!attr_done) {
        cudaFuncSetAttribute(eri_highl_kernel<MODE_DIGEST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HL_SMEM_BYTES);
        cudaFuncSetAttribute(eri_highl_kernel<MODE_DUMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HL_SMEM_BYTES);
        cudaFuncSetAttribute(eri_highl_kernel<MODE_SCHWARZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HL_SMEM_BYTES);
        attr_done = true;
    }
    if (grid <= 0) return cudaSuccess;
    if (mode == MODE_DIGEST) eri_highl_kernel<MODE_DIGEST><<<grid, HL_THREADS, HL_SMEM_BYTES, stream>>>(task, hl);
    else if (mode == MODE_DUMP) eri_highl_kernel<MODE_DUMP><<<grid, HL_THREADS, HL_SMEM_BYTES, stream>>>(task, hl);
    else eri_highl_kernel<MODE_SCHWARZ><<<grid, HL_THREADS, HL_SMEM_BYTES, stream>>>(task, hl);
    return cudaGetLastError();
}

}  // namespace ub200
