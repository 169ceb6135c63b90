// unomol_b200/csrc/pair_device.cu -- shell-pair / primitive-pair tables built ON THE DEVICE.
//
// Replaces the per-quartet recomputation of p, P, P-A and exp(-ab|AB|^2/p) in the reference's primitive loops
// (TwoElectronInts.cpp:439-460) by a once-per-geometry precompute (north star subsystem 1).  Same arithmetic and the
// same exact prune as the host path in engine.cu (build_pairs): a primitive pair whose SR*u*umax/sqrt(p) is below the
// reference's primitive cut can never pass `sr < 1e-12` against any partner (TwoElectronInts.cpp:478-479) and is
// dropped; a whole shell pair is rejected before any exp() when even its most diffuse primitive pair fails.
//   pass 1  pair_count_kernel : one thread per shell pair (i >= j): surviving primitive pairs -> count
//   scan    cub::DeviceScan   : primitive offsets and compact pair slots
//   pass 2  pair_fill_kernel  : writes the PrimPair records (sorted by u descending, insertion sort, <= 36..100
//                               records) and the ShellPair record of every kept pair
// The host only groups the kept pairs into (class, bucket, block) lists afterwards (engine.cu).
#include <cub/device/device_scan.cuh>
#include <vector>
#include "engine.h"

namespace ub200 {

struct DevBasis {
    const int *npr, *lv, *cen, *off, *poff;
    const double *alpha, *coef, *xyz, *amin;
    int ns;
    double umax, prim_cut;
};

__device__ __forceinline__ void pair_decode(long long pid, int &i, int &j) {
    i = (int)floor((sqrt(8.0 * (double)pid + 1.0) - 1.0) * 0.5);
    while ((long long)i * (i + 1) / 2 > pid) --i;
    while ((long long)(i + 1) * (i + 2) / 2 <= pid) ++i;
    j = (int)(pid - (long long)i * (i + 1) / 2);
}

// MODE 0: count survivors.  MODE 1: write records.
template <int MODE>
__global__ void pair_kernel(DevBasis B, long long npairs, int *__restrict__ count, const int *__restrict__ prim_off,
                            const int *__restrict__ slot, PrimPair *__restrict__ prims, ShellPair *__restrict__ pairs,
                            int *__restrict__ pair_cls) {
    const long long pid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= npairs) return;
    if (MODE == 1 && count[pid] == 0) return;
    // the fill pass must write exactly the count of the counting pass (the prune test is re-evaluated in a different template
    // instantiation; should code generation ever flip a borderline primitive, extra survivors are dropped here and a
    // shortfall shows up as sp.nprim < count, which the host checks)
    const int nmax = (MODE == 1) ? count[pid] : 0x7fffffff;
    int i, j;
    pair_decode(pid, i, j);
    int a = i, b = j;   // first shell = higher l (reference swap, TwoElectronInts.cpp:563-580)
    if (B.lv[i] < B.lv[j]) { a = j; b = i; }
    const double *A = B.xyz + 3 * B.cen[a], *Bc = B.xyz + 3 * B.cen[b];
    const double abx = A[0] - Bc[0], aby = A[1] - Bc[1], abz = A[2] - Bc[2];
    const double ab2 = abx * abx + aby * aby + abz * abz;
    if (MODE == 0) {
        const double pm = B.amin[a] + B.amin[b], mu = B.amin[a] * B.amin[b] / pm;
        if (SR_TERM * exp(-mu * ab2) / pm * B.umax / sqrt(pm) * 1.0000001 < B.prim_cut) { count[pid] = 0; return; }
    }
    const bool same = (a == b);   // the reference's pointer test al1==al2 (TwoElectronInts.cpp:444)
    const int na = B.npr[a], nb = B.npr[b];
    const double *ala = B.alpha + B.poff[a], *alb = B.alpha + B.poff[b];
    const double *coa = B.coef + B.poff[a], *cob = B.coef + B.poff[b];
    PrimPair *out = (MODE == 1) ? prims + prim_off[pid] : nullptr;
    int n = 0;
    double pmin = 1e300;
    for (int ia = 0; ia < na; ++ia) {
        const double axp = ala[ia];
        const int jend = same ? ia + 1 : nb;
        for (int ib = 0; ib < jend; ++ib) {
            const double bxp = alb[ib];
            const double p = axp + bxp, ip = 1.0 / p;
            const double u = exp(-axp * bxp * ab2 * ip) * ip;
            if (SR_TERM * u * B.umax / sqrt(p) * 1.0000001 < B.prim_cut) continue;
            if (n >= nmax) continue;
            if (MODE == 1) {
                PrimPair pp;
                pp.u = u; pp.p = p; pp.ip = ip;
                pp.c = coa[ia] * cob[ib] * ((same && ia != ib) ? 2.0 : 1.0);
                pp.P[0] = (axp * A[0] + bxp * Bc[0]) * ip; pp.P[1] = (axp * A[1] + bxp * Bc[1]) * ip; pp.P[2] = (axp * A[2] + bxp * Bc[2]) * ip;
                pp.PA[0] = pp.P[0] - A[0]; pp.PA[1] = pp.P[1] - A[1]; pp.PA[2] = pp.P[2] - A[2];
                // insertion into the list kept sorted by u, descending (stable: equal u keep generation order)
                int k = n;
                while (k > 0 && out[k - 1].u < u) { out[k] = out[k - 1]; --k; }
                out[k] = pp;
                pmin = fmin(pmin, p);
            }
            ++n;
        }
    }
    if (MODE == 0) { count[pid] = n; return; }
    ShellPair sp;
    sp.pmin = pmin; sp.umax = out[0].u; sp.spare = 0.0;
    sp.AB[0] = abx; sp.AB[1] = aby; sp.AB[2] = abz;
    sp.Q = 0.0;
    sp.offa = B.off[a]; sp.offb = B.off[b];
    sp.prim_off = prim_off[pid]; sp.nprim = n;
    sp.sha = a; sp.shb = b;
    sp.pairid = (int)pid;
    sp.pad = 0;
    pairs[slot[pid]] = sp;
    pair_cls[slot[pid]] = B.lv[a] * (B.lv[a] + 1) / 2 + B.lv[b];
}

__global__ void flag_kernel(const int *__restrict__ count, int *__restrict__ flag, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = count[i] > 0 ? 1 : 0;
}

#define PD_TRY(expr)                                  \
    do {                                              \
        cudaError_t _e = (expr);                      \
        if (_e != cudaSuccess) { rc = UNOMOL_E_CUDA; goto done; } \
    } while (0)

// Builds the tables on h->stream.  kept / cls: the kept shell pairs in canonical (i,j) order with prim_off, nprim,
// pmin, umax filled; *d_prims_out: device array of all primitive pairs (ownership passes to the caller).
int build_pair_tables_device(unomol_b200 *h, double prune_cut, std::vector<ShellPair> &kept, std::vector<int> &cls, PrimPair **d_prims_out,
                             long long *nprim_out) {
    const HostBasis &HB = h->basis;
    const int ns = HB.nshell;
    const long long np = (long long)ns * (ns + 1) / 2;
    cudaStream_t st = h->stream;
    int rc = UNOMOL_OK;
    // offsets, slots and the scan length are 32-bit: refuse instead of overflowing (2^31 shell pairs = 65 000 shells)
    if (np > 0x7fffffffLL) return UNOMOL_E_UNSUPPORTED;
    std::vector<double> amin(ns);
    double umax = 0.0;
    for (int s = 0; s < ns; ++s) {
        amin[s] = HB.alpha[HB.poff[s]];
        for (int k = 1; k < HB.npr[s]; ++k) amin[s] = std::min(amin[s], HB.alpha[HB.poff[s] + k]);
        umax = std::max(umax, 0.5 / amin[s]);
    }
    int *d_i = nullptr, *d_count = nullptr, *d_off = nullptr, *d_flag = nullptr, *d_slot = nullptr, *d_cls = nullptr;
    double *d_d = nullptr;
    void *d_tmp = nullptr;
    PrimPair *d_prims = nullptr;
    ShellPair *d_pairs = nullptr;
    size_t tmp_bytes = 0, tb2 = 0;
    const size_t nprim_basis = HB.alpha.size();
    int last_count = 0, last_off = 0, last_flag = 0, last_slot = 0;
    long long nprim = 0, nkept = 0;
    const unsigned blocks = (unsigned)((np + 255) / 256);
    DevBasis B;
    // basis arrays: 5 int arrays of ns, then alpha, coef, xyz, amin
    PD_TRY(cudaMalloc(&d_i, sizeof(int) * 5 * ns));
    PD_TRY(cudaMalloc(&d_d, sizeof(double) * (2 * nprim_basis + 3 * HB.ncen + ns)));
    PD_TRY(cudaMemcpyAsync(d_i, HB.npr.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, st));
    PD_TRY(cudaMemcpyAsync(d_i + ns, HB.lv.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, st));
    PD_TRY(cudaMemcpyAsync(d_i + 2 * ns, HB.cen.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, st));
    PD_TRY(cudaMemcpyAsync(d_i + 3 * ns, HB.off.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, st));
    PD_TRY(cudaMemcpyAsync(d_i + 4 * ns, HB.poff.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, st));
    PD_TRY(cudaMemcpyAsync(d_d, HB.alpha.data(), sizeof(double) * nprim_basis, cudaMemcpyHostToDevice, st));
    PD_TRY(cudaMemcpyAsync(d_d + nprim_basis, HB.coef.data(), sizeof(double) * nprim_basis, cudaMemcpyHostToDevice, st));
    PD_TRY(cudaMemcpyAsync(d_d + 2 * nprim_basis, HB.xyz.data(), sizeof(double) * 3 * HB.ncen, cudaMemcpyHostToDevice, st));
    PD_TRY(cudaMemcpyAsync(d_d + 2 * nprim_basis + 3 * HB.ncen, amin.data(), sizeof(double) * ns, cudaMemcpyHostToDevice, st));
    B.npr = d_i; B.lv = d_i + ns; B.cen = d_i + 2 * ns; B.off = d_i + 3 * ns; B.poff = d_i + 4 * ns;
    B.alpha = d_d; B.coef = d_d + nprim_basis; B.xyz = d_d + 2 * nprim_basis; B.amin = d_d + 2 * nprim_basis + 3 * HB.ncen;
    B.ns = ns; B.umax = umax; B.prim_cut = prune_cut;   // 0 with f/g shells: the reference's MD path has no primitive cut
    PD_TRY(cudaMalloc(&d_count, sizeof(int) * np));
    PD_TRY(cudaMalloc(&d_off, sizeof(int) * np));
    PD_TRY(cudaMalloc(&d_flag, sizeof(int) * np));
    PD_TRY(cudaMalloc(&d_slot, sizeof(int) * np));
    pair_kernel<0><<<blocks, 256, 0, st>>>(B, np, d_count, nullptr, nullptr, nullptr, nullptr, nullptr);
    flag_kernel<<<blocks, 256, 0, st>>>(d_count, d_flag, np);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_count, d_off, (int)np, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, d_flag, d_slot, (int)np, st);
    tmp_bytes = std::max(tmp_bytes, tb2);
    PD_TRY(cudaMalloc(&d_tmp, tmp_bytes));
    PD_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_count, d_off, (int)np, st));
    PD_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_flag, d_slot, (int)np, st));
    PD_TRY(cudaMemcpyAsync(&last_count, d_count + np - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    PD_TRY(cudaMemcpyAsync(&last_off, d_off + np - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    PD_TRY(cudaMemcpyAsync(&last_flag, d_flag + np - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    PD_TRY(cudaMemcpyAsync(&last_slot, d_slot + np - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    PD_TRY(cudaStreamSynchronize(st));
    nprim = (long long)last_off + last_count;
    if (last_off < 0 || nprim > 0x7fffffffLL) { rc = UNOMOL_E_UNSUPPORTED; goto done; }   // exclusive scan wrapped
    nkept = (long long)last_slot + last_flag;
    if (nprim > 0) {
        PD_TRY(cudaMalloc(&d_prims, sizeof(PrimPair) * nprim));
        PD_TRY(cudaMalloc(&d_pairs, sizeof(ShellPair) * nkept));
        PD_TRY(cudaMalloc(&d_cls, sizeof(int) * nkept));
        pair_kernel<1><<<blocks, 256, 0, st>>>(B, np, d_count, d_off, d_slot, d_prims, d_pairs, d_cls);
        kept.resize(nkept);
        cls.resize(nkept);
        PD_TRY(cudaMemcpyAsync(kept.data(), d_pairs, sizeof(ShellPair) * nkept, cudaMemcpyDeviceToHost, st));
        PD_TRY(cudaMemcpyAsync(cls.data(), d_cls, sizeof(int) * nkept, cudaMemcpyDeviceToHost, st));
        PD_TRY(cudaStreamSynchronize(st));
    } else {
        kept.clear();
        cls.clear();
    }
    *d_prims_out = d_prims;
    d_prims = nullptr;
    *nprim_out = nprim;
done:
    cudaFree(d_i); cudaFree(d_d); cudaFree(d_count); cudaFree(d_off); cudaFree(d_flag); cudaFree(d_slot);
    cudaFree(d_tmp); cudaFree(d_pairs); cudaFree(d_cls);
    if (d_prims) cudaFree(d_prims);
    return rc;
}

}  // namespace ub200
