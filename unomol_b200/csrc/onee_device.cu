// unomol_b200/csrc/onee_device.cu -- overlap, kinetic, core-Hamiltonian and dipole / quadrupole moment matrices ON THE DEVICE
// (SURVEY.md 8(f)4).  Replaces the reference's OneElectronInts (reference OneElectronInts.cpp:6-202) and MomentInts
// (Moments.cpp:5-187) for callers of the C ABI; the threaded host versions (host/OneElectron.hpp, host/Moments.hpp) remain
// as the cross-check.  O(N^2 ncen) nuclear attraction was ~2 s on 16 host threads at 2002 functions = five Fock builds
// per geometry; here it is one kernel, one thread per shell pair.
//
// Arithmetic (independent of the reference's McMurchie-Davidson tables; same values to rounding):
//   * overlap, per axis, Obara-Saika: s(0,0) = 1, s(i+1,j) = PA s(i,j) + (i s(i-1,j) + j s(i,j-1)) / 2p, same for j with PB;
//     3-D prefactor (pi/p)^(3/2) exp(-ab|AB|^2/p);
//   * kinetic energy on the second function: k(i,j) = -2 b^2 s(i,j+2) + b (2j+1) s(i,j) - j (j-1)/2 s(i,j-2);
//   * moments <x^k>, k <= 2, about the origin from the same table: <x> = s(i,j+1) + B_x s(i,j),
//     <x^2> = s(i,j+2) + 2 B_x s(i,j+1) + B_x^2 s(i,j);
//   * nuclear attraction by Rys quadrature, n = (la+lb)/2 + 1 nodes of rys_roots.cuh:
//     (a|1/r_C|b) = (2 pi / p) K_ab sum_nodes w prod_axes g(la_x, lb_x),  g(0,0) = 1,
//     g(i+1,0) = (PA - t^2 PC) g(i,0) + i (1 - t^2)/(2p) g(i-1,0),  g(i,j+1) = g(i+1,j) + (A-B) g(i,j).
//     The reference's Boys function is the bare asymptotic series for t > 20 (MD_Rfunction.hpp:2186-2193, relative error up
//     to 7e-11 just above 20).  The Gauss-Hermite limit of the Rys quadrature integrates exactly that series, so it is used
//     from X = 20 on: H agrees with the reference's to rounding, which the 1e-9 Eh energy parity rests on.
#include <cuda_runtime.h>
#include <vector>
#include "engine.h"
#include "rys_roots.cuh"

namespace ub200 {

struct OneeArgs {
    const int *npr, *lv, *cen, *off, *poff;
    const double *alpha, *coef, *xyz, *charge;
    int ns, ncen, nbf;
    RysTables rys;
    double *S, *T, *H;      // packed lower triangle
    double *M;              // 9 packed moment matrices (dx dy dz qxx qxy qxz qyy qyz qzz) or null
    int dpm_cen;            // >= 0: centre that carries the reference's positron charge model (GDPMInts), else -1
};

// The dipole-polarisation model of the reference's polarisation scan (reference GDPMInts.cpp:5-83): the positron is a fixed
// contraction of six s Gaussians at the scan centre, and H -= sum_i c_i (g_i | a b), a three-centre Coulomb integral.  In Rys
// form it is the nuclear-attraction recursion with t^2 -> t^2 p/(p+q) (p = the model exponent): a smeared point charge.
__constant__ double dpm_alf[6] = {6.8505018000, 4.0491646000, 3.5941062989, 1.2478274000, 0.7927690989, 0.3377107978};
__constant__ double dpm_cof[6] = {0.0766926784, 0.1483475912, 0.0462334641, 0.0717376427, 0.0447149792, 0.0069678529};

__device__ __forceinline__ void onee_comp(int l, int c, int *lmn) {
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= c) ++i;     // row i = l - lx holds i + 1 entries
    const int j = c - i * (i + 1) / 2;
    lmn[0] = l - i; lmn[1] = i - j; lmn[2] = j;
}
__device__ __forceinline__ double onee_cnorm(const int *lmn) {
    // per-component norm 1/sqrt(df[lx] df[ly] df[lz]) with the reference's OWN recurrence (AuxFunctions.hpp:49-64:
    // df[i] = df[i-1] * dx, dx *= 2i+1), which is (2l-1)!! only up to l = 2: df = 1, 1, 3, 45, 4725
    const double df[5] = {1.0, 1.0, 3.0, 45.0, 4725.0};
    return rsqrt(df[lmn[0]] * df[lmn[1]] * df[lmn[2]]);
}

template <int N>
__device__ __forceinline__ void onee_nodes(double X, double *t2, double *w, const RysTables &R) {
    if (X > 20.0) rys_hermite_limit_t2<N>(X, t2, w);   // = the reference's asymptotic Boys branch (see the header)
    else {
        RysTables Rx = R;
        Rx.rys2_exact = 1;                             // the two-root parity band belongs to the ERI path only
        rys_t2<N>(X, t2, w, Rx);
    }
}

// LM = largest shell angular momentum the instantiation handles (1: s/p, 2: d, 4: g)
template <int LM>
__global__ void __launch_bounds__(128) onee_kernel(const OneeArgs a) {
    constexpr int NC = (LM + 1) * (LM + 2) / 2, NJ = LM + 3, NI = LM + 1, NG = 2 * LM + 1;
    const long long pid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npairs = (long long)a.ns * (a.ns + 1) / 2;
    if (pid >= npairs) return;
    int ish = (int)floor((sqrt(8.0 * (double)pid + 1.0) - 1.0) * 0.5);
    while ((long long)ish * (ish + 1) / 2 > pid) --ish;
    while ((long long)(ish + 1) * (ish + 2) / 2 <= pid) ++ish;
    const int jsh = (int)(pid - (long long)ish * (ish + 1) / 2);
    const int la = a.lv[ish], lb = a.lv[jsh], na = (la + 1) * (la + 2) / 2, nb = (lb + 1) * (lb + 2) / 2, L = la + lb;
    const double *ra = a.xyz + 3 * a.cen[ish], *rb = a.xyz + 3 * a.cen[jsh];
    const double abv[3] = {ra[0] - rb[0], ra[1] - rb[1], ra[2] - rb[2]};
    const double ab2 = abv[0] * abv[0] + abv[1] * abv[1] + abv[2] * abv[2];
    double sv[NC * NC], tv[NC * NC], vv[NC * NC], mv[9][NC * NC];
    for (int k = 0; k < na * nb; ++k) {
        sv[k] = tv[k] = vv[k] = 0.0;
        for (int m = 0; m < 9; ++m) mv[m][k] = 0.0;
    }
    const int nroots = L / 2 + 1;
    for (int ip = 0; ip < a.npr[ish]; ++ip)
        for (int jp = 0; jp < a.npr[jsh]; ++jp) {
            const double ea = a.alpha[a.poff[ish] + ip], eb = a.alpha[a.poff[jsh] + jp], p = ea + eb, ip2 = 0.5 / p;
            const double c12 = a.coef[a.poff[ish] + ip] * a.coef[a.poff[jsh] + jp];
            const double kab = exp(-ea * eb / p * ab2);
            if (fabs(c12) * kab < 1e-20) continue;     // as host/OneElectron.hpp: contributes < 1e-17 to any element
            double P[3], PA[3], PB[3];
            for (int x = 0; x < 3; ++x) {
                P[x] = (ea * ra[x] + eb * rb[x]) / p;
                PA[x] = P[x] - ra[x];
                PB[x] = P[x] - rb[x];
            }
            // ---- overlap tables s[x][i][j], i <= la, j <= lb + 2
            double s[3][NI][NJ];
            for (int x = 0; x < 3; ++x) {
                s[x][0][0] = 1.0;
                for (int j = 0; j < lb + 2; ++j)
                    s[x][0][j + 1] = PB[x] * s[x][0][j] + (j > 0 ? j * ip2 * s[x][0][j - 1] : 0.0);
                for (int i = 0; i < la; ++i)
                    for (int j = 0; j <= lb + 2; ++j)
                        s[x][i + 1][j] = PA[x] * s[x][i][j] + (i > 0 ? i * ip2 * s[x][i - 1][j] : 0.0) + (j > 0 ? j * ip2 * s[x][i][j - 1] : 0.0);
            }
            const double pi_p = 3.14159265358979323846 / p;
            const double s3 = pi_p * sqrt(pi_p) * kab;
            for (int ia = 0; ia < na; ++ia) {
                int l1[3];
                onee_comp(la, ia, l1);
                for (int ib = 0; ib < nb; ++ib) {
                    int l2[3];
                    onee_comp(lb, ib, l2);
                    const double nf = onee_cnorm(l1) * onee_cnorm(l2) * c12 * s3;
                    double m0[3], m1[3], m2[3], kk[3];
                    for (int x = 0; x < 3; ++x) {
                        const int i = l1[x], j = l2[x];
                        const double B = rb[x];
                        m0[x] = s[x][i][j];
                        m1[x] = s[x][i][j + 1] + B * s[x][i][j];
                        m2[x] = s[x][i][j + 2] + 2.0 * B * s[x][i][j + 1] + B * B * s[x][i][j];
                        double k = -2.0 * eb * eb * s[x][i][j + 2] + eb * (2 * j + 1) * s[x][i][j];
                        if (j >= 2) k -= 0.5 * j * (j - 1) * s[x][i][j - 2];
                        kk[x] = k;
                    }
                    const int o = ia * nb + ib;
                    sv[o] += nf * m0[0] * m0[1] * m0[2];
                    tv[o] += nf * (kk[0] * m0[1] * m0[2] + m0[0] * kk[1] * m0[2] + m0[0] * m0[1] * kk[2]);
                    if (a.M) {
                        mv[0][o] += nf * m1[0] * m0[1] * m0[2];
                        mv[1][o] += nf * m0[0] * m1[1] * m0[2];
                        mv[2][o] += nf * m0[0] * m0[1] * m1[2];
                        mv[3][o] += nf * m2[0] * m0[1] * m0[2];
                        mv[4][o] += nf * m1[0] * m1[1] * m0[2];
                        mv[5][o] += nf * m1[0] * m0[1] * m1[2];
                        mv[6][o] += nf * m0[0] * m2[1] * m0[2];
                        mv[7][o] += nf * m0[0] * m1[1] * m1[2];
                        mv[8][o] += nf * m0[0] * m0[1] * m2[2];
                    }
                }
            }
            // ---- nuclear attraction: sum over centres and Rys nodes
            const double vpref = 2.0 * pi_p * kab * c12;
            for (int ic = 0; ic < a.ncen; ++ic) {
                const double Z = a.charge[ic];
                if (Z == 0.0) continue;
                const double *rc = a.xyz + 3 * ic;
                const double pc[3] = {P[0] - rc[0], P[1] - rc[1], P[2] - rc[2]};
                const double X = p * (pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2]);
                double t2[5], w[5];
                switch (nroots) {
                    case 1: onee_nodes<1>(X, t2, w, a.rys); break;
                    case 2: onee_nodes<2>(X, t2, w, a.rys); break;
                    case 3: onee_nodes<3>(X, t2, w, a.rys); break;
                    case 4: onee_nodes<4>(X, t2, w, a.rys); break;
                    default: onee_nodes<5>(X, t2, w, a.rys); break;
                }
                for (int ir = 0; ir < nroots; ++ir) {
                    const double tt = t2[ir], b1 = (1.0 - tt) * ip2;
                    double g[3][NG][NI];      // g[x][i][j], i <= L - j, j <= lb
                    for (int x = 0; x < 3; ++x) {
                        const double c0 = PA[x] - tt * pc[x];
                        g[x][0][0] = 1.0;
                        for (int i = 0; i < L; ++i) g[x][i + 1][0] = c0 * g[x][i][0] + (i > 0 ? i * b1 * g[x][i - 1][0] : 0.0);
                        for (int j = 0; j < lb; ++j)
                            for (int i = 0; i <= L - j - 1; ++i) g[x][i][j + 1] = g[x][i + 1][j] + abv[x] * g[x][i][j];
                    }
                    const double wz = -Z * vpref * w[ir];
                    for (int ia = 0; ia < na; ++ia) {
                        int l1[3];
                        onee_comp(la, ia, l1);
                        for (int ib = 0; ib < nb; ++ib) {
                            int l2[3];
                            onee_comp(lb, ib, l2);
                            vv[ia * nb + ib] += wz * g[0][l1[0]][l2[0]] * g[1][l1[1]][l2[1]] * g[2][l1[2]][l2[2]];
                        }
                    }
                }
            }
            // ---- positron charge model (GDPMInts): roots in the ERI path's mode (the reference calls Rys::Recur here)
            if (a.dpm_cen >= 0) {
                const double *rp = a.xyz + 3 * a.dpm_cen;
                const double pq[3] = {rp[0] - P[0], rp[1] - P[1], rp[2] - P[2]};      // (model centre) - (product centre)
                const double pq2 = pq[0] * pq[0] + pq[1] * pq[1] + pq[2] * pq[2];
                for (int im = 0; im < 6; ++im) {
                    const double pe = dpm_alf[im], txp = pe + p;
                    const double X = pe * p / txp * pq2;
                    // sr = 2 pi^(5/2) s34 / (pe q sqrt(pe + q))   (GDPMInts.cpp:53), q = p here
                    const double sr = SR_TERM * kab / (pe * p * sqrt(txp)) * dpm_cof[im] * c12;
                    double t2[5], w[5];
                    switch (nroots) {
                        case 1: rys_t2<1>(X, t2, w, a.rys); break;
                        case 2: rys_t2<2>(X, t2, w, a.rys); break;
                        case 3: rys_t2<3>(X, t2, w, a.rys); break;
                        case 4: rys_t2<4>(X, t2, w, a.rys); break;
                        default: rys_t2<5>(X, t2, w, a.rys); break;
                    }
                    for (int ir = 0; ir < nroots; ++ir) {
                        const double te = t2[ir] * pe / txp, b1 = (1.0 - te) * ip2;      // B1p of Rys::Recur (Rys.hpp:131)
                        double g[3][NG][NI];
                        for (int x = 0; x < 3; ++x) {
                            const double c0 = PA[x] + te * pq[x];                         // Cp = qc + p pq t^2/(p+q) (Rys.hpp:133)
                            g[x][0][0] = 1.0;
                            for (int i = 0; i < L; ++i) g[x][i + 1][0] = c0 * g[x][i][0] + (i > 0 ? i * b1 * g[x][i - 1][0] : 0.0);
                            for (int j = 0; j < lb; ++j)
                                for (int i = 0; i <= L - j - 1; ++i) g[x][i][j + 1] = g[x][i + 1][j] + abv[x] * g[x][i][j];
                        }
                        const double wz = -sr * w[ir];                                     // Hmat -= (GDPMInts.cpp:139)
                        for (int ia = 0; ia < na; ++ia) {
                            int l1[3];
                            onee_comp(la, ia, l1);
                            for (int ib = 0; ib < nb; ++ib) {
                                int l2[3];
                                onee_comp(lb, ib, l2);
                                vv[ia * nb + ib] += wz * g[0][l1[0]][l2[0]] * g[1][l1[1]][l2[1]] * g[2][l1[2]][l2[2]];
                            }
                        }
                    }
                }
            }
        }
    const size_t no2 = (size_t)a.nbf * (a.nbf + 1) / 2;
    for (int ia = 0; ia < na; ++ia) {
        int l1[3];
        onee_comp(la, ia, l1);
        const int ir = a.off[ish] + ia;
        for (int ib = 0; ib < nb; ++ib) {
            const int jr = a.off[jsh] + ib;
            if (jr > ir) continue;
            int l2[3];
            onee_comp(lb, ib, l2);
            const double nv = onee_cnorm(l1) * onee_cnorm(l2);     // the Rys sum carries no component norms yet
            const size_t ij = (size_t)ir * (ir + 1) / 2 + jr;
            const int o = ia * nb + ib;
            a.S[ij] = sv[o];
            a.T[ij] = tv[o];
            a.H[ij] = tv[o] + nv * vv[o];
            if (a.M)
                for (int m = 0; m < 9; ++m) a.M[m * no2 + ij] = mv[m][o];
        }
    }
}

}  // namespace ub200

using namespace ub200;

extern "C" int unomol_b200_one_electron(unomol_b200_t *h, const double *charge, double *S, double *T, double *H, double *M) {
    return unomol_b200_one_electron_dpm(h, charge, -1, S, T, H, M);
}

extern "C" int unomol_b200_one_electron_dpm(unomol_b200_t *h, const double *charge, int dpm_center, double *S, double *T, double *H,
                                            double *M) {
    if (!h || !charge || !S || !T || !H || dpm_center >= h->basis.ncen) return UNOMOL_E_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return UNOMOL_E_CUDA;
    const HostBasis &B = h->basis;
    const int ns = B.nshell, n = B.nbf;
    const size_t no2 = (size_t)n * (n + 1) / 2, nprim = B.alpha.size();
    const long long npairs = (long long)ns * (ns + 1) / 2;
    int *d_i = nullptr;
    double *d_d = nullptr, *d_out = nullptr;
    const size_t nout = (M ? 12 : 3) * no2;
    int rc = UNOMOL_OK;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<double> host(nout);
    OneeArgs a{};
    if (cudaMalloc(&d_i, sizeof(int) * 5 * ns) != cudaSuccess || cudaMalloc(&d_d, sizeof(double) * (2 * nprim + 4 * B.ncen)) != cudaSuccess ||
        cudaMalloc(&d_out, sizeof(double) * nout) != cudaSuccess) {
        rc = UNOMOL_E_NOMEM;
        goto done;
    }
    cudaMemcpyAsync(d_i, B.npr.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_i + ns, B.lv.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_i + 2 * ns, B.cen.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_i + 3 * ns, B.off.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_i + 4 * ns, B.poff.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_d, B.alpha.data(), sizeof(double) * nprim, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_d + nprim, B.coef.data(), sizeof(double) * nprim, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_d + 2 * nprim, B.xyz.data(), sizeof(double) * 3 * B.ncen, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_d + 2 * nprim + 3 * B.ncen, charge, sizeof(double) * B.ncen, cudaMemcpyHostToDevice, h->stream);
    a.npr = d_i; a.lv = d_i + ns; a.cen = d_i + 2 * ns; a.off = d_i + 3 * ns; a.poff = d_i + 4 * ns;
    a.alpha = d_d; a.coef = d_d + nprim; a.xyz = d_d + 2 * nprim; a.charge = d_d + 2 * nprim + 3 * B.ncen;
    a.ns = ns; a.ncen = B.ncen; a.nbf = n;
    a.rys = h->rys;
    a.S = d_out; a.T = d_out + no2; a.H = d_out + 2 * no2;
    a.M = M ? d_out + 3 * no2 : nullptr;
    a.dpm_cen = dpm_center;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, h->stream);
    {
        const unsigned blocks = (unsigned)((npairs + 127) / 128);
        if (B.maxl <= 1) onee_kernel<1><<<blocks, 128, 0, h->stream>>>(a);
        else if (B.maxl <= 2) onee_kernel<2><<<blocks, 128, 0, h->stream>>>(a);
        else onee_kernel<4><<<blocks, 128, 0, h->stream>>>(a);
    }
    cudaEventRecord(e1, h->stream);
    if (cudaGetLastError() != cudaSuccess) { rc = UNOMOL_E_CUDA; goto done; }
    cudaMemcpyAsync(host.data(), d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = UNOMOL_E_CUDA; goto done; }
    {
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        h->stats.onee_ms = ms;
    }
    memcpy(S, host.data(), sizeof(double) * no2);
    memcpy(T, host.data() + no2, sizeof(double) * no2);
    memcpy(H, host.data() + 2 * no2, sizeof(double) * no2);
    if (M) memcpy(M, host.data() + 3 * no2, sizeof(double) * 9 * no2);
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d_i); cudaFree(d_d); cudaFree(d_out);
    return rc;
}
