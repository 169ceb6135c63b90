// oracle/ref_harness.cc — TEST INFRASTRUCTURE (ours, not reference code).
//
// Thin C entry points around the UNMODIFIED reference classes so that tests/ and bench.py's
// cpu_baseline / --impl reference legs can drive the reference's own implementation of the hot path
// through ctypes.  Compiled by oracle/Makefile with -I/root/reference and -fno-access-control (only to
// READ Rys::roots/weights and TwoElectronInts::cache); linked with the reference objects into
// oracle/_ref/libunomol_ref.so.  Nothing in the product path links or loads this file.
//
// What each entry drives in the reference:
//   ref_rys_roots        -> Rys::calculate_roots            (Rys.hpp:145-164, Rys.cpp:314-2197)
//   ref_basis_*          -> Basis::Basis(patin.dat)          (Basis.hpp:181-255)
//   ref_quartet_block    -> calc_two_electron_ints_rys       (TwoElectronInts.cpp:420-509) on one ordered
//                           shell quartet, all Cartesian components (no canonical filter); for l_tot > 8
//                           calc_two_electron_ints_md (TwoElectronInts.cpp:269-418), the reference's own
//                           dispatch rule (TwoElectronInts.cpp:661-665)
//   ref_tints_*          -> TwoElectronInts ctor/calculate, formGmatrix x2 (TwoElectronInts.cpp:511-869)
//   ref_one_electron     -> OneElectronInts                  (OneElectronInts.cpp:127-202)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <cmath>
#include <vector>
#include <ctime>
#include "Basis.hpp"
#include "Rys.hpp"
#include "TwoElectronInts.hpp"
#include "OneElectronInts.hpp"

namespace unomol {
// exported by the reference's TwoElectronInts.o (external linkage, not declared in its header)
void calc_two_electron_ints_rys(const ShellQuartet& sq, const AuxFunctions& aux, Rys& rys, TwoInts* sints);
void calc_two_electron_ints_md(const ShellQuartet& sq, const AuxFunctions& aux, MDInts& mds, TwoInts* sints);
}

using namespace unomol;

static double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

extern "C" {

// ---- Rys roots/weights -------------------------------------------------------------------------
int ref_rys_roots(int nroots, double x, double* roots, double* weights) {
    static Rys* rys = nullptr;
    if (!rys) rys = new Rys(4);
    if (nroots < 1 || nroots > 5) return -1;  // rootN (6..9) hangs for 2<~X<~15
    rys->calculate_roots(x, nroots);
    for (int i = 0; i < nroots; ++i) {
        roots[i] = rys->roots[i];
        weights[i] = rys->weights[i];
    }
    return 0;
}

// Rys::rootN (Rys.cpp:231-312), 6..9 roots.  It crashes or hangs for 2 <~ X <~ 15: callers probe it
// in a child process with a timeout (tests/golden/generate_golden.py) and keep the points where it returns.
int ref_rys_rootN(int nroots, double x, double* roots, double* weights) {
    static Rys* rys = nullptr;
    if (!rys) rys = new Rys(4);
    if (nroots < 6 || nroots > 9) return -1;
    rys->calculate_roots(x, nroots);
    for (int i = 0; i < nroots; ++i) {
        roots[i] = rys->roots[i];
        weights[i] = rys->weights[i];
    }
    return 0;
}

// ---- Basis ---------------------------------------------------------------------------------------
void* ref_basis_open(const char* patin_path) { return new Basis(std::string(patin_path)); }
void ref_basis_close(void* b) { delete static_cast<Basis*>(b); }
int ref_basis_nshell(void* b) { return static_cast<Basis*>(b)->number_of_shells(); }
int ref_basis_norb(void* b) { return static_cast<Basis*>(b)->number_of_orbitals(); }
int ref_basis_ncen(void* b) { return static_cast<Basis*>(b)->number_of_centers(); }
int ref_basis_nelec(void* b) { return static_cast<Basis*>(b)->number_of_electrons(); }
double ref_basis_eps(void* b) { return static_cast<Basis*>(b)->scf_eps(); }
// per-shell: npr, L, centre, bf offset; coefficients AFTER Shell::normalize (Basis.hpp:56-75,246)
void ref_basis_shell(void* bp, int ish, int* npr, int* lv, int* cen, int* off, double* al, double* co) {
    const Basis& b = *static_cast<Basis*>(bp);
    const Shell* s = b.shell_ptr() + ish;
    *npr = s->number_of_prims();
    *lv = s->Lvalue();
    *cen = s->center();
    *off = b.offset(ish);
    for (int i = 0; i < *npr; ++i) {
        al[i] = s->alf(i);
        co[i] = s->cof(i);
    }
}
void ref_basis_center(void* bp, int ic, double* q, double* xyz) {
    const Basis& b = *static_cast<Basis*>(bp);
    const Center* c = b.center_ptr() + ic;
    *q = c->charge();
    xyz[0] = c->position(0);
    xyz[1] = c->position(1);
    xyz[2] = c->position(2);
}

// ---- one ordered shell quartet, every Cartesian component ---------------------------------------
// out[((ia*n2+ib)*n3+ic)*n4+id] = (ab|cd) with the reference's l1>=l2 / l3>=l4 swaps applied the way
// calculate() applies them (TwoElectronInts.cpp:563-580,604-621); returns number of values or <0.
int ref_quartet_block(void* bp, int ish, int jsh, int ksh, int lsh, double* out) {
    const Basis& basis = *static_cast<Basis*>(bp);
    const Shell* shell = basis.shell_ptr();
    const Center* center = basis.center_ptr();
    const AuxFunctions& aux(*basis.auxfun_ptr());
    static Rys* rys = nullptr;
    if (!rys) rys = new Rys(4);
    ShellQuartet sq(basis.maxLvalue() > 2 ? basis.maxLvalue() : 2);
    int sh[4] = {ish, jsh, ksh, lsh};
    int lv[4], nls[4];
    for (int t = 0; t < 4; ++t) {
        lv[t] = (shell + sh[t])->Lvalue();
        nls[t] = aux.number_of_lstates(lv[t]);
    }
    const bool use_md = lv[0] + lv[1] + lv[2] + lv[3] > 8;   // TwoElectronInts.cpp:661-665
    bool sw12 = lv[0] < lv[1], sw34 = lv[2] < lv[3];
    int s1 = sw12 ? jsh : ish, s2 = sw12 ? ish : jsh, s3 = sw34 ? lsh : ksh, s4 = sw34 ? ksh : lsh;
    const Shell *p1 = shell + s1, *p2 = shell + s2, *p3 = shell + s3, *p4 = shell + s4;
    sq.npr1 = p1->number_of_prims(); sq.lv1 = p1->Lvalue(); sq.al1 = p1->alf_ptr(); sq.co1 = p1->cof_ptr();
    sq.npr2 = p2->number_of_prims(); sq.lv2 = p2->Lvalue(); sq.al2 = p2->alf_ptr(); sq.co2 = p2->cof_ptr();
    sq.npr3 = p3->number_of_prims(); sq.lv3 = p3->Lvalue(); sq.al3 = p3->alf_ptr(); sq.co3 = p3->cof_ptr();
    sq.npr4 = p4->number_of_prims(); sq.lv4 = p4->Lvalue(); sq.al4 = p4->alf_ptr(); sq.co4 = p4->cof_ptr();
    sq.a = (center + p1->center())->r_vec();
    sq.b = (center + p2->center())->r_vec();
    sq.c = (center + p3->center())->r_vec();
    sq.d = (center + p4->center())->r_vec();
    sq.ab2 = dist_sqr(sq.a, sq.b);
    sq.cd2 = dist_sqr(sq.c, sq.d);
    int n = nls[0] * nls[1] * nls[2] * nls[3];
    std::vector<TwoInts> sints(n);
    int knt = 0;
    for (int ia = 0; ia < nls[0]; ++ia)
        for (int ib = 0; ib < nls[1]; ++ib)
            for (int ic = 0; ic < nls[2]; ++ic)
                for (int id = 0; id < nls[3]; ++id) {
                    sints[knt].val = 0.0;
                    unsigned int l12 = sw12 ? ((ib << 4) + ia) : ((ia << 4) + ib);
                    unsigned int l34 = sw34 ? ((id << 4) + ic) : ((ic << 4) + id);
                    sq.lstates[knt] = (l12 << 8) + l34;
                    sq.norms[knt] = aux.normalization_factor(lv[0], ia) * aux.normalization_factor(lv[1], ib) *
                                    aux.normalization_factor(lv[2], ic) * aux.normalization_factor(lv[3], id);
                    ++knt;
                }
    sq.len = knt;
    if (use_md) {
        static MDInts* mds = nullptr;
        if (!mds) mds = new MDInts(4);
        calc_two_electron_ints_md(sq, aux, *mds, sints.data());
    } else {
        calc_two_electron_ints_rys(sq, aux, *rys, sints.data());
    }
    for (int i = 0; i < n; ++i) out[i] = sints[i].val;
    return n;
}

// ---- a batch of shell quartets, timed from C (bench.py cpu_baseline / --impl reference) -----------
// The reference's formGMatrixKernel is `inline` in TwoElectronInts.cpp (:699-747) and not exported, so the RHF digestion
// of one stored integral is restated here (test infrastructure): same index algebra, same coincidence guards.
static inline void harness_digest_rhf(const double* P, double* G, double val, int i, int j, int k, int l) {
    int ii = i * (i + 1) / 2, ij = ii + j, ik = ii + k, il = ii + l, jk, jl;
    int kk = k * (k + 1) / 2, kl = kk + l;
    if (j >= k) {
        int jj = j * (j + 1) / 2;
        jk = jj + k;
        jl = jj + l;
    } else {
        jk = kk + j;
        jl = (j > l) ? j * (j + 1) / 2 + l : l * (l + 1) / 2 + j;
    }
    double da = val * 2.0 * P[ij], db = val * 2.0 * P[kl];
    double sjl = val * P[ik], sjk = val * P[il], sik = val * P[jl], sil = val * P[jk];
    if (k != l) {
        db = db + db;
        G[ik] -= sik;
        if (i != j && j >= k) G[jk] -= sjk;
    }
    G[il] -= sil;
    G[ij] += db;
    if (i != j && j >= l) G[jl] -= sjl;
    if (ij != kl) {
        if (i != j) da = da + da;
        if (j <= k) {
            G[jk] -= sjk;
            if (i == k && i != j) G[ik] -= sik;
            if (k != l && j <= l) G[jl] -= sjl;
        }
        G[kl] += da;
    }
}

// n shell quartets (shells[4q..4q+3] = two shell pairs in any order): the reference's calc_two_electron_ints_rys / _md on
// each, called exactly as calculate() calls it (ordering, swaps, 1e-14 storage threshold), and -- when digest != 0 -- the
// canonical-function filter of calculate() (:627-633) plus the digestion above into G (packed).  Thread-safe: scratch
// objects are per thread, G must be per thread.  Returns the number of integrals that passed the storage threshold.
long ref_quartet_batch(void* bp, long n, const int* shells, const double* P, double* G, int digest) {
    const Basis& basis = *static_cast<Basis*>(bp);
    const Shell* shell = basis.shell_ptr();
    const Center* center = basis.center_ptr();
    const AuxFunctions& aux(*basis.auxfun_ptr());
    thread_local Rys* rys = nullptr;
    thread_local MDInts* mds = nullptr;
    thread_local ShellQuartet* sqp = nullptr;
    thread_local TwoInts* sints = nullptr;
    if (!rys) {
        rys = new Rys(4);
        mds = new MDInts(4);
        sqp = new ShellQuartet(basis.maxLvalue() > 2 ? basis.maxLvalue() : 2);
        sints = new TwoInts[50625];
    }
    ShellQuartet& sq = *sqp;
    long stored = 0;
    for (long q = 0; q < n; ++q) {
        // canonical shell order of calculate(): ish >= jsh, (ish,jsh) >= (ksh,lsh), ksh >= lsh
        int a = shells[4 * q], b = shells[4 * q + 1], c = shells[4 * q + 2], d = shells[4 * q + 3];
        if (a < b) std::swap(a, b);
        if (c < d) std::swap(c, d);
        if (a < c || (a == c && b < d)) { std::swap(a, c); std::swap(b, d); }
        // calculate() visits (ish jsh | ish lsh) for BOTH orders of jsh != lsh; the function filter splits the block between them
        const int nrep = (digest && a == c && b != d) ? 2 : 1;
        for (int rep = 0; rep < nrep; ++rep) {
        const int ish = a, jsh = rep ? d : b, ksh = c, lsh = rep ? b : d;
        int sh[4] = {ish, jsh, ksh, lsh}, lv[4], nls[4], off[4];
        for (int t = 0; t < 4; ++t) {
            lv[t] = (shell + sh[t])->Lvalue();
            nls[t] = aux.number_of_lstates(lv[t]);
            off[t] = basis.offset(sh[t]);
        }
        const bool use_md = lv[0] + lv[1] + lv[2] + lv[3] > 8;
        const bool sw12 = lv[0] < lv[1], sw34 = lv[2] < lv[3];
        const int s1 = sw12 ? jsh : ish, s2 = sw12 ? ish : jsh, s3 = sw34 ? lsh : ksh, s4 = sw34 ? ksh : lsh;
        const Shell *p1 = shell + s1, *p2 = shell + s2, *p3 = shell + s3, *p4 = shell + s4;
        sq.npr1 = p1->number_of_prims(); sq.lv1 = p1->Lvalue(); sq.al1 = p1->alf_ptr(); sq.co1 = p1->cof_ptr();
        sq.npr2 = p2->number_of_prims(); sq.lv2 = p2->Lvalue(); sq.al2 = p2->alf_ptr(); sq.co2 = p2->cof_ptr();
        sq.npr3 = p3->number_of_prims(); sq.lv3 = p3->Lvalue(); sq.al3 = p3->alf_ptr(); sq.co3 = p3->cof_ptr();
        sq.npr4 = p4->number_of_prims(); sq.lv4 = p4->Lvalue(); sq.al4 = p4->alf_ptr(); sq.co4 = p4->cof_ptr();
        sq.a = (center + p1->center())->r_vec();
        sq.b = (center + p2->center())->r_vec();
        sq.c = (center + p3->center())->r_vec();
        sq.d = (center + p4->center())->r_vec();
        sq.ab2 = dist_sqr(sq.a, sq.b);
        sq.cd2 = dist_sqr(sq.c, sq.d);
        int knt = 0;
        for (int ia = 0; ia < nls[0]; ++ia) {
            const int ir = off[0] + ia;
            for (int ib = 0; ib < nls[1]; ++ib) {
                const int jr = off[1] + ib;
                if (digest && jr > ir) break;
                for (int ic = 0; ic < nls[2]; ++ic) {
                    const int kr = off[2] + ic;
                    if (digest && kr > ir) break;
                    for (int id = 0; id < nls[3]; ++id) {
                        const int lr = off[3] + id;
                        if (digest && (lr > kr || (ir == kr && lr > jr))) break;
                        sints[knt].val = 0.0;
                        sints[knt].i = ir; sints[knt].j = jr; sints[knt].k = kr; sints[knt].l = lr;
                        unsigned int l12 = sw12 ? ((ib << 4) + ia) : ((ia << 4) + ib);
                        unsigned int l34 = sw34 ? ((id << 4) + ic) : ((ic << 4) + id);
                        sq.lstates[knt] = (l12 << 8) + l34;
                        sq.norms[knt] = aux.normalization_factor(lv[0], ia) * aux.normalization_factor(lv[1], ib) *
                                        aux.normalization_factor(lv[2], ic) * aux.normalization_factor(lv[3], id);
                        ++knt;
                    }
                }
            }
        }
        sq.len = knt;
        if (!knt) continue;
        if (use_md) calc_two_electron_ints_md(sq, aux, *mds, sints);
        else calc_two_electron_ints_rys(sq, aux, *rys, sints);
        for (int t = 0; t < knt; ++t)
            if (std::fabs(sints[t].val) > 1.e-14) {      // storage threshold, TwoElectronInts.cpp:513,667-671
                ++stored;
                if (digest) harness_digest_rhf(P, G, sints[t].val, sints[t].i, sints[t].j, sints[t].k, sints[t].l);
            }
        }   // rep
    }
    return stored;
}

// ---- TwoElectronInts: the reference's stored-integral path --------------------------------------
struct RefTints {
    TwoElectronInts* t;
    double eri_seconds;
};
void* ref_tints_create(void* bp, int start_shell, const char* cache_name) {
    RefTints* r = new RefTints;
    double t0 = now_s();
    r->t = new TwoElectronInts(*static_cast<Basis*>(bp), start_shell, std::string(cache_name));
    r->eri_seconds = now_s() - t0;
    return r;
}
void ref_tints_destroy(void* rp) {
    RefTints* r = static_cast<RefTints*>(rp);
    delete r->t;
    delete r;
}
double ref_tints_eri_seconds(void* rp) { return static_cast<RefTints*>(rp)->eri_seconds; }
long ref_tints_count(void* rp) {
    RefTints* r = static_cast<RefTints*>(rp);
    return (long)(r->t->cache.total_size() / sizeof(TwoInts));
}
// copies the stored unique-integral records (val,i,j,k,l; TwoElectronInts.hpp:20-23) in storage order
long ref_tints_dump(void* rp, double* vals, int* ijkl, long cap) {
    RefTints* r = static_cast<RefTints*>(rp);
    long n = (long)(r->t->cache.total_size() / sizeof(TwoInts));
    if (n > cap) n = cap;
    r->t->cache.open_for_reading();
    const long BIN = 4096;
    std::vector<TwoInts> buf(BIN);
    long done = 0;
    while (done < n) {
        long m = (n - done < BIN) ? (n - done) : BIN;
        r->t->cache.read(buf.data(), m);
        for (long q = 0; q < m; ++q) {
            vals[done + q] = buf[q].val;
            ijkl[4 * (done + q) + 0] = buf[q].i;
            ijkl[4 * (done + q) + 1] = buf[q].j;
            ijkl[4 * (done + q) + 2] = buf[q].k;
            ijkl[4 * (done + q) + 3] = buf[q].l;
        }
        done += m;
    }
    r->t->cache.close();
    return n;
}
// G += 2J-K (packed lower-triangular); returns seconds spent inside formGmatrix
double ref_tints_form_g_rhf(void* rp, const double* P, double* G) {
    double t0 = now_s();
    static_cast<RefTints*>(rp)->t->formGmatrix(P, G);
    return now_s() - t0;
}
double ref_tints_form_g_uhf(void* rp, const double* PA, const double* PB, double* GA, double* GB) {
    double t0 = now_s();
    static_cast<RefTints*>(rp)->t->formGmatrix(PA, PB, GA, GB);
    return now_s() - t0;
}

// ---- one-electron matrices (packed): used to pin the host-side S/T/H of the SCF driver ----------
void ref_one_electron(void* bp, double* S, double* T, double* H) {
    OneElectronInts(*static_cast<Basis*>(bp), S, T, H);
}

}  // extern "C"
