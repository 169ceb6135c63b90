/* oracle/unomol_oracle.h -- TEST INFRASTRUCTURE.
 *
 * Plain-C CPU restatement of the reference hot path (ERIs over contracted Cartesian Gaussian shell
 * quartets by Rys quadrature -> unique-integral list -> J/K digestion into packed G), used only as the
 * CHECKER by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product path
 * (unomol_b200/, include/) never includes, links or loads anything from oracle/.
 *
 * Parity status: PINNED -- tests/test_oracle_vs_reference.py checks every function below against the
 * unmodified reference compiled into oracle/_ref (roots/weights on a dense X grid, every stored
 * integral of the small inputs, G matrices for random P) and tests/test_oracle_golden.py checks the
 * committed fixtures under tests/golden/ that were generated from the reference (generate_golden.py).
 */
#ifndef UNOMOL_ORACLE_H
#define UNOMOL_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAXL 4      /* Basis.hpp:222 */
#define ORACLE_MAXCART 15

typedef struct oracle_basis {
    int nshell, nbf, ncen, maxl, nelec, maxits;
    double eps;
    int int_flag[2], scf_flag[3], prt_flag[3];
    int *npr, *lv, *cen, *off, *poff; /* per shell; poff = offset into alpha/coef */
    double *alpha, *coef;             /* coef AFTER contraction normalisation (Basis.hpp:56-75) */
    double *coef_raw;                 /* as read from the file */
    double *xyz, *charge;             /* per centre */
    int nprim_total;
} oracle_basis;

/* patin.dat reader + Shell::normalize (Basis.hpp:181-255, 56-75).  NULL on failure. */
oracle_basis *oracle_basis_read(const char *path);
void oracle_basis_free(oracle_basis *b);
/* flat copies for ctypes */
void oracle_basis_dims(const oracle_basis *b, int *nshell, int *nbf, int *ncen, int *maxl, int *nelec, int *nprim);
void oracle_basis_copy(const oracle_basis *b, int *npr, int *lv, int *cen, int *off, int *poff, double *alpha,
                       double *coef, double *xyz, double *charge);
void oracle_basis_set_center(oracle_basis *b, int icen, double x, double y, double z);

/* Cartesian component tables (AuxFunctions.hpp:35-41,49-64): ncart, (lx,ly,lz), per-component norm */
int oracle_ncart(int l);
void oracle_cart(int l, int comp, int *lxyz);
double oracle_cart_norm(int l, int comp);

/* Rys roots r_i = t_i^2/(1-t_i^2) and weights, 1..5 roots (Rys.cpp:314-2197). returns 0 / -1 */
int oracle_rys_roots(int nroots, double x, double *r, double *w);

/* One ordered shell quartet, every Cartesian component, out[((a*n2+b)*n3+c)*n4+d]
 * (TwoElectronInts.cpp:420-509 + Rys.hpp:85-212; for l_tot > 8 the McMurchie-Davidson routine,
 * TwoElectronInts.cpp:9-418, by the reference's dispatch rule :661-665).  Returns the number of values. */
int oracle_quartet_block(const oracle_basis *b, int ish, int jsh, int ksh, int lsh, double *out);

/* The unique-integral list of TwoElectronInts::calculate (TwoElectronInts.cpp:511-697): same loop order,
 * canonical filter, |val|>thresh.  vals/ijkl may be NULL to count only.  Returns the number stored;
 * *ncalc (optional) receives the number computed. */
long oracle_unique_eris(const oracle_basis *b, int start_shell, double thresh, double *vals, int *ijkl,
                        long cap, long *ncalc);

/* formGMatrixKernel / formGMatrixKernel2 over a record list (TwoElectronInts.cpp:699-820). G accumulates. */
void oracle_form_g_rhf(long n, const double *vals, const int *ijkl, const double *P, double *G);
void oracle_form_g_uhf(long n, const double *vals, const int *ijkl, const double *PA, const double *PB,
                       double *GA, double *GB);

/* Integral-direct convenience: same loops, digest immediately, nothing stored.  sample_mod>1 processes
 * only shell quartets whose running index % sample_mod == sample_rem (bench cpu_baseline sampling).
 * Returns the number of shell quartets evaluated. */
long oracle_direct_g_rhf(const oracle_basis *b, double thresh, const double *P, double *G, long sample_mod,
                         long sample_rem, long *nprimq);

long oracle_direct_g_uhf(const oracle_basis *b, double thresh, const double *PA, const double *PB, double *GA, double *GB,
                         long sample_mod, long sample_rem, long *nprimq);

/* G_ij = sum_kl P_kl [2 (ij|kl) - (ik|jl)] for nelem basis-function pairs (ij[2e], ij[2e+1]); shells ksh % mod == rem only;
 * out[] accumulates.  Returns the number of shell-quartet blocks evaluated. */
long oracle_g_elements_rhf(const oracle_basis *b, double thresh, const double *P, int nelem, const int *ij, double *out,
                           int mod, int rem);

long oracle_direct_g_rhf_start(const oracle_basis *b, int start_shell, double thresh, const double *P, double *G, long sample_mod,
                               long sample_rem);
/* n shell quartets (two shell pairs each) evaluated + optionally digested (RHF) from C: bench.py cpu_baseline, port flavour */
long oracle_quartet_batch(const oracle_basis *b, long n, const int *shells, const double *P, double *G, int digest);

#ifdef __cplusplus
}
#endif
#endif
