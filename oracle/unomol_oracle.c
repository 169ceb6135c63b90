/* oracle/unomol_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.  See unomol_oracle.h.
 *
 * Restates, with flat arrays and no classes, the arithmetic contract of the reference hot path
 * (SURVEY.md section 9).  Every function cites the reference lines it follows.
 */
#include "unomol_oracle.h"
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * Cartesian component tables.  Order: lx = L..0, then ly = (L-lx)..0 (AuxFunctions.hpp:35-41).
 * Per-component factor 1/sqrt(df[lx] df[ly] df[lz]) with the reference's own df recurrence
 * (AuxFunctions.hpp:49-64): df[0]=1, dx=1; df[i]=df[i-1]*dx; dx*=(2i+1).  That recurrence equals
 * (2l-1)!! for l<=2 (1,1,3) and deviates for l>=3 (45, 4725); restated as written.
 * ---------------------------------------------------------------------------------------------- */
static int g_cart[ORACLE_MAXL + 1][ORACLE_MAXCART][3];
static double g_cnorm[ORACLE_MAXL + 1][ORACLE_MAXCART];
static int g_tables_ready = 0;

static void init_tables(void) {
    if (g_tables_ready) return;
    double df[ORACLE_MAXL + 1];
    df[0] = 1.0;
    double dx = 1.0;
    for (int i = 1; i <= ORACLE_MAXL; ++i) {
        df[i] = df[i - 1] * dx;
        dx *= (2 * i + 1);
    }
    for (int l = 0; l <= ORACLE_MAXL; ++l) {
        int k = 0;
        for (int lx = l; lx >= 0; --lx)
            for (int ly = l - lx; ly >= 0; --ly) {
                g_cart[l][k][0] = lx;
                g_cart[l][k][1] = ly;
                g_cart[l][k][2] = l - lx - ly;
                g_cnorm[l][k] = 1.0 / sqrt(df[lx] * df[ly] * df[l - lx - ly]);
                ++k;
            }
    }
    g_tables_ready = 1;
}

int oracle_ncart(int l) { return (l + 1) * (l + 2) / 2; }
void oracle_cart(int l, int comp, int *lxyz) {
    init_tables();
    lxyz[0] = g_cart[l][comp][0];
    lxyz[1] = g_cart[l][comp][1];
    lxyz[2] = g_cart[l][comp][2];
}
double oracle_cart_norm(int l, int comp) {
    init_tables();
    return g_cnorm[l][comp];
}

/* ------------------------------------------------------------------------------------------------
 * Basis: patin.dat reader (Basis.hpp:181-255, int_flag[0]==0 path only: no posin.bas augmentation)
 * and contracted-shell normalisation (Basis.hpp:56-75).
 * ---------------------------------------------------------------------------------------------- */
oracle_basis *oracle_basis_read(const char *path) {
    init_tables();
    FILE *fp = fopen(path, "r");
    if (!fp) return NULL;
    oracle_basis *b = (oracle_basis *)calloc(1, sizeof(oracle_basis));
    int ok = 1;
    ok &= fscanf(fp, "%d %d %d %d", &b->nshell, &b->nbf, &b->ncen, &b->maxl) == 4;
    ok &= fscanf(fp, "%d %d", &b->nelec, &b->maxits) == 2;
    ok &= fscanf(fp, "%lf", &b->eps) == 1;
    ok &= fscanf(fp, "%d %d", &b->int_flag[0], &b->int_flag[1]) == 2;
    ok &= fscanf(fp, "%d %d %d", &b->scf_flag[0], &b->scf_flag[1], &b->scf_flag[2]) == 3;
    ok &= fscanf(fp, "%d %d %d", &b->prt_flag[0], &b->prt_flag[1], &b->prt_flag[2]) == 3;
    if (!ok || b->maxl > ORACLE_MAXL || b->nshell <= 0) {
        fclose(fp);
        free(b);
        return NULL;
    }
    b->xyz = (double *)calloc(3 * (size_t)b->ncen, sizeof(double));
    b->charge = (double *)calloc((size_t)b->ncen, sizeof(double));
    for (int i = 0; i < b->ncen; ++i)
        ok &= fscanf(fp, "%lf %lf %lf %lf", &b->charge[i], &b->xyz[3 * i], &b->xyz[3 * i + 1], &b->xyz[3 * i + 2]) == 4;
    b->npr = (int *)calloc((size_t)b->nshell, sizeof(int));
    b->lv = (int *)calloc((size_t)b->nshell, sizeof(int));
    b->cen = (int *)calloc((size_t)b->nshell, sizeof(int));
    b->off = (int *)calloc((size_t)b->nshell, sizeof(int));
    b->poff = (int *)calloc((size_t)b->nshell, sizeof(int));
    size_t cap = 64, np = 0;
    b->alpha = (double *)malloc(cap * sizeof(double));
    b->coef_raw = (double *)malloc(cap * sizeof(double));
    int off = 0;
    for (int s = 0; s < b->nshell && ok; ++s) {
        ok &= fscanf(fp, "%d %d %d", &b->npr[s], &b->lv[s], &b->cen[s]) == 3;
        if (!ok || b->lv[s] > ORACLE_MAXL || b->npr[s] <= 0) { ok = 0; break; }
        b->off[s] = off;                      /* Basis.hpp:226-232 */
        off += oracle_ncart(b->lv[s]);
        b->poff[s] = (int)np;
        for (int k = 0; k < b->npr[s]; ++k) {
            if (np == cap) {
                cap *= 2;
                b->alpha = (double *)realloc(b->alpha, cap * sizeof(double));
                b->coef_raw = (double *)realloc(b->coef_raw, cap * sizeof(double));
            }
            ok &= fscanf(fp, "%lf %lf", &b->alpha[np], &b->coef_raw[np]) == 2;
            ++np;
        }
    }
    fclose(fp);
    if (!ok) {
        oracle_basis_free(b);
        return NULL;
    }
    b->nprim_total = (int)np;
    b->coef = (double *)malloc((np ? np : 1) * sizeof(double));
    /* Shell::normalize, Basis.hpp:56-75 (the double sum uses the RAW coefficients) */
    const double twofact = 2.8284271247461903, piterm = 5.568327996831707;
    for (int s = 0; s < b->nshell; ++s) {
        const double *al = b->alpha + b->poff[s];
        const double *co = b->coef_raw + b->poff[s];
        double lpow = 1.5 + b->lv[s], sum = 0.0;
        for (int i = 0; i < b->npr[s]; ++i)
            for (int j = 0; j < b->npr[s]; ++j) sum += co[i] * co[j] * pow(sqrt(al[i] * al[j]) / (al[i] + al[j]), lpow);
        sum *= twofact;
        sum = 1.0 / sqrt(sum);
        for (int i = 0; i < b->npr[s]; ++i) b->coef[b->poff[s] + i] = co[i] * sum * sqrt(pow(2 * al[i], lpow) / piterm);
    }
    /* eps floor, Basis.hpp:249-254 */
    double xeps = DBL_EPSILON * b->nbf * b->nbf * 0.5;
    if (b->eps < xeps) b->eps = xeps;
    return b;
}

void oracle_basis_free(oracle_basis *b) {
    if (!b) return;
    free(b->npr); free(b->lv); free(b->cen); free(b->off); free(b->poff);
    free(b->alpha); free(b->coef); free(b->coef_raw); free(b->xyz); free(b->charge);
    free(b);
}

void oracle_basis_dims(const oracle_basis *b, int *nshell, int *nbf, int *ncen, int *maxl, int *nelec, int *nprim) {
    *nshell = b->nshell; *nbf = b->nbf; *ncen = b->ncen; *maxl = b->maxl; *nelec = b->nelec; *nprim = b->nprim_total;
}

void oracle_basis_copy(const oracle_basis *b, int *npr, int *lv, int *cen, int *off, int *poff, double *alpha,
                       double *coef, double *xyz, double *charge) {
    memcpy(npr, b->npr, sizeof(int) * b->nshell);
    memcpy(lv, b->lv, sizeof(int) * b->nshell);
    memcpy(cen, b->cen, sizeof(int) * b->nshell);
    memcpy(off, b->off, sizeof(int) * b->nshell);
    memcpy(poff, b->poff, sizeof(int) * b->nshell);
    memcpy(alpha, b->alpha, sizeof(double) * b->nprim_total);
    memcpy(coef, b->coef, sizeof(double) * b->nprim_total);
    memcpy(xyz, b->xyz, sizeof(double) * 3 * b->ncen);
    memcpy(charge, b->charge, sizeof(double) * b->ncen);
}

void oracle_basis_set_center(oracle_basis *b, int icen, double x, double y, double z) {
    b->xyz[3 * icen] = x; b->xyz[3 * icen + 1] = y; b->xyz[3 * icen + 2] = z;
}

/* ------------------------------------------------------------------------------------------------
 * Rys 2-D recurrence and transfer (Rys.hpp:113-143,194-212 and :85-111,173-192).
 * ---------------------------------------------------------------------------------------------- */
#define GDIM 10 /* l12, l34 <= 8 */
#define ORACLE_MAXFUNC 50625 /* 15^4: (gg|gg) */
typedef struct {
    double G[3][5][GDIM][GDIM]; /* [axis][root][i][j] */
    double w[5];
    int nroots;
} rys_tables;

static const double BINOM[5][5] = {{1, 0, 0, 0, 0}, {1, 1, 0, 0, 0}, {1, 2, 1, 0, 0}, {1, 3, 3, 1, 0}, {1, 4, 6, 4, 1}};

/* RecurKernel, Rys.hpp:194-212 */
static void vrr_axis(double (*G)[GDIM], int l12, int l34, double B00, double B1, double B1p, double C, double Cp) {
    G[0][0] = 1.0;
    G[0][1] = Cp;
    G[1][0] = C;
    G[1][1] = B00 + C * Cp;
    if (l12 < 2 && l34 < 2) return;
    for (int j = 1; j < l34; ++j) {
        G[0][j + 1] = j * B1p * G[0][j - 1] + Cp * G[0][j];
        G[1][j + 1] = j * B1p * G[1][j - 1] + B00 * G[0][j] + Cp * G[1][j];
    }
    for (int i = 2; i <= l12; ++i) {
        G[i][0] = (i - 1) * B1 * G[i - 2][0] + C * G[i - 1][0];
        G[i][1] = i * B00 * G[i - 1][0] + Cp * G[i][0];
        for (int j = 1; j < l34; ++j) G[i][j + 1] = j * B1p * G[i][j - 1] + i * B00 * G[i - 1][j] + Cp * G[i][j];
    }
}

/* Rys::Recur, Rys.hpp:113-143 */
static void rys_recur(rys_tables *T, const double *p, const double *q, const double *pa, const double *qc, double pxp,
                      double qxp, double txp, int l12, int l34, int nroots) {
    double pq[3] = {p[0] - q[0], p[1] - q[1], p[2] - q[2]};
    double pq2 = pq[0] * pq[0] + pq[1] * pq[1] + pq[2] * pq[2];
    double x = pxp * qxp / txp * pq2;
    double r[5];
    oracle_rys_roots(nroots, x, r, T->w);
    T->nroots = nroots;
    for (int ir = 0; ir < nroots; ++ir) {
        double dr = r[ir] / (1.0 + r[ir]);
        double fff = dr / txp;
        double B00 = 0.5 * fff;
        double B1 = (0.5 - B00 * qxp) / pxp;
        double B1p = (0.5 - B00 * pxp) / qxp;
        for (int ax = 0; ax < 3; ++ax) {
            double C = pa[ax] - qxp * pq[ax] * fff;
            double Cp = qc[ax] + pxp * pq[ax] * fff;
            vrr_axis(T->G[ax][ir], l12, l34, B00, B1, B1p, C, Cp);
        }
    }
}

/* ShiftKernel, Rys.hpp:173-192 */
static double shift_axis(double (*G)[GDIM], double abx, double cdx, int l12, int l2, int l34, int l4) {
    double sum = 0.0, x12t = 1.0;
    for (int i = 0; i <= l2; ++i) {
        double x34t = BINOM[l2][i] * x12t;
        for (int j = 0; j <= l4; ++j) {
            sum += BINOM[l4][j] * x34t * G[l12 - i][l34 - j];
            x34t *= cdx;
        }
        x12t *= abx;
    }
    return sum;
}

/* Rys::Shift, Rys.hpp:85-111 */
static double rys_shift(rys_tables *T, const double *ab, const double *cd, const int *lv1, const int *lv2,
                        const int *lv3, const int *lv4) {
    double sum = 0.0;
    for (int ir = 0; ir < T->nroots; ++ir) {
        double prod = T->w[ir];
        double xyz = 1.0;
        for (int ax = 0; ax < 3; ++ax)
            xyz *= shift_axis(T->G[ax][ir], ab[ax], cd[ax], lv1[ax] + lv2[ax], lv2[ax], lv3[ax] + lv4[ax], lv4[ax]);
        sum += xyz * prod;
    }
    return sum;
}

/* ------------------------------------------------------------------------------------------------
 * One shell quartet in the reference's ShellQuartet view (TwoElectronInts.hpp:25-62): shells already
 * ordered l1>=l2, l3>=l4; a function list of packed component ids + norm products.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int s1, s2, s3, s4;  /* shell ids after the l-ordering swaps */
    int len;
    int comp[ORACLE_MAXFUNC][4];   /* component ids (for shells s1..s4, i.e. after swap) */
    double norm[ORACLE_MAXFUNC];
} quartet_view;

/* calc_two_electron_ints_rys, TwoElectronInts.cpp:420-509 */
static void eri_quartet_rys(const oracle_basis *b, const quartet_view *q, double *vals, long *nprimq) {
    const double SRterm = 34.9868366552497250, threshold = 1.e-12;
    const int s1 = q->s1, s2 = q->s2, s3 = q->s3, s4 = q->s4;
    const double *A = b->xyz + 3 * b->cen[s1], *B = b->xyz + 3 * b->cen[s2];
    const double *C = b->xyz + 3 * b->cen[s3], *D = b->xyz + 3 * b->cen[s4];
    const double *al1 = b->alpha + b->poff[s1], *co1 = b->coef + b->poff[s1];
    const double *al2 = b->alpha + b->poff[s2], *co2 = b->coef + b->poff[s2];
    const double *al3 = b->alpha + b->poff[s3], *co3 = b->coef + b->poff[s3];
    const double *al4 = b->alpha + b->poff[s4], *co4 = b->coef + b->poff[s4];
    const int lv1 = b->lv[s1], lv2 = b->lv[s2], lv3 = b->lv[s3], lv4 = b->lv[s4];
    double ab[3] = {A[0] - B[0], A[1] - B[1], A[2] - B[2]};
    double cd[3] = {C[0] - D[0], C[1] - D[1], C[2] - D[2]};
    double ab2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
    double cd2 = cd[0] * cd[0] + cd[1] * cd[1] + cd[2] * cd[2];
    const int lvt12 = lv1 + lv2, lvt34 = lv3 + lv4, nroots = (lvt12 + lvt34) / 2 + 1;
    static _Thread_local rys_tables T; /* per-thread scratch: the reference is single-threaded (TwoElectronInts.cpp:529-534), the test
                                        * harness runs several walks side by side (oracle.py: direct_g_threads) */
    for (int k = 0; k < q->len; ++k) vals[k] = 0.0;
    const int same12 = (s1 == s2), same34 = (s3 == s4); /* the reference's pointer test al1==al2 (:444,:466) */
    for (int i = 0; i < b->npr[s1]; ++i) {
        double axp = al1[i], c1 = co1[i], f12 = 1.0;
        int jend = b->npr[s2];
        if (same12) { f12 = 2.0; jend = i + 1; }
        for (int j = 0; j < jend; ++j) {
            if (i == j) f12 = 1.0;
            double c12 = c1 * f12 * co2[j];
            double bxp = al2[j], pxp = axp + bxp, abi = 1.0 / pxp;
            double s12 = exp(-axp * bxp * ab2 * abi);
            double p[3], pa[3];
            for (int t = 0; t < 3; ++t) {
                p[t] = (axp * A[t] + bxp * B[t]) * abi;
                pa[t] = p[t] - A[t];
            }
            for (int k = 0; k < b->npr[s3]; ++k) {
                double cxp = al3[k], c3 = co3[k], f34 = 1.0;
                int lend = b->npr[s4];
                if (same34) { f34 = 2.0; lend = k + 1; }
                for (int l = 0; l < lend; ++l) {
                    if (k == l) f34 = 1.0;
                    double c34 = c3 * f34 * co4[l];
                    double dxp = al4[l], qxp = cxp + dxp, cdi = 1.0 / qxp;
                    double s34 = exp(-cxp * dxp * cd2 * cdi);
                    double txp = pxp + qxp;
                    double sr = SRterm * s12 * s34 * abi * cdi / sqrt(txp);
                    if (sr < threshold) continue; /* :479 -- BEFORE the contraction coefficients */
                    sr *= c12 * c34;
                    double qq[3], qc[3];
                    for (int t = 0; t < 3; ++t) {
                        qq[t] = (cxp * C[t] + dxp * D[t]) * cdi;
                        qc[t] = qq[t] - C[t];
                    }
                    if (nprimq) ++*nprimq;
                    rys_recur(&T, p, qq, pa, qc, pxp, qxp, txp, lvt12, lvt34, nroots);
                    for (int kc = 0; kc < q->len; ++kc) {
                        const int *c = q->comp[kc];
                        double sum = rys_shift(&T, ab, cd, g_cart[lv1][c[0]], g_cart[lv2][c[1]], g_cart[lv3][c[2]],
                                               g_cart[lv4][c[3]]);
                        vals[kc] += sum * sr * q->norm[kc];
                    }
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * McMurchie-Davidson path for l_tot > 8 (f/g shells): calc_two_electron_ints_md and its one-/two-centre
 * variants (TwoElectronInts.cpp:9-418), MD_Dfunction::eval (MD_Dfunction.hpp:39-72), MD_Rfunction::eval
 * -> Fgamma + loop_eval (MD_Rfunction.hpp:49-70, 2185-2319).  The one- and two-centre variants are the
 * general routine with P-A = P-B = 0 (and P-Q taken from the centres); they are restated by setting
 * those differences to exactly zero when both shells of a pair sit on the same centre (the reference
 * tests pointer equality of the centre vectors, TwoElectronInts.cpp:274-275).  No primitive threshold on
 * this path.  Fgamma keeps the reference's semantics on purpose: power series + downward recursion for
 * t <= 20, bare asymptotic value + upward recursion for t > 20 (relative error up to 2.5e-10 just above 20).
 * ---------------------------------------------------------------------------------------------- */
#define MD_LDIM 5   /* l per shell <= 4 */
#define MD_TDIM 9   /* hermite index <= 8 per pair */
#define MD_RDIM 17  /* l_tot <= 16 */

/* MD_Rfunction::Fgamma, MD_Rfunction.hpp:2185-2206 */
static void md_fgamma(double *fm, double t, int m) {
    const double tcrit = 20.0, sqrtpi = 0.88622692545275801365, eps = 1.e-12;
    if (t > tcrit) {
        fm[0] = sqrtpi / sqrt(t);
        for (int i = 1; i <= m; i++) fm[i] = fm[i - 1] * (i - 0.5) / t;
        return;
    }
    double mphalf = m + 0.5, term = 0.5 / mphalf, sum = term;
    for (int i = 1; i <= 200; i++) {
        term *= (t / (mphalf + i));
        sum += term;
        if (term < eps) break;
    }
    double twot = 2.0 * t, expt = exp(-t);
    fm[m] = sum * expt;
    for (int i = m - 1; i >= 0; i--) fm[i] = (fm[i + 1] * twot + expt) / (i + i + 1.0);
}

/* MD_Dfunction::eval, MD_Dfunction.hpp:39-72: E[i][j][n], i <= l1 (first shell), j <= l2, n <= i+j */
static void md_ecoef(double (*E)[MD_LDIM][MD_TDIM], double abi, double ax, double bx, int l1, int l2) {
    for (int i = 0; i <= l1; i++)
        for (int j = 0; j <= l2; j++)
            for (int k = 0; k < MD_TDIM; k++) E[i][j][k] = 0.0;
    E[0][0][0] = 1.0;
    for (int j = 1; j <= l2; j++) {
        E[0][j][0] = bx * E[0][j - 1][0] + E[0][j - 1][1];
        for (int n = 1; n < j; n++) E[0][j][n] = abi * E[0][j - 1][n - 1] + bx * E[0][j - 1][n] + (n + 1) * E[0][j - 1][n + 1];
        E[0][j][j] = abi * E[0][j - 1][j - 1];
    }
    for (int i = 1; i <= l1; i++)
        for (int j = 0; j <= l2; j++) {
            int ipj = i + j;
            E[i][j][0] = ax * E[i - 1][j][0] + E[i - 1][j][1];
            for (int n = 1; n < ipj; n++) E[i][j][n] = abi * E[i - 1][j][n - 1] + ax * E[i - 1][j][n] + (n + 1) * E[i - 1][j][n + 1];
            E[i][j][ipj] = abi * E[i - 1][j][ipj - 1];
        }
}

/* MD_Rfunction::eval + loop_eval, MD_Rfunction.hpp:49-70, 2208-2319: R[lx][ly][lz] for lx+ly+lz <= ltot.
 * Same recurrences (z first, then y, then x; coefficient (l-1) on the l-2 term), written with one loop per
 * axis instead of the reference's unrolled l = 1, 2 cases. */
static _Thread_local double g_rz[MD_RDIM][MD_RDIM][MD_RDIM][MD_RDIM + 1];
static void md_rtensor(double (*R)[MD_RDIM][MD_RDIM], double sr, double t, double w, const double *pq, int ltot) {
    double *r0 = g_rz[0][0][0];
    md_fgamma(r0, t, ltot);
    double term = -(w + w), sterm = sr;
    for (int m = 0; m <= ltot; ++m) {
        r0[m] = sterm * r0[m];
        sterm *= term;
    }
    const double x = pq[0], y = pq[1], z = pq[2];
    for (int lz = 1; lz <= ltot; ++lz)
        for (int m = 0; m <= ltot - lz; ++m)
            g_rz[0][0][lz][m] = z * g_rz[0][0][lz - 1][m + 1] + (lz > 1 ? (lz - 1) * g_rz[0][0][lz - 2][m + 1] : 0.0);
    for (int ly = 1; ly <= ltot; ++ly)
        for (int lz = 0; lz <= ltot - ly; ++lz)
            for (int m = 0; m <= ltot - ly - lz; ++m)
                g_rz[0][ly][lz][m] = y * g_rz[0][ly - 1][lz][m + 1] + (ly > 1 ? (ly - 1) * g_rz[0][ly - 2][lz][m + 1] : 0.0);
    for (int lx = 1; lx <= ltot; ++lx)
        for (int ly = 0; ly <= ltot - lx; ++ly)
            for (int lz = 0; lz <= ltot - lx - ly; ++lz)
                for (int m = 0; m <= ltot - lx - ly - lz; ++m)
                    g_rz[lx][ly][lz][m] = x * g_rz[lx - 1][ly][lz][m + 1] + (lx > 1 ? (lx - 1) * g_rz[lx - 2][ly][lz][m + 1] : 0.0);
    for (int lx = 0; lx <= ltot; ++lx)
        for (int ly = 0; ly <= ltot - lx; ++ly)
            for (int lz = 0; lz <= ltot - lx - ly; ++lz) R[lx][ly][lz] = g_rz[lx][ly][lz][0];
}

/* calc_two_electron_ints_md, TwoElectronInts.cpp:269-418 */
static void eri_quartet_md(const oracle_basis *b, const quartet_view *q, double *vals, long *nprimq) {
    const double SRterm = 34.9868366552497250;
    const int s1 = q->s1, s2 = q->s2, s3 = q->s3, s4 = q->s4;
    const double *A = b->xyz + 3 * b->cen[s1], *B = b->xyz + 3 * b->cen[s2];
    const double *C = b->xyz + 3 * b->cen[s3], *D = b->xyz + 3 * b->cen[s4];
    const double *al1 = b->alpha + b->poff[s1], *co1 = b->coef + b->poff[s1];
    const double *al2 = b->alpha + b->poff[s2], *co2 = b->coef + b->poff[s2];
    const double *al3 = b->alpha + b->poff[s3], *co3 = b->coef + b->poff[s3];
    const double *al4 = b->alpha + b->poff[s4], *co4 = b->coef + b->poff[s4];
    const int lv1 = b->lv[s1], lv2 = b->lv[s2], lv3 = b->lv[s3], lv4 = b->lv[s4];
    const int lvt = lv1 + lv2 + lv3 + lv4;
    double ab2 = 0.0, cd2 = 0.0;
    for (int t = 0; t < 3; ++t) {
        ab2 += (A[t] - B[t]) * (A[t] - B[t]);
        cd2 += (C[t] - D[t]) * (C[t] - D[t]);
    }
    const int one12 = (b->cen[s1] == b->cen[s2]), one34 = (b->cen[s3] == b->cen[s4]);
    const int same12 = (s1 == s2), same34 = (s3 == s4);
    static double E12[3][MD_LDIM][MD_LDIM][MD_TDIM], E34[3][MD_LDIM][MD_LDIM][MD_TDIM], R[MD_RDIM][MD_RDIM][MD_RDIM];
    for (int k = 0; k < q->len; ++k) vals[k] = 0.0;
    for (int i = 0; i < b->npr[s1]; ++i) {
        double axp = al1[i], c1 = co1[i], f12 = 1.0;
        int jend = b->npr[s2];
        if (same12) { f12 = 2.0; jend = i + 1; }
        for (int j = 0; j < jend; ++j) {
            if (i == j) f12 = 1.0;
            double c12 = c1 * f12 * co2[j];
            double bxp = al2[j], pxp = axp + bxp, abi = 1.0 / pxp;
            double s12 = c12 * exp(-axp * bxp * ab2 * abi);
            double p[3];
            for (int t = 0; t < 3; ++t) p[t] = one12 ? A[t] : (axp * A[t] + bxp * B[t]) * abi;
            abi *= 0.5;
            for (int t = 0; t < 3; ++t) md_ecoef(E12[t], abi, p[t] - A[t], p[t] - B[t], lv1, lv2);
            for (int k = 0; k < b->npr[s3]; ++k) {
                double cxp = al3[k], c3 = co3[k], f34 = 1.0;
                int lend = b->npr[s4];
                if (same34) { f34 = 2.0; lend = k + 1; }
                for (int l = 0; l < lend; ++l) {
                    if (k == l) f34 = 1.0;
                    double c34 = c3 * f34 * co4[l];
                    double dxp = al4[l], qxp = cxp + dxp, cdi = 1.0 / qxp;
                    double s34 = c34 * exp(-cxp * dxp * cd2 * cdi);
                    double txp = pxp + qxp;
                    double sr = SRterm * s12 * s34 * 2. * abi * cdi / sqrt(txp);
                    double qq[3], pq[3];
                    for (int t = 0; t < 3; ++t) qq[t] = one34 ? C[t] : (cxp * C[t] + dxp * D[t]) * cdi;
                    cdi *= 0.5;
                    for (int t = 0; t < 3; ++t) md_ecoef(E34[t], cdi, qq[t] - C[t], qq[t] - D[t], lv3, lv4);
                    for (int t = 0; t < 3; ++t) pq[t] = p[t] - qq[t];
                    double pq2 = pq[0] * pq[0] + pq[1] * pq[1] + pq[2] * pq[2];
                    double w = pxp * qxp / txp, tt = w * pq2;
                    if (nprimq) ++*nprimq;
                    md_rtensor(R, sr, tt, w, pq, lvt);
                    for (int kc = 0; kc < q->len; ++kc) {
                        const int *c = q->comp[kc];
                        const int *v1 = g_cart[lv1][c[0]], *v2 = g_cart[lv2][c[1]], *v3 = g_cart[lv3][c[2]], *v4 = g_cart[lv4][c[3]];
                        const int l12 = v1[0] + v2[0], m12 = v1[1] + v2[1], n12 = v1[2] + v2[2];
                        const int l34 = v3[0] + v4[0], m34 = v3[1] + v4[1], n34 = v3[2] + v4[2];
                        double sum = 0;
                        for (int ix12 = 0; ix12 <= l12; ++ix12)
                            for (int iy12 = 0; iy12 <= m12; ++iy12)
                                for (int iz12 = 0; iz12 <= n12; ++iz12) {
                                    double v12 = E12[0][v1[0]][v2[0]][ix12] * E12[1][v1[1]][v2[1]][iy12] * E12[2][v1[2]][v2[2]][iz12];
                                    for (int ix34 = 0; ix34 <= l34; ++ix34)
                                        for (int iy34 = 0; iy34 <= m34; ++iy34) {
                                            const double v34 = v12 * E34[0][v3[0]][v4[0]][ix34] * E34[1][v3[1]][v4[1]][iy34];
                                            const double *rzp = R[ix12 + ix34][iy12 + iy34] + iz12;
                                            const double *dzp = E34[2][v3[2]][v4[2]];
                                            int sx = ((ix34 + iy34) % 2) ? -1 : 1;
                                            for (int iz34 = 0; iz34 <= n34; ++iz34) {
                                                sum += sx * v34 * dzp[iz34] * rzp[iz34];
                                                sx = -sx;
                                            }
                                        }
                                }
                        vals[kc] += sum * q->norm[kc];
                    }
                }
            }
        }
    }
}

int oracle_quartet_block(const oracle_basis *b, int ish, int jsh, int ksh, int lsh, double *out) {
    init_tables();
    int sh[4] = {ish, jsh, ksh, lsh}, lv[4], n[4];
    for (int t = 0; t < 4; ++t) {
        lv[t] = b->lv[sh[t]];
        n[t] = oracle_ncart(lv[t]);
    }
    int sw12 = lv[0] < lv[1], sw34 = lv[2] < lv[3]; /* TwoElectronInts.cpp:563,604 */
    static _Thread_local quartet_view q;
    q.s1 = sw12 ? jsh : ish; q.s2 = sw12 ? ish : jsh;
    q.s3 = sw34 ? lsh : ksh; q.s4 = sw34 ? ksh : lsh;
    int knt = 0;
    for (int ia = 0; ia < n[0]; ++ia)
        for (int ib = 0; ib < n[1]; ++ib)
            for (int ic = 0; ic < n[2]; ++ic)
                for (int id = 0; id < n[3]; ++id) {
                    q.comp[knt][0] = sw12 ? ib : ia; q.comp[knt][1] = sw12 ? ia : ib;
                    q.comp[knt][2] = sw34 ? id : ic; q.comp[knt][3] = sw34 ? ic : id;
                    q.norm[knt] = g_cnorm[lv[0]][ia] * g_cnorm[lv[1]][ib] * g_cnorm[lv[2]][ic] * g_cnorm[lv[3]][id];
                    ++knt;
                }
    q.len = knt;
    if (lv[0] + lv[1] + lv[2] + lv[3] > 8) eri_quartet_md(b, &q, out, NULL); /* dispatch: TwoElectronInts.cpp:661-665 */
    else eri_quartet_rys(b, &q, out, NULL);
    return knt;
}

/* shared shell-quartet enumeration of calculate() / directFormGMatrix (TwoElectronInts.cpp:541-653) */
typedef void (*quartet_sink)(void *ctx, const quartet_view *q, const int (*ijkl)[4], const double *vals);

static long walk_quartets(const oracle_basis *b, int start, long sample_mod, long sample_rem, quartet_sink sink,
                          void *ctx, long *ncalc, long *nprimq) {
    init_tables();
    static _Thread_local quartet_view q;
    static _Thread_local int ijkl[ORACLE_MAXFUNC][4];
    static _Thread_local double vals[ORACLE_MAXFUNC];
    long nq = 0, running = 0;
    for (int ish = start; ish < b->nshell; ++ish)
        for (int jsh = 0; jsh <= ish; ++jsh)
            for (int ksh = 0; ksh <= ish; ++ksh)
                for (int lsh = 0; lsh <= ksh; ++lsh) {
                    const int lv1 = b->lv[ish], lv2 = b->lv[jsh], lv3 = b->lv[ksh], lv4 = b->lv[lsh];
                    const int sw12 = lv1 < lv2, sw34 = lv3 < lv4;
                    int knt = 0;
                    /* canonical function filter with the reference's break semantics (:623-653) */
                    for (int ils = 0; ils < oracle_ncart(lv1); ++ils) {
                        int ir = b->off[ish] + ils;
                        for (int jls = 0; jls < oracle_ncart(lv2); ++jls) {
                            int jr = b->off[jsh] + jls;
                            if (jr > ir) break;
                            for (int kls = 0; kls < oracle_ncart(lv3); ++kls) {
                                int kr = b->off[ksh] + kls;
                                if (kr > ir) break;
                                for (int lls = 0; lls < oracle_ncart(lv4); ++lls) {
                                    int lr = b->off[lsh] + lls;
                                    if (lr > kr || (ir == kr && lr > jr)) break;
                                    ijkl[knt][0] = ir; ijkl[knt][1] = jr; ijkl[knt][2] = kr; ijkl[knt][3] = lr;
                                    q.comp[knt][0] = sw12 ? jls : ils; q.comp[knt][1] = sw12 ? ils : jls;
                                    q.comp[knt][2] = sw34 ? lls : kls; q.comp[knt][3] = sw34 ? kls : lls;
                                    q.norm[knt] = g_cnorm[lv1][ils] * g_cnorm[lv2][jls] * g_cnorm[lv3][kls] * g_cnorm[lv4][lls];
                                    ++knt;
                                }
                            }
                        }
                    }
                    if (!knt) continue;
                    if (sample_mod > 1 && (running++ % sample_mod) != sample_rem) continue;
                    q.s1 = sw12 ? jsh : ish; q.s2 = sw12 ? ish : jsh;
                    q.s3 = sw34 ? lsh : ksh; q.s4 = sw34 ? ksh : lsh;
                    q.len = knt;
                    if (ncalc) *ncalc += knt;
                    if (lv1 + lv2 + lv3 + lv4 > 8) eri_quartet_md(b, &q, vals, nprimq); /* :661-665 */
                    else eri_quartet_rys(b, &q, vals, nprimq);
                    sink(ctx, &q, ijkl, vals);
                    ++nq;
                }
    return nq;
}

typedef struct {
    double thresh, *vals;
    int *ijkl;
    long cap, n;
} store_ctx;

static void store_sink(void *vctx, const quartet_view *q, const int (*ijkl)[4], const double *vals) {
    store_ctx *c = (store_ctx *)vctx;
    for (int k = 0; k < q->len; ++k)
        if (fabs(vals[k]) > c->thresh) { /* :667-671 */
            if (c->vals && c->n < c->cap) {
                c->vals[c->n] = vals[k];
                memcpy(c->ijkl + 4 * c->n, ijkl[k], 4 * sizeof(int));
            }
            ++c->n;
        }
}

long oracle_unique_eris(const oracle_basis *b, int start_shell, double thresh, double *vals, int *ijkl, long cap,
                        long *ncalc) {
    store_ctx c = {thresh, vals, ijkl, cap, 0};
    if (ncalc) *ncalc = 0;
    walk_quartets(b, start_shell, 1, 0, store_sink, &c, ncalc, NULL);
    return c.n;
}

/* formGMatrixKernel, TwoElectronInts.cpp:699-747 */
static void digest_rhf(const double *P, double *G, double val, int i, int j, int k, int l) {
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

/* formGMatrixKernel2, TwoElectronInts.cpp:749-820 */
static void digest_uhf(const double *PA, const double *PB, double *GA, double *GB, double val, int i, int j, int k, int l) {
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
    double da = val * (PA[ij] + PB[ij]), db = val * (PA[kl] + PB[kl]);
    double sjlA = val * PA[ik], sjkA = val * PA[il], sikA = val * PA[jl], silA = val * PA[jk];
    double sjlB = val * PB[ik], sjkB = val * PB[il], sikB = val * PB[jl], silB = val * PB[jk];
    if (k != l) {
        db = db + db;
        GA[ik] -= sikA; GB[ik] -= sikB;
        if (i != j && j >= k) { GA[jk] -= sjkA; GB[jk] -= sjkB; }
    }
    GA[il] -= silA; GA[ij] += db;
    GB[il] -= silB; GB[ij] += db;
    if (i != j && j >= l) { GA[jl] -= sjlA; GB[jl] -= sjlB; }
    if (ij != kl) {
        if (i != j) da = da + da;
        if (j <= k) {
            GA[jk] -= sjkA; GB[jk] -= sjkB;
            if (i != j && i <= k) { GA[ik] -= sikA; GB[ik] -= sikB; }
            if (k != l && j <= l) { GA[jl] -= sjlA; GB[jl] -= sjlB; }
        }
        GA[kl] += da; GB[kl] += da;
    }
}

void oracle_form_g_rhf(long n, const double *vals, const int *ijkl, const double *P, double *G) {
    for (long q = 0; q < n; ++q) digest_rhf(P, G, vals[q], ijkl[4 * q], ijkl[4 * q + 1], ijkl[4 * q + 2], ijkl[4 * q + 3]);
}

void oracle_form_g_uhf(long n, const double *vals, const int *ijkl, const double *PA, const double *PB, double *GA,
                       double *GB) {
    for (long q = 0; q < n; ++q)
        digest_uhf(PA, PB, GA, GB, vals[q], ijkl[4 * q], ijkl[4 * q + 1], ijkl[4 * q + 2], ijkl[4 * q + 3]);
}

typedef struct {
    double thresh;
    const double *P;
    double *G;
} direct_ctx;

static void direct_sink(void *vctx, const quartet_view *q, const int (*ijkl)[4], const double *vals) {
    direct_ctx *c = (direct_ctx *)vctx;
    for (int k = 0; k < q->len; ++k)
        if (fabs(vals[k]) > c->thresh) digest_rhf(c->P, c->G, vals[k], ijkl[k][0], ijkl[k][1], ijkl[k][2], ijkl[k][3]);
}

long oracle_direct_g_rhf(const oracle_basis *b, double thresh, const double *P, double *G, long sample_mod,
                         long sample_rem, long *nprimq) {
    direct_ctx c = {thresh, P, G};
    if (nprimq) *nprimq = 0;
    return walk_quartets(b, 0, sample_mod, sample_rem, direct_sink, &c, NULL, nprimq);
}

typedef struct {
    double thresh;
    const double *PA, *PB;
    double *GA, *GB;
} direct_uhf_ctx;

static void direct_uhf_sink(void *vctx, const quartet_view *q, const int (*ijkl)[4], const double *vals) {
    direct_uhf_ctx *c = (direct_uhf_ctx *)vctx;
    for (int k = 0; k < q->len; ++k)
        if (fabs(vals[k]) > c->thresh)
            digest_uhf(c->PA, c->PB, c->GA, c->GB, vals[k], ijkl[k][0], ijkl[k][1], ijkl[k][2], ijkl[k][3]);
}

/* UHF twin of oracle_direct_g_rhf: the calculate() loops (TwoElectronInts.cpp:511-697) feeding formGMatrixKernel2 (:749-820) */
long oracle_direct_g_uhf(const oracle_basis *b, double thresh, const double *PA, const double *PB, double *GA, double *GB,
                         long sample_mod, long sample_rem, long *nprimq) {
    direct_uhf_ctx c = {thresh, PA, PB, GA, GB};
    if (nprimq) *nprimq = 0;
    return walk_quartets(b, 0, sample_mod, sample_rem, direct_uhf_sink, &c, NULL, nprimq);
}

/* A handful of G elements of a LARGE system without storing anything (bench.py parity block, tests):
 *   G_ij = sum_kl P_kl [ 2 (ij|kl) - (ik|jl) ]      (what formGmatrix amounts to for RHF)
 * for nelem basis-function pairs ij[2e], ij[2e+1]; every needed shell-quartet block is evaluated by the routine above
 * (reference primitive cut sr < 1e-12 included) and an integral enters only if |val| > thresh, the reference's storage
 * rule (TwoElectronInts.cpp:513,667-671).  Only shells ksh with ksh % mod == rem are visited (one call per host thread,
 * partial sums added by the caller).  Returns the number of shell-quartet blocks evaluated. */
long oracle_g_elements_rhf(const oracle_basis *b, double thresh, const double *P, int nelem, const int *ij, double *out,
                           int mod, int rem) {
    init_tables();
    static _Thread_local double blk[ORACLE_MAXFUNC];
    long nblk = 0;
    int *shell_of = (int *)malloc(sizeof(int) * b->nbf);
    for (int s = 0; s < b->nshell; ++s)
        for (int c = 0; c < oracle_ncart(b->lv[s]); ++c) shell_of[b->off[s] + c] = s;
    for (int e = 0; e < nelem; ++e) {
        const int i = ij[2 * e], j = ij[2 * e + 1];
        const int I = shell_of[i], J = shell_of[j], ci = i - b->off[I], cj = j - b->off[J];
        const int nI = oracle_ncart(b->lv[I]), nJ = oracle_ncart(b->lv[J]);
        double acc = 0.0;
        for (int K = rem; K < b->nshell; K += mod) {
            const int nK = oracle_ncart(b->lv[K]);
            for (int L = 0; L < b->nshell; ++L) {
                const int nL = oracle_ncart(b->lv[L]);
                /* Coulomb: (I J | K L), element (ci, cj, k, l) */
                oracle_quartet_block(b, I, J, K, L, blk);
                for (int k = 0; k < nK; ++k)
                    for (int l = 0; l < nL; ++l) {
                        const double v = blk[((ci * nJ + cj) * nK + k) * nL + l];
                        if (fabs(v) > thresh) {
                            const int a = b->off[K] + k, c = b->off[L] + l;
                            const int hi = a > c ? a : c, lo = a > c ? c : a;
                            acc += 2.0 * v * P[hi * (hi + 1) / 2 + lo];
                        }
                    }
                /* exchange: (I K | J L), element (ci, k, cj, l) */
                oracle_quartet_block(b, I, K, J, L, blk);
                for (int k = 0; k < nK; ++k)
                    for (int l = 0; l < nL; ++l) {
                        const double v = blk[((ci * nK + k) * nJ + cj) * nL + l];
                        if (fabs(v) > thresh) {
                            const int a = b->off[K] + k, c = b->off[L] + l;
                            const int hi = a > c ? a : c, lo = a > c ? c : a;
                            acc -= v * P[hi * (hi + 1) / 2 + lo];
                        }
                    }
                nblk += 2;
            }
        }
        (void)nI;
        out[e] += acc;
    }
    free(shell_of);
    return nblk;
}

/* Port twin of ref_harness.cc:ref_quartet_batch (bench.py cpu_baseline when oracle/_ref is absent): n shell quartets given as
 * two shell pairs each, evaluated in calculate()'s canonical order, storage threshold 1e-14, optional RHF digestion of the
 * canonical function quartets into G.  Returns the number of integrals above the threshold. */
long oracle_quartet_batch(const oracle_basis *b, long n, const int *shells, const double *P, double *G, int digest) {
    init_tables();
    static _Thread_local double blk[ORACLE_MAXFUNC];
    long stored = 0;
    for (long q = 0; q < n; ++q) {
        int a = shells[4 * q], bb = shells[4 * q + 1], c = shells[4 * q + 2], d = shells[4 * q + 3], t;
        if (a < bb) { t = a; a = bb; bb = t; }
        if (c < d) { t = c; c = d; d = t; }
        if (a < c || (a == c && bb < d)) { t = a; a = c; c = t; t = bb; bb = d; d = t; }
        /* calculate() visits (ish jsh | ish lsh) for BOTH orders of jsh != lsh; the function filter splits the block between them */
        const int nrep = (digest && a == c && bb != d) ? 2 : 1;
        for (int rep = 0; rep < nrep; ++rep) {
        if (rep) { t = bb; bb = d; d = t; }
        const int n1 = oracle_ncart(b->lv[a]), n2 = oracle_ncart(b->lv[bb]), n3 = oracle_ncart(b->lv[c]), n4 = oracle_ncart(b->lv[d]);
        oracle_quartet_block(b, a, bb, c, d, blk);
        for (int ia = 0; ia < n1; ++ia) {
            const int ir = b->off[a] + ia;
            for (int ib = 0; ib < n2; ++ib) {
                const int jr = b->off[bb] + ib;
                if (digest && jr > ir) break;
                for (int ic = 0; ic < n3; ++ic) {
                    const int kr = b->off[c] + ic;
                    if (digest && kr > ir) break;
                    for (int id = 0; id < n4; ++id) {
                        const int lr = b->off[d] + id;
                        if (digest && (lr > kr || (ir == kr && lr > jr))) break;
                        const double v = blk[((ia * n2 + ib) * n3 + ic) * n4 + id];
                        if (fabs(v) > 1e-14) {
                            ++stored;
                            if (digest) digest_rhf(P, G, v, ir, jr, kr, lr);
                        }
                    }
                }
            }
        }
        }   /* rep */
    }
    return stored;
}

/* direct RHF G over the quartets whose largest shell index is >= start_shell: what a TwoElectronInts built with
 * start_shell > 0 contributes (TwoElectronInts.cpp:541; the polarisation-potential scan, RHF.hpp:315,354) */
long oracle_direct_g_rhf_start(const oracle_basis *b, int start_shell, double thresh, const double *P, double *G, long sample_mod,
                               long sample_rem) {
    direct_ctx c = {thresh, P, G};
    return walk_quartets(b, start_shell, sample_mod, sample_rem, direct_sink, &c, NULL, NULL);
}
