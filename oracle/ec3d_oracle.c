/*
 * ec3d_oracle.c -- CPU oracle for the EC3D hot path.  TEST INFRASTRUCTURE ONLY (see ec3d_oracle.h).
 * PARITY UNPINNED by reference tests (the reference has none; no Fortran compiler here).
 *
 * Every function cites the reference lines (relative to /root/reference/src) it restates.
 * Loop order, operand order and the absence of FMA/reassociation follow gfortran -O2 on x86-64.
 */
#include "ec3d_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

double orc_dot(const double *a, const double *b, int64_t n);
static int g_dot_mode = 0;   /* see orc_set_dot_mode */

/* ------------------------------------------------------------------------------------------ */
/* gfortran intrinsics                                                                          */
/* ------------------------------------------------------------------------------------------ */

/* NORM2 as gfortran evaluates it (libgfortran norm2_r8 / the inline expansion in
 * trans-intrinsic.c): running scale + scaled sum of squares, one sequential pass. */
double orc_norm2(const double *x, int64_t n)
{
    if (g_dot_mode == 1) return sqrt(orc_dot(x, x, n));
    double result = 0.0, scale = 1.0;
    for (int64_t i = 0; i < n; ++i) {
        if (x[i] != 0.0) {
            double absX = fabs(x[i]);
            if (scale < absX) {
                double val = scale / absX;
                result = 1.0 + result * val * val;
                scale = absX;
            } else {
                double val = absX / scale;
                result += val * val;
            }
        }
    }
    return scale * sqrt(result);
}

/* Summation-order switch, FOR SENSITIVITY EXPERIMENTS ONLY.  0 (default) = the reference's order
 * (sequential).  1 = pairwise (recursive halving) dot products and sqrt(pairwise sum of squares)
 * norms: an equally valid evaluation of the same formulas, used by the tests to measure how much
 * the reference algorithm's output moves under reassociation of its reductions. */
void orc_set_dot_mode(int mode) { g_dot_mode = mode; }

static double pairwise_dot(const double *a, const double *b, int64_t n)
{
    if (n <= 8) {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
        return s;
    }
    int64_t h = n / 2;
    return pairwise_dot(a, b, h) + pairwise_dot(a + h, b + h, n - h);
}

/* DOT_PRODUCT: one sequential accumulator starting at 0. */
double orc_dot(const double *a, const double *b, int64_t n)
{
    if (g_dot_mode == 1) return pairwise_dot(a, b, n);
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* ------------------------------------------------------------------------------------------ */
/* solvers.f90                                                                                  */
/* ------------------------------------------------------------------------------------------ */

/* solvers.f90:54-61  sprsAx(i) = dot_product(valA(i1:i2), V(jcol(i1:i2))) */
void orc_sprsAx(const double *valA, const int32_t *irow, const int32_t *jcol, int32_t n,
                const double *v, double *y)
{
    for (int32_t i = 0; i < n; ++i) {
        int32_t i1 = irow[i], i2 = irow[i + 1] - 1; /* 1-based inclusive */
        double s = 0.0;
        for (int32_t m = i1; m <= i2; ++m) s += valA[m - 1] * v[jcol[m - 1] - 1];
        y[i] = s;
    }
}

/* solvers.f90:3-50 */
int orc_sprsBCGstabWR(const double *valA, const int32_t *irow, const int32_t *jcol, int32_t n,
                      const double *b, double *x, double tolerance, int32_t itmax, int32_t *iter)
{
    double alpha, beta, omega, rr0, rr0_new, Bnorm;
    size_t nb = (size_t)n * sizeof(double);
    double *R = malloc(nb), *R0 = malloc(nb), *P = malloc(nb), *AP = malloc(nb), *S = malloc(nb),
           *AS = malloc(nb);
    if (!R || !R0 || !P || !AP || !S || !AS) {
        free(R); free(R0); free(P); free(AP); free(S); free(AS);
        return -1;
    }
    *iter = 0;                                           /* :13 */
    orc_sprsAx(valA, irow, jcol, n, x, R);               /* :14 */
    for (int32_t j = 0; j < n; ++j) R[j] = b[j] - R[j];  /* :15-17 */
    memcpy(R0, R, nb);                                   /* :18 */
    memcpy(P, R, nb);                                    /* :19 */
    Bnorm = orc_norm2(b, n);                             /* :21 */
    if (Bnorm == 0.0) goto done;                         /* :23 */
    for (;;) {
        if (*iter > itmax) {                             /* :25-28 */
            printf(" %24.16E\n", orc_norm2(R, n));
            break;
        }
        *iter += 1;                                      /* :29 */
        orc_sprsAx(valA, irow, jcol, n, P, AP);          /* :30 */
        rr0 = orc_dot(R, R0, n);                         /* :31 */
        alpha = rr0 / orc_dot(AP, R0, n);                /* :32 */
        for (int32_t j = 0; j < n; ++j) S[j] = R[j] - alpha * AP[j]; /* :33 */
        if (orc_norm2(S, n) / Bnorm < tolerance) {       /* :34 */
            for (int32_t j = 0; j < n; ++j) x[j] = x[j] + alpha * P[j]; /* :36 */
            break;
        }
        orc_sprsAx(valA, irow, jcol, n, S, AS);          /* :39 */
        omega = orc_dot(AS, S, n) / orc_dot(AS, AS, n);  /* :40 */
        for (int32_t j = 0; j < n; ++j) x[j] = x[j] + alpha * P[j] + omega * S[j]; /* :41 */
        for (int32_t j = 0; j < n; ++j) R[j] = S[j] - omega * AS[j];              /* :42 */
        if ((orc_norm2(R, n) / Bnorm) < tolerance) break; /* :43 */
        rr0_new = orc_dot(R, R0, n);                     /* :44 */
        beta = (alpha / omega) * rr0_new / rr0;          /* :45 */
        for (int32_t j = 0; j < n; ++j) P[j] = R[j] + beta * (P[j] - omega * AP[j]); /* :46 */
        if ((fabs(rr0_new) / Bnorm) < tolerance) {       /* :47-49  the "restart" */
            memcpy(R0, R, nb);
            memcpy(P, R, nb);
        }
    }
done:
    free(R); free(R0); free(P); free(AP); free(S); free(AS);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Bridge solver: the SAME algorithm (solvers.f90:3-50) with EXACTLY ROUNDED inner products.      */
/*                                                                                                */
/* The CUDA path cannot reproduce the rounding errors of the reference's sequential DOT_PRODUCT / */
/* NORM2 (a serial recurrence); it sums the products in double-double instead, i.e. it returns    */
/* the correctly rounded value of the sum of its product terms.  This function does the same on   */
/* the CPU, with the same definition of the product terms:                                        */
/*   F(v,w)  dots fused into the stencil SpMV -- one term per x-adjacent cell pair (i even) and   */
/*           plane: fma chain over Ax(a),Ax(b),Ay(a),Ay(b),Az(a),Az(b),U(a),U(b) starting from 0  */
/*   B(v,w)  dots of the vector-update kernels -- one term per x-adjacent cell pair and           */
/*           component: fma(v_b, w_b, v_a*w_a)                                                    */
/* and the same reuse of (R,R0): rr0 of iteration i+1 is rr0_new of iteration i (||R||^2 after a  */
/* restart, ||R||^2 from the initial residual in iteration 1) instead of a recomputed dot product. */
/* Every vector operation and the SpMV are the reference's, unchanged.  The tests use it as the   */
/* bridge: CUDA == bridge bit for bit, and bridge vs orc_sprsBCGstabWR isolates what the          */
/* reference's own summation error does to its iterates.  Needs even sdx (the CUDA main path).    */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double hi, lo; } orc_dd;
static inline void dd_add(orc_dd *a, double b)
{
    double s = a->hi + b;
    double bb = s - a->hi;
    double e = (a->hi - (s - bb)) + (b - bb);
    a->hi = s;
    a->lo += e;
}

typedef struct {
    int32_t sdx, sdy, sdz;
    int64_t nC;
    const int32_t *geoC;
} xd_grid;

static double xd_F(const xd_grid *g, const double *v, const double *w)
{
    orc_dd acc = {0.0, 0.0};
    const int64_t nC = g->nC;
    for (int64_t c0 = 0; c0 < nC; c0 += 2) {            /* sdx even: (c0, c0+1) are x-adjacent, same j,k */
        double t = 0.0;
        for (int comp = 0; comp < 3; ++comp) {
            t = fma(v[comp * nC + c0], w[comp * nC + c0], t);
            t = fma(v[comp * nC + c0 + 1], w[comp * nC + c0 + 1], t);
        }
        const int32_t ga = g->geoC[c0], gb = g->geoC[c0 + 1];
        if (ga) t = fma(v[ga - 1], w[ga - 1], t);
        if (gb) t = fma(v[gb - 1], w[gb - 1], t);
        dd_add(&acc, t);
    }
    return acc.hi + acc.lo;
}

static double xd_B(const xd_grid *g, const double *v, const double *w)
{
    orc_dd acc = {0.0, 0.0};
    const int64_t nC = g->nC;
    for (int comp = 0; comp < 3; ++comp)
        for (int64_t c0 = 0; c0 < nC; c0 += 2)
            dd_add(&acc, fma(v[comp * nC + c0 + 1], w[comp * nC + c0 + 1], v[comp * nC + c0] * w[comp * nC + c0]));
    for (int64_t c0 = 0; c0 < nC; c0 += 2) {
        const int32_t ga = g->geoC[c0], gb = g->geoC[c0 + 1];
        if (!ga && !gb) continue;
        const double va = ga ? v[ga - 1] : 0.0, wa = ga ? w[ga - 1] : 0.0;
        const double vb = gb ? v[gb - 1] : 0.0, wb = gb ? w[gb - 1] : 0.0;
        dd_add(&acc, fma(vb, wb, va * wa));
    }
    return acc.hi + acc.lo;
}

int orc_sprsBCGstabWR_exact_dots(const double *valA, const int32_t *irow, const int32_t *jcol, int32_t n,
                                 const double *b, double *x, double tolerance, int32_t itmax, int32_t *iter,
                                 int32_t sdx, int32_t sdy, int32_t sdz, const int32_t *geoPHYS_C)
{
    if (sdx % 2 != 0) return -2;
    const xd_grid g = {sdx, sdy, sdz, (int64_t)sdx * sdy * sdz, geoPHYS_C};
    double alpha, beta, omega, rr0, rr0_next = 0.0, rr = 0.0;
    size_t nb = (size_t)n * sizeof(double);
    double *R = malloc(nb), *R0 = malloc(nb), *P = malloc(nb), *AP = malloc(nb), *S = malloc(nb),
           *AS = malloc(nb);
    if (!R || !R0 || !P || !AP || !S || !AS) {
        free(R); free(R0); free(P); free(AP); free(S); free(AS);
        return -1;
    }
    *iter = 0;
    orc_sprsAx(valA, irow, jcol, n, x, R);
    for (int32_t j = 0; j < n; ++j) R[j] = b[j] - R[j];
    memcpy(R0, R, nb);
    memcpy(P, R, nb);
    const double bb = xd_F(&g, b, b), rr_init = xd_F(&g, R, R);
    const double Bnorm = sqrt(bb);
    if (bb == 0.0) goto done;
    for (;;) {
        if (*iter > itmax) {
            printf(" %24.16E\n", sqrt(rr));
            break;
        }
        *iter += 1;
        orc_sprsAx(valA, irow, jcol, n, P, AP);
        rr0 = (*iter == 1) ? rr_init : rr0_next;
        alpha = rr0 / xd_F(&g, AP, R0);
        for (int32_t j = 0; j < n; ++j) S[j] = R[j] - alpha * AP[j];
        orc_sprsAx(valA, irow, jcol, n, S, AS);
        if (sqrt(xd_F(&g, S, S)) / Bnorm < tolerance) {
            for (int32_t j = 0; j < n; ++j) x[j] = x[j] + alpha * P[j];
            break;
        }
        omega = xd_F(&g, AS, S) / xd_F(&g, AS, AS);
        for (int32_t j = 0; j < n; ++j) x[j] = x[j] + alpha * P[j] + omega * S[j];
        for (int32_t j = 0; j < n; ++j) R[j] = S[j] - omega * AS[j];
        rr = xd_B(&g, R, R);
        if (sqrt(rr) / Bnorm < tolerance) break;
        const double rr0_new = xd_B(&g, R, R0);
        beta = (alpha / omega) * rr0_new / rr0;
        const int restart = (fabs(rr0_new) / Bnorm) < tolerance;
        rr0_next = restart ? rr : rr0_new;
        if (restart) {
            memcpy(R0, R, nb);
            memcpy(P, R, nb);
        } else {
            for (int32_t j = 0; j < n; ++j) P[j] = R[j] + beta * (P[j] - omega * AP[j]);
        }
    }
done:
    free(R); free(R0); free(P); free(AP); free(S); free(AS);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* utilites.f90:477-508  full_sort(a,b,n,1,1): repeated adjacent-swap passes, ascending column  */
/* ------------------------------------------------------------------------------------------ */
static void full_sort(int32_t *a, double *b, int n)
{
    for (;;) {
        int sorted = 1;
        for (int i = 0; i + 1 < n; ++i) {
            if (a[i] == a[i + 1]) continue;              /* :484-488 (l > m => cycle) */
            if (a[i] > a[i + 1]) {
                int32_t z = a[i]; a[i] = a[i + 1]; a[i + 1] = z;
                double t = b[i]; b[i] = b[i + 1]; b[i + 1] = t;
                sorted = 0;
            }
        }
        if (sorted) break;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* EC3D.f90:465-1049  gen_sparse_matrix                                                         */
/* ------------------------------------------------------------------------------------------ */

int64_t orc_nnz_upper_bound(const orc_grid *g)
{
    int64_t nC = (int64_t)g->sdx * g->sdy * g->sdz;
    return 3 * (7 * nC + 3 * (int64_t)g->nCells0) + 13 * (int64_t)g->nCells0;
}

#define VP(n, c) (g->valPHYS[((n) - 1) * 5 + ((c) - 1)])
#define GC(i, j, k) (g->geoPHYS_C[((i) - 1) + (int64_t)sdx * ((j) - 1) + (int64_t)sdx * sdy * ((k) - 1)])
#define BNDf(a, s) (g->BND[(a) - 1][(s) - 1])

#define SETROW4(c1, c2, c3, c4, v1, v2, v3, v4)                                   \
    do { Lx = 4; colX[0] = c1; colX[1] = c2; colX[2] = c3; colX[3] = c4;          \
         valX[0] = v1; valX[1] = v2; valX[2] = v3; valX[3] = v4; } while (0)
#define SETROW5(c1, c2, c3, c4, c5, v1, v2, v3, v4, v5)                           \
    do { Lx = 5; colX[0] = c1; colX[1] = c2; colX[2] = c3; colX[3] = c4; colX[4] = c5; \
         valX[0] = v1; valX[1] = v2; valX[2] = v3; valX[3] = v4; valX[4] = v5; } while (0)
#define SETROW6(c1, c2, c3, c4, c5, c6, v1, v2, v3, v4, v5, v6)                   \
    do { Lx = 6; colX[0] = c1; colX[1] = c2; colX[2] = c3; colX[3] = c4; colX[4] = c5; colX[5] = c6; \
         valX[0] = v1; valX[1] = v2; valX[2] = v3; valX[3] = v4; valX[4] = v5; valX[5] = v6; } while (0)
#define SETU7(c1, c2, c3, c4, c5, c6, c7, v1, v2, v3, v4, v5, v6, v7)             \
    do { Lfi = 7; colU[0] = c1; colU[1] = c2; colU[2] = c3; colU[3] = c4; colU[4] = c5; \
         colU[5] = c6; colU[6] = c7; valU[0] = v1; valU[1] = v2; valU[2] = v3; valU[3] = v4; \
         valU[4] = v5; valU[5] = v6; valU[6] = v7; } while (0)

int orc_gen_sparse_matrix(const orc_grid *g, orc_csr *out)
{
    const int32_t sdx = g->sdx, sdy = g->sdy, sdz = g->sdz;
    const int64_t nCells64 = (int64_t)sdx * sdy * sdz;
    if (3 * nCells64 + g->nCells0 + 1 > INT32_MAX) return 3;
    const int32_t nCells = (int32_t)nCells64, kdz = sdx * sdy;
    const int32_t nCellsGlob = 3 * nCells + g->nCells0;
    const double dt = g->dt;
    const double *delta = g->delta;

    int32_t colX[10], colY[10], colZ[10], colU[13];
    double valX[10], valY[10], valZ[10], valU[13];
    double s, a, b;

    /* The reference appends every entry to one of four linked lists and copies them out
     * backwards (:991-1029); the net effect is four consecutive CSR blocks in emission order.
     * Here each block is written to scratch storage and concatenated afterwards. */
    const int64_t capA = 7 * nCells64 + 3 * (int64_t)g->nCells0;
    const int64_t capU = 13 * (int64_t)g->nCells0;
    int32_t *jc[4];
    double *va[4];
    int64_t nz[4] = {0, 0, 0, 0};
    for (int q = 0; q < 4; ++q) {
        int64_t cap = (q < 3 ? capA : capU) + 1;
        jc[q] = malloc((size_t)cap * sizeof(int32_t));
        va[q] = malloc((size_t)cap * sizeof(double));
        if (!jc[q] || !va[q]) return -1;
    }
    int32_t *irow = out->irow;
    for (int32_t i = 0; i <= nCellsGlob; ++i) irow[i] = 1;   /* EC3D.f90:112 source=1 */

    int32_t num_bndX = 0, num_bndY = 0, num_bndZ = 0, num_bndUx = 0, num_bndUy = 0, num_bndUz = 0;
    int rc = 0;
    out->err_cell = 0; out->err_col = 0;

    const double sz = 1.0 / (delta[2] * delta[2]);            /* :496-501 */
    const double sy = 1.0 / (delta[1] * delta[1]);
    const double sx = 1.0 / (delta[0] * delta[0]);
    const double dsx = 0.5 / delta[0];
    const double dsy = 0.5 / delta[1];
    const double dsz = 0.5 / delta[2];

    int32_t nn = 0, countU = 0;
    for (int32_t k = 1; k <= sdz && !rc; ++k)
    for (int32_t j = 1; j <= sdy && !rc; ++j)
    for (int32_t i = 1; i <= sdx && !rc; ++i) {
        int32_t n = g->geoPHYS[nn];                           /* :509 */
        nn += 1;
        int Lx = 0, Ly = 0, Lz = 0, Lfi = 0;
        int nAx = 0, nAy = 0, nAz = 0, nFix = 0, nFiy = 0, nFiz = 0, kFi = 0;
        for (int m = 0; m < 10; ++m) { colX[m] = colY[m] = colZ[m] = 0; valX[m] = valY[m] = valZ[m] = 0.0; }
        for (int m = 0; m < 13; ++m) { colU[m] = 0; valU[m] = 0.0; }

        if (GC(i, j, k) != 0) { kFi = 1; countU += 1; }       /* :519-522 */

        const int32_t kim = nn - 1, kjm = nn - sdx, kkm = nn - kdz;   /* :524-525 */
        const int32_t kip = nn + 1, kjp = nn + sdx, kkp = nn + kdz;

        if (i == 1 || j == 1 || k == 1 || i == sdx || j == sdy || k == sdz) {     /* :528 */
            const int xin = (i > 1 && i < sdx), yin = (j > 1 && j < sdy), zin = (k > 1 && k < sdz);
            /* 8 corners :530-561 */
            if (i == 1 && j == 1 && k == 1)
                SETROW4(kip, kjp, kkp, nn, BNDf(1,2)*sx, BNDf(2,2)*sy, BNDf(3,2)*sz, (sx + sy + sz));
            else if (i == sdx && j == 1 && k == 1)
                SETROW4(kim, kjp, kkp, nn, BNDf(1,1)*sx, BNDf(2,2)*sy, BNDf(3,2)*sz, (sx + sy + sz));
            else if (i == 1 && j == sdy && k == 1)
                SETROW4(kip, kjm, kkp, nn, BNDf(1,2)*sx, BNDf(2,1)*sy, BNDf(3,2)*sz, (sx + sy + sz));
            else if (i == sdx && j == sdy && k == 1)
                SETROW4(kim, kjm, kkp, nn, BNDf(1,1)*sx, BNDf(2,1)*sy, BNDf(3,2)*sz, (sx + sy + sz));
            else if (i == 1 && j == 1 && k == sdz)
                SETROW4(kip, kjp, kkm, nn, BNDf(1,2)*sx, BNDf(2,2)*sy, BNDf(3,1)*sz, (sx + sy + sz));
            else if (i == sdx && j == 1 && k == sdz)
                SETROW4(kim, kjp, kkm, nn, BNDf(1,1)*sx, BNDf(2,2)*sy, BNDf(3,1)*sz, (sx + sy + sz));
            else if (i == 1 && j == sdy && k == sdz)
                SETROW4(kip, kjm, kkm, nn, BNDf(1,2)*sx, BNDf(2,1)*sy, BNDf(3,1)*sz, (sx + sy + sz));
            else if (i == sdx && j == sdy && k == sdz)
                SETROW4(kim, kjm, kkm, nn, BNDf(1,1)*sx, BNDf(2,1)*sy, BNDf(3,1)*sz, (sx + sy + sz));
            /* edges along x :565-580 */
            else if (xin && j == 1 && k == 1)
                SETROW5(kim, kip, kjp, kkp, nn, -sx, -sx, BNDf(2,2)*sy, BNDf(3,2)*sz, (2.0*sx + sy + sz));
            else if (xin && j == sdy && k == 1)
                SETROW5(kim, kip, kjm, kkp, nn, -sx, -sx, BNDf(2,1)*sy, BNDf(3,2)*sz, (2.0*sx + sy + sz));
            else if (xin && j == 1 && k == sdz)
                SETROW5(kim, kip, kjp, kkm, nn, -sx, -sx, BNDf(2,2)*sy, BNDf(3,1)*sz, (2.0*sx + sy + sz));
            else if (xin && j == sdy && k == sdz)
                SETROW5(kim, kip, kjm, kkm, nn, -sx, -sx, BNDf(2,1)*sy, BNDf(3,1)*sz, (2.0*sx + sy + sz));
            /* edges along y :583-598 */
            else if (i == 1 && yin && k == 1)
                SETROW5(kip, kjm, kjp, kkp, nn, BNDf(1,2)*sx, -sy, -sy, BNDf(3,2)*sz, (sx + 2.0*sy + sz));
            else if (i == sdx && yin && k == 1)
                SETROW5(kim, kjm, kjp, kkp, nn, BNDf(1,1)*sx, -sy, -sy, BNDf(3,2)*sz, (sx + 2.0*sy + sz));
            else if (i == 1 && yin && k == sdz)
                SETROW5(kip, kjm, kjp, kkm, nn, BNDf(1,2)*sx, -sy, -sy, BNDf(3,1)*sz, (sx + 2.0*sy + sz));
            else if (i == sdx && yin && k == sdz)
                SETROW5(kim, kjm, kjp, kkm, nn, BNDf(1,1)*sx, -sy, -sy, BNDf(3,1)*sz, (sx + 2.0*sy + sz));
            /* edges along z :601-616 */
            else if (i == 1 && j == 1 && zin)
                SETROW5(kip, kjp, kkm, kkp, nn, BNDf(1,2)*sx, BNDf(2,2)*sy, -sz, -sz, (sx + sy + 2.0*sz));
            else if (i == sdx && j == 1 && zin)
                SETROW5(kim, kjp, kkm, kkp, nn, BNDf(1,1)*sx, BNDf(2,2)*sy, -sz, -sz, (sx + sy + 2.0*sz));
            else if (i == 1 && j == sdy && zin)
                SETROW5(kip, kjm, kkm, kkp, nn, BNDf(1,2)*sx, BNDf(2,1)*sy, -sz, -sz, (sx + sy + 2.0*sz));
            else if (i == sdx && j == sdy && zin)
                SETROW5(kim, kjm, kkm, kkp, nn, BNDf(1,1)*sx, BNDf(2,1)*sy, -sz, -sz, (sx + sy + 2.0*sz));
            /* 6 faces :619-642 */
            else if (xin && yin && k == 1)
                SETROW6(kim, kip, kjm, kjp, kkp, nn, -sx, -sx, -sy, -sy, BNDf(3,2)*sz, (2.0*sx + 2.0*sy + sz));
            else if (xin && yin && k == sdz)
                SETROW6(kim, kip, kjm, kjp, kkm, nn, -sx, -sx, -sy, -sy, BNDf(3,1)*sz, (2.0*sx + 2.0*sy + sz));
            else if (i == 1 && yin && zin)
                SETROW6(kip, kjm, kjp, kkm, kkp, nn, BNDf(1,2)*sx, -sy, -sy, -sz, -sz, (sx + 2.0*sy + 2.0*sz));
            else if (i == sdx && yin && zin)
                SETROW6(kim, kjm, kjp, kkm, kkp, nn, BNDf(1,1)*sx, -sy, -sy, -sz, -sz, (sx + 2.0*sy + 2.0*sz));
            else if (xin && j == 1 && zin)
                SETROW6(kim, kip, kjp, kkm, kkp, nn, -sx, -sx, BNDf(2,2)*sy, -sz, -sz, (2.0*sx + sy + 2.0*sz));
            else if (xin && j == sdy && zin)
                SETROW6(kim, kip, kjm, kkm, kkp, nn, -sx, -sx, BNDf(2,1)*sy, -sz, -sz, (2.0*sx + sy + 2.0*sz));

            Ly = Lx; Lz = Lx;                                                /* :645-646 */
            for (int m = 0; m < 10; ++m) {
                colY[m] = nCells + colX[m]; colZ[m] = 2 * nCells + colX[m];
                valY[m] = valX[m]; valZ[m] = valX[m];
            }
        } else {
            Lx = 7;                                                          /* :649-654 */
            colX[0] = kim; colX[1] = kip; colX[2] = kjm; colX[3] = kjp; colX[4] = kkm; colX[5] = kkp; colX[6] = nn;
            valX[0] = -sx; valX[1] = -sx; valX[2] = -sy; valX[3] = -sy; valX[4] = -sz; valX[5] = -sz;
            valX[6] = 2.0 * (sx + sy + sz);
            Ly = Lx; Lz = Lx;
            for (int m = 0; m < 10; ++m) {
                colY[m] = nCells + colX[m]; colZ[m] = 2 * nCells + colX[m];
                valY[m] = valX[m]; valZ[m] = valX[m];
            }
            if (kFi != 0) {                                                  /* :656-711 */
                valX[0] = valX[0] - VP(n,3) / (2.0 * delta[0]);
                valX[1] = valX[1] + VP(n,3) / (2.0 * delta[0]);
                valX[2] = valX[2] - VP(n,4) / (2.0 * delta[1]);
                valX[3] = valX[3] + VP(n,4) / (2.0 * delta[1]);
                valX[4] = valX[4] - VP(n,5) / (2.0 * delta[2]);
                valX[5] = valX[5] + VP(n,5) / (2.0 * delta[2]);
                valX[6] = valX[6] + 2.0 * VP(n,2) / dt;
                for (int m = 0; m < 10; ++m) { valY[m] = valX[m]; valZ[m] = valX[m]; }

                /* geoPHYS_C(i+-2,..) is only evaluated on the branches that need it; the caller
                 * guarantees the conductor is >= 1 cell from the faces so i+-1 is in range, and
                 * a 0 / out-of-range reach shows up as column <= 0 below (reference: STOP). */
#define GCS(ii, jj, kk) (((ii) < 1 || (ii) > sdx || (jj) < 1 || (jj) > sdy || (kk) < 1 || (kk) > sdz) ? 0 : GC(ii, jj, kk))
                if (GCS(i + 1, j, k) == 0) {
                    colX[Lx] = GC(i, j, k);      valX[Lx] = -3.0 * VP(n,2) * dsx; Lx++;
                    colX[Lx] = GCS(i - 1, j, k); valX[Lx] = +4.0 * VP(n,2) * dsx; Lx++;
                    colX[Lx] = GCS(i - 2, j, k); valX[Lx] = -1.0 * VP(n,2) * dsx; Lx++;
                    nAx = 1;
                } else if (GCS(i - 1, j, k) == 0) {
                    colX[Lx] = GC(i, j, k);      valX[Lx] = +3.0 * VP(n,2) * dsx; Lx++;
                    colX[Lx] = GCS(i + 1, j, k); valX[Lx] = -4.0 * VP(n,2) * dsx; Lx++;
                    colX[Lx] = GCS(i + 2, j, k); valX[Lx] = +1.0 * VP(n,2) * dsx; Lx++;
                    nAx = 1;
                } else {
                    colX[Lx] = GCS(i + 1, j, k); valX[Lx] = -VP(n,2) * dsx; Lx++;
                    colX[Lx] = GCS(i - 1, j, k); valX[Lx] = +VP(n,2) * dsx; Lx++;
                }
                if (GCS(i, j + 1, k) == 0) {
                    colY[Ly] = GC(i, j, k);      valY[Ly] = -3.0 * VP(n,2) * dsy; Ly++;
                    colY[Ly] = GCS(i, j - 1, k); valY[Ly] = +4.0 * VP(n,2) * dsy; Ly++;
                    colY[Ly] = GCS(i, j - 2, k); valY[Ly] = -1.0 * VP(n,2) * dsy; Ly++;
                    nAy = 1;
                } else if (GCS(i, j - 1, k) == 0) {
                    colY[Ly] = GC(i, j, k);      valY[Ly] = +3.0 * VP(n,2) * dsy; Ly++;
                    colY[Ly] = GCS(i, j + 1, k); valY[Ly] = -4.0 * VP(n,2) * dsy; Ly++;
                    colY[Ly] = GCS(i, j + 2, k); valY[Ly] = +1.0 * VP(n,2) * dsy; Ly++;
                    nAy = 1;
                } else {
                    colY[Ly] = GCS(i, j + 1, k); valY[Ly] = -VP(n,2) * dsy; Ly++;
                    colY[Ly] = GCS(i, j - 1, k); valY[Ly] = +VP(n,2) * dsy; Ly++;
                }
                if (GCS(i, j, k + 1) == 0) {
                    colZ[Lz] = GC(i, j, k);      valZ[Lz] = -3.0 * VP(n,2) * dsz; Lz++;
                    colZ[Lz] = GCS(i, j, k - 1); valZ[Lz] = +4.0 * VP(n,2) * dsz; Lz++;
                    colZ[Lz] = GCS(i, j, k - 2); valZ[Lz] = -1.0 * VP(n,2) * dsz; Lz++;
                    nAz = 1;
                } else if (GCS(i, j, k - 1) == 0) {
                    colZ[Lz] = GC(i, j, k);      valZ[Lz] = +3.0 * VP(n,2) * dsz; Lz++;
                    colZ[Lz] = GCS(i, j, k + 1); valZ[Lz] = -4.0 * VP(n,2) * dsz; Lz++;
                    colZ[Lz] = GCS(i, j, k + 2); valZ[Lz] = +1.0 * VP(n,2) * dsz; Lz++;
                    nAz = 1;
                } else {
                    colZ[Lz] = GCS(i, j, k + 1); valZ[Lz] = -VP(n,2) * dsz; Lz++;
                    colZ[Lz] = GCS(i, j, k - 1); valZ[Lz] = +VP(n,2) * dsz; Lz++;
                }
            }
        }
        /* ---- emit X, Y, Z rows :715-756 ---- */
        full_sort(colX, valX, Lx);
        for (int m = 0; m < Lx; ++m) {
            if (colX[m] <= 0) { rc = 1; out->err_cell = nn; out->err_col = colX[m]; break; }
            jc[0][nz[0]] = colX[m]; va[0][nz[0]] = valX[m]; nz[0]++;
        }
        if (rc) break;
        irow[nn] = irow[nn - 1] + Lx;                                       /* irow(nn+1) */
        full_sort(colY, valY, Ly);
        for (int m = 0; m < Ly; ++m) {
            if (colY[m] <= 0) { rc = 1; out->err_cell = nn; out->err_col = colY[m]; break; }
            jc[1][nz[1]] = colY[m]; va[1][nz[1]] = valY[m]; nz[1]++;
        }
        if (rc) break;
        irow[nCells + nn] = irow[nCells + nn - 1] + Ly;
        full_sort(colZ, valZ, Lz);
        for (int m = 0; m < Lz; ++m) {
            if (colZ[m] <= 0) { rc = 1; out->err_cell = nn; out->err_col = colZ[m]; break; }
            jc[2][nz[2]] = colZ[m]; va[2][nz[2]] = valZ[m]; nz[2]++;
        }
        if (rc) break;
        irow[2 * nCells + nn] = irow[2 * nCells + nn - 1] + Lz;

        if (g->size_PHYS_C != 0) {                                          /* :758-760 */
            if (nAx == 1) out->cel_bndX[num_bndX++] = nn;
            if (nAy == 1) out->cel_bndY[num_bndY++] = nn + nCells;
            if (nAz == 1) out->cel_bndZ[num_bndZ++] = nn + 2 * nCells;
        }

        /* ---- U row :766-957 ---- */
        if (kFi != 0) {
            const int32_t nc = GC(i, j, k);
            const int32_t nim = GCS(i - 1, j, k), nip = GCS(i + 1, j, k);
            const int32_t njm = GCS(i, j - 1, k), njp = GCS(i, j + 1, k);
            const int32_t nkm = GCS(i, j, k - 1), nkp = GCS(i, j, k + 1);
            const int32_t nY = nCells + nn, nZ = 2 * nCells + nn;
            s = 2.0 * (sx + sy + sz);
            /* 8 corners :773-812 */
            if (nim == 0 && njm == 0 && nkm == 0) {
                a = -2.0 / (dt * delta[0]); b = -2.0 / (dt * delta[1]);
                SETU7(nip, njp, nkp, nc, nn, nY, nZ, -2.0*sx, -2.0*sy, -2.0*sz, s, a, b, -2.0 / (dt * delta[2]));
                nFix = nFiy = nFiz = 1;
            } else if (nip == 0 && njm == 0 && nkm == 0) {
                a = +2.0 / (dt * delta[0]); b = -2.0 / (dt * delta[1]);
                SETU7(nim, njp, nkp, nc, nn, nY, nZ, -2.0*sx, -2.0*sy, -2.0*sz, s, a, b, -2.0 / (dt * delta[2]));
                nFix = nFiy = nFiz = 1;
            } else if (nim == 0 && njp == 0 && nkm == 0) {
                a = -2.0 / (dt * delta[0]); b = +2.0 / (dt * delta[1]);
                SETU7(nip, njm, nkp, nc, nn, nY, nZ, -2.0*sx, -2.0*sy, -2.0*sz, s, a, b, -2.0 / (dt * delta[2]));
                nFix = nFiy = nFiz = 1;
            } else if (nip == 0 && njp == 0 && nkm == 0) {
                a = +2.0 / (dt * delta[0]); b = +2.0 / (dt * delta[1]);
                SETU7(nim, njm, nkp, nc, nn, nY, nZ, -2.0*sx, -2.0*sy, -2.0*sz, s, a, b, -2.0 / (dt * delta[2]));
                nFix = nFiy = nFiz = 1;
            } else if (nim == 0 && njm == 0 && nkp == 0) {
                a = -2.0 / (dt * delta[0]); b = -2.0 / (dt * delta[1]);
                SETU7(nip, njp, nkm, nc, nn, nY, nZ, -2.0*sx, -2.0*sy, -2.0*sz, s, a, b, +2.0 / (dt * delta[2]));
                nFix = nFiy = nFiz = 1;
            } else if (nip == 0 && njm == 0 && nkp == 0) {
                a = +2.0 / (dt * delta[0]); b = -2.0 / (dt * delta[1]);
                SETU7(nim, njp, nkm, nc, nn, nY, nZ, -2.0*sx, -2.0*sy, -2.0*sz, s, a, b, +2.0 / (dt * delta[2]));
                nFix = nFiy = nFiz = 1;
            } else if (nim == 0 && njp == 0 && nkp == 0) {
                /* :803-807 -- signs of a and b are as written in the reference (anomaly B1) */
                a = +2.0 / (dt * delta[0]); b = -2.0 / (dt * delta[1]);
                SETU7(nip, njm, nkm, nc, nn, nY, nZ, -2.0*sx, -2.0*sy, -2.0*sz, s, a, b, +2.0 / (dt * delta[2]));
                nFix = nFiy = nFiz = 1;
            } else if (nip == 0 && njp == 0 && nkp == 0) {
                a = +2.0 / (dt * delta[0]); b = +2.0 / (dt * delta[1]);
                SETU7(nim, njm, nkm, nc, nn, nY, nZ, -2.0*sx, -2.0*sy, -2.0*sz, s, a, b, +2.0 / (dt * delta[2]));
                nFix = nFiy = nFiz = 1;
            }
            /* edges along X :815-834 */
            else if (njp == 0 && nkm == 0) {
                a = +2.0 / (dt * delta[1]); b = -2.0 / (dt * delta[2]);
                SETU7(nip, nim, njm, nkp, nc, nY, nZ, -1.0*sx, -sx, -2.0*sy, -2.0*sz, s, a, b);
                nFiy = 1; nFiz = 1;
            } else if (njm == 0 && nkm == 0) {
                a = -2.0 / (dt * delta[1]); b = -2.0 / (dt * delta[2]);
                SETU7(nip, nim, njp, nkp, nc, nY, nZ, -sx, -sx, -2.0*sy, -2.0*sz, s, a, b);
                nFiy = 1; nFiz = 1;
            } else if (njp == 0 && nkp == 0) {
                a = +2.0 / (dt * delta[1]); b = +2.0 / (dt * delta[2]);
                SETU7(nip, nim, njm, nkm, nc, nY, nZ, -sx, -sx, -2.0*sy, -2.0*sz, s, a, b);
                nFiy = 1; nFiz = 1;
            } else if (njm == 0 && nkp == 0) {
                a = -2.0 / (dt * delta[1]); b = +2.0 / (dt * delta[2]);
                SETU7(nip, nim, njp, nkm, nc, nY, nZ, -sx, -sx, -2.0*sy, -2.0*sz, s, a, b);
                nFiy = 1; nFiz = 1;
            }
            /* edges along Y :837-856 */
            else if (nip == 0 && nkm == 0) {
                a = +2.0 / (dt * delta[0]); b = -2.0 / (dt * delta[2]);
                SETU7(nim, njm, njp, nkp, nc, nn, nZ, -2.0*sx, -sy, -sy, -2.0*sz, s, a, b);
                nFix = 1; nFiz = 1;
            } else if (nim == 0 && nkm == 0) {
                a = -2.0 / (dt * delta[0]); b = -2.0 / (dt * delta[2]);
                SETU7(nip, njm, njp, nkp, nc, nn, nZ, -2.0*sx, -sy, -sy, -2.0*sz, s, a, b);
                nFix = 1; nFiz = 1;
            } else if (nip == 0 && nkp == 0) {
                a = +2.0 / (dt * delta[0]); b = +2.0 / (dt * delta[2]);
                SETU7(nim, njm, njp, nkm, nc, nn, nZ, -2.0*sx, -sy, -sy, -2.0*sz, s, a, b);
                nFix = 1; nFiz = 1;
            } else if (nim == 0 && nkp == 0) {
                a = -2.0 / (dt * delta[0]); b = +2.0 / (dt * delta[2]);
                SETU7(nip, njm, njp, nkm, nc, nn, nZ, -2.0*sx, -sy, -sy, -2.0*sz, s, a, b);
                nFix = 1; nFiz = 1;
            }
            /* edges along Z :859-878 */
            else if (nim == 0 && njm == 0) {
                a = -2.0 / (dt * delta[0]); b = -2.0 / (dt * delta[1]);
                SETU7(nip, njp, nkp, nkm, nc, nn, nY, -2.0*sx, -2.0*sy, -sz, -sz, s, a, b);
                nFix = 1; nFiy = 1;
            } else if (nip == 0 && njm == 0) {
                a = +2.0 / (dt * delta[0]); b = -2.0 / (dt * delta[1]);
                SETU7(nim, njp, nkp, nkm, nc, nn, nY, -2.0*sx, -2.0*sy, -sz, -sz, s, a, b);
                nFix = 1; nFiy = 1;
            } else if (nim == 0 && njp == 0) {
                a = -2.0 / (dt * delta[0]); b = +2.0 / (dt * delta[1]);
                SETU7(nip, njm, nkp, nkm, nc, nn, nY, -2.0*sx, -2.0*sy, -sz, -sz, s, a, b);
                nFix = 1; nFiy = 1;
            } else if (nip == 0 && njp == 0) {
                a = +2.0 / (dt * delta[0]); b = +2.0 / (dt * delta[1]);
                SETU7(nim, njm, nkm, nkp, nc, nn, nY, -2.0*sx, -2.0*sy, -sz, -sz, s, a, b);
                nFix = 1; nFiy = 1;
            }
            /* 6 faces :881-916 */
            else if (nim == 0 && njp != 0 && njm != 0 && nkp != 0 && nkm != 0) {
                a = -2.0 / (dt * delta[0]);
                SETU7(nip, njm, njp, nkm, nkp, nc, nn, -2.0*sx, -sy, -sy, -sz, -sz, s, a);
                nFix = 1;
            } else if (nip == 0 && njp != 0 && njm != 0 && nkp != 0 && nkm != 0) {
                a = +2.0 / (dt * delta[0]);
                SETU7(nim, njm, njp, nkm, nkp, nc, nn, -2.0*sx, -sy, -sy, -sz, -sz, s, a);
                nFix = 1;
            } else if (njp == 0 && nip != 0 && nim != 0 && nkp != 0 && nkm != 0) {
                a = +2.0 / (dt * delta[1]);
                SETU7(nim, nip, njm, nkm, nkp, nc, nY, -sx, -sx, -2.0*sy, -sz, -sz, s, a);
                nFiy = 1;
            } else if (njm == 0 && nip != 0 && nim != 0 && nkp != 0 && nkm != 0) {
                a = -2.0 / (dt * delta[1]);
                SETU7(nim, nip, njp, nkm, nkp, nc, nY, -sx, -sx, -2.0*sy, -sz, -sz, s, a);
                nFiy = 1;
            } else if (nkp == 0 && nip != 0 && nim != 0 && njp != 0 && njm != 0) {
                a = +2.0 / (dt * delta[2]);
                SETU7(nim, nip, njm, njp, nkm, nc, nZ, -sx, -sx, -sy, -sy, -2.0*sz, s, a);
                nFiz = 1;
            } else if (nkm == 0 && nip != 0 && nim != 0 && njp != 0 && njm != 0) {
                a = -2.0 / (dt * delta[2]);
                SETU7(nim, nip, njm, njp, nkp, nc, nZ, -sx, -sx, -sy, -sy, -2.0*sz, s, a);
                nFiz = 1;
            } else {                                                        /* :917-922 */
                Lfi = 13;
                colU[0] = nim; colU[1] = nip; colU[2] = njm; colU[3] = njp; colU[4] = nkm; colU[5] = nkp;
                colU[6] = nc; colU[7] = kip; colU[8] = kim; colU[9] = nCells + kjp; colU[10] = nCells + kjm;
                colU[11] = 2 * nCells + kkp; colU[12] = 2 * nCells + kkm;
                valU[0] = -sx; valU[1] = -sx; valU[2] = -sy; valU[3] = -sy; valU[4] = -sz; valU[5] = -sz; valU[6] = s;
                const double hd = 0.5 / dt;
                valU[7]  = hd * (-1.0 / delta[0]); valU[8]  = hd * (1.0 / delta[0]);
                valU[9]  = hd * (-1.0 / delta[1]); valU[10] = hd * (1.0 / delta[1]);
                valU[11] = hd * (-1.0 / delta[2]); valU[12] = hd * (1.0 / delta[2]);
            }
            /* duplicate-column check :924-936 */
            int32_t k0 = 0;
            for (int k1 = 0; k1 < Lfi - 1 && !k0; ++k1)
                for (int k2 = k1 + 1; k2 < Lfi; ++k2)
                    if (colU[k1] == colU[k2]) { k0 = colU[k2]; break; }
            if (k0 != 0) { rc = 2; out->err_cell = nn; out->err_col = k0; break; }

            if (nFix == 1) out->cel_bndUx[num_bndUx++] = nc;                 /* :938-940 */
            if (nFiy == 1) out->cel_bndUy[num_bndUy++] = nc;
            if (nFiz == 1) out->cel_bndUz[num_bndUz++] = nc;

            full_sort(colU, valU, Lfi);                                     /* :942 */
            for (int m = 0; m < Lfi; ++m) {
                if (colU[m] <= 0) { rc = 1; out->err_cell = nn; out->err_col = colU[m]; break; }
                jc[3][nz[3]] = colU[m]; va[3][nz[3]] = valU[m]; nz[3]++;
            }
            if (rc) break;
            irow[3 * nCells + countU] = irow[3 * nCells + countU - 1] + Lfi; /* :955 */
        }
    }

    if (!rc) {
        int64_t tot = nz[0] + nz[1] + nz[2] + nz[3];
        if (tot + 1 > INT32_MAX) rc = 3;
    }
    if (!rc) {
        /* irow block offsets :973-986 */
        int32_t m;
        irow[nCells] = (int32_t)nz[0] + 1;
        for (int32_t i = nCells + 2; i <= 2 * nCells; ++i) irow[i - 1] += (int32_t)nz[0];
        m = (int32_t)(nz[0] + nz[1]);
        irow[2 * nCells] = m + 1;
        for (int32_t i = 2 * nCells + 2; i <= 3 * nCells; ++i) irow[i - 1] += m;
        m = (int32_t)(nz[0] + nz[1] + nz[2]);
        irow[3 * nCells] = m + 1;
        for (int32_t i = 3 * nCells + 2; i <= nCellsGlob + 1; ++i) irow[i - 1] += m;
        /* lists -> jcol/valA :991-1029 */
        int64_t off = 0;
        for (int q = 0; q < 4; ++q) {
            memcpy(out->jcol + off, jc[q], (size_t)nz[q] * sizeof(int32_t));
            memcpy(out->valA + off, va[q], (size_t)nz[q] * sizeof(double));
            off += nz[q];
        }
    }
    out->num_nzX = nz[0]; out->num_nzY = nz[1]; out->num_nzZ = nz[2]; out->num_nzU = nz[3];
    out->num_nz = nz[0] + nz[1] + nz[2] + nz[3];
    out->num_bndX = num_bndX; out->num_bndY = num_bndY; out->num_bndZ = num_bndZ;
    out->num_bndUx = num_bndUx; out->num_bndUy = num_bndUy; out->num_bndUz = num_bndUz;
    for (int q = 0; q < 4; ++q) { free(jc[q]); free(va[q]); }
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* EC3D.f90:157-186  motion preparation                                                         */
/* ------------------------------------------------------------------------------------------ */
void orc_motion_prepare(orc_sources *s, const double delta[3], double dt)
{
    s->flag_move = 0;
    for (int32_t i = 0; i < s->numfun; ++i) {
        for (int a = 0; a < 3; ++a) s->Distance[3 * i + a] = 0.0;
        for (int a = 0; a < 3; ++a) {
            if (s->num_Vmech[3 * i + a] == 0 && s->move[3 * i + a] != 0) {
                s->shift[3 * i + a] = s->vel_Vmech[3 * i + a] * dt / delta[a];
                s->flag_move = 1;
            } else if (s->move[3 * i + a] != 0) {
                s->flag_move = 1;
            }
        }
    }
    s->movestop[0] = s->movestop[1] = s->movestop[2] = 1;                   /* :238 */
}

/* EC3D.f90:1052-1062 */
static void motion_calc(orc_sources *s, int32_t n, const double *vmech_vely, const double delta[3],
                        double dt)
{
    for (int i = 0; i < 3; ++i) {
        if (s->num_Vmech[3 * n + i] == 0) {
            s->Distance[3 * n + i] = s->Distance[3 * n + i] + s->movestop[0] * s->shift[3 * n + i];
            s->length[3 * n + i] = (int32_t)lround(s->Distance[3 * n + i]);
        } else {
            s->Distance[3 * n + i] = s->Distance[3 * n + i] +
                                     vmech_vely[s->num_Vmech[3 * n + i] - 1] * dt / delta[i];
            s->length[3 * n + i] = (int32_t)lround(s->Distance[3 * n + i]);
        }
    }
}

/* EC3D.f90:1064-1114  (default REAL = single precision in the two ceiling() expressions) */
static int32_t new_m(const orc_grid *g, orc_sources *s, int32_t n, int32_t m)
{
    const int32_t sdx = g->sdx, sdy = g->sdy, sdz = g->sdz;
    int32_t L = (int32_t)ceilf((float)m / ((float)(sdx * sdy)));
    int32_t Lnew = L + s->length[3 * n + 2];
    if (Lnew > sdz - 2) { s->movestop[2] = 0; Lnew = sdz - 2; }
    else if (Lnew < 2) { s->movestop[2] = 0; Lnew = 2; }
    else if (s->movestop[2] == 0 && (Lnew < sdz - 2 || Lnew > 2)) s->movestop[2] = 1;
    int32_t nij = (L == 1) ? m : m - (L - 1) * sdx * sdy;
    int32_t j = (int32_t)ceilf((float)nij / (float)sdx);
    int32_t jnew = j + s->length[3 * n + 1];
    if (jnew > sdy - 2) { s->movestop[1] = 0; jnew = sdy - 2; }
    else if (jnew < 2) { s->movestop[1] = 0; jnew = 2; }
    else if (s->movestop[1] == 0 && (jnew < sdy - 2 || jnew > 2)) s->movestop[1] = 1;
    int32_t i = nij - (j - 1) * sdx;
    int32_t inew = i + s->length[3 * n + 0];
    if (inew > sdx - 2) { s->movestop[0] = 0; inew = sdx - 2; }
    else if (inew < 2) { s->movestop[0] = 0; inew = 2; }
    else if (s->movestop[0] == 0 && (inew < sdx - 2 || inew > 2)) s->movestop[0] = 1;
    return inew + sdx * (jnew - 1) + sdx * sdy * (Lnew - 1);
}

/* EC3D.f90:275-367 */
int orc_scatter_sources(const orc_grid *g, orc_sources *s, const orc_conductors *c,
                        const double *fun_vely, const double *vmech_vely, double *Jaf,
                        double *Jafbuf, int32_t *new_nodes)
{
    const int32_t nCells = g->sdx * g->sdy * g->sdz;
    const int64_t nGlob = 3 * (int64_t)nCells + g->nCells0;
    int64_t cnt = 0;
    if (s->flag_move == 1) {
        if (c->size_PHYS_C != 0) {                                           /* :277-292 */
            memset(Jafbuf, 0, (size_t)nGlob * sizeof(double));
            for (int32_t m = 0; m < c->size_PHYS_C; ++m)
                for (int32_t n = c->nod_ptr[m]; n < c->nod_ptr[m + 1]; ++n) {
                    int32_t L = c->nod[n], k = L + nCells, nl = L + 2 * nCells;
                    Jafbuf[L - 1] = Jaf[L - 1];
                    Jafbuf[k - 1] = Jaf[k - 1];
                    Jafbuf[nl - 1] = Jaf[nl - 1];
                }
            memcpy(Jaf, Jafbuf, (size_t)nGlob * sizeof(double));
        } else {
            memset(Jaf, 0, (size_t)nGlob * sizeof(double));                 /* :294 */
        }
        for (int32_t n = 0; n < s->numfun; ++n) {                           /* :301-339 */
            motion_calc(s, n, vmech_vely, g->delta, g->dt);
            double a = fun_vely[n];
            int32_t off = (s->ex[n] == 'X') ? 0 : (s->ex[n] == 'Y') ? nCells : (s->ex[n] == 'Z') ? 2 * nCells : -1;
            if (off < 0) return 1;                                          /* STOP :337 */
            for (int32_t k = s->nod_ptr[n]; k < s->nod_ptr[n + 1]; ++k) {
                int32_t m = s->nods[k] - off;
                m = new_m(g, s, n, m);
                if (new_nodes) new_nodes[cnt] = m;
                cnt++;
                Jaf[m + off - 1] = a;
            }
        }
    } else {
        for (int32_t n = 0; n < s->numfun; ++n) {                           /* :345-365 */
            double a = fun_vely[n];
            if (s->ex[n] != 'X' && s->ex[n] != 'Y' && s->ex[n] != 'Z') return 1; /* STOP :363 */
            for (int32_t k = s->nod_ptr[n]; k < s->nod_ptr[n + 1]; ++k) Jaf[s->nods[k] - 1] = a;
        }
    }
    return 0;
}

/* EC3D.f90:370-404 */
void orc_rhs_pre(const orc_grid *g, const orc_conductors *c, const orc_csr *A, const double *Uaf,
                 double *Jaf)
{
    const int32_t nCells = g->sdx * g->sdy * g->sdz;
    if (c->size_PHYS_C == 0) return;
    for (int32_t m = 0; m < c->size_PHYS_C; ++m) {
        double a = c->valdom[m];
        for (int32_t q = c->nod_ptr[m]; q < c->nod_ptr[m + 1]; ++q) {
            int32_t n = q - c->nod_ptr[m] + 1;                               /* index within domain m */
            int32_t L = c->nod[q], k = L + nCells, nl = L + 2 * nCells;
            Jaf[L - 1] = a * Uaf[L - 1] + Jaf[L - 1];
            Jaf[k - 1] = a * Uaf[k - 1] + Jaf[k - 1];
            Jaf[nl - 1] = a * Uaf[nl - 1] + Jaf[nl - 1];
            double s = 0.0;
            for (int32_t i = A->irow[3 * nCells + n - 1]; i <= A->irow[3 * nCells + n] - 1; ++i) {
                int32_t kk = A->jcol[i - 1];
                if (kk < 3 * nCells + 1) s = s + A->valA[i - 1] * Uaf[kk - 1];
            }
            Jaf[3 * nCells + n - 1] = s;
        }
    }
    for (int32_t i = 0; i < A->num_bndUx; ++i) Jaf[A->cel_bndUx[i] - 1] = 0.0;  /* :396-402 */
    for (int32_t i = 0; i < A->num_bndUy; ++i) Jaf[A->cel_bndUy[i] - 1] = 0.0;
    for (int32_t i = 0; i < A->num_bndUz; ++i) Jaf[A->cel_bndUz[i] - 1] = 0.0;
    for (int32_t i = 0; i < A->num_bndX; ++i) Jaf[A->cel_bndX[i] - 1] = 0.0;
    for (int32_t i = 0; i < A->num_bndY; ++i) Jaf[A->cel_bndY[i] - 1] = 0.0;
    for (int32_t i = 0; i < A->num_bndZ; ++i) Jaf[A->cel_bndZ[i] - 1] = 0.0;
}

/* EC3D.f90:412-433 */
void orc_rhs_post(const orc_grid *g, const orc_conductors *c, const orc_csr *A, double *Uaf,
                  double *Jaf)
{
    const int32_t nCells = g->sdx * g->sdy * g->sdz;
    if (c->size_PHYS_C == 0) return;
    for (int32_t m = 0; m < c->size_PHYS_C; ++m) {
        double a = c->valdom[m];
        for (int32_t q = c->nod_ptr[m]; q < c->nod_ptr[m + 1]; ++q) {
            int32_t L = c->nod[q], k = L + nCells, nl = L + 2 * nCells;
            Jaf[L - 1] = a * Uaf[L - 1] - Jaf[L - 1];
            Jaf[k - 1] = a * Uaf[k - 1] - Jaf[k - 1];
            Jaf[nl - 1] = a * Uaf[nl - 1] - Jaf[nl - 1];
        }
    }
    for (int32_t i = 0; i < A->num_bndX; ++i) Jaf[A->cel_bndX[i] - 1] = 0.0;
    for (int32_t i = 0; i < A->num_bndY; ++i) Jaf[A->cel_bndY[i] - 1] = 0.0;
    for (int32_t i = 0; i < A->num_bndZ; ++i) Jaf[A->cel_bndZ[i] - 1] = 0.0;
    for (int32_t i = 0; i < A->num_bndX; ++i) Uaf[A->cel_bndX[i] - 1] = 0.0;
    for (int32_t i = 0; i < A->num_bndY; ++i) Uaf[A->cel_bndY[i] - 1] = 0.0;
    for (int32_t i = 0; i < A->num_bndZ; ++i) Uaf[A->cel_bndZ[i] - 1] = 0.0;
}

/* ------------------------------------------------------------------------------------------ */
/* output post-processing, utilites.f90:222-290                                                 */
/* ------------------------------------------------------------------------------------------ */
void orc_vtk_fields(int32_t sdx, int32_t sdy, int32_t sdz, const double delta[3], const double *Uaf,
                    const double *Jaf, const int32_t *geoPHYS_C, int32_t size_PHYS_C, float *fieldA,
                    float *eddy, float *source, float *fieldB)
{
    const int64_t kdz = (int64_t)sdx * sdy, nCells = kdz * sdz;
    const double s = -0.07957747154594766788444e7;                /* utilites.f90:239 */
    int64_t m = 0;                                                /* 1-based running cell number */
    for (int k = 1; k <= sdz; ++k)
        for (int j = 1; j <= sdy; ++j)
            for (int i = 1; i <= sdx; ++i) {
                m = m + 1;
                const int64_t o = 3 * (m - 1);
                const int n = geoPHYS_C[m - 1];
                if (fieldA) {                                     /* :222-233 */
                    fieldA[o] = (float)Uaf[m - 1];
                    fieldA[o + 1] = (float)Uaf[nCells + m - 1];
                    fieldA[o + 2] = (float)Uaf[2 * nCells + m - 1];
                }
                if (eddy) {                                       /* :237-250 */
                    if (size_PHYS_C != 0 && n != 0) {
                        eddy[o] = (float)(s * Jaf[m - 1]);
                        eddy[o + 1] = (float)(s * Jaf[nCells + m - 1]);
                        eddy[o + 2] = (float)(s * Jaf[2 * nCells + m - 1]);
                    } else {
                        eddy[o] = eddy[o + 1] = eddy[o + 2] = 0.0f;
                    }
                }
                if (source) {                                     /* :253-274 */
                    if (size_PHYS_C == 0 || n == 0) {
                        source[o] = (float)Jaf[m - 1];
                        source[o + 1] = (float)Jaf[nCells + m - 1];
                        source[o + 2] = (float)Jaf[2 * nCells + m - 1];
                    } else {
                        source[o] = source[o + 1] = source[o + 2] = 0.0f;
                    }
                }
                if (fieldB) {                                     /* :276-290 */
                    int64_t nim = m - 1, njm = m - sdx, nkm = m - kdz, nip = m + 1, njp = m + sdx, nkp = m + kdz;
                    if (i == 1) nim = m;
                    if (i == sdx) nip = m;
                    if (j == 1) njm = m;
                    if (j == sdy) njp = m;
                    if (k == 1) nkm = m;
                    if (k == sdz) nkp = m;
                    const double *A0 = Uaf - 1, *A1 = Uaf + nCells - 1, *A2 = Uaf + 2 * nCells - 1;   /* 1-based views */
                    const double sxm = 0.5 * (A2[njp] - A2[njm]) / delta[1] - 0.5 * (A1[nkp] - A1[nkm]) / delta[2];
                    const double sym = 0.5 * (A0[nkp] - A0[nkm]) / delta[2] - 0.5 * (A2[nip] - A2[nim]) / delta[0];
                    const double szm = 0.5 * (A1[nip] - A1[nim]) / delta[0] - 0.5 * (A0[njp] - A0[njm]) / delta[1];
                    fieldB[o] = (float)sxm;
                    fieldB[o + 1] = (float)sym;
                    fieldB[o + 2] = (float)szm;
                }
            }
}
