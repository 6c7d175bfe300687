/* Plain-C stand-ins for the five BLAS/LAPACK entry points the reference Fortran calls
 * (solvshift/src/solpol2.f:168: ILAENV, DGETRF, DGETRI, DDOT, DGEMM), with the Fortran calling
 * convention, so the transpiled reference in oracle/_ref links without a BLAS.
 * TEST INFRASTRUCTURE ONLY (oracle). Column-major storage throughout. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

double ddot_(int *n, double *x, int *incx, double *y, int *incy)
{
    double s = 0.0;
    for (int i = 0; i < *n; ++i) s += x[(size_t)i * *incx] * y[(size_t)i * *incy];
    return s;
}

int ilaenv_(int *ispec, const char *name, const char *opts, int *n1, int *n2, int *n3, int *n4)
{
    (void)ispec; (void)name; (void)opts; (void)n1; (void)n2; (void)n3; (void)n4;
    return 64;
}

void dgemm_(const char *ta, const char *tb, int *m, int *n, int *k, double *alpha, double *a, int *lda,
            double *b, int *ldb, double *beta, double *c, int *ldc)
{
    const int M = *m, N = *n, K = *k;
    const int TA = (ta[0] == 'T' || ta[0] == 't'), TB = (tb[0] == 'T' || tb[0] == 't');
    for (int j = 0; j < N; ++j) {
        double *cj = c + (size_t)j * *ldc;
        if (*beta == 0.0) memset(cj, 0, sizeof(double) * M);
        else for (int i = 0; i < M; ++i) cj[i] *= *beta;
        for (int l = 0; l < K; ++l) {
            const double bv = *alpha * (TB ? b[j + (size_t)l * *ldb] : b[l + (size_t)j * *ldb]);
            if (bv == 0.0) continue;
            if (!TA) {
                const double *al = a + (size_t)l * *lda;
                for (int i = 0; i < M; ++i) cj[i] += al[i] * bv;
            } else {
                for (int i = 0; i < M; ++i) cj[i] += a[l + (size_t)i * *lda] * bv;
            }
        }
    }
}

/* LU with partial pivoting, LAPACK layout and 1-based pivots */
void dgetrf_(int *m, int *n, double *a, int *lda, int *ipiv, int *info)
{
    const int M = *m, N = *n, L = *lda;
    *info = 0;
    const int mn = M < N ? M : N;
    for (int j = 0; j < mn; ++j) {
        int p = j; double big = fabs(a[j + (size_t)j * L]);
        for (int i = j + 1; i < M; ++i) { double v = fabs(a[i + (size_t)j * L]); if (v > big) { big = v; p = i; } }
        ipiv[j] = p + 1;
        if (big == 0.0) { if (*info == 0) *info = j + 1; continue; }
        if (p != j) for (int c = 0; c < N; ++c) { double t = a[j + (size_t)c * L]; a[j + (size_t)c * L] = a[p + (size_t)c * L]; a[p + (size_t)c * L] = t; }
        const double inv = 1.0 / a[j + (size_t)j * L];
        for (int i = j + 1; i < M; ++i) a[i + (size_t)j * L] *= inv;
        for (int c = j + 1; c < N; ++c) {
            const double f = a[j + (size_t)c * L];
            if (f == 0.0) continue;
            double *ac = a + (size_t)c * L; const double *aj = a + (size_t)j * L;
            for (int i = j + 1; i < M; ++i) ac[i] -= aj[i] * f;
        }
    }
}

/* inverse from the LU factors: solve A X = I column by column */
void dgetri_(int *n, double *a, int *lda, int *ipiv, double *work, int *lwork, int *info)
{
    (void)work; (void)lwork;
    const int N = *n, L = *lda;
    *info = 0;
    for (int j = 0; j < N; ++j) if (a[j + (size_t)j * L] == 0.0) { *info = j + 1; return; }
    double *x = (double *)malloc(sizeof(double) * (size_t)N * N);
    for (int c = 0; c < N; ++c) {
        double *xc = x + (size_t)c * N;
        for (int i = 0; i < N; ++i) xc[i] = 0.0;
        xc[c] = 1.0;
        for (int i = 0; i < N; ++i) { int p = ipiv[i] - 1; if (p != i) { double t = xc[i]; xc[i] = xc[p]; xc[p] = t; } }
        for (int j = 0; j < N; ++j) {          /* L y = P e_c  (unit lower) */
            const double v = xc[j]; if (v == 0.0) continue;
            const double *aj = a + (size_t)j * L;
            for (int i = j + 1; i < N; ++i) xc[i] -= aj[i] * v;
        }
        for (int j = N - 1; j >= 0; --j) {     /* U x = y */
            const double *aj = a + (size_t)j * L;
            xc[j] /= aj[j];
            const double v = xc[j];
            for (int i = 0; i < j; ++i) xc[i] -= aj[i] * v;
        }
    }
    for (int c = 0; c < N; ++c) memcpy(a + (size_t)c * L, x + (size_t)c * N, sizeof(double) * N);
    free(x);
}
