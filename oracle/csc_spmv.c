/* CPU baseline kernels of the oracle (TEST INFRASTRUCTURE / reported baseline only).
 *
 * csc_mul_serial restates what Julia's stdlib does for `mul!(y, A::SparseMatrixCSC, x)` - the
 * per-iteration product a user of the reference runs on the matrix returned by create_A
 * (reference src/model/model.jl:225-246): zero y, then for every column j scatter
 * nzval[k]*x[j] into y[rowval[k]].  Single-threaded, like the stdlib routine.
 * csr_mul_omp is the "all host cores" variant (row-parallel on the transposed storage); it is what
 * bench.py --impl reference times.  Indices are 0-based int64; values are C99 double complex.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void csc_mul_serial(int64_t n, const int64_t *colptr, const int64_t *rowval, const double complex *nzval,
                    const double complex *x, double complex *y) {
    memset(y, 0, (size_t)n * sizeof(double complex));
    for (int64_t j = 0; j < n; ++j) {
        const double complex xj = x[j];
        for (int64_t k = colptr[j]; k < colptr[j + 1]; ++k) y[rowval[k]] += nzval[k] * xj;
    }
}

void csr_mul_omp(int64_t n, const int64_t *rowptr, const int64_t *colval, const double complex *nzval,
                 const double complex *x, double complex *y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double complex acc = 0;
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) acc += nzval[k] * x[colval[k]];
        y[i] = acc;
    }
}

/* Unpreconditioned BiCGSTAB on the CSR copy, all host cores - the stand-in for "the reference's solve" (the reference
 * stops at create_linsys, src/model/model.jl:209-220; its README leaves the solve to the user, README.md:27-33).
 * Runs exactly `iters` iterations without a convergence exit (iterations/s measurement, like fdfd_bench_solve) and
 * returns ||b - A x|| / ||b|| from the recurrence.  work: 6 n complex.  x holds x0 on entry. */
static double complex zdot(int64_t n, const double complex *a, const double complex *b) {   /* conj(a) . b */
    double re = 0, im = 0;
#pragma omp parallel for schedule(static) reduction(+ : re, im)
    for (int64_t i = 0; i < n; ++i) {
        const double complex t = conj(a[i]) * b[i];
        re += creal(t);
        im += cimag(t);
    }
    return re + im * I;
}

double bicgstab_csr_omp(int64_t n, const int64_t *rowptr, const int64_t *colval, const double complex *nzval,
                        const double complex *b, double complex *x, int iters, double complex *work) {
    double complex *r = work, *rh = work + n, *p = work + 2 * n, *v = work + 3 * n, *s = work + 4 * n, *t = work + 5 * n;
    csr_mul_omp(n, rowptr, colval, nzval, x, r);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) { r[i] = b[i] - r[i]; rh[i] = r[i]; p[i] = 0; v[i] = 0; }
    double complex rho = 1, alpha = 1, om = 1;
    const double bn = sqrt(creal(zdot(n, b, b)));
    for (int it = 0; it < iters; ++it) {
        const double complex rho1 = zdot(n, rh, r);
        const double complex beta = (rho1 / rho) * (alpha / om);
        rho = rho1;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - om * v[i]);
        csr_mul_omp(n, rowptr, colval, nzval, p, v);
        alpha = rho / zdot(n, rh, v);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) s[i] = r[i] - alpha * v[i];
        csr_mul_omp(n, rowptr, colval, nzval, s, t);
        om = zdot(n, t, s) / zdot(n, t, t);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) { x[i] += alpha * p[i] + om * s[i]; r[i] = s[i] - om * t[i]; }
    }
    return sqrt(creal(zdot(n, r, r))) / (bn > 0 ? bn : 1.0);
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the all-cores baseline sets its team size explicitly. */
void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
