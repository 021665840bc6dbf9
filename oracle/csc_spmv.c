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
