/* Shim so the reference's `#include <lapacke.h>` resolves to the LP64 OpenBLAS/LAPACKE that
 * ships inside SciPy (symbols carry a `scipy_` prefix). TEST INFRASTRUCTURE ONLY: used solely to
 * compile the unmodified reference sources into oracle/_ref/. Declares only what the reference calls
 * (parallel-final/lib/eigen.cu:12-20, serial/lib/eigen.cc:14, serial/lib/lanczos.cc:202-207). */
#ifndef LZ_ORACLE_SHIM_LAPACKE_H
#define LZ_ORACLE_SHIM_LAPACKE_H
#ifdef __cplusplus
extern "C" {
#endif
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
lapack_int scipy_LAPACKE_dstevd(int layout, char jobz, lapack_int n, double* d, double* e, double* z, lapack_int ldz);
lapack_int scipy_LAPACKE_sstevd(int layout, char jobz, lapack_int n, float* d, float* e, float* z, lapack_int ldz);
lapack_int scipy_LAPACKE_dgeqrf(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau);
lapack_int scipy_LAPACKE_dorgqr(int layout, lapack_int m, lapack_int n, lapack_int k, double* a, lapack_int lda, const double* tau);
#ifdef __cplusplus
}
#endif
#define LAPACKE_dstevd scipy_LAPACKE_dstevd
#define LAPACKE_sstevd scipy_LAPACKE_sstevd
#define LAPACKE_dgeqrf scipy_LAPACKE_dgeqrf
#define LAPACKE_dorgqr scipy_LAPACKE_dorgqr
#endif
