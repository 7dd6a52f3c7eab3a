/* Shim so the reference's `#include <cblas.h>` / `"cblas.h"` resolves to SciPy's bundled LP64 OpenBLAS
 * (`scipy_` symbol prefix). TEST INFRASTRUCTURE ONLY (see lapacke.h in this directory). Declares only the
 * entry points the reference calls (parallel-final/lib/multiplyOut.cu:37-46, helpers.cu:121-139,
 * serial/lib/multiplyOut.cc:30-33). */
#ifndef LZ_ORACLE_SHIM_CBLAS_H
#define LZ_ORACLE_SHIM_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef unsigned long CBLAS_INDEX;
void scipy_cblas_dgemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int m, int n, double alpha, const double* a, int lda,
                       const double* x, int incx, double beta, double* y, int incy);
void scipy_cblas_sgemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int m, int n, float alpha, const float* a, int lda,
                       const float* x, int incx, float beta, float* y, int incy);
void scipy_cblas_dgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int m, int n, int k, double alpha,
                       const double* a, int lda, const double* b, int ldb, double beta, double* c, int ldc);
void scipy_cblas_sgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int m, int n, int k, float alpha,
                       const float* a, int lda, const float* b, int ldb, float beta, float* c, int ldc);
void scipy_openblas_set_num_threads(int);
#ifdef __cplusplus
}
#endif
#define cblas_dgemv scipy_cblas_dgemv
#define cblas_sgemv scipy_cblas_sgemv
#define cblas_dgemm scipy_cblas_dgemm
#define cblas_sgemm scipy_cblas_sgemm
#define openblas_set_num_threads scipy_openblas_set_num_threads
#endif
