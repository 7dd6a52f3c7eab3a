/* lanczos_oracle.c — CPU restatement of the reference's Lanczos e^A·x path in plain C.
 *
 * TEST INFRASTRUCTURE ONLY. This file is the parity oracle: tests/ compare the CUDA path against it, and it is itself
 * pinned against the UNMODIFIED reference compiled into oracle/_ref (tests/golden/*.npz were produced by oracle/_ref
 * binaries via tests/golden/make_golden.py; tests/test_oracle.py replays them, and re-runs oracle/_ref live when it
 * is present). The product library never links or loads this file.
 *
 * Every loop keeps the reference's operation order (sequential left-to-right sums, true division, separate
 * multiply and add — compile with -ffp-contract=off) so alpha/beta agree with the reference bit for bit on x86-64.
 *
 * Third-party arithmetic on the path that is NOT under /root/reference: LAPACKE_dstevd (LAPACK divide and conquer,
 * version unpinned — the reference links -llapacke, parallel-final/Makefile:6; here it resolves to the OpenBLAS
 * bundled with SciPy) and cblas_dgemv. lzo_tridiag_eig restates the eigenproblem with the published implicit-shift
 * QL algorithm (EISPACK tql2 / LAPACK dsteqr family); eigenvectors are unique only up to sign, so parity is asserted
 * on sign-free quantities (eigenvalues, the coefficient vector, ans), never vector by vector.
 */
#include "lanczos_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* reference parallel-final/lib/SPMV.cc:19-28 (same in serial/lib/SPMV.cc): value-less CSR gather-sum. */
void lzo_spmv(uint32_t n, const uint32_t* row_offset, const uint32_t* col_idx, const double* in, double* out) {
  for (uint32_t i = 0; i < n; ++i) out[i] = 0.0;
  for (uint32_t i = 0; i < n; ++i)
    for (uint32_t j = row_offset[i]; j < row_offset[i + 1]; j++) out[i] += in[col_idx[j]];
}

/* reference parallel-final/lib/cu_lanczos.h:18-24 (norm) */
double lzo_norm(const double* v, uint32_t n) {
  double s = 0.0;
  for (uint32_t i = 0; i < n; i++) s += v[i] * v[i];
  return sqrt(s);
}

/* reference parallel-final/lib/lanczos.cu:99-107 (inner_prod) */
double lzo_inner_prod(const double* v, const double* w, uint32_t n) {
  double s = 0.0;
  for (uint32_t i = 0; i < n; i++) s += v[i] * w[i];
  return s;
}

/* mode 0: reference parallel-final/lib/lanczos.cu:17-60 (lanczosDecomp<T>::decompose), == serial/lib/lanczos.cc:9-56
 * mode 1: reference serial/lib/lanczos.cc:58-132 (decompose_with_arnoldi)
 * mode 2: full CGS2 reorthogonalisation (ours) */
static int lanczos_impl(uint32_t n, const uint32_t* ro, const uint32_t* ci, uint32_t k, const double* x, double* alpha,
                        double* beta, double* Q, int mode) {
  double* v = (double*)malloc(sizeof(double) * n);
  double* Qraw = (double*)malloc(sizeof(double) * 2 * (size_t)n);
  double* Qs[2] = {Qraw, Qraw + n};
  double* Qcol = NULL; /* vector-contiguous copy, modes 1 and 2 (lanczos.cc:65-67) */
  double* h = NULL;
  if (mode) Qcol = (double*)malloc(sizeof(double) * (size_t)n * k);
  if (mode == 2) h = (double*)malloc(sizeof(double) * k);
  uint32_t i = 0;
  double x_norm = lzo_norm(x, n);
  for (uint32_t t = 0; t < n; t++) Qs[i][t] = x[t] / x_norm;

  for (uint32_t j = 0; j < k; j++) {
    lzo_spmv(n, ro, ci, Qs[i], v);                                    /* v = A q_j            lanczos.cu:32 */
    if (mode == 1 && j % 2 == 0 && j > 2) {                           /* lanczos.cc:85-91 */
      for (uint32_t t = 0; t + 1 < j; t++) {
        double dot = lzo_inner_prod(v, Qcol + (size_t)t * n, n);
        for (uint32_t r = 0; r < n; r++) v[r] -= dot * Qcol[(size_t)t * n + r];
      }
    }
    alpha[j] = lzo_inner_prod(v, Qs[i], n);                           /* lanczos.cu:34 */
    for (uint32_t t = 0; t < n; t++) v[t] -= alpha[j] * Qs[i][t];     /* lanczos.cu:36-37 */
    if (j > 0)
      for (uint32_t t = 0; t < n; t++) v[t] -= beta[j - 1] * Qs[1 - i][t]; /* lanczos.cu:39-43 */
    if (mode) memcpy(Qcol + (size_t)j * n, Qs[i], sizeof(double) * n);
    if (mode == 2) {
      for (int pass = 0; pass < 2; pass++) {
        for (uint32_t t = 0; t <= j; t++) h[t] = lzo_inner_prod(v, Qcol + (size_t)t * n, n);
        for (uint32_t t = 0; t <= j; t++)
          for (uint32_t r = 0; r < n; r++) v[r] -= h[t] * Qcol[(size_t)t * n + r];
      }
    }
    if (j + 1 < k) {
      beta[j] = lzo_norm(v, n);                                       /* lanczos.cu:47 */
      for (uint32_t t = 0; t < n; t++) Qs[1 - i][t] = v[t] / beta[j]; /* lanczos.cu:48-49 */
    }
    for (uint32_t t = 0; t < n; t++) Q[j + (size_t)t * k] = Qs[i][t]; /* lanczos.cu:53-54 (row-major n x k) */
    i = 1 - i;
  }
  free(v); free(Qraw); free(Qcol); free(h);
  return 0;
}

int lzo_lanczos(uint32_t n, const uint32_t* ro, const uint32_t* ci, uint32_t k, const double* x, double* alpha,
                double* beta, double* Q) { return lanczos_impl(n, ro, ci, k, x, alpha, beta, Q, 0); }
int lzo_lanczos_arnoldi(uint32_t n, const uint32_t* ro, const uint32_t* ci, uint32_t k, const double* x, double* alpha,
                        double* beta, double* Q) { return lanczos_impl(n, ro, ci, k, x, alpha, beta, Q, 1); }
int lzo_lanczos_fullreorth(uint32_t n, const uint32_t* ro, const uint32_t* ci, uint32_t k, const double* x,
                           double* alpha, double* beta, double* Q) { return lanczos_impl(n, ro, ci, k, x, alpha, beta, Q, 2); }

/* Restates what reference parallel-final/lib/eigen.cu:17-21 obtains from LAPACKE_dstevd(LAPACK_ROW_MAJOR,'V',k,d,e,Z,k):
 * eigenvalues ascending in d, Z[i*k+j] = component i of eigenvector j, e destroyed. Implicit-shift QL (tql2). */
int lzo_tridiag_eig(uint32_t k, double* d, double* e_in, double* Z) {
  int n = (int)k;
  double* e = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i + 1 < n; i++) e[i] = e_in[i];
  if (n > 0) e[n - 1] = 0.0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) Z[(size_t)i * n + j] = (i == j) ? 1.0 : 0.0;
  const double eps = 2.220446049250313e-16;
  double f = 0.0, tst1 = 0.0;
  int fail = 0;
  for (int l = 0; l < n && !fail; l++) {
    double t = fabs(d[l]) + fabs(e[l]);
    if (t > tst1) tst1 = t;
    int m = l;
    while (m < n) { if (fabs(e[m]) <= eps * tst1) break; m++; }
    if (m >= n) m = n - 1;   /* only when e holds NaN (Lanczos breakdown upstream): e[n-1] = 0 stops the scan otherwise */
    if (m > l) {
      int iter = 0;
      do {
        if (++iter > 60) { fail = l + 1; break; }
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; i++) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c, el1 = e[l + 1], s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; i--) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * p;
          r = hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int q = 0; q < n; q++) {
            h = Z[(size_t)q * n + i + 1];
            Z[(size_t)q * n + i + 1] = s * Z[(size_t)q * n + i] + c * h;
            Z[(size_t)q * n + i] = c * Z[(size_t)q * n + i] - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (fabs(e[l]) > eps * tst1);
    }
    d[l] = d[l] + f;
    e[l] = 0.0;
  }
  /* ascending order, as dstevd returns */
  for (int i = 0; i + 1 < n; i++) {
    int kk = i;
    double p = d[i];
    for (int j = i + 1; j < n; j++) if (d[j] < p) { kk = j; p = d[j]; }
    if (kk != i) {
      d[kk] = d[i]; d[i] = p;
      for (int q = 0; q < n; q++) { double t = Z[(size_t)q * n + i]; Z[(size_t)q * n + i] = Z[(size_t)q * n + kk]; Z[(size_t)q * n + kk] = t; }
    }
  }
  for (int i = 0; i + 1 < n; i++) e_in[i] = 0.0; /* dstevd leaves e destroyed */
  free(e);
  return fail;
}

/* reference parallel-final/lib/multiplyOut.cu:25-49. The two cblas_dgemv calls are restated as plain loops
 * (row-major NoTrans = dot per row; row-major Trans = axpy per row of Q). */
void lzo_multout(uint32_t n, uint32_t k, const double* eigvals, const double* Z, const double* Q, double x_norm,
                 int qtrans, double* ans, double* coeff_out) {
  double* f = (double*)malloc(sizeof(double) * k);
  double* tmp = (double*)malloc(sizeof(double) * k);
  for (uint32_t j = 0; j < k; j++) f[j] = exp(eigvals[j]);                 /* :30 */
  for (uint32_t j = 0; j < k; j++) f[j] *= x_norm * Z[j];                  /* :33  (first row of eigenvectors) */
  for (uint32_t i = 0; i < k; i++) {                                       /* :40  tmp = Z f */
    double s = 0.0;
    for (uint32_t j = 0; j < k; j++) s += Z[(size_t)i * k + j] * f[j];
    tmp[i] = s;
  }
  if (qtrans) {                                                            /* :44  Q is k x n */
    for (uint32_t i = 0; i < n; i++) ans[i] = 0.0;
    for (uint32_t j = 0; j < k; j++)
      for (uint32_t i = 0; i < n; i++) ans[i] += tmp[j] * Q[(size_t)j * n + i];
  } else {                                                                 /* :46  Q is n x k */
    for (uint32_t i = 0; i < n; i++) {
      double s = 0.0;
      for (uint32_t j = 0; j < k; j++) s += Q[(size_t)i * k + j] * tmp[j];
      ans[i] = s;
    }
  }
  if (coeff_out) memcpy(coeff_out, tmp, sizeof(double) * k);
  free(f); free(tmp);
}

/* reference parallel-final/main.cu:83-93: lanczosDecomp -> eigenDecomp -> multOut(L,E,A,false) */
int lzo_expv(uint32_t n, const uint32_t* ro, const uint32_t* ci, uint32_t k, const double* x, int reorth, double* ans,
             double* alpha_out, double* beta_out) {
  double* alpha = (double*)malloc(sizeof(double) * k);
  double* beta = (double*)malloc(sizeof(double) * (k > 1 ? k - 1 : 1));
  double* Q = (double*)malloc(sizeof(double) * (size_t)n * k);
  double* Z = (double*)malloc(sizeof(double) * (size_t)k * k);
  lanczos_impl(n, ro, ci, k, x, alpha, beta, Q, reorth);
  if (alpha_out) memcpy(alpha_out, alpha, sizeof(double) * k);
  if (beta_out && k > 1) memcpy(beta_out, beta, sizeof(double) * (k - 1));
  int rc = lzo_tridiag_eig(k, alpha, beta, Z);          /* eigen.h:26 copies alpha into eigenvalues */
  lzo_multout(n, k, alpha, Z, Q, lzo_norm(x, n), 0, ans, NULL);
  free(alpha); free(beta); free(Q); free(Z);
  return rc;
}

/* reference parallel-final/lib/check_ans.cu:12-29 */
void lzo_check_ans(uint32_t n, const double* a, const double* b, double* max_abs, uint32_t* max_idx, double* norm_diff,
                   double* rel) {
  double mx = -1.0, s = 0.0;
  uint32_t mi = 0;
  for (uint32_t i = 0; i < n; i++) {
    double dlt = fabs(a[i] - b[i]);
    if (dlt > mx) { mx = dlt; mi = i; }
    s += dlt * dlt;
  }
  *max_abs = mx; *max_idx = mi; *norm_diff = sqrt(s); *rel = sqrt(s) / lzo_norm(b, n);
}

/* Ranking used by the parity tests (the reference never ranks): argsort(-y), ties -> lower index. */
typedef struct { double y; uint32_t i; } lzo_pair;
static int cmp_pair(const void* a, const void* b) {
  const lzo_pair* p = (const lzo_pair*)a; const lzo_pair* q = (const lzo_pair*)b;
  if (p->y > q->y) return -1;
  if (p->y < q->y) return 1;
  return (p->i > q->i) - (p->i < q->i);
}
void lzo_top_k(uint32_t n, const double* y, uint32_t top, uint32_t* idx_out) {
  lzo_pair* a = (lzo_pair*)malloc(sizeof(lzo_pair) * (size_t)n);
  for (uint32_t i = 0; i < n; i++) { a[i].y = y[i]; a[i].i = i; }
  qsort(a, n, sizeof(lzo_pair), cmp_pair);
  for (uint32_t t = 0; t < top && t < n; t++) idx_out[t] = a[t].i;
  free(a);
}
