/* lanczos_oracle.h — CPU restatement of the reference's e^A·x path (TEST INFRASTRUCTURE ONLY).
 * See lanczos_oracle.c for the per-function reference citations. Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this; the product (liblzb200.so) never does. */
#ifndef LANCZOS_ORACLE_H
#define LANCZOS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
void lzo_spmv(uint32_t n, const uint32_t* row_offset, const uint32_t* col_idx, const double* in, double* out);
double lzo_norm(const double* v, uint32_t n);
double lzo_inner_prod(const double* v, const double* w, uint32_t n);
/* Q: row-major n x k (Q[j + i*k] = component i of q_j). beta has k-1 entries. Returns 0. */
int lzo_lanczos(uint32_t n, const uint32_t* row_offset, const uint32_t* col_idx, uint32_t k, const double* x,
                double* alpha, double* beta, double* Q);
/* Arnoldi-assisted variant of serial/: MGS against q_0..q_{j-2} when j%2==0 && j>2. */
int lzo_lanczos_arnoldi(uint32_t n, const uint32_t* row_offset, const uint32_t* col_idx, uint32_t k, const double* x,
                        double* alpha, double* beta, double* Q);
/* Full reorthogonalisation (classical Gram-Schmidt twice against q_0..q_j every step): the specification our
 * LZ_REORTH_FULL mode is checked against. Not in the reference. */
int lzo_lanczos_fullreorth(uint32_t n, const uint32_t* row_offset, const uint32_t* col_idx, uint32_t k, const double* x,
                           double* alpha, double* beta, double* Q);
/* d[k] diag in / eigenvalues ascending out; e[k-1] sub-diagonal in / destroyed; Z row-major k x k out with
 * Z[i*k+j] = component i of eigenvector j. Returns 0, or l+1 if eigenvalue l did not converge. */
int lzo_tridiag_eig(uint32_t k, double* d, double* e, double* Z);
/* ans = x_norm * Q * (Z * (exp(lambda) .* Z[0,:])).  Q row-major n x k (qtrans=0) or k x n (qtrans=1). */
void lzo_multout(uint32_t n, uint32_t k, const double* eigvals, const double* Z, const double* Q, double x_norm,
                 int qtrans, double* ans, double* coeff_out);
/* Whole pipeline. reorth: 0 plain, 1 arnoldi-assisted (serial/), 2 full CGS2. alpha_out/beta_out may be NULL. */
int lzo_expv(uint32_t n, const uint32_t* row_offset, const uint32_t* col_idx, uint32_t k, const double* x, int reorth,
             double* ans, double* alpha_out, double* beta_out);
/* check_ans: max |a-b| and its index, ||a-b||_2, ||a-b||_2 / ||b||_2 */
void lzo_check_ans(uint32_t n, const double* a, const double* b, double* max_abs, uint32_t* max_idx, double* norm_diff,
                   double* rel);
/* argsort(-y) with lowest-index-first tie-break, first `top` entries. */
void lzo_top_k(uint32_t n, const double* y, uint32_t top, uint32_t* idx_out);
#ifdef __cplusplus
}
#endif
#endif
