/* oracle/ref_final.cu — driver around the UNMODIFIED reference `parallel-final/lib` (TEST INFRASTRUCTURE ONLY).
 *
 * Built by oracle/Makefile into oracle/_ref/ref_final together with the reference's own objects, compiled from
 * the sources where they lie under /root/reference. Nothing from the reference is copied into this repository;
 * this file only *calls* the reference's public objects in the same order as parallel-final/main.cu:83-127:
 *     lanczosDecomp<double> L(A,k,x,cuda)  ->  eigenDecomp<double> E(L)  ->  multOut(L,E,A,Qtrans)
 * and dumps alpha, beta, ans in binary so the parity tests and the C restatement (lanczos_oracle.c) can be pinned
 * against the reference's actual output. It is also the `--impl reference` / cpu_baseline leg of bench.py.
 *
 * Access to private members: the reference gives no accessor for alpha/beta/ans or for installing a prebuilt CSR
 * (its only loader is the std::set text reader, adjMatrix.cc:21-46, which needs ~48 B/entry). We include the
 * standard headers first and then re-define `private` for the reference headers only. Class layout is unaffected.
 *
 * Inputs
 *   --csr FILE   binary CSR: "LZCSR1\0\0", u64 n, u64 nnz, u32 row_offset[n+1], u32 col_idx[nnz]   (injected)
 *   --mtx FILE   reference text format ("n n E" then E lines "col row", 1-based)  (reference's own loader)
 *   -k K         Krylov dimension
 *   --cuda       run the reference's cu_decompose() path (needs a GPU) with Qtrans=true, as main.cu:115-127
 *   --iters M    time only an M-step decomposition (bounded CPU sample for bench.py), skip eig/multOut
 *   --reps R     repeat the timed part R times (prints one JSON line per repetition)
 *   --x FILE     starting vector (n float64, raw); default all ones (main.cu:79)
 *   --float      run the reference's single-precision instantiation lanczosDecomp<float> / eigenDecomp<float> / multOut<float>
 *                (cu_lanczos.cu:144, eigen.cu:23, multiplyOut.cu:52; switch described at parallel-final/README.md:21); outputs are
 *                widened to float64 for the dumps
 *   --out PREFIX write PREFIX.alpha.f64 / .beta.f64 / .ans.f64 (raw little-endian float64)
 */
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include <sys/time.h>
#include <unistd.h>
#include <cuda_runtime.h>
#include "cublas_v2.h"

#define private public
#include "lib/adjMatrix.h"
#include "lib/cu_lanczos.h"
#include "lib/eigen.h"
#include "lib/multiplyOut.h"
#include "lib/check_ans.h"
#undef private

static double now_s() {
  timeval t;
  gettimeofday(&t, NULL);
  return t.tv_sec + 1e-6 * t.tv_usec;
}

static void dump(const std::string& path, const double* p, size_t n) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { perror(path.c_str()); exit(2); }
  fwrite(p, sizeof(double), n, f);
  fclose(f);
}

static void load_csr(const char* path, adjMatrix& A) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  char magic[8];
  uint64_t n = 0, nnz = 0;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "LZCSR1\0\0", 8) != 0) { fprintf(stderr, "bad CSR magic\n"); exit(2); }
  if (fread(&n, 8, 1, f) != 1 || fread(&nnz, 8, 1, f) != 1) { fprintf(stderr, "bad CSR header\n"); exit(2); }
  A.row_offset = new unsigned[n + 1];
  A.col_idx = new unsigned[nnz];
  if (fread(A.row_offset, 4, n + 1, f) != n + 1 || fread(A.col_idx, 4, nnz, f) != nnz) { fprintf(stderr, "short CSR file\n"); exit(2); }
  fclose(f);
  A.n = (unsigned)n;
  A.edge_count = (unsigned)(nnz / 2);
  A.matrix_type = 'f';
}

template <typename T>
static void run_reps(adjMatrix& A, unsigned n, unsigned k, unsigned iters, unsigned reps, bool cuda, const std::vector<double>& xd, const std::string& out,
                     double t_load) {
  std::vector<T> x(xd.begin(), xd.end());
  const char* prec = sizeof(T) == 4 ? "f32" : "f64";
  for (unsigned r = 0; r < reps; r++) {
    if (iters) {
      double s = now_s();
      lanczosDecomp<T> L(A, iters, x.data(), cuda);
      if (cuda) cudaDeviceSynchronize();
      double t = now_s() - s;
      printf("{\"mode\":\"iters\",\"precision\":\"%s\",\"n\":%u,\"nnz\":%llu,\"iters\":%u,\"cuda\":%d,\"lanczos_s\":%.6f,\"iters_per_s\":%.6f,\"load_s\":%.3f}\n",
             prec, n, 2ull * A.get_edges(), iters, (int)cuda, t, iters / t, t_load);
      fflush(stdout);
      continue;
    }
    double s = now_s();
    lanczosDecomp<T> L(A, k, x.data(), cuda);
    if (cuda) cudaDeviceSynchronize();
    double e1 = now_s();
    std::vector<double> alpha(L.alpha, L.alpha + k), beta(L.beta, L.beta + (k - 1));  /* dstevd destroys L.beta */
    eigenDecomp<T> E(L);
    double e2 = now_s();
    multOut(L, E, A, cuda);           /* Qtrans == cuda, as main.cu:93 and :127 */
    double e3 = now_s();
    printf("{\"mode\":\"full\",\"precision\":\"%s\",\"n\":%u,\"nnz\":%llu,\"k\":%u,\"cuda\":%d,\"lanczos_s\":%.6f,\"eig_s\":%.6f,\"multout_s\":%.6f,\"total_s\":%.6f,\"load_s\":%.3f}\n",
           prec, n, 2ull * A.get_edges(), k, (int)cuda, e1 - s, e2 - e1, e3 - e2, e3 - s, t_load);
    fflush(stdout);
    if (!out.empty() && r + 1 == reps) {
      std::vector<double> ans(L.ans, L.ans + n);
      dump(out + ".alpha.f64", alpha.data(), k);
      dump(out + ".beta.f64", beta.data(), k - 1);
      dump(out + ".ans.f64", ans.data(), n);
    }
    if (cuda) { cudaFree(L.Q_d); L.Q_d = nullptr; }   /* the reference never frees Q_d and would cudaFree(ans) */
  }
}

int main(int argc, char** argv) {
  std::string csr, mtx, out, xfile;
  unsigned k = 20, iters = 0, reps = 1;
  bool cuda = false, use_float = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
    if (a == "--csr") csr = next();
    else if (a == "--mtx") mtx = next();
    else if (a == "--out") out = next();
    else if (a == "--x") xfile = next();
    else if (a == "-k") k = (unsigned)atoi(next());
    else if (a == "--iters") iters = (unsigned)atoi(next());
    else if (a == "--reps") reps = (unsigned)atoi(next());
    else if (a == "--cuda") cuda = true;
    else if (a == "--float") use_float = true;
    else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
  }
  if (csr.empty() == mtx.empty()) { fprintf(stderr, "exactly one of --csr / --mtx\n"); return 2; }

  adjMatrix A;
  double t0 = now_s();
  if (!csr.empty()) {
    load_csr(csr.c_str(), A);
  } else {
    std::ifstream fs(mtx);
    if (fs.fail()) { fprintf(stderr, "cannot open %s\n", mtx.c_str()); return 2; }
    unsigned n, edges;
    fs >> n >> n >> edges;            /* parallel-final/main.cu:62 */
    adjMatrix B(n, edges, fs);
    A = std::move(B);
  }
  double t_load = now_s() - t0;
  unsigned n = A.get_n();

  std::vector<double> x(n, 1.0);      /* parallel-final/main.cu:79 */
  if (!xfile.empty()) {
    FILE* f = fopen(xfile.c_str(), "rb");
    if (!f || fread(x.data(), 8, n, f) != n) { fprintf(stderr, "cannot read x\n"); return 2; }
    fclose(f);
  }

  std::cout.setstate(std::ios_base::failbit);   /* silence the reference's chatter (free_mem prints, memory line) */
  if (use_float) run_reps<float>(A, n, k, iters, reps, cuda, x, out, t_load);
  else run_reps<double>(A, n, k, iters, reps, cuda, x, out, t_load);
  return 0;
}
