/* oracle/ref_serial.cc — driver around the UNMODIFIED reference `serial/lib` (TEST INFRASTRUCTURE ONLY).
 *
 * Secondary oracle: the reference's CPU-only tree, called exactly as serial/main.cc:57-88 does
 * (adjMatrix(n,E,fs) -> lanczosDecomp L(A,k,ones[,arnoldi]) -> eigenDecomp E(L) -> multOut(L,E,A)), dumping
 * alpha, beta and ans in binary. Linked with zero_new.cc because serial/lib/multiplyOut.cc:27-33 accumulates
 * with beta=1 onto buffers it never initialises. Reference sources are compiled where
 * they lie; none are copied. `private` is re-defined for the reference headers only, after the std headers.
 *
 *   ref_serial --mtx FILE -k K [--arnoldi] [--out PREFIX]
 */
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <random>
#include <set>
#include <string>
#include <vector>
#include <sys/time.h>

#define private public
#include "lib/adjMatrix.h"
#include "lib/lanczos.h"
#include "lib/eigen.h"
#include "lib/multiplyOut.h"
#undef private

static double now_s() { timeval t; gettimeofday(&t, NULL); return t.tv_sec + 1e-6 * t.tv_usec; }
static void dump(const std::string& path, const double* p, size_t n) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { perror(path.c_str()); exit(2); }
  fwrite(p, sizeof(double), n, f);
  fclose(f);
}

int main(int argc, char** argv) {
  std::string mtx, out;
  long unsigned k = 20;
  bool arnoldi = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--mtx" && i + 1 < argc) mtx = argv[++i];
    else if (a == "--out" && i + 1 < argc) out = argv[++i];
    else if (a == "-k" && i + 1 < argc) k = strtoul(argv[++i], 0, 10);
    else if (a == "--arnoldi") arnoldi = true;
    else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
  }
  std::ifstream fs(mtx);
  if (fs.fail()) { fprintf(stderr, "cannot open %s\n", mtx.c_str()); return 2; }
  long unsigned n, edges;
  fs >> n >> n >> edges;                          /* serial/main.cc:60 */
  adjMatrix A(n, edges, fs);
  std::vector<double> x(n, 1.0);                  /* serial/main.cc:79 */
  std::cout.setstate(std::ios_base::failbit);     /* the Arnoldi variant announces itself on stdout */
  double s = now_s();
  lanczosDecomp L(A, k, x.data(), arnoldi);
  double e1 = now_s();
  std::vector<double> alpha(L.alpha, L.alpha + k), beta(L.beta, L.beta + (k - 1));
  eigenDecomp E(L);
  double e2 = now_s();
  multOut(L, E, A);
  double e3 = now_s();
  printf("{\"mode\":\"serial\",\"n\":%lu,\"k\":%lu,\"arnoldi\":%d,\"lanczos_s\":%.6f,\"eig_s\":%.6f,\"multout_s\":%.6f}\n",
         n, k, (int)arnoldi, e1 - s, e2 - e1, e3 - e2);
  if (!out.empty()) {
    dump(out + ".alpha.f64", alpha.data(), k);
    dump(out + ".beta.f64", beta.data(), k - 1);
    dump(out + ".ans.f64", L.ans, n);
  }
  return 0;
}
