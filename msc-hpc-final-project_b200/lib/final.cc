// final.cc — driver mirroring the reference's parallel-final/main.cu:34-161 on the B200-native library.
//
//   ./final -k 30 -f NAME                  reads ../data/NAME/NAME.mtx (main.cu:54), or --path FILE for an explicit file
//   ./final -k 50 --graph rmat --scale 20 [--ef 8] [--seed 1]     seeded generators (BASELINE.json configs)
//   ./final -k 20 --graph er -n 10000 -e 50000 [--seed S]
//   ./final -k 100 --graph band -n 1048576 [--seed S]
//   ./final -k 30 --graph ba -n 100000 -b 20                      Barabasi-Albert, as main.cu:67-71
//   options: --reorth (full reorthogonalisation)  --gpus N (row-sharded, one host thread per GPU, NCCL)
//            --check FILE (raw float64 e^A x to compare with, e.g. written by oracle/_ref/ref_final --out)
//            --write FILE (one value per line, write_ans)   --top M (print the M most central vertices, device-side ranking)
//
// Same object sequence as the reference: adjMatrix -> lanczosDecomp<double>(A,k,x,cuda=true) -> eigenDecomp<double> ->
// multOut(L,E,A,true), same TIMING / ERROR CHECKING tables. The reference's serial arm (main.cu:83-99) is not here: this
// library has no CPU path; the serial code is the test oracle (oracle/), whose answer can be passed with --check.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <thread>
#include <vector>
#include <sys/time.h>

#include "adjMatrix.h"
#include "check_ans.h"
#include "cu_lanczos.h"
#include "eigen.h"
#include "helpers.h"
#include "multiplyOut.h"
#include "write_ans.h"

#define WIDTH 81

static double seconds_between(timeval s, timeval e) { return e.tv_sec - s.tv_sec + (e.tv_usec - s.tv_usec) / 1000000.0; }

// Row-sharded run over `gpus` devices: one host thread per GPU, each driving its own lz_ctx (C ABI directly).
static int run_multi(adjMatrix& A, unsigned k, int reorth, int gpus, std::vector<double>& ans, double& t_lanczos, double& t_mult) {
  unsigned char uid[LZ_NCCL_UID_BYTES];
  if (lz_nccl_unique_id(uid) != LZ_OK) { std::cerr << lz_last_error() << '\n'; return 1; }
  std::vector<int> rc(gpus, 0);
  std::vector<std::string> err(gpus);
  std::vector<double> tl(gpus, 0.0), tm(gpus, 0.0);
  std::vector<std::thread> th;
  for (int r = 0; r < gpus; r++) {
    th.emplace_back([&, r]() {
      lz_ctx* c = nullptr;
      auto bad = [&](int code) { rc[r] = code; err[r] = lz_last_error(); if (c) lz_destroy(c); };
      int e;
      if ((e = lz_create_dist(r, r, gpus, uid, &c)) != LZ_OK) return bad(e);
      if ((e = lz_csr_upload(c, A.get_n(), A.get_row_offset(), A.get_col_idx())) != LZ_OK) return bad(e);
      if ((e = lz_set_start_vector(c, nullptr)) != LZ_OK) return bad(e);
      if ((e = lz_sync(c)) != LZ_OK) return bad(e);
      timeval s, e1, e2;
      gettimeofday(&s, NULL);
      if ((e = lz_lanczos_run(c, k, reorth)) != LZ_OK) return bad(e);
      if ((e = lz_sync(c)) != LZ_OK) return bad(e);
      gettimeofday(&e1, NULL);
      if ((e = lz_tridiag_expv(c)) != LZ_OK) return bad(e);
      if ((e = lz_multout(c)) != LZ_OK) return bad(e);
      std::vector<double> y(A.get_n());
      if ((e = lz_get_ans(c, y.data())) != LZ_OK) return bad(e);
      gettimeofday(&e2, NULL);
      tl[r] = seconds_between(s, e1); tm[r] = seconds_between(e1, e2);
      if (r == 0) ans.swap(y);
      lz_destroy(c);
    });
  }
  for (auto& t : th) t.join();
  for (int r = 0; r < gpus; r++)
    if (rc[r]) { std::cerr << "rank " << r << ": " << err[r] << '\n'; return 1; }
  t_lanczos = *std::max_element(tl.begin(), tl.end());
  t_mult = *std::max_element(tm.begin(), tm.end());
  return 0;
}

int main(int argc, char** argv) {
  unsigned n{10'000}, deg{5}, edges{n * 10}, krylov_dim{100}, width{17};
  bool verbose{true};
  std::string filename{"bn1000000e9999944"}, graph{"file"}, path, check_file, write_file;
  unsigned scale = 16, ef = 8;
  uint64_t seed = 1;
  int reorth = LZ_REORTH_NONE, gpus = 1;
  unsigned top = 0;

  // long options first; what is left goes to the reference's getopt string
  std::vector<char*> rest{argv[0]};
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&]() -> const char* { if (i + 1 >= argc) { std::cerr << "missing value for " << a << '\n'; exit(2); } return argv[++i]; };
    if (a == "--graph") graph = val();
    else if (a == "--scale") scale = (unsigned)atoi(val());
    else if (a == "--ef") ef = (unsigned)atoi(val());
    else if (a == "--seed") seed = strtoull(val(), nullptr, 10);
    else if (a == "--path") path = val();
    else if (a == "--check") check_file = val();
    else if (a == "--write") write_file = val();
    else if (a == "--gpus") gpus = atoi(val());
    else if (a == "--reorth") reorth = LZ_REORTH_FULL;
    else if (a == "--top") top = (unsigned)atoi(val());
    else rest.push_back(argv[i]);
  }
  if (parseArguments((int)rest.size(), rest.data(), filename, krylov_dim, verbose, n, deg, edges) != 0) return 2;

  timeval start, end;
  gettimeofday(&start, NULL);
  adjMatrix A;
  adjMatrix::generator_seed = seed;
  if (graph == "file") {
    std::string filepath = path.empty() ? "../data/" + filename + "/" + filename + ".mtx" : path;
    std::cout << "Going to open file: " << filepath << std::endl;
    std::ifstream fs;
    fs.open(filepath);
    if (fs.fail()) { std::cerr << "File opening failed\n"; return 1; }
    fs >> n >> n >> edges;
    adjMatrix B(n, edges, fs);
    A = std::move(B);
  } else if (graph == "rmat") {
    A = adjMatrix::rmat(scale, ef, seed);
  } else if (graph == "er") {
    adjMatrix B(n, edges);
    A = std::move(B);
  } else if (graph == "band") {
    A = adjMatrix::banded(n, seed);
  } else if (graph == "ba") {
    adjMatrix B(n, deg, 'b');
    A = std::move(B);
  } else {
    std::cerr << "unknown --graph " << graph << '\n';
    return 2;
  }
  n = A.get_n();
  edges = A.get_edges();
  gettimeofday(&end, NULL);
  std::cout << "\nTime elapsed to build adjacency matrix with n = " << n << " edges = " << edges << ":\n\t" << seconds_between(start, end)
            << " seconds\n\n";
  std::cout << "Running Lanczos algorithm for krylov_dim " << krylov_dim << (reorth ? " with full reorthogonalisation" : "") << " on "
            << gpus << " GPU(s)\n\n";

  std::vector<double> x_double(n, 1);   // main.cu:79
  std::vector<double> result;
  double gpu_time_lanczos = 0, gpu_time_mult = 0, gpu_time_whole = 0;
  lz_timings tm{};
  if (gpus > 1) {
    if (run_multi(A, krylov_dim, reorth, gpus, result, gpu_time_lanczos, gpu_time_mult)) return 1;
    gpu_time_whole = gpu_time_lanczos + gpu_time_mult;
  } else {
    timeval s, e1, e2, e3;
    gettimeofday(&s, NULL);
    lanczosDecomp<double> cu_L(A, krylov_dim, &x_double[0], /*cuda=*/true, reorth);   // main.cu:115
    gettimeofday(&e1, NULL);
    eigenDecomp<double> cu_E(cu_L);                                                   // main.cu:124
    gettimeofday(&e2, NULL);
    multOut(cu_L, cu_E, A, true);                                                     // main.cu:127
    gettimeofday(&e3, NULL);
    gpu_time_lanczos = seconds_between(s, e1);
    gpu_time_mult = seconds_between(e2, e3);
    gpu_time_whole = seconds_between(s, e3);
    tm = cu_L.timings();
    result.assign(cu_L.get_ans_ptr(), cu_L.get_ans_ptr() + n);
    if (!write_file.empty()) write_ans(write_file, cu_L);                             // main.cu:159
    if (top) {
      std::vector<unsigned> ti(top);
      std::vector<double> tv(top);
      const unsigned cnt = cu_L.top_k(top, ti.data(), tv.data());
      std::cout << "TOP " << cnt << " vertices by total communicability (rank vertex value)\n" << std::setprecision(17);
      for (unsigned i = 0; i < cnt; i++) std::cout << "rank " << i + 1 << ' ' << ti[i] << ' ' << tv[i] << '\n';
      std::cout << std::setprecision(6);
    }
  }

  std::cout << std::setfill('~') << std::setw(WIDTH) << '\n' << std::setfill(' ');
  std::cout << "TIMING (host wall clock, seconds; Lanczos includes CSR upload and relabelling as the reference's includes its cudaMalloc + H2D)\n";
  std::cout << std::setfill('~') << std::setw(WIDTH) << '\n' << std::setfill(' ');
  std::cout << std::setw(width) << std::left << "Lanczos" << std::right << std::setw(width) << gpu_time_lanczos << "\n\n";
  std::cout << std::setw(width) << std::left << "Multiply Out" << std::right << std::setw(width) << gpu_time_mult << "\n\n";
  std::cout << std::setw(width) << std::left << "Entire algorithm" << std::right << std::setw(width) << gpu_time_whole << "\n\n";
  if (gpus == 1)
    std::cout << "device time: Lanczos loop " << tm.lanczos_ms << " ms (" << krylov_dim / (tm.lanczos_ms * 1e-3) << " iterations/s), tridiagonal "
              << tm.tridiag_ms << " ms, multOut " << tm.multout_ms << " ms, kernels launched " << tm.kernel_launches << "\n\n";

  std::cout << std::setfill('~') << std::setw(WIDTH) << '\n' << std::setfill(' ');
  std::cout << "ERROR CHECKING\n";
  std::cout << std::setfill('~') << std::setw(WIDTH) << '\n' << std::setfill(' ');
  bool finite = true;
  for (double v : result) finite &= std::isfinite(v);
  std::cout << "result finite: " << (finite ? "yes" : "NO") << ", ||ans||_2 = " << std::setprecision(17) << norm(result.data(), n) << '\n';
  int status = finite ? 0 : 3;
  if (!check_file.empty()) {
    std::vector<double> ref(n);
    FILE* f = fopen(check_file.c_str(), "rb");
    if (!f || fread(ref.data(), 8, n, f) != n) { std::cerr << "cannot read " << check_file << '\n'; return 1; }
    fclose(f);
    check_result r = compare_ans(ref.data(), result.data(), n);
    std::cout << "Max difference of " << std::setprecision(10) << r.max_abs << " found at index " << r.max_idx << '\n';
    std::cout << std::setw(30) << std::left << "Total norm of differences" << "=" << std::right << std::setprecision(20) << std::setw(30)
              << r.total_norm << '\n';
    std::cout << std::setw(30) << std::left << "Relative norm of differences" << "=" << std::right << std::setprecision(20)
              << std::setw(30) << r.relative_norm << std::endl;
    if (!(r.relative_norm < 1e-9)) status = 4;
  }
  if (gpus > 1 && !write_file.empty()) {
    std::ofstream fs(write_file);
    fs << std::setprecision(17);
    for (double v : result) fs << v << '\n';
  }
  std::cout << std::setfill('~') << std::setw(WIDTH) << '\n' << std::setfill(' ');
  return status;
}
