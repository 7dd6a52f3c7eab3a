// helpers.h — host mirror of the reference's helpers (parallel-final/lib/helpers.h, helpers.cu:14-63): getopt parsing
// with the same flags (-k -f -b -n -e -v) and a wall-clock stopwatch pair with the cuda_start_timer/cuda_end_timer
// names (device time is reported by lz_timings).
#ifndef LZ_HELPERS_H
#define LZ_HELPERS_H

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <unistd.h>

inline int parseArguments(int argc, char* argv[], std::string& filename, unsigned& krylov_dim, bool& verbose, unsigned& n,
                          unsigned& bar_deg, unsigned& E) {
  int c;
  optind = 1;
  while ((c = getopt(argc, argv, "k:f:b:n:e:v")) != -1) {
    switch (c) {
      case 'f': filename = optarg; break;
      case 'k': krylov_dim = (unsigned)atoi(optarg); break;
      case 'b': bar_deg = (unsigned)atoi(optarg); break;
      case 'n': n = (unsigned)atoi(optarg); break;
      case 'e': E = (unsigned)atoi(optarg); break;
      case 'v': verbose = true; break;
      default: fprintf(stderr, "Invalid option given\n"); return -1;
    }
  }
  return 0;
}

using lz_stopwatch = std::chrono::steady_clock::time_point;
inline void cuda_start_timer(lz_stopwatch& start, lz_stopwatch& end) { start = end = std::chrono::steady_clock::now(); }
inline float cuda_end_timer(lz_stopwatch& start, lz_stopwatch& end) {
  end = std::chrono::steady_clock::now();
  return std::chrono::duration<float>(end - start).count();
}
#endif
