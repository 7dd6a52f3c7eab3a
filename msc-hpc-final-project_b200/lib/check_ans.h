// check_ans.h — host mirror of check_ans(L1, L2) (parallel-final/lib/check_ans.cu:12-29): max |difference| and where,
// ||difference||_2, and ||difference||_2 / ||L2.ans||_2 (the parity measure of BASELINE.json).
#ifndef LZ_CHECK_ANS_H
#define LZ_CHECK_ANS_H

#include <algorithm>
#include <cmath>
#include <iomanip>
#include <iostream>
#include <vector>

#include "cu_lanczos.h"

struct check_result {
  double max_abs, total_norm, relative_norm;
  unsigned max_idx;
};

template <typename T, typename U>
check_result compare_ans(const T* a, const U* b, unsigned n) {
  check_result r{0.0, 0.0, 0.0, 0u};
  double s = 0.0, nb = 0.0;
  for (unsigned i = 0; i < n; i++) {
    const double d = std::abs((double)a[i] - (double)b[i]);
    if (d > r.max_abs) { r.max_abs = d; r.max_idx = i; }
    s += d * d;
    nb += (double)b[i] * (double)b[i];
  }
  r.total_norm = std::sqrt(s);
  r.relative_norm = r.total_norm / std::sqrt(nb);
  return r;
}

template <typename T, typename U>
void check_ans(lanczosDecomp<T>& L1, lanczosDecomp<U>& L2) {
  const unsigned n = L1.A.get_n();
  const check_result r = compare_ans(L1.ans, L2.ans, n);
  std::cout << "\nMax difference of " << r.max_abs << " (Relative difference: " << r.max_abs / L2.ans[r.max_idx] << ") "
            << "found at index:\n" << std::setw(15) << "serial_ans[" << r.max_idx << "] = " << std::setprecision(10) << std::setw(15)
            << L1.ans[r.max_idx] << "\n" << std::setw(15) << "cuda_ans[" << r.max_idx << "] = " << std::setprecision(10) << std::setw(15)
            << L2.ans[r.max_idx] << '\n' << std::endl;
  std::cout << std::setw(30) << std::left << "Total norm of differences" << "=" << std::right << std::setprecision(20) << std::setw(30)
            << r.total_norm << std::endl;
  std::cout << std::setw(30) << std::left << "Relative norm of differences" << "=" << std::right << std::setprecision(20)
            << std::setw(30) << r.relative_norm << std::endl;
}
#endif
