// write_ans.h — host mirror of write_ans (parallel-final/lib/write_ans.h:9-16): one value per line. The reference
// prints with the stream's default 6 significant digits; we print 17 so a written answer can be compared to 1e-9.
#ifndef LZ_WRITE_ANS_H
#define LZ_WRITE_ANS_H

#include <cassert>
#include <fstream>
#include <iomanip>
#include <string>

#include "cu_lanczos.h"

template <typename T>
void write_ans(std::string filename, lanczosDecomp<T>& L) {
  std::ofstream fs;
  fs.open(filename);
  assert(!fs.fail());
  fs << std::setprecision(17);
  for (unsigned i = 0; i < L.A.get_n(); i++) fs << L.ans[i] << '\n';
}
#endif
