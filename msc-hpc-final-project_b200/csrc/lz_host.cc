// lz_host.cc — host-side parts of the C ABI that need no GPU: error text, deterministic graph generators, CSR file I/O.
// Mirrors the reference's host loader / generators (parallel-final/lib/adjMatrix.cc:21-69, make_graph.cc:21-113) in
// function, not in method: edges become 64-bit (row<<32|col) keys, sorted and de-duplicated in one pass.
#include "lz_internal.h"
#include "lz_gen.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <vector>

static thread_local char g_err[512] = "";

int lz_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

extern "C" const char* lz_last_error(void) { return g_err; }
extern "C" int lz_version(void) { return 100; }
extern "C" void lz_free_host(void* p) { free(p); }

// keys: (row << 32 | col), unsorted, may contain duplicates. Consumed (sorted in place).
int lz_build_csr_from_keys(uint64_t n, uint64_t* keys, uint64_t nkeys, uint64_t* nnz_out,
                           uint32_t** row_offset_out, uint32_t** col_idx_out) {
  std::sort(keys, keys + nkeys);
  uint64_t nnz = std::unique(keys, keys + nkeys) - keys;
  if (nnz > 0xFFFFFFFFull) return lz_fail(LZ_ERR_ARG, "nnz %llu does not fit 32-bit row offsets", (unsigned long long)nnz);
  uint32_t* ro = (uint32_t*)calloc(n + 1, sizeof(uint32_t));
  uint32_t* ci = (uint32_t*)malloc((nnz ? nnz : 1) * sizeof(uint32_t));
  if (!ro || !ci) { free(ro); free(ci); return lz_fail(LZ_ERR_ALLOC, "host allocation of CSR failed"); }
  for (uint64_t i = 0; i < nnz; i++) {
    ro[(keys[i] >> 32) + 1]++;
    ci[i] = (uint32_t)keys[i];
  }
  for (uint64_t i = 0; i < n; i++) ro[i + 1] += ro[i];
  *nnz_out = nnz; *row_offset_out = ro; *col_idx_out = ci;
  return LZ_OK;
}

extern "C" int lz_graph_generate_host(const lz_graph_spec* spec, uint64_t* n_out, uint64_t* nnz_out,
                                      uint32_t** row_offset_out, uint32_t** col_idx_out) {
  if (!spec || !n_out || !nnz_out || !row_offset_out || !col_idx_out) return lz_fail(LZ_ERR_ARG, "null argument");
  lz_gen_params p;
  if (lz_gen_prepare(spec, &p)) return lz_fail(LZ_ERR_ARG, "bad graph spec (kind %u)", spec->kind);
  try {
    std::vector<uint64_t> keys;
    keys.reserve(2 * (p.m + 1));
    for (uint64_t e = 0; e <= p.m; e++) {
      uint32_t u, v;
      lz_gen_edge(p, e, &u, &v);
      if (u == v) continue;
      keys.push_back(((uint64_t)u << 32) | v);
      keys.push_back(((uint64_t)v << 32) | u);
    }
    *n_out = p.n;
    return lz_build_csr_from_keys(p.n, keys.data(), keys.size(), nnz_out, row_offset_out, col_idx_out);
  } catch (const std::exception& ex) {          // no exception crosses the C ABI
    return lz_fail(LZ_ERR_ALLOC, "lz_graph_generate_host: %s", ex.what());
  }
}

extern "C" int lz_csr_from_edges(uint64_t n, uint64_t n_edges, const uint32_t* u, const uint32_t* v, uint64_t* nnz_out,
                                 uint32_t** row_offset_out, uint32_t** col_idx_out) {
  if (!nnz_out || !row_offset_out || !col_idx_out || (n_edges && (!u || !v))) return lz_fail(LZ_ERR_ARG, "null argument");
  if (n == 0 || n > 0xFFFFFFFFull) return lz_fail(LZ_ERR_ARG, "bad vertex count %llu", (unsigned long long)n);
  try {
    std::vector<uint64_t> keys;
    keys.reserve(2 * n_edges);
    for (uint64_t i = 0; i < n_edges; i++) {
      if (u[i] >= n || v[i] >= n) return lz_fail(LZ_ERR_ARG, "edge %llu has a vertex out of range", (unsigned long long)i);
      if (u[i] == v[i]) continue;
      keys.push_back(((uint64_t)u[i] << 32) | v[i]);
      keys.push_back(((uint64_t)v[i] << 32) | u[i]);
    }
    return lz_build_csr_from_keys(n, keys.data(), keys.size(), nnz_out, row_offset_out, col_idx_out);
  } catch (const std::exception& ex) {
    return lz_fail(LZ_ERR_ALLOC, "lz_csr_from_edges: %s", ex.what());
  }
}

// Next line that carries data: skips blank lines and '%' comment lines (MatrixMarket banner / comments). Returns false at EOF.
static bool next_data_line(FILE* f, char* buf, size_t cap) {
  while (fgets(buf, (int)cap, f)) {
    const char* p = buf;
    while (*p == ' ' || *p == '\t') p++;
    if (*p == '%' || *p == '\n' || *p == '\r' || *p == 0) continue;
    return true;
  }
  return false;
}

extern "C" int lz_csr_read_text(const char* path, uint64_t* n_out, uint64_t* nnz_out, uint32_t** row_offset_out,
                                uint32_t** col_idx_out) {
  if (!path || !n_out || !nnz_out || !row_offset_out || !col_idx_out) return lz_fail(LZ_ERR_ARG, "null argument");
  FILE* f = fopen(path, "r");
  if (!f) return lz_fail(LZ_ERR_IO, "cannot open %s", path);
  // The reference's format is a MatrixMarket coordinate file with the banner stripped (serial/README.md:9): "n n E", then E lines
  // "col row". The same reader therefore also takes the unstripped file: '%' lines are skipped and anything after the two indices
  // of an entry line (a value column) is ignored — the matrix is pattern-only either way.
  char line[512];
  unsigned long long n = 0, n2 = 0, e = 0;
  if (!next_data_line(f, line, sizeof line) || sscanf(line, "%llu %llu %llu", &n, &n2, &e) != 3 || n == 0 || n != n2 || n > 0xFFFFFFFFull) {
    fclose(f);
    return lz_fail(LZ_ERR_IO, "%s: bad header (expected 'n n E')", path);
  }
  try {
    std::vector<uint64_t> keys;
    // the header is untrusted input: do not reserve what it claims beyond what the file can possibly hold (>= 4 bytes per edge)
    long here = ftell(f);
    fseek(f, 0, SEEK_END);
    const unsigned long long max_edges = (unsigned long long)(ftell(f) > here ? ftell(f) - here : 0) / 4 + 1;
    fseek(f, here, SEEK_SET);
    keys.reserve(2 * (e < max_edges ? e : max_edges));
    for (unsigned long long i = 0; i < e; i++) {
      unsigned long long col, row;
      if (!next_data_line(f, line, sizeof line) || sscanf(line, "%llu %llu", &col, &row) != 2) {
        fclose(f);
        return lz_fail(LZ_ERR_IO, "%s: short edge list (%llu of %llu)", path, i, e);
      }
      if (col < 1 || row < 1 || col > n || row > n) { fclose(f); return lz_fail(LZ_ERR_IO, "%s: vertex out of range on edge %llu", path, i); }
      --col; --row;                                   // files are 1-based (adjMatrix.cc:31-34)
      keys.push_back(((uint64_t)row << 32) | col);    // both triangles, duplicates collapse in the builder
      keys.push_back(((uint64_t)col << 32) | row);
    }
    fclose(f);
    f = nullptr;
    *n_out = n;
    return lz_build_csr_from_keys(n, keys.data(), keys.size(), nnz_out, row_offset_out, col_idx_out);
  } catch (const std::exception& ex) {
    if (f) fclose(f);
    return lz_fail(LZ_ERR_ALLOC, "lz_csr_read_text: %s", ex.what());
  }
}

extern "C" int lz_csr_write_text(const char* path, uint64_t n, const uint32_t* row_offset, const uint32_t* col_idx) {
  if (!path || !row_offset || !col_idx) return lz_fail(LZ_ERR_ARG, "null argument");
  FILE* f = fopen(path, "w");
  if (!f) return lz_fail(LZ_ERR_IO, "cannot create %s", path);
  uint64_t upper = 0;
  for (uint64_t i = 0; i < n; i++)
    for (uint32_t j = row_offset[i]; j < row_offset[i + 1]; j++) upper += (i < col_idx[j]);
  fprintf(f, "%llu %llu %llu\n", (unsigned long long)n, (unsigned long long)n, (unsigned long long)upper);
  for (uint64_t i = 0; i < n; i++)                  // same "col row" line order as write_matrix_to_file (adjMatrix.cc:62-66)
    for (uint32_t j = row_offset[i]; j < row_offset[i + 1]; j++)
      if (i < col_idx[j]) fprintf(f, "%u %llu\n", col_idx[j] + 1, (unsigned long long)(i + 1));
  if (fclose(f)) return lz_fail(LZ_ERR_IO, "write to %s failed", path);
  return LZ_OK;
}

static const char kMagic[8] = {'L', 'Z', 'C', 'S', 'R', '1', 0, 0};

extern "C" int lz_csr_write_bin(const char* path, uint64_t n, const uint32_t* row_offset, const uint32_t* col_idx) {
  if (!path || !row_offset || !col_idx) return lz_fail(LZ_ERR_ARG, "null argument");
  FILE* f = fopen(path, "wb");
  if (!f) return lz_fail(LZ_ERR_IO, "cannot create %s", path);
  uint64_t nnz = row_offset[n];
  bool ok = fwrite(kMagic, 1, 8, f) == 8 && fwrite(&n, 8, 1, f) == 1 && fwrite(&nnz, 8, 1, f) == 1 &&
            fwrite(row_offset, 4, n + 1, f) == n + 1 && fwrite(col_idx, 4, nnz, f) == nnz;
  if (fclose(f) || !ok) return lz_fail(LZ_ERR_IO, "write to %s failed", path);
  return LZ_OK;
}

extern "C" int lz_csr_read_bin(const char* path, uint64_t* n_out, uint64_t* nnz_out, uint32_t** row_offset_out,
                               uint32_t** col_idx_out) {
  if (!path || !n_out || !nnz_out || !row_offset_out || !col_idx_out) return lz_fail(LZ_ERR_ARG, "null argument");
  FILE* f = fopen(path, "rb");
  if (!f) return lz_fail(LZ_ERR_IO, "cannot open %s", path);
  char magic[8];
  uint64_t n = 0, nnz = 0;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kMagic, 8) || fread(&n, 8, 1, f) != 1 || fread(&nnz, 8, 1, f) != 1) {
    fclose(f);
    return lz_fail(LZ_ERR_IO, "%s: not an LZCSR1 file", path);
  }
  uint32_t* ro = (uint32_t*)malloc((n + 1) * 4);
  uint32_t* ci = (uint32_t*)malloc((nnz ? nnz : 1) * 4);
  if (!ro || !ci) { free(ro); free(ci); fclose(f); return lz_fail(LZ_ERR_ALLOC, "host allocation of CSR failed"); }
  bool ok = fread(ro, 4, n + 1, f) == n + 1 && fread(ci, 4, nnz, f) == nnz;
  fclose(f);
  if (!ok) { free(ro); free(ci); return lz_fail(LZ_ERR_IO, "%s: truncated", path); }
  *n_out = n; *nnz_out = nnz; *row_offset_out = ro; *col_idx_out = ci;
  return LZ_OK;
}
