// adjMatrix.cc — see adjMatrix.h. Reference counterpart: parallel-final/lib/adjMatrix.cc, make_graph.cc.
#include "adjMatrix.h"

#include <cstdlib>
#include <vector>

uint64_t adjMatrix::generator_seed = 1;

static void die(const char* what) {
  std::cerr << what << ": " << lz_last_error() << '\n';
  std::abort();  // the reference asserts on I/O problems (main.cu:61) — same severity, with a message
}

void adjMatrix::release() {
  lz_free_host(row_offset);
  lz_free_host(col_idx);
  row_offset = col_idx = nullptr;
}

void adjMatrix::adopt(uint64_t nn, uint64_t nnz, uint32_t* ro, uint32_t* ci) {
  release();
  n = (unsigned)nn;
  edge_count = (unsigned)(nnz / 2);
  row_offset = ro;
  col_idx = ci;
}

// File body: edge_count lines "col row", 1-based (adjMatrix.cc:29-34). Both orientations are stored, duplicates collapse.
void adjMatrix::populate_sparse_matrix(std::ifstream& f) {
  std::vector<uint32_t> u(edge_count), v(edge_count);
  for (unsigned i = 0; i < edge_count; i++) {
    unsigned col = 0, row = 0;
    f >> col >> row;
    if (f.fail() || col < 1 || row < 1 || col > n || row > n) {
      std::cerr << "adjMatrix: bad edge on line " << i + 2 << " of the input file\n";
      std::abort();
    }
    u[i] = row - 1;
    v[i] = col - 1;
  }
  uint64_t nnz = 0;
  uint32_t *ro = nullptr, *ci = nullptr;
  if (lz_csr_from_edges(n, edge_count, u.data(), v.data(), &nnz, &ro, &ci) != LZ_OK) die("adjMatrix(file)");
  adopt(n, nnz, ro, ci);
}

void adjMatrix::generate_sparse_matrix(const char c) {
  switch (c) {
    case 'b': barabasi(barabasi_degree); break;
    case 'r': random_adj(); break;
    default:
      std::cerr << "adjMatrix: unknown generator '" << c << "'\n";   // the reference's 's' (stencil) case is an empty stub
      std::abort();
  }
}

// G(n,m)-style: edge_count candidate pairs drawn uniformly (make_graph.cc:21-47), deterministic under generator_seed.
void adjMatrix::random_adj() {
  lz_graph_spec s{};
  s.kind = LZ_GRAPH_ER; s.n = n; s.param_a = edge_count; s.seed = generator_seed;
  *this = from_spec(s);
  matrix_type = 'r';
}

// Barabasi-Albert preferential attachment with minimum degree m (make_graph.cc:57-113), in O(E) with the
// repeated-endpoints urn instead of the reference's O(n * E) scan, and replayable.
void adjMatrix::barabasi(const unsigned m_in) {
  const unsigned m = (m_in > n - 1) ? n - 1 : (m_in < 1 ? 1 : m_in);
  std::vector<uint32_t> u, v, urn;
  u.reserve((size_t)n * m); v.reserve((size_t)n * m); urn.reserve(2 * (size_t)n * m);
  for (unsigned a = 0; a <= m; a++)          // seed clique on m + 1 vertices
    for (unsigned b = a + 1; b <= m; b++) { u.push_back(a); v.push_back(b); urn.push_back(a); urn.push_back(b); }
  uint64_t state = generator_seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
  auto next = [&]() {  // splitmix64
    uint64_t z = (state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  };
  std::vector<uint32_t> picked;
  for (unsigned node = m + 1; node < n; node++) {
    picked.clear();
    while (picked.size() < m) {
      uint32_t t = urn[next() % urn.size()];
      bool dup = false;
      for (uint32_t p : picked) dup |= (p == t);
      if (!dup) picked.push_back(t);
    }
    for (uint32_t t : picked) { u.push_back(node); v.push_back(t); urn.push_back(node); urn.push_back(t); }
  }
  uint64_t nnz = 0;
  uint32_t *ro = nullptr, *ci = nullptr;
  if (lz_csr_from_edges(n, u.size(), u.data(), v.data(), &nnz, &ro, &ci) != LZ_OK) die("adjMatrix(barabasi)");
  adopt(n, nnz, ro, ci);
}

adjMatrix adjMatrix::from_spec(const lz_graph_spec& spec) {
  adjMatrix A;
  uint64_t nn = 0, nnz = 0;
  uint32_t *ro = nullptr, *ci = nullptr;
  if (lz_graph_generate_host(&spec, &nn, &nnz, &ro, &ci) != LZ_OK) die("adjMatrix(generator)");
  A.adopt(nn, nnz, ro, ci);
  A.matrix_type = spec.kind == LZ_GRAPH_RMAT ? 'm' : (spec.kind == LZ_GRAPH_BAND ? 'd' : 'r');
  return A;
}

adjMatrix adjMatrix::rmat(unsigned scale, unsigned edge_factor, uint64_t seed) {
  lz_graph_spec s{};
  s.kind = LZ_GRAPH_RMAT; s.scale = scale; s.param_a = edge_factor; s.seed = seed;
  return from_spec(s);
}

adjMatrix adjMatrix::banded(unsigned N, uint64_t seed) {
  lz_graph_spec s{};
  s.kind = LZ_GRAPH_BAND; s.n = N; s.seed = seed;
  return from_spec(s);
}

void adjMatrix::write_matrix_to_file(const std::string& path) {
  if (lz_csr_write_text(path.c_str(), n, row_offset, col_idx) != LZ_OK) die("write_matrix_to_file");
}

void adjMatrix::write_matrix_to_file() {  // same naming scheme as adjMatrix.cc:53-58
  std::string type{matrix_type};
  std::string filename = "../data/generated/" + type + "n" + std::to_string(n) + "e" + std::to_string(edge_count);
  std::cout << "Filename: " << filename << '\n';
  write_matrix_to_file(filename);
}

void adjMatrix::print_full() const {
  for (unsigned i = 0; i < n; i++) {
    unsigned j = row_offset[i];
    for (unsigned c = 0; c < n; c++) {
      if (j < row_offset[i + 1] && col_idx[j] == c) { std::cout << "1 "; j++; }
      else std::cout << "0 ";
    }
    std::cout << '\n';
  }
}

std::ostream& operator<<(std::ostream& os, const adjMatrix& A) {  // adjMatrix.cc:115-129
  os << "JA\n";
  for (unsigned i = 0; i < A.edge_count * 2; ++i) os << A.col_idx[i] << " ";
  os << "\nIA\n";
  for (unsigned i = 0; i < A.n + 1; ++i) os << A.row_offset[i] << " ";
  return os;
}
