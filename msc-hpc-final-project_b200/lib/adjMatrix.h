// adjMatrix.h — host mirror of the reference's adjacency-matrix class (parallel-final/lib/adjMatrix.h:19-117).
// Same constructors, accessors and friends; storage is the same value-less CSR (row_offset / col_idx, 32-bit).
// Construction goes through the C ABI (include/lz.h): edges are sorted/de-duplicated as 64-bit keys instead of a
// std::set<Edge> (adjMatrix.cc:21-46), and the generators are seeded and replayable (make_graph.cc:21-113 draws from
// std::random_device). Nothing here touches the GPU; lanczosDecomp uploads the CSR.
#ifndef LZ_ADJ_MATRIX_H
#define LZ_ADJ_MATRIX_H

#include <cassert>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <string>

#include "../../include/lz.h"

template <typename T> class eigenDecomp;
template <typename T> class lanczosDecomp;

class adjMatrix {
 private:
  unsigned* row_offset = nullptr;  // JA  (malloc'ed by the C ABI; released with lz_free_host)
  unsigned* col_idx = nullptr;     // IA
  unsigned n = 0;                  // number of nodes
  unsigned edge_count = 0;         // number of undirected edges (after de-duplication, as adjMatrix.cc:44)
  unsigned barabasi_degree = 0;
  char matrix_type = 'f';          // 'b' Barabasi-Albert, 'r' random G(n,m), 'f' read from file, 'm' R-MAT, 'd' banded

  void populate_sparse_matrix(std::ifstream&);
  void generate_sparse_matrix(const char c);
  void random_adj();
  void barabasi(const unsigned m);
  void adopt(uint64_t nn, uint64_t nnz, uint32_t* ro, uint32_t* ci);
  void release();

 public:
  static uint64_t generator_seed;  // seed used by the generating constructors (the reference is unseeded)

  adjMatrix() = default;
  adjMatrix(const unsigned N, const unsigned E, std::ifstream& f) : n{N}, edge_count{E}, matrix_type{'f'} {
    populate_sparse_matrix(f);  // reads E lines "col row" (1-based) that follow the "n n E" header
  }
  adjMatrix(const unsigned N, const unsigned m, const char c) : n{N}, barabasi_degree{m}, matrix_type{c} {
    generate_sparse_matrix(c);  // 'b': Barabasi-Albert with minimum degree m
  }
  adjMatrix(const unsigned N, const unsigned E) : n{N}, edge_count{E}, matrix_type{'r'} {
    generate_sparse_matrix('r');  // G(n,m)-style random graph
  }
  // Seeded generators used by the BASELINE configs (no reference equivalent)
  static adjMatrix rmat(unsigned scale, unsigned edge_factor, uint64_t seed);
  static adjMatrix banded(unsigned N, uint64_t seed);
  static adjMatrix from_spec(const lz_graph_spec& spec);

  adjMatrix(const adjMatrix&) = delete;
  adjMatrix& operator=(const adjMatrix&) = delete;
  adjMatrix(adjMatrix&& rhs) noexcept { *this = std::move(rhs); }
  adjMatrix& operator=(adjMatrix&& rhs) noexcept {  // move assignment, as the reference's (adjMatrix.h:80-93)
    if (this != &rhs) {
      release();
      row_offset = rhs.row_offset; col_idx = rhs.col_idx; n = rhs.n; edge_count = rhs.edge_count;
      matrix_type = rhs.matrix_type; barabasi_degree = rhs.barabasi_degree;
      rhs.row_offset = nullptr; rhs.col_idx = nullptr;
    }
    return *this;
  }
  ~adjMatrix() { release(); }

  unsigned get_n() const { return n; }
  unsigned get_edges() const { return edge_count; }
  const unsigned* get_row_offset() const { return row_offset; }
  const unsigned* get_col_idx() const { return col_idx; }

  void write_matrix_to_file();                       // "../data/generated/<type>n<N>e<E>", reference text format
  void write_matrix_to_file(const std::string& path);
  void print_full() const;

  friend std::ostream& operator<<(std::ostream&, const adjMatrix&);
  template <typename T> friend void multOut(lanczosDecomp<T>&, eigenDecomp<T>&, adjMatrix&, bool);
  template <typename T> friend class lanczosDecomp;
};
#endif
