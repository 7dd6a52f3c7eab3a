// lz_gen.h — counter-based synthetic-graph edge generators, shared verbatim by the host builder (lz_host.cc) and the
// device builder (lz_graph.cu) so both produce bit-identical CSR. Integer arithmetic only.
//
// These replace adjMatrix::random_adj / adjMatrix::barabasi (reference parallel-final/lib/make_graph.cc:21-113), which
// draw from std::random_device (not replayable) and insert into a std::set (O(E log E), ~48 B per entry).
// Every generator maps a candidate index e in [0, lz_gen_candidates(spec)) to one undirected edge (u,v) or to
// "none" (u == v). The builder symmetrises, sorts, and removes duplicates.
#pragma once
#include <stdint.h>
#include "../../include/lz.h"

#if defined(__CUDACC__)
#define LZ_HD __host__ __device__ __forceinline__
#else
#define LZ_HD inline
#endif

LZ_HD uint64_t lz_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// Random 64-bit word number `w` of stream `idx` under `seed`.
LZ_HD uint64_t lz_rand64(uint64_t seed, uint64_t idx, uint64_t w) {
  return lz_mix64(lz_mix64(seed ^ (idx * 0xD6E8FEB86659FD93ull)) + w * 0xA0761D6478BD642Full);
}
// Unbiased-enough reduction of a 64-bit word onto [0, n) (multiply-high).
LZ_HD uint64_t lz_reduce(uint64_t r, uint64_t n) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(r, n);
#else
  return (uint64_t)(((unsigned __int128)r * n) >> 64);
#endif
}
// Bijection on [0, 2^bits): multiply by odd constants / xor-shift / add, all invertible modulo 2^bits.
LZ_HD uint64_t lz_relabel_pow2(uint64_t v, uint32_t bits, uint64_t seed) {
  const uint64_t mask = (bits >= 64) ? ~0ull : ((1ull << bits) - 1);
  const uint32_t h = bits > 1 ? bits / 2 : 1;
  const uint64_t k0 = lz_mix64(seed ^ 0x5851F42D4C957F2Dull) | 1ull;
  const uint64_t k1 = lz_mix64(seed ^ 0x14057B7EF767814Full) | 1ull;
  const uint64_t c0 = lz_mix64(seed ^ 0x2545F4914F6CDD1Dull);
  v = (v * k0 + c0) & mask;
  v ^= v >> h;
  v = (v * k1) & mask;
  v ^= v >> h;
  v = (v * k0 + (c0 >> 7)) & mask;
  v ^= v >> h;
  return v & mask;
}

struct lz_gen_params {
  uint32_t kind, scale;
  uint64_t n, m, seed;       // m = number of candidate edges
  uint32_t t_a, t_ab, t_abc; // RMAT quadrant thresholds on a 32-bit scale
  uint64_t band;             // BAND: b = ceil(sqrt(n))
  uint32_t t_keep;           // BAND: keep threshold (0.98 * 2^32)
};

inline uint64_t lz_isqrt_ceil(uint64_t n) {
  uint64_t r = 0;
  while (r * r < n) r++;
  return r;
}

// Returns 0 on success; fills p. Host only.
inline int lz_gen_prepare(const lz_graph_spec* s, lz_gen_params* p) {
  p->kind = s->kind; p->scale = s->scale; p->seed = s->seed; p->band = 0; p->t_keep = 0;
  p->t_a = p->t_ab = p->t_abc = 0;
  if (s->kind == LZ_GRAPH_RMAT) {
    if (s->scale < 2 || s->scale > 31) return -1;
    p->n = 1ull << s->scale;
    p->m = s->param_a * p->n;
    double a = s->rmat_a, b = s->rmat_b, c = s->rmat_c;
    if (a == 0.0 && b == 0.0 && c == 0.0) { a = 0.45; b = 0.15; c = 0.15; }
    if (a <= 0 || b < 0 || c < 0 || a + b + c >= 1.0) return -1;
    p->t_a = (uint32_t)(a * 4294967296.0);
    p->t_ab = (uint32_t)((a + b) * 4294967296.0);
    p->t_abc = (uint32_t)((a + b + c) * 4294967296.0);
  } else if (s->kind == LZ_GRAPH_ER) {
    if (s->n < 2 || s->n > 0xFFFFFFFFull) return -1;
    p->n = s->n; p->m = s->param_a;
  } else if (s->kind == LZ_GRAPH_BAND) {
    if (s->n < 4 || s->n > 0xFFFFFFFFull) return -1;
    p->n = s->n;
    p->band = lz_isqrt_ceil(s->n);
    p->m = 2 * s->n + s->n / 64;
    p->t_keep = (uint32_t)(0.98 * 4294967296.0);
  } else {
    return -1;
  }
  return 0;
}

// Candidate e -> (u,v). u == v means "no edge". One extra candidate index (e == m) always yields (0, n-1) so the last
// vertex is never isolated (the reference text loader mis-handles trailing empty rows, adjMatrix.cc:36-43).
LZ_HD void lz_gen_edge(const lz_gen_params& p, uint64_t e, uint32_t* u_out, uint32_t* v_out) {
  uint64_t u = 0, v = 0;
  if (e >= p.m) {
    u = 0; v = p.n - 1;
  } else if (p.kind == LZ_GRAPH_RMAT) {
    for (uint32_t l = 0; l < p.scale; l += 2) {
      uint64_t r = lz_rand64(p.seed, e, l >> 1);
      uint32_t r0 = (uint32_t)r, r1 = (uint32_t)(r >> 32);
      uint32_t ub = (r0 >= p.t_ab), vb = (r0 >= p.t_a && r0 < p.t_ab) || (r0 >= p.t_abc);
      u = (u << 1) | ub; v = (v << 1) | vb;
      if (l + 1 < p.scale) {
        ub = (r1 >= p.t_ab); vb = (r1 >= p.t_a && r1 < p.t_ab) || (r1 >= p.t_abc);
        u = (u << 1) | ub; v = (v << 1) | vb;
      }
    }
    u = lz_relabel_pow2(u, p.scale, p.seed);
    v = lz_relabel_pow2(v, p.scale, p.seed);
  } else if (p.kind == LZ_GRAPH_ER) {
    u = lz_reduce(lz_rand64(p.seed, e, 0), p.n);
    v = lz_reduce(lz_rand64(p.seed, e, 1), p.n);
  } else { // BAND
    if (e < p.n) {                       // (i, i+1)
      uint64_t i = e;
      bool keep = (uint32_t)lz_rand64(p.seed, e, 0) < p.t_keep;
      if (keep && i + 1 < p.n) { u = i; v = i + 1; }
    } else if (e < 2 * p.n) {            // (i, i+b)
      uint64_t i = e - p.n;
      bool keep = (uint32_t)lz_rand64(p.seed, e, 0) < p.t_keep;
      if (keep && i + p.band < p.n) { u = i; v = i + p.band; }
    } else {                             // random chord
      u = lz_reduce(lz_rand64(p.seed, e, 0), p.n);
      v = lz_reduce(lz_rand64(p.seed, e, 1), p.n);
    }
  }
  *u_out = (uint32_t)u; *v_out = (uint32_t)v;
}
