// lz_graph.cu — adjacency matrix on the device: synthetic-graph construction, relabelling, sharding, SpMV launch plan.
//
// Replaces, for the B200 path, the host loader adjMatrix::populate_sparse_matrix (reference
// parallel-final/lib/adjMatrix.cc:21-46; std::set, 8-55 s per graph in the published runs) and the per-GPU offset
// rebasing of parallel-two-cards/lib/cu_lanczos.cu:21-27,62-69. CUB is used for the (setup-time) sorts and scans; the
// hot-path kernels are in lz_kernels.cu.
//
// Internal vertex order. Vertices are sorted by degree (descending, ties by original id) and dealt cyclically to the
// `world` ranks: the vertex at sorted position s lives on rank s % world at local row s / world and gets the new global
// id (s % world) * n_loc + s / world. Consequences: (1) every rank owns the same number of rows with a near-identical
// degree profile, so nnz, vector work and basis memory are balanced without a separate partitioner; (2) local rows are
// sorted by length, so each SpMV degree bin is a contiguous row range served by one lanes-per-row variant with no
// intra-warp imbalance; (3) high-degree (= most frequently gathered) entries of x are contiguous, which keeps the hot
// part of x resident in L2; (4) ownership blocks are equal-sized and contiguous in the new numbering, so the Krylov
// vector is exchanged with a plain ncclAllGather.
#include "lz_ctx.h"
#include "lz_gen.h"

#include <cub/cub.cuh>

namespace {

__global__ void k_gen_keys(lz_gen_params p, uint64_t* __restrict__ keys) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e > p.m) return;
  uint32_t u, v;
  lz_gen_edge(p, e, &u, &v);
  uint64_t a = ~0ull, b = ~0ull;   // sentinel sorts last
  if (u != v) { a = ((uint64_t)u << 32) | v; b = ((uint64_t)v << 32) | u; }
  keys[2 * e] = a;
  keys[2 * e + 1] = b;
}

// ro[r] = number of keys with row < r  (keys sorted, unique, no sentinel)
__global__ void k_row_offsets_from_keys(const uint64_t* __restrict__ keys, uint64_t nnz, uint64_t n, uint32_t* __restrict__ ro) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  uint64_t target = r << 32, lo = 0, hi = nnz;
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  ro[r] = (uint32_t)lo;
}

__global__ void k_low32(const uint64_t* __restrict__ keys, uint64_t nnz, uint32_t* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) out[i] = (uint32_t)keys[i];
}

// sort key for the degree ordering: descending degree == ascending (~deg); value = original id
__global__ void k_degree_keys(const uint32_t* __restrict__ ro, uint64_t n, uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  key[v] = ~(ro[v + 1] - ro[v]);
  val[v] = (uint32_t)v;
}

__global__ void k_fill_u32(uint32_t* p, uint64_t n, uint32_t v) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// sorted position s -> new id; builds old2new and new2old
__global__ void k_relabel(const uint32_t* __restrict__ sorted_old, uint64_t n, uint32_t world, uint64_t n_loc,
                          uint32_t* __restrict__ old2new, uint32_t* __restrict__ new2old) {
  uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  uint32_t old = sorted_old[s];
  uint32_t nw = (uint32_t)((s % world) * n_loc + s / world);
  old2new[old] = nw;
  new2old[nw] = old;
}

// local row l of rank r is sorted position l*world + r
__global__ void k_local_lengths(const uint32_t* __restrict__ sorted_old, const uint32_t* __restrict__ ro, uint64_t n, uint32_t world,
                                uint32_t rank, uint64_t n_loc, uint32_t* __restrict__ len) {
  uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_loc) return;
  uint64_t s = l * world + rank;
  uint32_t d = 0;
  if (s < n) { uint32_t old = sorted_old[s]; d = ro[old + 1] - ro[old]; }
  len[l] = d;
}

// one warp per local row: keys[row_ptr[l] + j] = (l << 32) | old2new[ci[ro[old] + j]]
__global__ void k_local_keys(const uint32_t* __restrict__ sorted_old, const uint32_t* __restrict__ ro, const uint32_t* __restrict__ ci,
                             const uint32_t* __restrict__ old2new, const uint32_t* __restrict__ row_ptr, uint64_t n, uint32_t world,
                             uint32_t rank, uint64_t n_loc, uint64_t* __restrict__ keys) {
  uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t lane = threadIdx.x & 31;
  if (warp >= n_loc) return;
  uint64_t s = warp * world + rank;
  if (s >= n) return;
  uint32_t old = sorted_old[s];
  uint32_t b = ro[old], e = ro[old + 1], dst = row_ptr[warp];
  for (uint32_t j = b + lane; j < e; j += 32) keys[dst + (j - b)] = (warp << 32) | old2new[ci[j]];
}

// first local row whose length is <= thr[t]  (lengths are non-increasing)
__global__ void k_bin_bounds(const uint32_t* __restrict__ row_ptr, uint32_t n_loc, const uint32_t* __restrict__ thr, uint32_t nthr,
                             uint32_t* __restrict__ out) {
  uint32_t t = threadIdx.x;
  if (t >= nthr) return;
  uint32_t lo = 0, hi = n_loc;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    uint32_t len = row_ptr[mid + 1] - row_ptr[mid];
    if (len > thr[t]) lo = mid + 1; else hi = mid;
  }
  out[t] = lo;
}

struct IsEmpty { __host__ __device__ uint64_t operator()(uint32_t k) const { return k == 0xFFFFFFFFu ? 1ull : 0ull; } };

inline unsigned grid_for(uint64_t items, unsigned block) { return (unsigned)((items + block - 1) / block); }

struct DevBuf {   // frees on scope exit unless released
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  template <class T> T* as() { return (T*)p; }
  void* release() { void* q = p; p = nullptr; return q; }
};

int bits_for(uint64_t v) { int b = 0; while ((1ull << b) < v && b < 63) b++; return b < 1 ? 1 : b; }

}  // namespace

void lz_free_graph(lz_ctx* c) {
  cudaFree(c->orig_ro); cudaFree(c->orig_ci); cudaFree(c->row_ptr); cudaFree(c->col); cudaFree(c->new2old);
  c->orig_ro = c->orig_ci = c->row_ptr = c->col = c->new2old = nullptr;
  c->n = c->nnz = c->n_loc = c->nnz_loc = 0;
}

static void make_plan(const uint32_t* bounds /* [5]: first row with len <= 32,16,8,4,2 */, uint32_t n_loc, lz_spmv_plan* plan) {
  // len > 32 -> 32 lanes, (16,32] -> 16, (8,16] -> 8, (4,8] -> 4, (2,4] -> 2, <= 2 -> 1 lane per row
  const uint32_t lg[6] = {5, 4, 3, 2, 1, 0};
  uint32_t begin = 0, blocks = 0;
  plan->nbins = 0;
  for (int b = 0; b < 6; b++) {
    uint32_t end = (b < 5) ? bounds[b] : n_loc;
    if (end > begin) {
      lz_spmv_bin& bin = plan->bin[plan->nbins++];
      bin.row_begin = begin; bin.row_end = end; bin.log2_lanes = lg[b]; bin.block_begin = blocks;
      uint64_t threads = (uint64_t)(end - begin) << lg[b];
      blocks += (uint32_t)((threads + 255) / 256);
    }
    begin = end;
  }
  plan->nblocks = blocks;
}

// Takes ownership of ro_d / ci_d (original-order CSR on the device).
int lz_ingest_device_csr(lz_ctx* c, uint64_t n, uint64_t nnz, uint32_t* ro_d, uint32_t* ci_d) {
  lz_free_graph(c);
  c->orig_ro = ro_d; c->orig_ci = ci_d;
  c->n = n; c->nnz = nnz;
  const uint32_t world = (uint32_t)c->world, rank = (uint32_t)c->rank;
  const uint64_t n_loc = (n + world - 1) / world, n_pad = n_loc * world;
  if (n_pad > 0xFFFFFFFEull) return lz_fail(LZ_ERR_ARG, "n = %llu too large for 32-bit vertex ids", (unsigned long long)n);
  c->n_loc = n_loc;
  cudaStream_t st = c->stream;

  DevBuf key_in, key_out, val_in, sorted_old, old2new, len, tmp;
  LZ_CUDA(cudaMalloc(&key_in.p, n * 4)); LZ_CUDA(cudaMalloc(&key_out.p, n * 4));
  LZ_CUDA(cudaMalloc(&val_in.p, n * 4)); LZ_CUDA(cudaMalloc(&sorted_old.p, n * 4));
  LZ_CUDA(cudaMalloc(&old2new.p, n * 4)); LZ_CUDA(cudaMalloc(&len.p, (n_loc + 1) * 4));
  LZ_CUDA(cudaMalloc((void**)&c->new2old, n_pad * 4));
  LZ_CUDA(cudaMalloc((void**)&c->row_ptr, (n_loc + 1) * 4));

  // 1. degree ordering (stable radix sort => ties keep ascending original id)
  k_degree_keys<<<grid_for(n, 256), 256, 0, st>>>(ro_d, n, key_in.as<uint32_t>(), val_in.as<uint32_t>());
  size_t tb = 0;
  LZ_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key_in.as<uint32_t>(), key_out.as<uint32_t>(), val_in.as<uint32_t>(),
                                          sorted_old.as<uint32_t>(), (int64_t)n, 0, 32, st));
  LZ_CUDA(cudaMalloc(&tmp.p, tb ? tb : 1));
  LZ_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key_in.as<uint32_t>(), key_out.as<uint32_t>(), val_in.as<uint32_t>(),
                                          sorted_old.as<uint32_t>(), (int64_t)n, 0, 32, st));
  // max degree / isolated vertices from the sorted keys (key = ~deg, ascending)
  uint32_t kfirst = 0;
  LZ_CUDA(cudaMemcpyAsync(&kfirst, key_out.p, 4, cudaMemcpyDeviceToHost, st));

  // 2. relabel
  k_fill_u32<<<grid_for(n_pad, 256), 256, 0, st>>>(c->new2old, n_pad, 0xFFFFFFFFu);
  k_relabel<<<grid_for(n, 256), 256, 0, st>>>(sorted_old.as<uint32_t>(), n, world, n_loc, old2new.as<uint32_t>(), c->new2old);

  // 3. local row pointer
  k_local_lengths<<<grid_for(n_loc, 256), 256, 0, st>>>(sorted_old.as<uint32_t>(), ro_d, n, world, rank, n_loc, len.as<uint32_t>());
  LZ_CUDA(cudaMemsetAsync(len.as<uint32_t>() + n_loc, 0, 4, st));
  {
    DevBuf t2; size_t b2 = 0;
    LZ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b2, len.as<uint32_t>(), c->row_ptr, (int64_t)(n_loc + 1), st));
    LZ_CUDA(cudaMalloc(&t2.p, b2 ? b2 : 1));
    LZ_CUDA(cub::DeviceScan::ExclusiveSum(t2.p, b2, len.as<uint32_t>(), c->row_ptr, (int64_t)(n_loc + 1), st));
    LZ_CUDA(cudaStreamSynchronize(st));
  }
  // number of isolated vertices: rows with ~deg == 0xFFFFFFFF are at the end of key_out -> count via binary search on host-side copy is
  // overkill; use the original row offsets instead (empty rows = n - #rows with ro[v+1] > ro[v]); computed with a device reduction below.
  uint32_t nnz_loc32 = 0;
  LZ_CUDA(cudaMemcpy(&nnz_loc32, c->row_ptr + n_loc, 4, cudaMemcpyDeviceToHost));
  {
    // the scan is 32-bit: guard against wrap-around by summing lengths in 64 bit
    DevBuf t3, s64; size_t b3 = 0;
    LZ_CUDA(cudaMalloc(&s64.p, 8));
    cub::TransformInputIterator<uint64_t, cub::CastOp<uint64_t>, const uint32_t*> it(len.as<uint32_t>(), cub::CastOp<uint64_t>());
    LZ_CUDA(cub::DeviceReduce::Sum(nullptr, b3, it, s64.as<uint64_t>(), (int64_t)n_loc, st));
    LZ_CUDA(cudaMalloc(&t3.p, b3 ? b3 : 1));
    LZ_CUDA(cub::DeviceReduce::Sum(t3.p, b3, it, s64.as<uint64_t>(), (int64_t)n_loc, st));
    uint64_t total = 0;
    LZ_CUDA(cudaMemcpyAsync(&total, s64.p, 8, cudaMemcpyDeviceToHost, st));
    LZ_CUDA(cudaStreamSynchronize(st));
    if (total > 0xFFFFFFFFull) return lz_fail(LZ_ERR_ARG, "local nnz %llu does not fit 32-bit row offsets", (unsigned long long)total);
    c->nnz_loc = total;
  }
  c->max_degree = ~kfirst;

  // 4. local column lists in the new numbering, ascending within each row (one 64-bit radix sort)
  {
    DevBuf k_in, k_out, t4; size_t b4 = 0;
    uint64_t m = c->nnz_loc;
    LZ_CUDA(cudaMalloc(&k_in.p, (m ? m : 1) * 8)); LZ_CUDA(cudaMalloc(&k_out.p, (m ? m : 1) * 8));
    LZ_CUDA(cudaMalloc((void**)&c->col, (m ? m : 1) * 4));
    if (m) {
      k_local_keys<<<grid_for(n_loc * 32, 256), 256, 0, st>>>(sorted_old.as<uint32_t>(), ro_d, ci_d, old2new.as<uint32_t>(), c->row_ptr, n,
                                                                world, rank, n_loc, k_in.as<uint64_t>());
      int end_bit = 32 + bits_for(n_loc);
      LZ_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, b4, k_in.as<uint64_t>(), k_out.as<uint64_t>(), (int64_t)m, 0, end_bit, st));
      LZ_CUDA(cudaMalloc(&t4.p, b4 ? b4 : 1));
      LZ_CUDA(cub::DeviceRadixSort::SortKeys(t4.p, b4, k_in.as<uint64_t>(), k_out.as<uint64_t>(), (int64_t)m, 0, end_bit, st));
      k_low32<<<grid_for(m, 256), 256, 0, st>>>(k_out.as<uint64_t>(), m, c->col);
    }
    LZ_CUDA(cudaStreamSynchronize(st));
  }

  // 5. SpMV plan from the (non-increasing) local row lengths
  {
    const uint32_t thr_h[6] = {32, 16, 8, 4, 2, 0};
    DevBuf thr_d, out_d;
    uint32_t out_h[6];
    LZ_CUDA(cudaMalloc(&thr_d.p, sizeof(thr_h))); LZ_CUDA(cudaMalloc(&out_d.p, sizeof(out_h)));
    LZ_CUDA(cudaMemcpyAsync(thr_d.p, thr_h, sizeof(thr_h), cudaMemcpyHostToDevice, st));
    k_bin_bounds<<<1, 32, 0, st>>>(c->row_ptr, (uint32_t)n_loc, thr_d.as<uint32_t>(), 6, out_d.as<uint32_t>());
    LZ_CUDA(cudaMemcpyAsync(out_h, out_d.p, sizeof(out_h), cudaMemcpyDeviceToHost, st));
    LZ_CUDA(cudaStreamSynchronize(st));
    make_plan(out_h, (uint32_t)n_loc, &c->plan_auto);
    // warp-per-row for everything
    c->plan_warp.nbins = 1;
    c->plan_warp.bin[0] = {0u, (uint32_t)n_loc, 5u, 0u};
    c->plan_warp.nblocks = (uint32_t)((n_loc * 32 + 255) / 256);
    c->plan = (c->spmv_variant == LZ_SPMV_WARP) ? c->plan_warp : c->plan_auto;
    // isolated vertices: local rows with length 0 start at out_h[5]; every rank sees ~1/world of them. Global count from the
    // sorted degree keys: positions s with deg == 0 are the tail; first such s = lower_bound over all ranks -> computed on rank-agnostic data:
    // empty_global = n - (#vertices with deg > 0). Use the degree keys (ascending ~deg): deg == 0 <=> key == 0xFFFFFFFF.
    DevBuf cnt; size_t b5 = 0; DevBuf t5;
    LZ_CUDA(cudaMalloc(&cnt.p, 8));
    // count keys equal to 0xFFFFFFFF via a transform-reduce
    cub::TransformInputIterator<uint64_t, IsEmpty, const uint32_t*> it(key_out.as<uint32_t>(), IsEmpty());
    LZ_CUDA(cub::DeviceReduce::Sum(nullptr, b5, it, cnt.as<uint64_t>(), (int64_t)n, st));
    LZ_CUDA(cudaMalloc(&t5.p, b5 ? b5 : 1));
    LZ_CUDA(cub::DeviceReduce::Sum(t5.p, b5, it, cnt.as<uint64_t>(), (int64_t)n, st));
    LZ_CUDA(cudaMemcpyAsync(&c->empty_rows, cnt.p, 8, cudaMemcpyDeviceToHost, st));
    LZ_CUDA(cudaStreamSynchronize(st));
  }
  LZ_CUDA(cudaGetLastError());
  return LZ_OK;
}

extern "C" int lz_csr_upload(lz_ctx* c, uint64_t n, const uint32_t* row_offset, const uint32_t* col_idx) {
  if (!c || !row_offset || !col_idx || n == 0) return lz_fail(LZ_ERR_ARG, "lz_csr_upload: null argument or n == 0");
  LZ_CUDA(cudaSetDevice(c->device));
  uint64_t nnz = row_offset[n];
  uint32_t *ro_d = nullptr, *ci_d = nullptr;
  LZ_CUDA(cudaMalloc((void**)&ro_d, (n + 1) * 4));
  if (cudaMalloc((void**)&ci_d, (nnz ? nnz : 1) * 4) != cudaSuccess) { cudaFree(ro_d); return lz_fail(LZ_ERR_ALLOC, "device allocation of col_idx failed"); }
  // the three H2D copies of cu_lanczos.cu:88-90, minus the vector
  cudaError_t e1 = cudaMemcpyAsync(ro_d, row_offset, (n + 1) * 4, cudaMemcpyHostToDevice, c->stream);
  cudaError_t e2 = cudaMemcpyAsync(ci_d, col_idx, nnz * 4, cudaMemcpyHostToDevice, c->stream);
  if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaFree(ro_d); cudaFree(ci_d); return lz_fail(LZ_ERR_CUDA, "H2D copy of CSR failed"); }
  c->have_x = c->have_tridiag = c->have_coef = c->have_ans = false;
  return lz_ingest_device_csr(c, n, nnz, ro_d, ci_d);
}

extern "C" int lz_graph_generate(lz_ctx* c, const lz_graph_spec* spec) {
  if (!c || !spec) return lz_fail(LZ_ERR_ARG, "lz_graph_generate: null argument");
  LZ_CUDA(cudaSetDevice(c->device));
  lz_gen_params p;
  if (lz_gen_prepare(spec, &p)) return lz_fail(LZ_ERR_ARG, "bad graph spec (kind %u)", spec->kind);
  cudaStream_t st = c->stream;
  const uint64_t nkeys = 2 * (p.m + 1);
  DevBuf k_in, k_out, tmp, nsel;
  LZ_CUDA(cudaMalloc(&k_in.p, nkeys * 8)); LZ_CUDA(cudaMalloc(&k_out.p, nkeys * 8)); LZ_CUDA(cudaMalloc(&nsel.p, 8));
  k_gen_keys<<<grid_for(p.m + 1, 256), 256, 0, st>>>(p, k_in.as<uint64_t>());
  size_t tb = 0;
  int end_bit = 64;   // sentinel uses all bits
  LZ_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, k_in.as<uint64_t>(), k_out.as<uint64_t>(), (int64_t)nkeys, 0, end_bit, st));
  LZ_CUDA(cudaMalloc(&tmp.p, tb ? tb : 1));
  LZ_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, k_in.as<uint64_t>(), k_out.as<uint64_t>(), (int64_t)nkeys, 0, end_bit, st));
  size_t ub = 0;
  LZ_CUDA(cub::DeviceSelect::Unique(nullptr, ub, k_out.as<uint64_t>(), k_in.as<uint64_t>(), nsel.as<int64_t>(), (int64_t)nkeys, st));
  if (ub > tb) { cudaFree(tmp.release()); LZ_CUDA(cudaMalloc(&tmp.p, ub)); }
  LZ_CUDA(cub::DeviceSelect::Unique(tmp.p, ub, k_out.as<uint64_t>(), k_in.as<uint64_t>(), nsel.as<int64_t>(), (int64_t)nkeys, st));
  int64_t nuniq = 0;
  uint64_t last = 0;
  LZ_CUDA(cudaMemcpyAsync(&nuniq, nsel.p, 8, cudaMemcpyDeviceToHost, st));
  LZ_CUDA(cudaStreamSynchronize(st));
  if (nuniq > 0) {
    LZ_CUDA(cudaMemcpy(&last, k_in.as<uint64_t>() + (nuniq - 1), 8, cudaMemcpyDeviceToHost));
    if (last == ~0ull) nuniq--;   // drop the sentinel
  }
  uint64_t nnz = (uint64_t)nuniq;
  if (nnz > 0xFFFFFFFFull) return lz_fail(LZ_ERR_ARG, "nnz %llu does not fit 32-bit row offsets", (unsigned long long)nnz);
  cudaFree(k_out.release());
  uint32_t *ro_d = nullptr, *ci_d = nullptr;
  LZ_CUDA(cudaMalloc((void**)&ro_d, (p.n + 1) * 4));
  if (cudaMalloc((void**)&ci_d, (nnz ? nnz : 1) * 4) != cudaSuccess) { cudaFree(ro_d); return lz_fail(LZ_ERR_ALLOC, "device allocation of col_idx failed"); }
  k_row_offsets_from_keys<<<grid_for(p.n + 1, 256), 256, 0, st>>>(k_in.as<uint64_t>(), nnz, p.n, ro_d);
  if (nnz) k_low32<<<grid_for(nnz, 256), 256, 0, st>>>(k_in.as<uint64_t>(), nnz, ci_d);
  LZ_CUDA(cudaStreamSynchronize(st));
  cudaFree(k_in.release()); cudaFree(tmp.release());
  c->have_x = c->have_tridiag = c->have_coef = c->have_ans = false;
  return lz_ingest_device_csr(c, p.n, nnz, ro_d, ci_d);
}

extern "C" int lz_graph_info_get(lz_ctx* c, lz_graph_info* out) {
  if (!c || !out) return lz_fail(LZ_ERR_ARG, "null argument");
  if (!c->row_ptr) return lz_fail(LZ_ERR_ARG, "no graph loaded");
  out->n = c->n; out->nnz = c->nnz; out->n_local = c->n_loc; out->nnz_local = c->nnz_loc;
  out->max_degree = c->max_degree; out->pad_ = 0; out->empty_rows = c->empty_rows;
  return LZ_OK;
}

extern "C" int lz_csr_download(lz_ctx* c, uint32_t* row_offset_out, uint32_t* col_idx_out) {
  if (!c || !row_offset_out || !col_idx_out) return lz_fail(LZ_ERR_ARG, "null argument");
  if (!c->orig_ro) return lz_fail(LZ_ERR_ARG, "no graph loaded");
  LZ_CUDA(cudaSetDevice(c->device));
  LZ_CUDA(cudaMemcpyAsync(row_offset_out, c->orig_ro, (c->n + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
  LZ_CUDA(cudaMemcpyAsync(col_idx_out, c->orig_ci, c->nnz * 4, cudaMemcpyDeviceToHost, c->stream));
  LZ_CUDA(cudaStreamSynchronize(c->stream));
  return LZ_OK;
}
