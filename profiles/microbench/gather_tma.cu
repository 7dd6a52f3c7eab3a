// gather_tma.cu — round-2 follow-up to gather_paths.cu: are there Blackwell paths around the
// ~0.9 line requests / clock / SM that bound random 8-byte gathers through LSU, TEX and cp.async (LDGSTS)?
//   A  cp.async.bulk (UBLKCP) 16-byte copies global -> shared, one per gather, completion on an mbarrier
//   B  TMA tile::gather4 (UTMALDG): x viewed as a [n/2][2] fp64 tensor, one instruction fetches 4 random 16-byte rows
//   C  B issued concurrently with LSU gathers in the same warps (is the TMA request path independent of the L1-miss path?)
//   D  gathers from distributed shared memory: a cluster of 8 / 16 CTAs holds a table in its shared memories (ld.shared::cluster)
//   E  local shared-memory gathers (the staging upper bound)
//   F  scattered red.global.add.f64 into an L2-resident footprint (the A = A^T scatter formulation of the SpMV)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_tma gather_tma.cu   (driver API via cudaGetDriverEntryPoint)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
__global__ void fill_idx(uint32_t* idx, uint64_t cnt, uint32_t range) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (uint64_t)gridDim.x * blockDim.x)
    idx[i] = (uint32_t)__umul64hi(mix64(i), (uint64_t)range);
}
__global__ void fill_x(double* x, uint64_t n, double v) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) x[i] = v;
}

constexpr int U = 8;
constexpr int kThreads = 256;
__device__ __forceinline__ void load_idx(const uint32_t* idx, uint64_t i, uint32_t (&c)[U]) {
  uint4 a = __ldcs(reinterpret_cast<const uint4*>(idx) + i * 2), b = __ldcs(reinterpret_cast<const uint4*>(idx) + i * 2 + 1);
  c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- baseline: LSU gathers (same loop shape as the others) ----------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_lsu(const uint32_t* __restrict__ idx, uint64_t cnt, const double* __restrict__ x, double* out) {
  double acc = 0.0;
  const uint64_t nvec = cnt / U;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c[U];
    load_idx(idx, i, c);
    double v[U];
#pragma unroll
    for (int u = 0; u < U; u++) v[u] = __ldg(x + c[u]);
#pragma unroll
    for (int u = 0; u < U; u++) acc += v[u];
  }
  if (acc == 12345.678) out[0] = acc;
}

// ---- A: cp.async.bulk 16 B per gather ------------------------------------------------------------------------------------
// every thread copies the aligned 16-byte pair that holds x[c] into its own slot and picks the half it wants
__global__ void __launch_bounds__(kThreads) k_bulk16(const uint32_t* __restrict__ idx, uint64_t cnt, const double* __restrict__ x, double* out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* stage = reinterpret_cast<double2*>(smem_raw);           // [U][kThreads]
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  double acc = 0.0;
  const uint64_t nvec = cnt / U;
  const uint64_t rounds = (nvec + (uint64_t)gridDim.x * blockDim.x - 1) / ((uint64_t)gridDim.x * blockDim.x);
  uint32_t parity = 0;
  for (uint64_t r = 0; r < rounds; r++) {
    uint64_t i = (r * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nvec) i = nvec - 1;                                    // keep the byte count uniform
    uint32_t c[U];
    load_idx(idx, i, c);
    if (threadIdx.x == 0) mbar_expect(&bar, kThreads * U * 16);
#pragma unroll
    for (int u = 0; u < U; u++)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(smem_u32(&stage[u * kThreads + threadIdx.x])),
                   "l"(x + (c[u] & ~1u)), "r"(smem_u32(&bar))
                   : "memory");
    mbar_wait(&bar, parity);
    parity ^= 1;
#pragma unroll
    for (int u = 0; u < U; u++) {
      const double2 v = stage[u * kThreads + threadIdx.x];
      acc += (c[u] & 1u) ? v.y : v.x;
    }
  }
  if (acc == 12345.678) out[0] = acc;
}

// ---- B / C: TMA tile::gather4 ------------------------------------------------------------------------------------------
// x as a 2-D fp64 tensor [n/2 rows][2]; one instruction brings 4 rows (4 x 16 B) into a 64-byte destination.
// LSU_PER: additional plain LSU gathers per thread and round, issued between the TMA issue and the wait (mode C).
template <int G4_PER, int LSU_PER>
__global__ void __launch_bounds__(kThreads) k_gather4(const uint32_t* __restrict__ idx, uint64_t cnt, const double* __restrict__ x,
                                                      const __grid_constant__ CUtensorMap tmap, double* out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // 128-byte slot per gather4 (destination alignment of tiled TMA), 64 bytes used
  unsigned char* stage = smem_raw;
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  constexpr int PER = 4 * G4_PER + LSU_PER;     // gathers per thread and round; multiple of 8 by construction below
  static_assert(PER % U == 0, "whole index vectors");
  double acc = 0.0;
  const uint64_t nvec = cnt / PER;
  const uint64_t rounds = (nvec + (uint64_t)gridDim.x * blockDim.x - 1) / ((uint64_t)gridDim.x * blockDim.x);
  uint32_t parity = 0;
  for (uint64_t r = 0; r < rounds; r++) {
    uint64_t i = (r * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nvec) i = nvec - 1;
    uint32_t c[PER];
#pragma unroll
    for (int v = 0; v < PER / U; v++) load_idx(idx, i * (PER / U) + v, *reinterpret_cast<uint32_t(*)[U]>(&c[v * U]));
    if (G4_PER > 0 && threadIdx.x == 0) mbar_expect(&bar, kThreads * G4_PER * 64);
#pragma unroll
    for (int g = 0; g < G4_PER; g++) {
      const uint32_t dst = smem_u32(stage + ((size_t)g * kThreads + threadIdx.x) * 128);
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
          "l"(&tmap), "r"(0), "r"((int)(c[4 * g] >> 1)), "r"((int)(c[4 * g + 1] >> 1)), "r"((int)(c[4 * g + 2] >> 1)),
          "r"((int)(c[4 * g + 3] >> 1)), "r"(smem_u32(&bar))
          : "memory");
    }
    double v[LSU_PER > 0 ? LSU_PER : 1];
#pragma unroll
    for (int u = 0; u < LSU_PER; u++) v[u] = __ldg(x + c[4 * G4_PER + u]);
#pragma unroll
    for (int u = 0; u < LSU_PER; u++) acc += v[u];
    if (G4_PER > 0) {
      mbar_wait(&bar, parity);
      parity ^= 1;
#pragma unroll
      for (int g = 0; g < G4_PER; g++) {
        const double* d = reinterpret_cast<const double*>(stage + ((size_t)g * kThreads + threadIdx.x) * 128);
#pragma unroll
        for (int t = 0; t < 4; t++) acc += d[2 * t + (c[4 * g + t] & 1u)];
      }
    }
  }
  if (acc == 12345.678) out[0] = acc;
}

// ---- D: distributed shared memory gathers -----------------------------------------------------------------------------------
// Every CTA of a cluster fills `tab_n` doubles of its shared memory; gathers pick (cta rank, offset) from the index stream.
template <bool REMOTE>
__global__ void __launch_bounds__(512) k_dsmem(const uint32_t* __restrict__ idx, uint64_t cnt, uint32_t tab_n, uint32_t csize, double* out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* tab = reinterpret_cast<double*>(smem_raw);
  for (uint32_t i = threadIdx.x; i < tab_n; i += blockDim.x) tab[i] = 1.0;
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  cl.sync();
  const uint32_t base = smem_u32(tab);
  double acc = 0.0;
  const uint64_t nvec = cnt / U;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c[U];
    load_idx(idx, i, c);
    double v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t off = c[u] % tab_n, rk = (c[u] / tab_n) % csize;
      if (REMOTE) {
        uint32_t ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(base + off * 8), "r"(rk));
        asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v[u]) : "r"(ra));
      } else {
        v[u] = tab[off];
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) acc += v[u];
  }
  cl.sync();
  if (acc == 12345.678) out[0] = acc;
}

// ---- F: scattered reductions -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_red(const uint32_t* __restrict__ idx, uint64_t cnt, double* y) {
  const uint64_t nvec = cnt / U;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c[U];
    load_idx(idx, i, c);
#pragma unroll
    for (int u = 0; u < U; u++) atomicAdd(y + c[u], 1.0);   // result unused -> RED.E.ADD.F64
  }
}

template <class F> float time_it(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const uint64_t cnt = 1ull << 27;
  uint32_t* idx; double* x; double* out; double* y;
  const uint64_t xn = 1ull << 24;
  CK(cudaMalloc(&idx, cnt * 4)); CK(cudaMalloc(&x, (xn + 16) * 8)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&y, xn * 8));
  fill_x<<<1024, 256>>>(x, xn + 16, 1.0);
  fill_x<<<1024, 256>>>(y, xn, 0.0);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double ghz = clk_khz * 1e-6;
  printf("SMs %d, clock %.3f GHz (max); %llu gathers per launch, best of 5\n", sms, ghz, (unsigned long long)cnt);

  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qr));
  if (!encode || qr != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled unavailable\n"); return 1; }

#define REPORT(name, grid, ms) printf("   %-58s grid %5d  %8.3f ms  %7.1f Ggather/s  (%.2f per SM-clk)\n", name, (int)(grid), ms, cnt / (ms) / 1e6, cnt / (ms) / 1e6 / ghz / sms)

  for (int lg : {20, 23, 24}) {
    const uint32_t range = 1u << lg;
    fill_idx<<<4096, 256>>>(idx, cnt, range);
    CK(cudaDeviceSynchronize());
    printf("-- random 8-byte gathers from a %.0f MB footprint\n", range * 8.0 / 1e6);
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {2, range / 2};
    cuuint64_t gstride[1] = {16};
    cuuint32_t box[2] = {2, 1}, estr[2] = {1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, x, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 1; }
    for (int mul : {4, 8}) {
      const int grid = sms * mul;
      { float ms = time_it([&] { k_lsu<<<grid, kThreads>>>(idx, cnt, x, out); }); REPORT("LSU ld.global.nc 8 B", grid, ms); }
    }
    for (int mul : {2, 4}) {
      const int grid = sms * mul;
      { const size_t sm = (size_t)U * kThreads * 16;
        CK(cudaFuncSetAttribute(k_bulk16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        float ms = time_it([&] { k_bulk16<<<grid, kThreads, sm>>>(idx, cnt, x, out); }); REPORT("A  cp.async.bulk 16 B per gather (UBLKCP)", grid, ms); }
      { const size_t sm = (size_t)2 * kThreads * 128;
        CK(cudaFuncSetAttribute(k_gather4<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        float ms = time_it([&] { k_gather4<2, 0><<<grid, kThreads, sm>>>(idx, cnt, x, tmap, out); }); REPORT("B  TMA tile::gather4, 2 per thread (UTMALDG)", grid, ms); }
      { const size_t sm = (size_t)1 * kThreads * 128;
        CK(cudaFuncSetAttribute(k_gather4<1, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        float ms = time_it([&] { k_gather4<1, 12><<<grid, kThreads, sm>>>(idx, cnt, x, tmap, out); }); REPORT("C  1 gather4 (4) + 12 LSU gathers per thread", grid, ms); }
      { const size_t sm = (size_t)2 * kThreads * 128;
        CK(cudaFuncSetAttribute(k_gather4<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        float ms = time_it([&] { k_gather4<2, 8><<<grid, kThreads, sm>>>(idx, cnt, x, tmap, out); }); REPORT("C  2 gather4 (8) + 8 LSU gathers per thread", grid, ms); }
      { const size_t sm = 128;
        float ms = time_it([&] { k_gather4<0, 8><<<grid, kThreads, sm>>>(idx, cnt, x, tmap, out); }); REPORT("   (same loop, 8 LSU gathers only)", grid, ms); }
    }
  }

  // D / E: shared-memory tables
  fill_idx<<<4096, 256>>>(idx, cnt, 0xFFFFFFFFu);
  CK(cudaDeviceSynchronize());
  printf("-- gathers from shared memory (table of 16 Ki doubles = 128 KB per CTA, 512 threads, 1 CTA per SM)\n");
  {
    const uint32_t tab_n = 16384;
    const size_t sm = (size_t)tab_n * 8;
    CK(cudaFuncSetAttribute(k_dsmem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    CK(cudaFuncSetAttribute(k_dsmem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    CK(cudaFuncSetAttribute(k_dsmem<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(k_dsmem<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int csize : {1, 2, 4, 8, 16}) {
      cudaLaunchConfig_t cfg = {};
      int grid = (sms / csize) * csize;
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = sm;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int maxc = 0;
      cudaError_t qe = cudaOccupancyMaxActiveClusters(&maxc, k_dsmem<true>, &cfg);
      if (qe != cudaSuccess || maxc == 0) { cudaGetLastError(); printf("   cluster size %d: not launchable (%s)\n", csize, cudaGetErrorString(qe)); continue; }
      if (grid > maxc * csize) grid = maxc * csize, cfg.gridDim = dim3(grid);
      char name[96];
      if (csize == 1) {
        float ms = time_it([&] { CK(cudaLaunchKernelEx(&cfg, k_dsmem<false>, (const uint32_t*)idx, cnt, tab_n, (uint32_t)csize, out)); });
        snprintf(name, sizeof name, "E  local shared memory (ld.shared)");
        REPORT(name, grid, ms);
      }
      float ms = time_it([&] { CK(cudaLaunchKernelEx(&cfg, k_dsmem<true>, (const uint32_t*)idx, cnt, tab_n, (uint32_t)csize, out)); });
      snprintf(name, sizeof name, "D  ld.shared::cluster, cluster of %d (%d KB table, %d active clusters)", csize, csize * 128, grid / csize);
      REPORT(name, grid, ms);
    }
  }

  // F: scattered reductions
  for (int lg : {20, 23}) {
    fill_idx<<<4096, 256>>>(idx, cnt, 1u << lg);
    CK(cudaDeviceSynchronize());
    printf("-- scattered red.global.add.f64 into a %.0f MB footprint\n", (1u << lg) * 8.0 / 1e6);
    const int grid = sms * 8;
    float ms = time_it([&] { k_red<<<grid, kThreads>>>(idx, cnt, y); });
    REPORT("F  red.global.add.f64", grid, ms);
  }
  return 0;
}
