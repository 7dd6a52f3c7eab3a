// gather_paths.cu — follow-up to gather_bench.cu: is the 1-gather-per-clock-per-SM ceiling a property of the LSU path only?
// Compares, for random 8-byte gathers: ld.global (LSU), tex1Dfetch<int2> (TEX path), cp.async 8 B (LDGSTS), gathers whose
// lane pairs share a 128-B line / a 32-B sector, 16-byte gathers, an L1-resident footprint, and half the SMs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_paths gather_paths.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
// share: 1 = independent; 2/4: groups of `share` consecutive indices fall in the same `gran`-element aligned group
__global__ void fill_idx(uint32_t* idx, uint64_t cnt, uint32_t range, uint32_t share, uint32_t gran, uint32_t align2) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (uint64_t)gridDim.x * blockDim.x) {
    // element i belongs to lane (i / 8) % 32 under the U=8 uint4x2 layout below; sharing is between consecutive lanes
    uint64_t vec = i / 8, u = i % 8;
    uint64_t lane_grp = vec / share;
    uint32_t base = (uint32_t)(__umul64hi(mix64(lane_grp * 8 + u), (uint64_t)range));
    if (share > 1) base = (base / gran) * gran + (uint32_t)(mix64(i) % gran);
    if (align2) base &= ~1u;
    idx[i] = base;
  }
}
__global__ void fill_x(double* x, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) x[i] = 1.0;
}

constexpr int U = 8;
__device__ __forceinline__ void load_idx(const uint32_t* idx, uint64_t i, uint32_t (&c)[U]) {
  uint4 a = __ldcs(reinterpret_cast<const uint4*>(idx) + i * 2), b = __ldcs(reinterpret_cast<const uint4*>(idx) + i * 2 + 1);
  c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
}

// MODE 0: ldg 8 B   1: tex1Dfetch int2   2: ldg 16 B (double2)   3: cp.async 8 B into smem
template <int MODE>
__global__ void __launch_bounds__(256) gather(const uint32_t* __restrict__ idx, uint64_t cnt, const double* __restrict__ x,
                                              cudaTextureObject_t tex, double* out) {
  __shared__ double stage[MODE == 3 ? 256 * U : 1];
  double acc = 0.0;
  const uint64_t nvec = cnt / U;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c[U];
    load_idx(idx, i, c);
    if (MODE == 0) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = __ldg(x + c[u]);
#pragma unroll
      for (int u = 0; u < U; u++) acc += v[u];
    } else if (MODE == 1) {
      int2 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = tex1Dfetch<int2>(tex, (int)c[u]);
#pragma unroll
      for (int u = 0; u < U; u++) acc += __hiloint2double(v[u].y, v[u].x);
    } else if (MODE == 2) {
      double2 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = __ldg(reinterpret_cast<const double2*>(x + c[u]));
#pragma unroll
      for (int u = 0; u < U; u++) acc += v[u].x + v[u].y;
    } else {
#pragma unroll
      for (int u = 0; u < U; u++) {
        uint32_t s = (uint32_t)__cvta_generic_to_shared(&stage[u * 256 + threadIdx.x]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(x + c[u]) : "memory");
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int u = 0; u < U; u++) acc += stage[u * 256 + threadIdx.x];
    }
  }
  if (acc == 12345.678) out[0] = acc;
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

int main() {
  const uint64_t cnt = 1ull << 27;
  uint32_t* idx; double* x; double* out;
  const uint64_t xn = 1ull << 24;
  CK(cudaMalloc(&idx, cnt * 4)); CK(cudaMalloc(&x, (xn + 16) * 8)); CK(cudaMalloc(&out, 64));
  fill_x<<<1024, 256>>>(x, xn + 16);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("SMs %d  maxTexture1DLinear %d\n", sms, prop.maxTexture1DLinear);
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = x;
  rd.res.linear.desc = cudaCreateChannelDesc<int2>(); rd.res.linear.sizeInBytes = xn * 8;
  cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType; td.addressMode[0] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
  cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));

  struct Cfg { int lg; uint32_t share, gran, align2; const char* what; };
  const Cfg cfgs[] = {
    {13, 1, 1, 0, "64 KB footprint (L1-resident)"},
    {20, 1, 1, 0, "8 MB footprint"},
    {23, 1, 1, 0, "67 MB footprint"},
    {23, 2, 16, 0, "67 MB, lane pairs share a 128-B line"},
    {23, 4, 16, 0, "67 MB, lane quads share a 128-B line"},
    {23, 2, 4, 0, "67 MB, lane pairs share a 32-B sector"},
    {23, 1, 1, 1, "67 MB, even indices (for 16-B gathers)"},
  };
  for (const Cfg& c : cfgs) {
    fill_idx<<<4096, 256>>>(idx, cnt, 1u << c.lg, c.share, c.gran, c.align2);
    CK(cudaDeviceSynchronize());
    printf("-- %s\n", c.what);
    for (int grid_mul : {8}) {
      int grid = sms * grid_mul;
#define RUN(MODE, name, GRID) { float ms = time_it([&] { gather<MODE><<<GRID, 256>>>(idx, cnt, x, tex, out); }); \
        printf("   %-22s grid %5d  %8.3f ms  %7.1f Ggather/s  (%.2f per SM-clk @1.965 GHz, all SMs)\n", name, GRID, ms, cnt / ms / 1e6, cnt / ms / 1e6 / 1.965 / sms); }
      RUN(0, "ld.global.nc 8B", grid)
      RUN(1, "tex1Dfetch<int2>", grid)
      RUN(3, "cp.async 8B -> smem", grid)
      if (c.align2) RUN(2, "ld.global.nc 16B", grid)
    }
    if (c.lg == 23 && c.share == 1 && !c.align2) {
      RUN(0, "ld 8B, 74 CTAs (half)", 74)
      RUN(0, "ld 8B, 148 CTAs", 148)
      RUN(1, "tex, 74 CTAs (half)", 74)
      RUN(1, "tex, 148 CTAs", 148)
    }
  }
  return 0;
}
