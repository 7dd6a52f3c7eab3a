// gather_bench.cu — what random 8-byte gather rates does a B200 sustain? Decides the SpMV design (direct gather through
// L1/L2 vs shared-memory staged gathers). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
__global__ void fill_idx(uint32_t* idx, uint64_t cnt, uint32_t range) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (uint64_t)gridDim.x * blockDim.x)
    idx[i] = (uint32_t)(__umul64hi(mix64(i), (uint64_t)range));
}
__global__ void fill_x(double* x, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) x[i] = 1.0;
}

template <int MODE> __device__ __forceinline__ double ld(const double* p) {
  double v;
  if (MODE == 0) v = __ldg(p);
  else if (MODE == 1) v = __ldcg(p);
  else if (MODE == 2) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// U independent gathers per thread per iteration; indices read as coalesced uint4
template <int U, int MODE>
__global__ void __launch_bounds__(256) gather(const uint32_t* __restrict__ idx, uint64_t cnt, const double* __restrict__ x, double* out) {
  double acc = 0.0;
  const uint64_t nvec = cnt / U;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c[U];
    if (U >= 4) {
#pragma unroll
      for (int u = 0; u < U / 4; u++) {
        uint4 v = __ldcs(reinterpret_cast<const uint4*>(idx) + i * (U / 4) + u);
        c[4 * u] = v.x; c[4 * u + 1] = v.y; c[4 * u + 2] = v.z; c[4 * u + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; u++) c[u] = __ldcs(idx + i * U + u);
    }
    double v[U];
#pragma unroll
    for (int u = 0; u < U; u++) v[u] = ld<MODE>(x + c[u]);
#pragma unroll
    for (int u = 0; u < U; u++) acc += v[u];
  }
  if (acc == 12345.678) out[0] = acc;
}

// shared-memory staged gather: CTA copies a slice of x (SL doubles) into smem, then gathers with 16-bit local indices
template <int U>
__global__ void __launch_bounds__(256) gather_smem(const uint16_t* __restrict__ idx, uint64_t cnt_per_cta, const double* __restrict__ x,
                                                   int slice, double* out, double* expand /* nullable: write gathered values */) {
  extern __shared__ double sx[];
  const double* xs = x + (uint64_t)blockIdx.x % 64 * slice;
  for (int i = threadIdx.x; i < slice; i += 256) sx[i] = xs[i];
  __syncthreads();
  const uint16_t* my = idx + (uint64_t)blockIdx.x * cnt_per_cta;
  double* ex = expand ? expand + (uint64_t)blockIdx.x * cnt_per_cta : nullptr;
  double acc = 0.0;
  const uint64_t nvec = cnt_per_cta / U;
  for (uint64_t i = threadIdx.x; i < nvec; i += 256) {
    uint16_t c[U];
    if (U == 8) {
      uint4 v = __ldcs(reinterpret_cast<const uint4*>(my) + i);
      c[0] = v.x & 0xffff; c[1] = v.x >> 16; c[2] = v.y & 0xffff; c[3] = v.y >> 16; c[4] = v.z & 0xffff; c[5] = v.z >> 16; c[6] = v.w & 0xffff; c[7] = v.w >> 16;
    } else {
#pragma unroll
      for (int u = 0; u < U; u++) c[u] = my[i * U + u];
    }
    double v[U];
#pragma unroll
    for (int u = 0; u < U; u++) v[u] = sx[c[u] % slice];
    if (ex) {
#pragma unroll
      for (int u = 0; u < U; u += 2) __stcs(reinterpret_cast<double2*>(ex + i * U + u), make_double2(v[u], v[u + 1]));
    } else {
#pragma unroll
      for (int u = 0; u < U; u++) acc += v[u];
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
  const uint64_t cnt = 1ull << 28;   // 2.68e8 gathers, as one C3 SpMV
  uint32_t* idx; double* x; double* out; double* expand;
  CK(cudaMalloc(&idx, cnt * 4)); CK(cudaMalloc(&x, (1ull << 27) * 8)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&expand, cnt * 8));
  fill_x<<<1024, 256>>>(x, 1ull << 27);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  for (int lg : {20, 22, 23, 24, 25, 27}) {
    uint32_t range = 1u << lg;
    fill_idx<<<4096, 256>>>(idx, cnt, range);
    CK(cudaDeviceSynchronize());
    int grid = sms * 8;
#define RUN(U, MODE, name) { float ms = time_it([&] { gather<U, MODE><<<grid, 256>>>(idx, cnt, x, out); }); \
      printf("x=2^%d doubles (%6.0f MB)  U=%2d %-14s %8.3f ms  %7.1f Ggather/s\n", lg, range * 8.0 / 1e6, U, name, ms, cnt / ms / 1e6); }
    RUN(1, 0, "ldg") RUN(4, 0, "ldg") RUN(8, 0, "ldg") RUN(16, 0, "ldg")
    RUN(8, 1, "ldcg") RUN(8, 2, "L1::no_alloc") RUN(8, 3, "L1::evict_last")
    grid = sms * 32;
    RUN(8, 0, "ldg grid x32")
  }
  // shared-memory staged
  uint16_t* idx16 = reinterpret_cast<uint16_t*>(idx);
  fill_idx<<<4096, 256>>>(idx, cnt / 2, 0xffffffffu);   // random 16-bit pairs
  CK(cudaDeviceSynchronize());
  for (int slice : {8192, 16384, 24576}) {
    for (int ctas_per_sm : {1, 2}) {
      if (slice * 8 * ctas_per_sm > 220 * 1024) continue;
      int grid = sms * ctas_per_sm * 4;
      uint64_t per = (cnt / grid) & ~63ull;
      CK(cudaFuncSetAttribute(gather_smem<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, slice * 8));
      float ms = time_it([&] { gather_smem<8><<<grid, 256, slice * 8>>>(idx16, per, x, slice, out, nullptr); });
      printf("smem slice %5d doubles grid %4d: gather-only   %8.3f ms  %7.1f Ggather/s\n", slice, grid, ms, per * grid / ms / 1e6);
      ms = time_it([&] { gather_smem<8><<<grid, 256, slice * 8>>>(idx16, per, x, slice, out, expand); });
      printf("smem slice %5d doubles grid %4d: gather+expand %8.3f ms  %7.1f Ggather/s  (%.0f GB/s written+read)\n", slice, grid, ms,
             per * grid / ms / 1e6, per * grid * 10.0 / ms / 1e6);
    }
  }
  return 0;
}
