// Micro-benchmark behind the map layout decisions: random gather
// rate on B200 as a function of footprint and access size.
//   nvcc -arch=sm_100a -O3 -o gather_bench gather_bench.cu && ./gather_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

template <int BYTES, int ILP>
__global__ void gather(const char* base, uint64_t n_items, uint64_t n_access, unsigned long long* sink, uint64_t seed) {
  uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
  unsigned long long acc = 0;
  for (uint64_t i = tid; i < n_access; i += stride * ILP) {
    uint4 v[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      uint64_t idx = ((mix(i + u * stride + seed) & 0xffffffffull) * n_items) >> 32;
      const char* p = base + idx * BYTES;
      if (BYTES == 4) v[u].x = __ldg(reinterpret_cast<const unsigned*>(p));
      else if (BYTES == 32) { v[u] = __ldg(reinterpret_cast<const uint4*>(p)); acc += __ldg(reinterpret_cast<const uint4*>(p) + 1).x; }
      else { v[u] = __ldg(reinterpret_cast<const uint4*>(p)); acc += __ldg(reinterpret_cast<const uint4*>(p) + 3).x; }
    }
#pragma unroll
    for (int u = 0; u < ILP; ++u) acc += v[u].x;
  }
  if (acc == 0x1234567) *sink = acc;
}

template <int BYTES, int ILP>
void run(const char* d, size_t footprint, unsigned long long* sink) {
  uint64_t n_items = footprint / BYTES, n_access = 1ull << 25;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather<BYTES, ILP><<<148 * 8, 256>>>(d, n_items, n_access, sink, 1);
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) gather<BYTES, ILP><<<148 * 8, 256>>>(d, n_items, n_access, sink, 7 + r);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  printf("footprint %7.0f MB  access %2d B  ilp %d : %7.2f G access/s  %7.1f GB/s useful\n", footprint / 1048576.0,
         BYTES, ILP, n_access / ms / 1e6, n_access * (double)BYTES / ms / 1e6);
}

int main() {
  size_t maxb = 8ull << 30;
  char* d; cudaMalloc(&d, maxb); cudaMemset(d, 1, maxb);
  unsigned long long* sink; cudaMalloc(&sink, 8);
  for (size_t mb : {32, 64, 128, 256, 512, 1024, 2048, 4096, 8192}) {
    size_t fp = mb << 20;
    run<4, 4>(d, fp, sink);
    run<64, 4>(d, fp, sink);
  }
  run<4, 1>(d, 128ull << 20, sink);
  run<64, 1>(d, 2048ull << 20, sink);
  run<64, 8>(d, 2048ull << 20, sink);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
