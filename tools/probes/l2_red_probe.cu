// What bounds the hash-table gradient scatter: L2 reduction ops per 32-byte SECTOR request, per lane op, or per line?
// 43 MB table (L2 resident), each active lane issues red.global.add.v2.f32 (8 B, one F=2 entry) or .v4 (16 B, a pair).
//   mode 0: 32 lanes, v2, 32 random sectors                 (odd-x corners today)
//   mode 1: 32 lanes, v2, lane pairs in the SAME sector      (x-neighbours issued by a lane pair)
//   mode 2: 16 lanes, v4, 16 random sectors                 (even-x pairs today: same data volume as mode 1)
//   mode 3: 32 lanes, v2, lane pairs same LINE different sector
//   mode 4: 32 lanes, v2, lane quads in the same sector (4 entries of one sector)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
template <int MODE>
__global__ void probe(float* __restrict__ table, uint32_t n_sectors, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
      const uint32_t sec = mix(tid * 2654435761u + it * 40503u) % n_sectors;
      atomicAdd(reinterpret_cast<float2*>(table + (size_t)sec * 8) + (mix(tid + it) & 3), make_float2(1.f, 1.f));
    } else if (MODE == 1) {
      const uint32_t sec = mix((tid >> 1) * 2654435761u + it * 40503u) % n_sectors;
      atomicAdd(reinterpret_cast<float2*>(table + (size_t)sec * 8) + ((lane & 1) + (mix(tid >> 1) & 2)), make_float2(1.f, 1.f));
    } else if (MODE == 2) {
      if (lane & 1) continue;
      const uint32_t sec = mix((tid >> 1) * 2654435761u + it * 40503u) % n_sectors;
      atomicAdd(reinterpret_cast<float4*>(table + (size_t)sec * 8) + (mix(tid >> 1) & 1), make_float4(1.f, 1.f, 1.f, 1.f));
    } else if (MODE == 3) {
      const uint32_t line = mix((tid >> 1) * 2654435761u + it * 40503u) % (n_sectors / 4);
      atomicAdd(reinterpret_cast<float2*>(table + (size_t)line * 32) + ((lane & 1) * 8 + (mix(tid >> 1) & 3)), make_float2(1.f, 1.f));
    } else {
      const uint32_t sec = mix((tid >> 2) * 2654435761u + it * 40503u) % n_sectors;
      atomicAdd(reinterpret_cast<float2*>(table + (size_t)sec * 8) + (lane & 3), make_float2(1.f, 1.f));
    }
  }
}
int main() {
  const uint32_t n_sectors = (43u << 20) / 32;
  float* table;
  cudaMalloc(&table, (size_t)n_sectors * 32); cudaMemset(table, 0, (size_t)n_sectors * 32);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 64, blocks = 148 * 32, threads = 256;
  for (int mode = 0; mode < 5; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(a);
      if (mode == 0) probe<0><<<blocks, threads>>>(table, n_sectors, iters);
      if (mode == 1) probe<1><<<blocks, threads>>>(table, n_sectors, iters);
      if (mode == 2) probe<2><<<blocks, threads>>>(table, n_sectors, iters);
      if (mode == 3) probe<3><<<blocks, threads>>>(table, n_sectors, iters);
      if (mode == 4) probe<4><<<blocks, threads>>>(table, n_sectors, iters);
      cudaEventRecord(b); cudaEventSynchronize(b);
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double lanes = (double)blocks * threads * iters;
    printf("mode %d: %.3f ms   %.1f G entry-updates/s (8-byte entries)\n", mode, ms, lanes / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
