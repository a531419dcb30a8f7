// Next-round probe for DESIGN.md §8 item 2: how does the L2 reduction rate depend on (a) the table's working set and
// (b) a concurrent stream of read-once rows through the same L2?  The street-shape scatter runs at 83 G sector
// requests/s on a 64 MB live set with 3.2 GB of rows streaming past it per chunk sweep; l2_red_probe.cu measures 220 G/s on
// a quiet 43 MB table.  Each thread issues `iters` 8-byte reductions (red.global.add.v2.f32) to random sectors of a table of
// `mb` megabytes; with stream_words > 0 it also reads that many 16-byte words per reduction from a large read-once buffer
// (ld.global.nc), as the scatter's x / dL/dy rows do.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/probes/l2_red_ws_probe.cu -o /tmp/l2_red_ws_probe && /tmp/l2_red_ws_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int STREAM>
__global__ void probe(float* __restrict__ table, uint32_t n_sectors, const float4* __restrict__ rows, uint32_t n_rows, int iters, float* sink) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  for (int it = 0; it < iters; it++) {
    float w = 1.f;
    if (STREAM > 0) {
#pragma unroll
      for (int k = 0; k < STREAM; k++) {            // contiguous per warp: a streamed row, read once
        const float4 v = __ldg(rows + ((size_t)(it * STREAM + k) * gridDim.x * blockDim.x + tid) % n_rows);
        w += v.x;
      }
    }
    const uint32_t sec = mix(tid * 2654435761u + it * 40503u) % n_sectors;
    atomicAdd(reinterpret_cast<float2*>(table + (size_t)sec * 8) + (mix(tid + it) & 3), make_float2(w, w));
    acc += w;
  }
  if (acc == 12345.678f) *sink = acc;
}

int main() {
  const int sizes_mb[] = {16, 32, 43, 64, 96, 128, 192, 363, 768};
  const size_t rows_bytes = (size_t)4 << 30;       // 4 GB read-once buffer, far beyond the L2
  float4* rows; cudaMalloc(&rows, rows_bytes); cudaMemset(rows, 0, rows_bytes);
  float* sink; cudaMalloc(&sink, 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 64, blocks = 148 * 32, threads = 256;
  for (int mb : sizes_mb) {
    const uint32_t n_sectors = (uint32_t)(((size_t)mb << 20) / 32);
    float* table; cudaMalloc(&table, (size_t)n_sectors * 32); cudaMemset(table, 0, (size_t)n_sectors * 32);
    for (int stream = 0; stream <= 2; stream++) {
      float ms = 0.f;
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(a);
        if (stream == 0) probe<0><<<blocks, threads>>>(table, n_sectors, rows, (uint32_t)(rows_bytes / 16), iters, sink);
        if (stream == 1) probe<1><<<blocks, threads>>>(table, n_sectors, rows, (uint32_t)(rows_bytes / 16), iters, sink);
        if (stream == 2) probe<2><<<blocks, threads>>>(table, n_sectors, rows, (uint32_t)(rows_bytes / 16), iters, sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
      }
      const double reqs = (double)blocks * threads * iters;
      printf("table %4d MB, %d streamed 16-B words per reduction: %8.3f ms  %6.1f G reductions/s  (+%5.1f GB/s of rows)\n",
             mb, stream, ms, reqs / ms / 1e6, reqs * stream * 16 / ms / 1e6);
    }
    cudaFree(table);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
