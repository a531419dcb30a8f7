// Is the L1 gather rate bound by distinct 128-byte LINES or by 32-byte SECTORS per warp instruction?
// Each lane loads 8 bytes (an F=2 fp32 hash-table entry) from a 43 MB table (L2 resident).
//   mode 0: 32 lanes -> 32 random lines                       (today's fine-level gather: one corner per instruction)
//   mode 1: lane pairs -> same random line, different sectors (16 lines, 32 sectors)
//   mode 2: lane octets -> same random line, 4 sectors x 2    (4 lines... 8 lanes/line, 16 sectors)
//   mode 3: lane pairs -> same sector (16 lines, 16 sectors)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/l1_probe tools/probes/l1_probe.cu && gpurun_out/l1_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
template <int MODE>
__global__ void probe(const float2* __restrict__ table, uint32_t n_lines, int iters, float* out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int it = 0; it < iters; it++) {
    uint32_t key, within;   // line selector, entry within the 16-entry line
    if (MODE == 0) { key = tid; within = mix(tid * 31 + it) & 15; }
    else if (MODE == 1) { key = tid >> 1; within = ((lane & 1) * 8 + (mix(key + it) & 3)) & 15; }     // different sectors (0-1 vs 2-3)
    else if (MODE == 2) { key = tid >> 3; within = ((lane & 7) * 2) & 15; }
    else { key = tid >> 1; within = (mix(key + it) & 12) | ((lane & 1) * 2 + 0); }                     // same sector (4 entries per sector)
    const uint32_t line = mix(key * 2654435761u + it * 40503u) % n_lines;
    const float2 v = __ldg(table + (size_t)line * 16 + within);
    acc += v.x + v.y;
  }
  if (acc == 123.456f) out[0] = acc;
}
int main() {
  const uint32_t n_lines = (43u << 20) / 128;
  float2* table; float* out;
  cudaMalloc(&table, (size_t)n_lines * 128); cudaMemset(table, 0, (size_t)n_lines * 128); cudaMalloc(&out, 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 64, blocks = 148 * 64, threads = 256;
  for (int mode = 0; mode < 4; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(a);
      if (mode == 0) probe<0><<<blocks, threads>>>(table, n_lines, iters, out);
      if (mode == 1) probe<1><<<blocks, threads>>>(table, n_lines, iters, out);
      if (mode == 2) probe<2><<<blocks, threads>>>(table, n_lines, iters, out);
      if (mode == 3) probe<3><<<blocks, threads>>>(table, n_lines, iters, out);
      cudaEventRecord(b); cudaEventSynchronize(b);
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double loads = (double)blocks * threads * iters;
    printf("mode %d: %.3f ms  %.1f G lane-loads/s  %.2f warp-instr/clk/SM (1.9 GHz)\n", mode, ms, loads / ms / 1e6, loads / 32 / (ms * 1e-3) / 148 / 1.9e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
