// Library-level entry points of libngp_b200.so: version, last-error string, device checks.
#include "common.cuh"
#include <stdio.h>
#include <string.h>

namespace ngp {
static thread_local char g_err[512] = "";
int set_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}
int set_error_msg(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return -1;
}
}  // namespace ngp

NGP_API int ngp_abi_version(void) { return 1; }

// Text of the last non-zero status returned on this thread ("" when none).
NGP_API const char* ngp_last_error(void) { return ngp::g_err; }

// 0 when the current device is a compute-capability-10.x part (the only target this library is
// built for), else a non-zero status with ngp_last_error() explaining why.  There is no fallback.
NGP_API int ngp_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return ngp::set_error(e, "ngp_check_device");
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return ngp::set_error(e, "ngp_check_device");
  if (p.major != 10) {
    char buf[256];
    snprintf(buf, sizeof(buf), "ngp_b200 is built for sm_100a only; device %d is sm_%d%d (%s)", dev, p.major, p.minor, p.name);
    return ngp::set_error_msg(buf);
  }
  return 0;
}

// ---- diagnostics -------------------------------------------------------------------------------------------------
// Measures the rate at which the L2 executes reduction SECTOR requests, the resource that bounds the hash-table gradient
// scatter (hashgrid.cu): every lane pair issues one red.global.add.v2.f32 pair into the same random 32-byte sector of
// `table` (= one request, the scatter's access pattern for x-neighbours).  bench.py times this launch in the same run to
// obtain the denominator of `roofline.limiter`; returns the number of sector requests issued, or < 0 on error.
namespace ngp {
__device__ __forceinline__ uint32_t probe_mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void __launch_bounds__(256) l2_red_probe_kernel(float* __restrict__ table, uint32_t n_sectors, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  for (int it = 0; it < iters; it++) {
    const uint32_t sec = probe_mix((tid >> 1) * 2654435761u + it * 40503u) % n_sectors;
    atomicAdd(reinterpret_cast<float2*>(table + (size_t)sec * 8) + ((lane & 1) + (probe_mix(tid >> 1) & 2)), make_float2(0.f, 0.f));
  }
}
}  // namespace ngp

NGP_API int64_t ngp_probe_l2_reduction(float* table, int64_t table_bytes, int iters, void* stream) {
  if (table_bytes < 32 || iters < 1) return ngp::set_error_msg("ngp_probe_l2_reduction: need table_bytes >= 32 and iters >= 1");
  const int blocks = ngp::kSMs * 32, threads = 256;
  ngp::l2_red_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(table, (uint32_t)(table_bytes / 32), iters);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -(int64_t)ngp::set_error(e, "ngp_probe_l2_reduction");
  return (int64_t)blocks * threads / 2 * iters;
}
