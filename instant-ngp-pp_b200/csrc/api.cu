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
