// Real spherical-harmonics direction encoding, degree 1..4 (tcnn "SphericalHarmonics").
// Replaces tcnn.Encoding({"otype":"SphericalHarmonics","degree":4|3}) as used by
// models/networks.py:78-85 (dir_encoder) and :128-135 (skybox_dir_encoder).  Input v in [0,1]^3
// is mapped to (x,y,z) = 2v-1 first (tcnn convention); the reference feeds (normalize(d)+1)/2
// (models/networks.py:221-222) so (x,y,z) is the unit view direction.  No parameters; the view
// direction never requires a gradient on this path (rays_d is a leaf without grad).
#include "common.cuh"

namespace ngp {

__device__ __forceinline__ void sh_eval(float x, float y, float z, int degree, float* o) {
  const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  o[0] = 0.28209479177387814f;
  if (degree <= 1) return;
  o[1] = -0.48860251190291987f * y;
  o[2] = 0.48860251190291987f * z;
  o[3] = -0.48860251190291987f * x;
  if (degree <= 2) return;
  o[4] = 1.0925484305920792f * xy;
  o[5] = -1.0925484305920792f * yz;
  o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = -1.0925484305920792f * xz;
  o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  if (degree <= 3) return;
  o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
  o[10] = 2.8906114426405538f * xy * z;
  o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
  o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
  o[14] = 1.4453057213202769f * z * (x2 - y2);
  o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

__global__ void __launch_bounds__(256) sh_fw_kernel(const float* __restrict__ v, int degree, int64_t n,
                                                    float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = fmaf(2.f, __ldg(v + 3 * i), -1.f), y = fmaf(2.f, __ldg(v + 3 * i + 1), -1.f),
              z = fmaf(2.f, __ldg(v + 3 * i + 2), -1.f);
  float o[16];
  sh_eval(x, y, z, degree, o);
  const int nd = degree * degree;
  float* dst = out + i * nd;
  if (nd == 16) {
#pragma unroll
    for (int k = 0; k < 16; k += 4) *(float4*)(dst + k) = make_float4(o[k], o[k + 1], o[k + 2], o[k + 3]);
  } else {
    for (int k = 0; k < nd; k++) dst[k] = o[k];
  }
}

}  // namespace ngp

using namespace ngp;

// out (N, degree^2) f32 = SH(2v-1), degree in 1..4.
NGP_API int ngp_sh_fw(const float* v, int degree, int64_t n, float* out, void* stream) {
  if (n <= 0) return 0;
  if (degree < 1 || degree > 4) return set_error_msg("ngp_sh_fw: degree must be in 1..4");
  sh_fw_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(v, degree, n, out);
  NGP_LAUNCH_CHECK("ngp_sh_fw");
  return 0;
}
