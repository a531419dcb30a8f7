// y = -normalize(x * s)  on (N,3) rows, forward and backward: the two normal outputs of the reference field,
//     normals_raw  = -F.normalize(d sigma / d x, p=2, dim=-1, eps=1e-6)      (models/networks.py:209)
//     normals_pred = -F.normalize(norm_pred_header(feat), p=2, dim=-1, eps=1e-6)   (networks.py:222-223)
// with the unit-cube -> world scale 1/(xyz_max - xyz_min) of the analytic gradient folded in as s.  In torch ops each of them
// is norm, clamp_min, expand, div, neg forward and ~8 elementwise / reduction kernels backward on (S,3) tensors: 5 ms of a
// 97 ms playground-shaped step (profiles/r02d_step_profile_playground.txt: aten::div, mul, neg, sum, linalg_vector_norm).
// One thread per row; a warp touches 384 contiguous bytes per tensor.
#include "common.cuh"

namespace ngp {

__global__ void __launch_bounds__(256) neg_normalize_fw_kernel(const float* __restrict__ x, float sx, float sy, float sz, float eps, int64_t n,
                                                               float* __restrict__ y, float* __restrict__ inv_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = __ldg(x + 3 * i) * sx, b = __ldg(x + 3 * i + 1) * sy, c = __ldg(x + 3 * i + 2) * sz;
  const float nrm = sqrtf(a * a + b * b + c * c);
  const float inv = 1.0f / fmaxf(nrm, eps);                       // F.normalize: v / max(||v||, eps)
  y[3 * i] = -a * inv; y[3 * i + 1] = -b * inv; y[3 * i + 2] = -c * inv;
  inv_out[i] = nrm > eps ? inv : -inv;                            // sign bit = "clamped": the Jacobian is then -I/eps, not the projector
}

// gx = s * d(-normalize(v))/dv^T gy  with v = x*s:  unclamped rows  -(gy - yh (yh . gy)) / ||v||  with yh = v/||v|| = -y ;  clamped rows -gy / eps
__global__ void __launch_bounds__(256) neg_normalize_bw_kernel(const float* __restrict__ gy, const float* __restrict__ y, const float* __restrict__ inv_in,
                                                               float sx, float sy, float sz, int64_t n, float* __restrict__ gx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g0 = __ldg(gy + 3 * i), g1 = __ldg(gy + 3 * i + 1), g2 = __ldg(gy + 3 * i + 2);
  const float iv = __ldg(inv_in + i);
  float d0, d1, d2;
  if (iv >= 0.f) {
    const float y0 = __ldg(y + 3 * i), y1 = __ldg(y + 3 * i + 1), y2 = __ldg(y + 3 * i + 2);
    const float dot = y0 * g0 + y1 * g1 + y2 * g2;                // (-yh) . gy
    d0 = -(g0 - y0 * dot) * iv; d1 = -(g1 - y1 * dot) * iv; d2 = -(g2 - y2 * dot) * iv;
  } else {
    d0 = g0 * iv; d1 = g1 * iv; d2 = g2 * iv;                     // iv = -1/eps
  }
  gx[3 * i] = d0 * sx; gx[3 * i + 1] = d1 * sy; gx[3 * i + 2] = d2 * sz;
}

}  // namespace ngp

using namespace ngp;

// y (N,3) = -normalize(x * scale, eps) ; inv (N): 1 / max(||x*scale||, eps), negated on clamped rows (saved for the backward).
// scale: three HOST floats (the field's 1 / (xyz_max - xyz_min); pass 1,1,1 for none).
NGP_API int ngp_neg_normalize_fw(const float* x, float scale_x, float scale_y, float scale_z, float eps, int64_t n, float* y, float* inv,
                                 void* stream) {
  if (n <= 0) return 0;
  neg_normalize_fw_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, scale_x, scale_y, scale_z, eps, n, y, inv);
  NGP_LAUNCH_CHECK("ngp_neg_normalize_fw");
  return 0;
}
// gx (N,3) = (d y / d x)^T gy from the forward's y and inv.
NGP_API int ngp_neg_normalize_bw(const float* gy, const float* y, const float* inv, float scale_x, float scale_y, float scale_z, int64_t n,
                                 float* gx, void* stream) {
  if (n <= 0) return 0;
  neg_normalize_bw_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(gy, y, inv, scale_x, scale_y, scale_z, n, gx);
  NGP_LAUNCH_CHECK("ngp_neg_normalize_bw");
  return 0;
}
