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

// The per-sample inputs of the Ref-NeRF regularisers (models/rendering.py:243-246):
//     normals_diff = (normals_raw - normals_pred)^2                                  (S,3)
//     normals_ori  = clamp(sum(normals_raw * normalize(dirs)), min=0)^2              (S)
// as one kernel per direction instead of sub, pow, normalize (5 ops), mul, sum, clamp, pow and their ~20 backward kernels.
__global__ void __launch_bounds__(256) refloss_prep_fw_kernel(const float* __restrict__ nraw, const float* __restrict__ npred,
                                                              const float* __restrict__ dirs, int64_t n, float* __restrict__ diff,
                                                              float* __restrict__ ori) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float r0 = __ldg(nraw + 3 * i), r1 = __ldg(nraw + 3 * i + 1), r2 = __ldg(nraw + 3 * i + 2);
  const float p0 = __ldg(npred + 3 * i), p1 = __ldg(npred + 3 * i + 1), p2 = __ldg(npred + 3 * i + 2);
  const float d0 = __ldg(dirs + 3 * i), d1 = __ldg(dirs + 3 * i + 1), d2 = __ldg(dirs + 3 * i + 2);
  const float inv = 1.0f / fmaxf(sqrtf(d0 * d0 + d1 * d1 + d2 * d2), 1e-6f);
  const float e0 = r0 - p0, e1 = r1 - p1, e2 = r2 - p2;
  diff[3 * i] = e0 * e0; diff[3 * i + 1] = e1 * e1; diff[3 * i + 2] = e2 * e2;
  const float c = fmaxf((r0 * d0 + r1 * d1 + r2 * d2) * inv, 0.f);
  ori[i] = c * c;
}
__global__ void __launch_bounds__(256) refloss_prep_bw_kernel(const float* __restrict__ nraw, const float* __restrict__ npred,
                                                              const float* __restrict__ dirs, const float* __restrict__ gdiff,
                                                              const float* __restrict__ gori, int64_t n, float* __restrict__ graw,
                                                              float* __restrict__ gpred) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float r0 = __ldg(nraw + 3 * i), r1 = __ldg(nraw + 3 * i + 1), r2 = __ldg(nraw + 3 * i + 2);
  const float p0 = __ldg(npred + 3 * i), p1 = __ldg(npred + 3 * i + 1), p2 = __ldg(npred + 3 * i + 2);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (gdiff) {
    a0 = 2.f * (r0 - p0) * __ldg(gdiff + 3 * i); a1 = 2.f * (r1 - p1) * __ldg(gdiff + 3 * i + 1); a2 = 2.f * (r2 - p2) * __ldg(gdiff + 3 * i + 2);
  }
  if (gpred) { gpred[3 * i] = -a0; gpred[3 * i + 1] = -a1; gpred[3 * i + 2] = -a2; }
  if (gori) {
    const float d0 = __ldg(dirs + 3 * i), d1 = __ldg(dirs + 3 * i + 1), d2 = __ldg(dirs + 3 * i + 2);
    const float inv = 1.0f / fmaxf(sqrtf(d0 * d0 + d1 * d1 + d2 * d2), 1e-6f);
    const float c = (r0 * d0 + r1 * d1 + r2 * d2) * inv;
    if (c > 0.f) { const float k = 2.f * c * __ldg(gori + i) * inv; a0 = fmaf(k, d0, a0); a1 = fmaf(k, d1, a1); a2 = fmaf(k, d2, a2); }
  }
  if (graw) { graw[3 * i] = a0; graw[3 * i + 1] = a1; graw[3 * i + 2] = a2; }
}

}  // namespace ngp

using namespace ngp;

// normals_diff (S,3), normals_ori (S) of models/rendering.py:243-246 from normals_raw, normals_pred, dirs (all (S,3)).
NGP_API int ngp_refloss_prep_fw(const float* normals_raw, const float* normals_pred, const float* dirs, int64_t n, float* normals_diff,
                                float* normals_ori, void* stream) {
  if (n <= 0) return 0;
  refloss_prep_fw_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(normals_raw, normals_pred, dirs, n, normals_diff, normals_ori);
  NGP_LAUNCH_CHECK("ngp_refloss_prep_fw");
  return 0;
}
// upstream g_diff (S,3) | NULL, g_ori (S) | NULL -> g_raw (S,3) | NULL, g_pred (S,3) | NULL
NGP_API int ngp_refloss_prep_bw(const float* normals_raw, const float* normals_pred, const float* dirs, const float* g_diff, const float* g_ori,
                                int64_t n, float* g_raw, float* g_pred, void* stream) {
  if (n <= 0) return 0;
  refloss_prep_bw_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(normals_raw, normals_pred, dirs, g_diff, g_ori, n, g_raw, g_pred);
  NGP_LAUNCH_CHECK("ngp_refloss_prep_bw");
  return 0;
}

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
