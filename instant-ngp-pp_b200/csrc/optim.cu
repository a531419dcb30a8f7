// Fused dense Adam + global-norm gradient clipping for the flat hash-table / MLP parameter vectors.
// Replaces torch.optim.Adam(eps=1e-8) + Lightning's gradient_clip_val=50 (reference train.py:244-251,
// 435): the reference's optimiser moves >= 28 B/param/step through a chain of
// multi-tensor kernels; here one pass reads p,g,m,v and writes p,m,v (28 B/param, the floor for dense
// Adam), with the clip coefficient read from a device scalar so that no host sync is needed.
#include "common.cuh"

namespace ngp {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ out) {
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = (((uintptr_t)g) & 15) == 0 ? (n >> 2) : 0;        // float4 path only for 16-byte aligned views
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += g[i] * g[i];
  acc = warp_sum(acc);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = s[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

// coef = min(1, max_norm / (sqrt(sumsq) + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ coef) {
  const float norm = sqrtf(*sumsq);
  *coef = fminf(1.0f, max_norm / (norm + 1e-6f));
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float b1, float b2, float eps, float step_size,
                                      float inv_sqrt_bc2) {
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);     // torch.optim.Adam: denom = sqrt(v)/sqrt(bc2) + eps
}

// LAZY: entries whose gradient is exactly zero are left alone (p, m, v untouched) — tiny-cuda-nn's Adam for encoding
// parameters (hash-table entries no sample of the batch touched keep their moments instead of coasting on them).  The
// skipped entries cost 4 B (the gradient read) instead of 28 B.  Not torch.optim.Adam's rule, hence opt-in.
template <bool LAZY>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float b1, float b2, float eps,
                                                   float step_size, float inv_sqrt_bc2, const float* __restrict__ gscale) {
  const float gs = gscale ? __ldg(gscale) : 1.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0;
  const int64_t n4 = aligned ? (n >> 2) : 0;                             // parameter views at odd offsets take the scalar loop
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
    if (LAZY && G.x == 0.f && G.y == 0.f && G.z == 0.f && G.w == 0.f) continue;
    float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    if (!LAZY || G.x != 0.f) adam1(P.x, G.x * gs, M.x, V.x, b1, b2, eps, step_size, inv_sqrt_bc2);
    if (!LAZY || G.y != 0.f) adam1(P.y, G.y * gs, M.y, V.y, b1, b2, eps, step_size, inv_sqrt_bc2);
    if (!LAZY || G.z != 0.f) adam1(P.z, G.z * gs, M.z, V.z, b1, b2, eps, step_size, inv_sqrt_bc2);
    if (!LAZY || G.w != 0.f) adam1(P.w, G.w * gs, M.w, V.w, b1, b2, eps, step_size, inv_sqrt_bc2);
    reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (LAZY && g[i] == 0.f) continue;
    adam1(p[i], g[i] * gs, m[i], v[i], b1, b2, eps, step_size, inv_sqrt_bc2);
  }
}

static inline int flat_grid(int64_t n) {
  int64_t b = ceil_div(ceil_div(n, 4), 256);
  const int64_t cap = (int64_t)kSMs * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace ngp

using namespace ngp;

// accum[0] += sum(g^2)   (caller zeroes accum before the first tensor of a step)
NGP_API int ngp_grad_sumsq(const float* g, int64_t n, float* accum, void* stream) {
  if (n <= 0) return 0;
  sumsq_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(g, n, accum);
  NGP_LAUNCH_CHECK("ngp_grad_sumsq");
  return 0;
}

// coef[0] = min(1, max_norm/(sqrt(sumsq[0])+1e-6)) — device resident, consumed by ngp_adam_step
NGP_API int ngp_clip_coef(const float* sumsq, float max_norm, float* coef, void* stream) {
  clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq, max_norm, coef);
  NGP_LAUNCH_CHECK("ngp_clip_coef");
  return 0;
}

// One dense Adam update (torch.optim.Adam semantics, no weight decay / amsgrad), step >= 1.
// grad_scale: optional device scalar multiplied into g (clip coefficient and/or 1/world_size).
NGP_API int ngp_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                          float beta1, float beta2, float eps, int step, const float* grad_scale, void* stream) {
  if (n <= 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  adam_kernel<false><<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, beta1, beta2, eps,
                                                                    (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), grad_scale);
  NGP_LAUNCH_CHECK("ngp_adam_step");
  return 0;
}

// Same update restricted to the entries with a non-zero gradient (tiny-cuda-nn's rule for encoding parameters): the others keep
// p, exp_avg and exp_avg_sq.  `step` is the tensor-wide step count (bias correction is not tracked per entry, as in tcnn).
NGP_API int ngp_adam_step_lazy(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                               float beta1, float beta2, float eps, int step, const float* grad_scale, void* stream) {
  if (n <= 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  adam_kernel<true><<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, beta1, beta2, eps,
                                                                   (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), grad_scale);
  NGP_LAUNCH_CHECK("ngp_adam_step_lazy");
  return 0;
}
