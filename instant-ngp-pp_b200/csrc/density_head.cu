// Density net with analytic normals: the elementwise stages of
//     z1 = W1 e + b1 ; a1 = softplus(z1) ; z2 = w2 . a1 + b2 ; sigma = softplus(z2)
//     g_e = d sigma / d e = W1^T (sigmoid(z2) * sigmoid(z1) * w2)
// and of the backward of BOTH outputs (sigma and g_e — the second is the double backward the reference gets from
// torch.autograd.grad(..., create_graph=True), models/networks.py:54-59,186-196).  The reference keeps this net in
// torch: per sample ~30 elementwise passes over (S, 128) tensors (Softplus, its backward and double backward, the
// broadcasts and sums around them: 70 ms of a 190 ms step at 14 M samples, profiles/r01e_step_profile_playground_after.txt).
// Here each direction is ONE pass: a warp owns a row of W = 128*k floats (float4 per lane), row reductions are
// shuffles, column sums (bias / w2 gradients) are kept in registers over a warp's rows and leave once per CTA.
// The four GEMMs with W1 stay with the caller (cuBLAS TF32, tensor pipe).  HBM bound: fw 2*W*4 B/sample, bw 3*W*4.
//
// With s1 = sigmoid(z1), s2 = sigmoid(z2), t = s2 * s1 * w2 (so that g_e = t W1), upstream (dsigma, dg_e) and
// v = dg_e W1^T:
//     uv   = s1 * v                          ds2 = uv . w2
//     dz2  = ds2 * s2 (1 - s2) + dsigma * s2
//     dz1  = w2 * (uv * s2 * (1 - s1) + dz2 * s1)
//     dw2  = sum_rows (s2 * uv + dz2 * a1)   db1 = sum_rows dz1    (db2 = sum dz2, dW1 = t^T dg_e + dz1^T e, de = dz1 W1: caller)
// torch's softplus is linear above threshold 20; sigmoid(z > 20) rounds to 1 in fp32, so the derivatives agree.
// (sigma at z2 > 20: z2 + log(1+e^-z2) = z2 in fp32, the same value.)
#include "common.cuh"

namespace ngp {

constexpr int kHeadMaxK = 4;   // W <= 512

template <int K>
__global__ void __launch_bounds__(256) density_head_fw_kernel(const float* __restrict__ z1, const float* __restrict__ w2,
                                                              const float* __restrict__ b2, int64_t n,
                                                              float* __restrict__ sigma, float* __restrict__ s2_out,
                                                              float* __restrict__ t) {
  constexpr int W = 128 * K;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 w[K];
#pragma unroll
  for (int k = 0; k < K; k++) w[k] = __ldg(reinterpret_cast<const float4*>(w2) + k * 32 + lane);
  const float bias = __ldg(b2);
  for (int64_t r = warp; r < n; r += n_warps) {
    float4 s[K];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const float4 z = __ldcs(reinterpret_cast<const float4*>(z1 + r * W) + k * 32 + lane);
      const SpSg a = softplus_sigmoid(z.x), b = softplus_sigmoid(z.y), c = softplus_sigmoid(z.z), d = softplus_sigmoid(z.w);
      dot += a.sp * w[k].x + b.sp * w[k].y + c.sp * w[k].z + d.sp * w[k].w;
      s[k] = make_float4(a.sg, b.sg, c.sg, d.sg);
    }
    const SpSg h = softplus_sigmoid(warp_sum(dot) + bias);
    const float sg = h.sg;
    if (lane == 0) { sigma[r] = h.sp; s2_out[r] = sg; }
#pragma unroll
    for (int k = 0; k < K; k++)
      __stcs(reinterpret_cast<float4*>(t + r * W) + k * 32 + lane,
             make_float4(sg * s[k].x * w[k].x, sg * s[k].y * w[k].y, sg * s[k].z * w[k].z, sg * s[k].w * w[k].w));
  }
}

template <int K>
__global__ void __launch_bounds__(256) density_head_bw_kernel(const float* __restrict__ z1, const float* __restrict__ v,
                                                              const float* __restrict__ s2_in, const float* __restrict__ dsigma,
                                                              const float* __restrict__ w2, int64_t n,
                                                              float* __restrict__ dz1, float* __restrict__ dz2_out,
                                                              float* __restrict__ dw2, float* __restrict__ db1) {
  constexpr int W = 128 * K;
  __shared__ float red[2 * W];
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 w[K], aw[K], ab[K];     // w2 ; running column sums of the w2 / b1 gradients
#pragma unroll
  for (int k = 0; k < K; k++) {
    w[k] = __ldg(reinterpret_cast<const float4*>(w2) + k * 32 + lane);
    aw[k] = make_float4(0.f, 0.f, 0.f, 0.f); ab[k] = aw[k];
  }
  for (int i = threadIdx.x; i < 2 * W; i += blockDim.x) red[i] = 0.f;
  for (int64_t r = warp; r < n; r += n_warps) {
    float4 s[K], a[K], uv[K];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const float4 z = __ldcs(reinterpret_cast<const float4*>(z1 + r * W) + k * 32 + lane);
      const float4 vv = v ? __ldcs(reinterpret_cast<const float4*>(v + r * W) + k * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      const SpSg p0 = softplus_sigmoid(z.x), p1 = softplus_sigmoid(z.y), p2 = softplus_sigmoid(z.z), p3 = softplus_sigmoid(z.w);
      s[k] = make_float4(p0.sg, p1.sg, p2.sg, p3.sg);
      a[k] = make_float4(p0.sp, p1.sp, p2.sp, p3.sp);
      uv[k] = make_float4(s[k].x * vv.x, s[k].y * vv.y, s[k].z * vv.z, s[k].w * vv.w);
      dot += uv[k].x * w[k].x + uv[k].y * w[k].y + uv[k].z * w[k].z + uv[k].w * w[k].w;
    }
    const float ds2 = warp_sum(dot);
    const float sg = __ldg(s2_in + r);
    const float dz2 = ds2 * sg * (1.f - sg) + (dsigma ? __ldg(dsigma + r) : 0.f) * sg;
    if (lane == 0) dz2_out[r] = dz2;
#pragma unroll
    for (int k = 0; k < K; k++) {
      float4 d;
      d.x = w[k].x * (uv[k].x * sg * (1.f - s[k].x) + dz2 * s[k].x);
      d.y = w[k].y * (uv[k].y * sg * (1.f - s[k].y) + dz2 * s[k].y);
      d.z = w[k].z * (uv[k].z * sg * (1.f - s[k].z) + dz2 * s[k].z);
      d.w = w[k].w * (uv[k].w * sg * (1.f - s[k].w) + dz2 * s[k].w);
      __stcs(reinterpret_cast<float4*>(dz1 + r * W) + k * 32 + lane, d);
      ab[k].x += d.x; ab[k].y += d.y; ab[k].z += d.z; ab[k].w += d.w;
      aw[k].x += sg * uv[k].x + dz2 * a[k].x; aw[k].y += sg * uv[k].y + dz2 * a[k].y;
      aw[k].z += sg * uv[k].z + dz2 * a[k].z; aw[k].w += sg * uv[k].w + dz2 * a[k].w;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int c = (k * 32 + lane) * 4;
    atomicAdd(&red[c], aw[k].x); atomicAdd(&red[c + 1], aw[k].y); atomicAdd(&red[c + 2], aw[k].z); atomicAdd(&red[c + 3], aw[k].w);
    atomicAdd(&red[W + c], ab[k].x); atomicAdd(&red[W + c + 1], ab[k].y); atomicAdd(&red[W + c + 2], ab[k].z); atomicAdd(&red[W + c + 3], ab[k].w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < W; i += blockDim.x) { atomicAdd(dw2 + i, red[i]); atomicAdd(db1 + i, red[W + i]); }
}

}  // namespace ngp

using namespace ngp;

static int head_grid(int64_t n) {
  const int64_t blocks = ceil_div(n, 8);                       // 8 warps (rows) per CTA pass
  return (int)(blocks < (int64_t)kSMs * 8 ? blocks : (int64_t)kSMs * 8);
}

// Forward elementwise stage of the density net (models/networks.py:54-59 xyz_net + sigma_act, and the first half of
// the autograd normals, networks.py:186-196).  z1 (N,W) = e W1^T + b1 (caller's GEMM), w2 (W), b2 (1): all device.
// Out: sigma (N) = softplus(w2 . softplus(z1) + b2), s2 (N) = sigmoid of the same pre-activation,
// t (N,W) = s2 * sigmoid(z1) * w2  (d sigma / d e = t W1).  W must be a multiple of 128, <= 512.
NGP_API int ngp_density_head_fw(const float* z1, const float* w2, const float* b2, int64_t n, int width, float* sigma,
                                float* s2, float* t, void* stream) {
  if (n <= 0) return 0;
  if (width % 128 != 0 || width < 128 || width > 128 * kHeadMaxK) return set_error_msg("ngp_density_head_fw: width must be 128, 256, 384 or 512");
  cudaStream_t st = (cudaStream_t)stream;
  switch (width / 128) {
    case 1: density_head_fw_kernel<1><<<head_grid(n), 256, 0, st>>>(z1, w2, b2, n, sigma, s2, t); break;
    case 2: density_head_fw_kernel<2><<<head_grid(n), 256, 0, st>>>(z1, w2, b2, n, sigma, s2, t); break;
    case 3: density_head_fw_kernel<3><<<head_grid(n), 256, 0, st>>>(z1, w2, b2, n, sigma, s2, t); break;
    default: density_head_fw_kernel<4><<<head_grid(n), 256, 0, st>>>(z1, w2, b2, n, sigma, s2, t); break;
  }
  NGP_LAUNCH_CHECK("ngp_density_head_fw");
  return 0;
}

// Backward elementwise stage for upstream (dsigma (N) | NULL, v (N,W) = dL/dg_e W1^T | NULL): dz1 (N,W), dz2 (N), and
// += into dw2 (W), db1 (W) (caller zeroes).  See the header comment for the formulas.
NGP_API int ngp_density_head_bw(const float* z1, const float* v, const float* s2, const float* dsigma, const float* w2,
                                int64_t n, int width, float* dz1, float* dz2, float* dw2, float* db1, void* stream) {
  if (n <= 0) return 0;
  if (width % 128 != 0 || width < 128 || width > 128 * kHeadMaxK) return set_error_msg("ngp_density_head_bw: width must be 128, 256, 384 or 512");
  cudaStream_t st = (cudaStream_t)stream;
  switch (width / 128) {
    case 1: density_head_bw_kernel<1><<<head_grid(n), 256, 0, st>>>(z1, v, s2, dsigma, w2, n, dz1, dz2, dw2, db1); break;
    case 2: density_head_bw_kernel<2><<<head_grid(n), 256, 0, st>>>(z1, v, s2, dsigma, w2, n, dz1, dz2, dw2, db1); break;
    case 3: density_head_bw_kernel<3><<<head_grid(n), 256, 0, st>>>(z1, v, s2, dsigma, w2, n, dz1, dz2, dw2, db1); break;
    default: density_head_bw_kernel<4><<<head_grid(n), 256, 0, st>>>(z1, v, s2, dsigma, w2, n, dz1, dz2, dw2, db1); break;
  }
  NGP_LAUNCH_CHECK("ngp_density_head_bw");
  return 0;
}
