// Shared device helpers for the ngp_b200 C-ABI kernels (sm_100a only).
//
// Arithmetic in this header follows the reference's per-ray scalar recurrences
// (models/csrc/raymarching.cu:4-60) because per-ray sample counts / t / dt are a bit-exact
// parity target.  Every floating-point step that the reference build contracts into an FMA
// is spelt with an explicit __fmaf_rn here, and every step it does NOT contract uses
// __fmul_rn/__fadd_rn, so the result does not depend on this file's optimisation flags
// (see DESIGN.md "bit-exact marching" for the SASS evidence the sequence was read from).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#ifndef NGP_API
#define NGP_API extern "C" __attribute__((visibility("default")))
#endif

namespace ngp {

constexpr float kSqrt3 = 1.73205080757f;   // raymarching.cu:4
constexpr int kSMs = 148;                  // B200

int set_error(cudaError_t e, const char* where);   // api.cu
int set_error_msg(const char* msg);

// internal launchers shared between translation units (not exported): the `_dn` forms take the element count from device
// memory (n is then the bound the launch is sized for) — used by the test-time wavefront round, ngp_render_round_compact
int mlp_fw_launch(int n_seg, const float* const* seg_ptr, const int* seg_width, const int* seg_kind, const int64_t* seg_stride,
                  const float* params, int width, int n_hidden, int n_out, int act_hidden, int act_out, int64_t n,
                  const int32_t* n_dev, float* out, int64_t out_stride, float* aux_exp_out, void* stream);
int hashgrid_fw_tiles_launch(const float* x, const float* aabb, const void* table, int table_dtype, int n_levels, int n_features,
                             int log2_hashmap_size, int base_resolution, float per_level_scale, int64_t n, const int32_t* n_dev,
                             void* y_tiles, void* stream);

#define NGP_LAUNCH_CHECK(where)                                   \
  do {                                                            \
    cudaError_t _e = cudaGetLastError();                          \
    if (_e != cudaSuccess) return ngp::set_error(_e, where);      \
  } while (0)

__host__ __device__ __forceinline__ int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- morton (raymarching.cu:35-60)
__host__ __device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
__host__ __device__ __forceinline__ uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) {
  return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t morton3D_invert(uint32_t x) {
  x = x & 0x49249249u;
  x = (x | (x >> 2)) & 0xc30c30c3u;
  x = (x | (x >> 4)) & 0x0f00f00fu;
  x = (x | (x >> 8)) & 0xff0000ffu;
  x = (x | (x >> 16)) & 0x0000ffffu;
  return x;
}

// ---------------------------------------------------------------- marching scalar helpers
// Loop-invariant pieces of calc_dt (raymarching.cu:11-13):
//   clamp(t*esf, SQRT3/max_samples, SQRT3*2*scale/grid_size), clamp(f,a,b)=fmaxf(a,fminf(f,b))
struct DtParams {
  float esf, dt_min, dt_max;
};
__host__ __device__ __forceinline__ DtParams make_dt_params(float esf, int max_samples, int grid_size,
                                                            float scale_arg) {
  DtParams p;
  p.esf = esf;
#ifdef __CUDA_ARCH__
  p.dt_min = __fdiv_rn(kSqrt3, (float)max_samples);
  p.dt_max = __fdiv_rn(__fmul_rn(kSqrt3 * 2, scale_arg), (float)grid_size);
#else
  p.dt_min = kSqrt3 / (float)max_samples;
  p.dt_max = ((kSqrt3 * 2) * scale_arg) / (float)grid_size;
#endif
  return p;
}
__device__ __forceinline__ float calc_dt(float t, const DtParams& p) {
  return fmaxf(p.dt_min, fminf(__fmul_rn(t, p.esf), p.dt_max));
}
// frexpf exponent of a finite float (0 -> 0), matching CUDA's frexpf incl. denormals.
__device__ __forceinline__ int frexp_exponent(float v) {
  int e;
  frexpf(v, &e);
  return e;
}
__device__ __forceinline__ int mip_from_pos(float x, float y, float z, int cascades) {
  const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
  return min(cascades - 1, max(0, frexp_exponent(mx) + 1));
}
__device__ __forceinline__ int mip_from_dt(float dt, int grid_size, int cascades) {
  return min(cascades - 1, max(0, frexp_exponent(__fmul_rn(dt, (float)grid_size))));
}

// ---------------------------------------------------------------- density-net activations (density_head.cu, density_net.cu)
// softplus and sigmoid of the same z from ONE exponential: e = exp(-|z|) in (0, 1],
//     sigmoid(z) = z >= 0 ? 1/(1+e) : e/(1+e) ;  softplus(z) = max(z, 0) + log(1 + e)
// (no overflow for any z; above torch's linear threshold 20 the log term is < 2.1e-9 and vanishes in fp32 exactly as
// torch's switch to the identity does).  MUFU.EX2 + MUFU.RCP + MUFU.LG2: with expf / log1pf / an IEEE division per
// element the kernels were ALU bound at 42 % (fw) / 56 % (bw) of the HBM rate (profiles/r01e_step_profile_playground_after.txt).
struct SpSg { float sp, sg; };
__device__ __forceinline__ SpSg softplus_sigmoid(float z) {
  const float e = __expf(-fabsf(z));
  float r;                                     // 1 + e is in (1, 2]: MUFU.RCP alone (1 ulp); __frcp_rn's IEEE fix-up was 45 % of the
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));      // density-net kernels' instructions (profiles/r02d_ncu_density_net_*)
  SpSg o;
  o.sg = z >= 0.f ? r : e * r;
  o.sp = fmaxf(z, 0.f) + __logf(1.f + e);
  return o;
}
__device__ __forceinline__ float softplus1(float z) { return softplus_sigmoid(z).sp; }
__device__ __forceinline__ float sigmoid1(float z) { return softplus_sigmoid(z).sg; }

// ---------------------------------------------------------------- warp utilities
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ngp
