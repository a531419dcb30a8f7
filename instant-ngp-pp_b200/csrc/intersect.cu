// Ray / AABB and ray / sphere intersection.
// Replaces reference models/csrc/intersection.cu:5-100 (ray_aabb_intersect_cu) and :103-197
// (ray_sphere_intersect_cu).  The reference launches a (rays x voxels) grid with a per-ray atomic
// hit counter, pre-fills three outputs with -1/0 and then runs torch::sort + 2 gathers on the hit
// axis.  Here one thread owns one ray: it walks the voxel list, writes every output slot itself
// (no pre-fill pass, no atomics, deterministic order) and orders its own hits, so the whole
// operator is ONE launch.  The only configuration the reference ever uses (1 voxel, max_hits 1,
// models/rendering.py:28-29) takes the specialised branch below.
#include "common.cuh"

namespace ngp {

struct Hit { float t1, t2; bool hit; };

// intersection.cu:5-22 + :48-52.  Operation order (c-h)-o, (c+h)-o, *inv_d is the float3 operator
// order of helper_math.h; none of these is an FMA candidate.
__device__ __forceinline__ Hit aabb_hit(float ox, float oy, float oz, float ix, float iy, float iz,
                                        float cx, float cy, float cz, float hx, float hy, float hz) {
  const float tminx = __fmul_rn(__fsub_rn(__fsub_rn(cx, hx), ox), ix);
  const float tminy = __fmul_rn(__fsub_rn(__fsub_rn(cy, hy), oy), iy);
  const float tminz = __fmul_rn(__fsub_rn(__fsub_rn(cz, hz), oz), iz);
  const float tmaxx = __fmul_rn(__fsub_rn(__fadd_rn(cx, hx), ox), ix);
  const float tmaxy = __fmul_rn(__fsub_rn(__fadd_rn(cy, hy), oy), iy);
  const float tmaxz = __fmul_rn(__fsub_rn(__fadd_rn(cz, hz), oz), iz);
  const float t1 = fmaxf(fmaxf(fminf(tminx, tmaxx), fminf(tminy, tmaxy)), fminf(tminz, tmaxz));
  const float t2 = fminf(fminf(fmaxf(tminx, tmaxx), fmaxf(tminy, tmaxy)), fmaxf(tminz, tmaxz));
  Hit h;
  if (t1 > t2) { h.t1 = -1.f; h.t2 = -1.f; } else { h.t1 = t1; h.t2 = t2; }
  h.hit = h.t2 > 0.f;
  return h;
}

// intersection.cu:103-121
__device__ __forceinline__ Hit sphere_hit(float ox, float oy, float oz, float dx, float dy, float dz,
                                          float cx, float cy, float cz, float radius) {
  const float cox = ox - cx, coy = oy - cy, coz = oz - cz;
  const float a = dx * dx + dy * dy + dz * dz;
  const float half_b = dx * cox + dy * coy + dz * coz;
  const float c = (cox * cox + coy * coy + coz * coz) - radius * radius;
  const float disc = half_b * half_b - a * c;
  Hit h;
  if (disc < 0) { h.t1 = -1.f; h.t2 = -1.f; }
  else {
    const float sq = sqrtf(disc);
    h.t1 = (-half_b - sq) / a;
    h.t2 = (-half_b + sq) / a;
  }
  h.hit = h.t2 > 0.f;
  return h;
}

// Stable insertion sort of the row's slots by t1 ascending, -1 fillers included — this is what
// torch::sort on hits_t[..., 0] followed by the two gathers does (intersection.cu:94-97).
__device__ __forceinline__ void sort_row(float* __restrict__ row_t, int64_t* __restrict__ row_i, int max_hits) {
  for (int i = 1; i < max_hits; i++) {
    const float k1 = row_t[2 * i], k2 = row_t[2 * i + 1];
    const int64_t ki = row_i[i];
    int j = i - 1;
    while (j >= 0 && row_t[2 * j] > k1) {
      row_t[2 * j + 2] = row_t[2 * j]; row_t[2 * j + 3] = row_t[2 * j + 1]; row_i[j + 1] = row_i[j];
      j--;
    }
    row_t[2 * j + 2] = k1; row_t[2 * j + 3] = k2; row_i[j + 1] = ki;
  }
}

template <bool kSphere>
__global__ void __launch_bounds__(256) intersect_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ centers,
    const float* __restrict__ extent,  // (V,3) half sizes or (V) radii
    int64_t n_rays, int64_t n_vox, int max_hits, int32_t* __restrict__ hit_cnt, float* __restrict__ hits_t,
    int64_t* __restrict__ hits_idx) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rays; r += stride) {
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const float ix = __fdiv_rn(1.0f, dx), iy = __fdiv_rn(1.0f, dy), iz = __fdiv_rn(1.0f, dz);
    float* row_t = hits_t + r * (int64_t)max_hits * 2;
    int64_t* row_i = hits_idx + r * (int64_t)max_hits;
    int cnt = 0;
    for (int64_t v = 0; v < n_vox; v++) {
      Hit h;
      if constexpr (kSphere)
        h = sphere_hit(ox, oy, oz, dx, dy, dz, __ldg(centers + 3 * v), __ldg(centers + 3 * v + 1),
                       __ldg(centers + 3 * v + 2), __ldg(extent + v));
      else
        h = aabb_hit(ox, oy, oz, ix, iy, iz, __ldg(centers + 3 * v), __ldg(centers + 3 * v + 1),
                     __ldg(centers + 3 * v + 2), __ldg(extent + 3 * v), __ldg(extent + 3 * v + 1),
                     __ldg(extent + 3 * v + 2));
      if (h.hit) {
        if (cnt < max_hits) {
          row_t[2 * cnt] = fmaxf(h.t1, 0.0f);
          row_t[2 * cnt + 1] = h.t2;
          row_i[cnt] = v;
        }
        cnt++;
      }
    }
    for (int k = min(cnt, max_hits); k < max_hits; k++) { row_t[2 * k] = -1.f; row_t[2 * k + 1] = -1.f; row_i[k] = -1; }
    hit_cnt[r] = cnt;
    if (max_hits > 1) sort_row(row_t, row_i, max_hits);
  }
}

static inline int ray_grid(int64_t n) {
  int64_t b = ceil_div(n, 256);
  const int64_t cap = (int64_t)kSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// Ray generation: rays_d = R(c2w) . direction, rays_o = t(c2w)  (datasets/ray_utils.py:49-72 get_rays, with the two
// gathers self.poses[img_idxs] / self.directions[pix_idxs] of train.py:136-137 fused in).  Thread per ray; the pose is
// 48 bytes that all rays of an image share (L1), the direction row is a random 12-byte read for a training batch and
// a streamed one for a full frame; six coalesced floats out.  The matrix-vector product is spelled as two FMAs onto
// one product per component (x*R0 first), left to right like a row-by-column dot product.
__global__ void __launch_bounds__(256) get_rays_kernel(const float* __restrict__ directions, const float* __restrict__ poses,
                                                       const int64_t* __restrict__ img_idx, const int64_t* __restrict__ pix_idx,
                                                       int64_t n, float* __restrict__ rays_o, float* __restrict__ rays_d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* P = poses + (img_idx ? __ldg(img_idx + i) : 0) * 12;       // (3,4) row-major camera-to-world
  const float* d = directions + (pix_idx ? __ldg(pix_idx + i) : i) * 3;
  const float dx = __ldg(d), dy = __ldg(d + 1), dz = __ldg(d + 2);
#pragma unroll
  for (int r = 0; r < 3; r++) {
    rays_d[3 * i + r] = fmaf(dz, __ldg(P + 4 * r + 2), fmaf(dy, __ldg(P + 4 * r + 1), dx * __ldg(P + 4 * r)));
    rays_o[3 * i + r] = __ldg(P + 4 * r + 3);
  }
}

}  // namespace ngp

using namespace ngp;

// rays_o, rays_d (R,3) = get_rays(directions[pix_idx], poses[img_idx])  — datasets/ray_utils.py:49-72 as called from
// train.py:136-156.  directions (P,3) camera-space (un-normalised, ray_utils.py:39-40), poses (V,3,4) camera-to-world.
// img_idx == NULL: every ray uses poses[0] (the test-time call with one (3,4) pose); pix_idx == NULL: ray i uses
// directions[i] (a full frame).  Indices are i64 (torch's index dtype) and are NOT range-checked, like torch's gather
// in release builds.
NGP_API int ngp_get_rays(const float* directions, const float* poses, const int64_t* img_idx, const int64_t* pix_idx,
                         int64_t n_rays, float* rays_o, float* rays_d, void* stream) {
  if (n_rays <= 0) return 0;
  get_rays_kernel<<<ray_grid(n_rays), 256, 0, (cudaStream_t)stream>>>(directions, poses, img_idx, pix_idx, n_rays, rays_o, rays_d);
  NGP_LAUNCH_CHECK("ngp_get_rays");
  return 0;
}

// Replaces vren.ray_aabb_intersect (binding.cpp:4-16 -> intersection.cu:59-100).
// Outputs: hit_cnt (R) i32, hits_t (R,max_hits,2) f32 [-1 = empty slot], hits_voxel_idx (R,max_hits) i64.
NGP_API int ngp_ray_aabb_intersect(const float* rays_o, const float* rays_d, const float* centers,
                                   const float* half_sizes, int64_t n_rays, int64_t n_voxels, int max_hits,
                                   int32_t* hit_cnt, float* hits_t, int64_t* hits_voxel_idx, void* stream) {
  if (n_rays <= 0) return 0;
  if (max_hits < 1) return set_error_msg("ngp_ray_aabb_intersect: max_hits must be >= 1");
  intersect_kernel<false><<<ray_grid(n_rays), 256, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, centers, half_sizes, n_rays, n_voxels, max_hits, hit_cnt, hits_t, hits_voxel_idx);
  NGP_LAUNCH_CHECK("ngp_ray_aabb_intersect");
  return 0;
}

// Replaces vren.ray_sphere_intersect (binding.cpp:19-31 -> intersection.cu:156-197).
NGP_API int ngp_ray_sphere_intersect(const float* rays_o, const float* rays_d, const float* centers,
                                     const float* radii, int64_t n_rays, int64_t n_spheres, int max_hits,
                                     int32_t* hit_cnt, float* hits_t, int64_t* hits_sphere_idx, void* stream) {
  if (n_rays <= 0) return 0;
  if (max_hits < 1) return set_error_msg("ngp_ray_sphere_intersect: max_hits must be >= 1");
  intersect_kernel<true><<<ray_grid(n_rays), 256, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, centers, radii, n_rays, n_spheres, max_hits, hit_cnt, hits_t, hits_sphere_idx);
  NGP_LAUNCH_CHECK("ngp_ray_sphere_intersect");
  return 0;
}

// hits_t[(hits_t[:,0,0] >= 0) & (hits_t[:,0,0] < near), 0, 0] = near   (models/rendering.py:29-30: rays that start inside the box begin at
// the near plane) — in place, one launch instead of the boolean-mask index_put (compare, compare, and, nonzero / where, copy).
namespace ngp {
__global__ void __launch_bounds__(256) near_clamp_kernel(float* __restrict__ hits_t, int64_t n, int64_t row_stride, float near_t) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t1 = hits_t[i * row_stride];
  if (t1 >= 0.f && t1 < near_t) hits_t[i * row_stride] = near_t;
}
}  // namespace ngp
NGP_API int ngp_near_clamp(float* hits_t, int64_t n_rays, int64_t row_stride, float near_distance, void* stream) {
  if (n_rays <= 0) return 0;
  ngp::near_clamp_kernel<<<(unsigned)ngp::ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(hits_t, n_rays, row_stride, near_distance);
  NGP_LAUNCH_CHECK("ngp_near_clamp");
  return 0;
}
