// Occupancy-grid maintenance kernels: morton3D, morton3D_invert, packbits.
// Replaces reference models/csrc/raymarching.cu:62-161 (morton3D_cu, morton3D_invert_cu,
// packbits_cu).  All three are HBM-streaming integer kernels; they are written as
// grid-stride loops over a fixed, SM-multiple grid with 16-byte vector accesses.
#include "common.cuh"

namespace ngp {

// coords (N,3) int32 AoS -> indices (N) int32.  Each thread handles 4 consecutive cells:
// 3 x int4 loads (48 B) -> 1 x int4 store, fully coalesced across the warp.
__global__ void __launch_bounds__(256) morton3D_kernel(const int32_t* __restrict__ coords, int64_t n,
                                                       int32_t* __restrict__ indices) {
  // 16-byte vector path only for 16-byte aligned views; an offset view (coords[1:]) takes the scalar loop like the
  // reference's scalar accessors would (raymarching.cu:62-70)
  const bool aligned = ((((uintptr_t)coords) | ((uintptr_t)indices)) & 15) == 0;
  const int64_t n4 = aligned ? (n >> 2) : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
    const int4* src = reinterpret_cast<const int4*>(coords) + q * 3;
    const int4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
    int4 o;
    o.x = (int32_t)morton3D(a.x, a.y, a.z);
    o.y = (int32_t)morton3D(a.w, b.x, b.y);
    o.z = (int32_t)morton3D(b.z, b.w, c.x);
    o.w = (int32_t)morton3D(c.y, c.z, c.w);
    reinterpret_cast<int4*>(indices)[q] = o;
  }
  // tail (n % 4 cells; everything when unaligned)
  for (int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
    indices[t] = (int32_t)morton3D(coords[3 * t], coords[3 * t + 1], coords[3 * t + 2]);
}

__global__ void __launch_bounds__(256) morton3D_invert_kernel(const int32_t* __restrict__ indices, int64_t n,
                                                              int32_t* __restrict__ coords) {
  const bool aligned = ((((uintptr_t)coords) | ((uintptr_t)indices)) & 15) == 0;
  const int64_t n4 = aligned ? (n >> 2) : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
    const int4 i = __ldg(reinterpret_cast<const int4*>(indices) + q);
    // `ind >> k` is an arithmetic shift of a signed int in the reference (raymarching.cu:97-100);
    // the invert mask keeps only bits < 30 so the sign fill never survives.
    int4 a, b, c;
    a.x = morton3D_invert((uint32_t)(i.x >> 0)); a.y = morton3D_invert((uint32_t)(i.x >> 1)); a.z = morton3D_invert((uint32_t)(i.x >> 2));
    a.w = morton3D_invert((uint32_t)(i.y >> 0)); b.x = morton3D_invert((uint32_t)(i.y >> 1)); b.y = morton3D_invert((uint32_t)(i.y >> 2));
    b.z = morton3D_invert((uint32_t)(i.z >> 0)); b.w = morton3D_invert((uint32_t)(i.z >> 1)); c.x = morton3D_invert((uint32_t)(i.z >> 2));
    c.y = morton3D_invert((uint32_t)(i.w >> 0)); c.z = morton3D_invert((uint32_t)(i.w >> 1)); c.w = morton3D_invert((uint32_t)(i.w >> 2));
    int4* dst = reinterpret_cast<int4*>(coords) + q * 3;
    dst[0] = a; dst[1] = b; dst[2] = c;
  }
  for (int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const int32_t ind = indices[t];
    coords[3 * t + 0] = morton3D_invert((uint32_t)(ind >> 0));
    coords[3 * t + 1] = morton3D_invert((uint32_t)(ind >> 1));
    coords[3 * t + 2] = morton3D_invert((uint32_t)(ind >> 2));
  }
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

// One thread packs 32 consecutive cells into one 32-bit word: eight 16-byte loads in flight per
// thread (128 B), one 4-byte store, consecutive threads on consecutive 128-byte lines.  Bit order is the
// reference's (bit i of byte n <-> cell 8n+i, raymarching.cu:133-140, strict '>'); the comparison is
// done in the type the reference compares in (scalar_t > float ==> float for half/float, double for
// double).  A scalar tail handles byte counts that are not a multiple of 4.
template <typename T> __device__ __forceinline__ bool above(T v, float thr) {
  if constexpr (sizeof(T) == 8) return v > (double)thr;
  else return to_f<T>(v) > thr;
}

template <typename T>
__global__ void __launch_bounds__(256) packbits_kernel(const T* __restrict__ grid, int64_t n_bytes, float thr,
                                                       const float* __restrict__ thr_dev, uint8_t* __restrict__ bitfield) {
  if (thr_dev) thr = __ldg(thr_dev);          // threshold produced on the device (occupancy update): no host read-back
  const int64_t n_words = n_bytes >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool aligned = (((uintptr_t)grid) & 15) == 0 && (((uintptr_t)bitfield) & 3) == 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
    const T* src = grid + (w << 5);
    uint32_t bits = 0;
    if constexpr (sizeof(T) == 4) {
      if (aligned) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = __ldg(reinterpret_cast<const float4*>(src) + k);
#pragma unroll
        for (int k = 0; k < 8; k++) {
          bits |= (v[k].x > thr ? 1u : 0u) << (4 * k);
          bits |= (v[k].y > thr ? 1u : 0u) << (4 * k + 1);
          bits |= (v[k].z > thr ? 1u : 0u) << (4 * k + 2);
          bits |= (v[k].w > thr ? 1u : 0u) << (4 * k + 3);
        }
        reinterpret_cast<uint32_t*>(bitfield)[w] = bits;
        continue;
      }
    }
#pragma unroll 8
    for (int k = 0; k < 32; k++) bits |= (above<T>(src[k], thr) ? 1u : 0u) << k;
    if (aligned) reinterpret_cast<uint32_t*>(bitfield)[w] = bits;
    else { for (int b = 0; b < 4; b++) bitfield[(w << 2) + b] = (uint8_t)(bits >> (8 * b)); }
  }
  // tail bytes
  const int64_t tb = (n_words << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tb < n_bytes) {
    uint8_t b8 = 0;
    for (int k = 0; k < 8; k++) b8 |= above<T>(grid[8 * tb + k], thr) ? (uint8_t)(1u << k) : 0;
    bitfield[tb] = b8;
  }
}

static inline int stream_grid(int64_t work_items, int per_block) {
  int64_t b = ceil_div(work_items, per_block);
  const int64_t cap = (int64_t)kSMs * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace ngp

using namespace ngp;

// Replaces vren.morton3D (binding.cpp:46-50 -> raymarching.cu:72-88).
NGP_API int ngp_morton3D(const int32_t* coords, int64_t n, int32_t* indices, void* stream) {
  if (n <= 0) return 0;
  morton3D_kernel<<<stream_grid((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(coords, n, indices);
  NGP_LAUNCH_CHECK("ngp_morton3D");
  return 0;
}

// Replaces vren.morton3D_invert (binding.cpp:53-57 -> raymarching.cu:103-119).
NGP_API int ngp_morton3D_invert(const int32_t* indices, int64_t n, int32_t* coords, void* stream) {
  if (n <= 0) return 0;
  morton3D_invert_kernel<<<stream_grid((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(indices, n, coords);
  NGP_LAUNCH_CHECK("ngp_morton3D_invert");
  return 0;
}

// Replaces vren.packbits (binding.cpp:35-43 -> raymarching.cu:143-161).
// dtype: 0 = float32, 1 = float16, 2 = float64 (the reference's AT_DISPATCH_FLOATING_TYPES_AND_HALF).
NGP_API int ngp_packbits(const void* density_grid, int dtype, int64_t n_bytes, float density_threshold,
                         uint8_t* density_bitfield, void* stream) {
  if (n_bytes <= 0) return 0;
  const int grid = stream_grid((n_bytes + 3) / 4, 256);
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case 0: packbits_kernel<float><<<grid, 256, 0, s>>>((const float*)density_grid, n_bytes, density_threshold, nullptr, density_bitfield); break;
    case 1: packbits_kernel<__half><<<grid, 256, 0, s>>>((const __half*)density_grid, n_bytes, density_threshold, nullptr, density_bitfield); break;
    case 2: packbits_kernel<double><<<grid, 256, 0, s>>>((const double*)density_grid, n_bytes, density_threshold, nullptr, density_bitfield); break;
    default: return set_error_msg("ngp_packbits: dtype must be 0 (f32), 1 (f16) or 2 (f64)");
  }
  NGP_LAUNCH_CHECK("ngp_packbits");
  return 0;
}

// Same as ngp_packbits with the threshold read from DEVICE memory (one float): the occupancy update
// (models/networks.py:404-408: thr = min(mean density, density_threshold)) stays on the stream, no .item().
NGP_API int ngp_packbits_dthr(const float* density_grid, int64_t n_bytes, const float* density_threshold_dev,
                              uint8_t* density_bitfield, void* stream) {
  if (n_bytes <= 0) return 0;
  const int grid = stream_grid((n_bytes + 3) / 4, 256);
  packbits_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(density_grid, n_bytes, 0.f, density_threshold_dev, density_bitfield);
  NGP_LAUNCH_CHECK("ngp_packbits_dthr");
  return 0;
}
