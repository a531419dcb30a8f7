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

// ============================================================================================================
// Occupancy-grid update as ONE kernel chain.  Replaces the reference's
// NGP.sample_uniform_and_occupied_cells + update_density_grid (models/networks.py:308-333,379-408): per cascade
// `randint`, `morton3D`, `nonzero`, `randint`, fancy index, `morton3D_invert`, 2 `cat`s, position arithmetic,
// `rand_like`, a fancy-index scatter, then `where`/`maximum` over the whole grid, a boolean-mask mean with
// `.item()` and `packbits` — ~20 torch ops and one host sync per cascade.  Here:
//   ngp_occupancy_sample : packbits(grid > thr) -> per-cascade exclusive scan of the word popcounts ->
//                          one thread per draw (counter-based RNG): uniform cell, or the r-th occupied cell by
//                          binary search over the 65 536 word prefixes + __fns inside the word; jittered position
//   [density(xyzs) — the field's own kernels, all cascades in one batch]
//   ngp_occupancy_update : decay pass -> atomicMax scatter (non-negative floats order like ints; duplicates of
//                          a cell keep the LARGEST density instead of torch's unspecified winner) -> two-stage
//                          mean of the positive cells, last block writes thr = min(mean, threshold) ->
//                          packbits with the device-resident threshold.
// No host synchronisation, no allocation: all scratch lives in a caller-owned workspace.
namespace ngp {

constexpr int kG = 128;                         // grid_size (models/networks.py:31)
constexpr int kCells = kG * kG * kG;            // per cascade
constexpr int kWords = kCells / 32;             // 65 536 mask words per cascade
constexpr int kMeanBlocks = 296;

struct OccWs {                                  // layout of the workspace (all 16-byte aligned)
  uint32_t* mask;      // [Cc * kWords]
  uint32_t* prefix;    // [Cc * kWords] exclusive prefix of popc(mask) inside the cascade
  uint32_t* total;     // [Cc] (padded to 16)
  double* part;        // [kMeanBlocks * 2] partial (sum, count)
  uint32_t* ticket;    // [1]
  float* thr;          // [1] (+ mean, count for inspection)
};
static inline int64_t occ_ws_bytes(int cascades) {
  return (int64_t)cascades * kWords * 8 + 64 + kMeanBlocks * 16 + 64;
}
static inline OccWs occ_ws(void* ws, int cascades) {
  OccWs w;
  uint8_t* p = (uint8_t*)ws;
  w.mask = (uint32_t*)p; p += (int64_t)cascades * kWords * 4;
  w.prefix = (uint32_t*)p; p += (int64_t)cascades * kWords * 4;
  w.total = (uint32_t*)p; p += 64;
  w.part = (double*)p; p += kMeanBlocks * 16;
  w.ticket = (uint32_t*)p; w.thr = (float*)(p + 16);
  return w;
}

// one block of 1024 threads per cascade; thread t owns words [64t, 64t+64)
__global__ void __launch_bounds__(1024) occ_scan_kernel(const uint32_t* __restrict__ mask, uint32_t* __restrict__ prefix,
                                                        uint32_t* __restrict__ total) {
  __shared__ uint32_t warp_sums[32];
  const uint32_t* m = mask + (int64_t)blockIdx.x * kWords + threadIdx.x * 64;
  uint32_t* out = prefix + (int64_t)blockIdx.x * kWords + threadIdx.x * 64;
  uint32_t cnt[64];
  uint32_t s = 0;
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(m) + q);
    cnt[4 * q] = __popc(v.x); cnt[4 * q + 1] = __popc(v.y); cnt[4 * q + 2] = __popc(v.z); cnt[4 * q + 3] = __popc(v.w);
    s += cnt[4 * q] + cnt[4 * q + 1] + cnt[4 * q + 2] + cnt[4 * q + 3];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = warp_sums[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += n; }
    warp_sums[lane] = wi - w;                                   // exclusive
    if (lane == 31) total[blockIdx.x] = wi;
  }
  __syncthreads();
  uint32_t run = warp_sums[warp] + inc - s;
#pragma unroll
  for (int q = 0; q < 16; q++) {
    uint4 o4;
    o4.x = run; run += cnt[4 * q]; o4.y = run; run += cnt[4 * q + 1]; o4.z = run; run += cnt[4 * q + 2]; o4.w = run; run += cnt[4 * q + 3];
    reinterpret_cast<uint4*>(out)[q] = o4;
  }
}

// counter-based generator: splitmix64 of (seed, counter) — every draw is a pure function of its index, so the
// result does not depend on the launch geometry
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint32_t bits24) { return (float)bits24 * (1.0f / 16777216.0f); }   // [0, 1)

// draws per cascade: warmup -> every cell once (n = kCells, Morton order); else M uniform cells then M occupied ones
__global__ void __launch_bounds__(256) occ_draw_kernel(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ prefix,
                                                       const uint32_t* __restrict__ total, int cascades, float scale, int64_t M,
                                                       int warmup, uint64_t seed, int32_t* __restrict__ indices,
                                                       float* __restrict__ xyzs) {
  const int64_t n = warmup ? (int64_t)kCells : 2 * M;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * cascades) return;
  const int c = (int)(i / n);
  const int64_t j = i - (int64_t)c * n;
  const uint64_t r0 = mix64(seed + (uint64_t)(2 * i) * 0x9E3779B97F4A7C15ull);
  const uint64_t r1 = mix64(seed + (uint64_t)(2 * i + 1) * 0x9E3779B97F4A7C15ull);
  int32_t idx;
  uint32_t cx, cy, cz;
  if (warmup) {
    idx = (int32_t)j;
  } else if (j < M) {
    cx = (uint32_t)(r1 >> 40) & 127u; cy = (uint32_t)(r1 >> 47) & 127u; cz = (uint32_t)(r1 >> 54) & 127u;
    idx = (int32_t)morton3D(cx, cy, cz);
  } else {
    const uint32_t tot = __ldg(total + c);
    if (tot == 0) {
      idx = -1;                      // the reference's index list is empty for this cascade (networks.py:325-329)
    } else {
      const uint32_t r = (uint32_t)(((r1 >> 32) * (uint64_t)tot) >> 32);          // rank in [0, tot)
      const uint32_t* pf = prefix + (int64_t)c * kWords;
      uint32_t lo = 0, hi = kWords;                                               // largest w with pf[w] <= r
      while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(pf + mid) <= r) lo = mid; else hi = mid; }
      const uint32_t word = __ldg(mask + (int64_t)c * kWords + lo);
      const uint32_t bit = __fns(word, 0, (int)(r - __ldg(pf + lo)) + 1);
      idx = (int32_t)(lo * 32u + bit);
    }
  }
  if (warmup || j >= M) {
    const uint32_t m = idx < 0 ? 0u : (uint32_t)idx;
    cx = morton3D_invert(m); cy = morton3D_invert(m >> 1); cz = morton3D_invert(m >> 2);
  }
  // cell centre in world units and a uniform jitter inside the cell (networks.py:391-395)
  const float s = fminf(exp2f((float)(c - 1)), scale);
  const float hgs = s / (float)kG;
  const float sp = s - hgs;
  const float ux = u01((uint32_t)r0 & 0xFFFFFFu), uy = u01((uint32_t)(r0 >> 24) & 0xFFFFFFu), uz = u01((uint32_t)(r0 >> 40) & 0xFFFFFFu);
  float* o = xyzs + 3 * i;
  o[0] = ((float)cx / (float)(kG - 1) * 2.f - 1.f) * sp + (ux * 2.f - 1.f) * hgs;
  o[1] = ((float)cy / (float)(kG - 1) * 2.f - 1.f) * sp + (uy * 2.f - 1.f) * hgs;
  o[2] = ((float)cz / (float)(kG - 1) * 2.f - 1.f) * sp + (uz * 2.f - 1.f) * hgs;
  indices[i] = idx;
}

// grid = grid < 0 ? grid : grid * decay  (decay per cell from count_grid when eroding, networks.py:397-399)
__global__ void __launch_bounds__(256) occ_decay_kernel(float* __restrict__ grid, int64_t n4, float decay, const float* __restrict__ count_grid) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
    float4 v = reinterpret_cast<float4*>(grid)[q];
    float4 d = make_float4(decay, decay, decay, decay);
    if (count_grid) {
      const float4 cg = __ldg(reinterpret_cast<const float4*>(count_grid) + q);
      d.x = fminf(fmaxf(powf(decay, 1.f / cg.x), 0.1f), 0.95f); d.y = fminf(fmaxf(powf(decay, 1.f / cg.y), 0.1f), 0.95f);
      d.z = fminf(fmaxf(powf(decay, 1.f / cg.z), 0.1f), 0.95f); d.w = fminf(fmaxf(powf(decay, 1.f / cg.w), 0.1f), 0.95f);
    }
    v.x = v.x < 0.f ? v.x : v.x * d.x; v.y = v.y < 0.f ? v.y : v.y * d.y;
    v.z = v.z < 0.f ? v.z : v.z * d.z; v.w = v.w < 0.f ? v.w : v.w * d.w;
    reinterpret_cast<float4*>(grid)[q] = v;
  }
}

// grid[c, idx] = max(grid[c, idx], density) for the cells that are not marked invisible (< 0)
__global__ void __launch_bounds__(256) occ_scatter_max_kernel(float* __restrict__ grid, const int32_t* __restrict__ indices,
                                                              const float* __restrict__ density, int64_t n_per_cascade, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t idx = __ldg(indices + i);
  if (idx < 0) return;
  float* cell = grid + (i / n_per_cascade) * (int64_t)kCells + idx;
  if (*cell < 0.f) return;                                        // invisible cells never change, so this read cannot race
  const float d = fmaxf(__ldg(density + i), 0.f);                 // also maps NaN to 0
  atomicMax(reinterpret_cast<int*>(cell), __float_as_int(d));     // non-negative floats order like their bit patterns
}

// mean of the cells > 0 (networks.py:404); the last block to finish adds the partials in block order and writes
// thr = min(mean, density_threshold) (no positive cell: the threshold itself — the bitfield is all zero either way)
__global__ void __launch_bounds__(256) occ_mean_kernel(const float* __restrict__ grid, int64_t n4, float thr_max, double* __restrict__ part,
                                                       uint32_t* __restrict__ ticket, float* __restrict__ thr_out) {
  __shared__ double ssum[8], scnt[8];
  __shared__ bool last;
  double sum = 0.0; uint32_t cnt = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(grid) + q);
    float s = 0.f;
    if (v.x > 0.f) { s += v.x; cnt++; } if (v.y > 0.f) { s += v.y; cnt++; }
    if (v.z > 0.f) { s += v.z; cnt++; } if (v.w > 0.f) { s += v.w; cnt++; }
    sum += (double)s;
  }
  double c = (double)cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
  if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = sum; scnt[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; w++) { a += ssum[w]; b += scnt[w]; }
    part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double a = 0.0, b = 0.0;
    for (unsigned k = 0; k < gridDim.x; k++) { a += __ldcg(part + 2 * k); b += __ldcg(part + 2 * k + 1); }
    const float mean = b > 0.0 ? (float)(a / b) : thr_max;
    thr_out[0] = fminf(mean, thr_max);
    thr_out[1] = mean; thr_out[2] = (float)b;
    *ticket = 0;                                                  // ready for the next update
  }
}

}  // namespace ngp

// Bytes of scratch ngp_occupancy_sample / ngp_occupancy_update need for `cascades` cascades of the 128^3 grid.  The caller
// zero-fills it ONCE after allocation (the kernels leave it ready for the next call).
NGP_API int64_t ngp_occupancy_workspace_bytes(int cascades) { return cascades > 0 ? occ_ws_bytes(cascades) : -1; }

// Cell sampling of the occupancy update (models/networks.py:294-333,388-395).  warmup != 0: every cell of every cascade
// (n_per_cascade = 128^3, Morton order); else M uniform cells followed by M draws (with replacement) among the cells with
// density_grid > density_threshold (n_per_cascade = 2M).  Outputs: indices (cascades * n_per_cascade) int32 Morton cell
// index (-1: the cascade has no occupied cell, the draw is void), xyzs (cascades * n_per_cascade, 3) jittered world
// positions.  `seed` selects the random stream (every draw is a pure function of (seed, its index)).
NGP_API int ngp_occupancy_sample(const float* density_grid, int cascades, float scale, float density_threshold, int64_t M, int warmup,
                                 uint64_t seed, void* workspace, int32_t* indices, float* xyzs, void* stream) {
  if (cascades <= 0 || (!warmup && M <= 0)) return set_error_msg("ngp_occupancy_sample: cascades and M must be positive");
  if ((((uintptr_t)density_grid) | ((uintptr_t)workspace)) & 15) return set_error_msg("ngp_occupancy_sample: density_grid / workspace must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  OccWs w = occ_ws(workspace, cascades);
  if (!warmup) {
    const int64_t n_bytes = (int64_t)cascades * kCells / 8;
    packbits_kernel<float><<<stream_grid(n_bytes / 4, 256), 256, 0, s>>>(density_grid, n_bytes, density_threshold, nullptr, (uint8_t*)w.mask);
    occ_scan_kernel<<<cascades, 1024, 0, s>>>(w.mask, w.prefix, w.total);
  }
  const int64_t n = (warmup ? (int64_t)kCells : 2 * M) * cascades;
  occ_draw_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(w.mask, w.prefix, w.total, cascades, scale, M, warmup, seed, indices, xyzs);
  NGP_LAUNCH_CHECK("ngp_occupancy_sample");
  return 0;
}

// The rest of update_density_grid (models/networks.py:397-408) on the sampled cells' densities:
//   density_grid = where(grid < 0, grid, max(grid * decay, scatter(indices -> density)))   (count_grid != NULL: erode, per-cell decay)
//   density_bitfield = packbits(density_grid, min(mean(density_grid[density_grid > 0]), density_threshold))
// in place, no host read-back.  Duplicate draws of a cell keep the largest density.
NGP_API int ngp_occupancy_update(float* density_grid, int cascades, const int32_t* indices, const float* density, int64_t n_per_cascade,
                                 float decay, const float* count_grid, float density_threshold, void* workspace,
                                 uint8_t* density_bitfield, void* stream) {
  if (cascades <= 0 || n_per_cascade <= 0) return set_error_msg("ngp_occupancy_update: cascades and n_per_cascade must be positive");
  if ((((uintptr_t)density_grid) | ((uintptr_t)workspace) | ((uintptr_t)count_grid)) & 15) return set_error_msg("ngp_occupancy_update: grids / workspace must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  OccWs w = occ_ws(workspace, cascades);
  const int64_t cells = (int64_t)cascades * kCells;
  occ_decay_kernel<<<stream_grid(cells / 4, 256), 256, 0, s>>>(density_grid, cells / 4, decay, count_grid);
  const int64_t n = n_per_cascade * cascades;
  occ_scatter_max_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(density_grid, indices, density, n_per_cascade, n);
  occ_mean_kernel<<<kMeanBlocks, 256, 0, s>>>(density_grid, cells / 4, density_threshold, w.part, w.ticket, w.thr);
  const int64_t n_bytes = cells / 8;
  packbits_kernel<float><<<stream_grid(n_bytes / 4, 256), 256, 0, s>>>(density_grid, n_bytes, 0.f, w.thr, density_bitfield);
  NGP_LAUNCH_CHECK("ngp_occupancy_update");
  return 0;
}
