// Occupancy-bitfield ray marching, training and test time.
// Replaces reference models/csrc/raymarching.cu:166-332 (raymarching_train_cu) and :335-454
// (raymarching_test_cu).
//
// Reference structure: one kernel, thread per ray, two serial passes joined by two independent
// global atomics (nondeterministic, mutually inconsistent sample order), behind a host-side zero
// fill of N_rays*max_samples rows of every output (8.6 GB at 2^18 rays).
//
// Here (training): three launches, no atomics, no fill, deterministic ray-index order
//   1. march_count   : the ONE serial march per ray -> n_samples[r], and the (t, dt) of its first
//                      kScratch samples into a per-ray scratch row; block sum of counts
//   2. block_scan    : one CTA scans the per-block sums -> block offsets, total (device resident)
//   3. march_emit    : block-local exclusive scan of counts + block offset = start_idx; a group of
//                      8 lanes per ray turns the scratch row into rays_a / xyzs (= fma(t,d,o), the same
//                      rounding as in the loop) / dirs / deltas / ts with coalesced stores.  Only
//                      rays with more than kScratch samples are re-marched (by one lane).
// The per-ray float recurrence (t += dt, cell lookup, empty-cell skip) is the reference's, step by
// step, with the FMA contractions of the reference build made explicit (common.cuh).
#include "common.cuh"
#include <stdlib.h>

namespace ngp {

constexpr int kMarchBlock = 256;
constexpr int kScratch = 256;    // test-time wavefront: samples per slot and round recorded in the scratch row (n_next <= 256)
constexpr int kTrainRow = 1024;  // training: scratch row = max_samples (models/rendering.py:9) entries of 8 B, so pass 2 never
                                 // re-marches (a 256-entry row left the tails of the longest rays — 407 samples in the Lego-shaped
                                 // scene, most rays of the street-shaped one — to ONE serial thread each: the emit kernel's critical
                                 // path).  2 GiB of address space at 2^18 rays; only touched entries cost bandwidth.

struct MarchParams {
  float dt0, mb0, mb0_inv;   // kSimple constants: dt = clamp(0, dt_min, dt_max), mip_bound = min(0.5, scale), 1/mip_bound
  const uint8_t* __restrict__ bitfield;
  int cascades;
  int grid_size;
  float scale;      // used for mip_bound (raymarching.cu:211)
  DtParams dt;      // calc_dt constants (built with `scale` for train, `cascades` for test — raymarching.cu:370)
  float grid_f;     // (float)grid_size
  float grid_inv;   // 1.0f/grid_size
  float grid_m1;    // grid_size-1.0f
  uint32_t grid3;
  // empty-ray culling (training, kSimple only; see coarse_occupancy_kernel): 32^3 bits, bit = "some occupied cell within
  // the 4^3 block or its 26 neighbour blocks"; NULL = off.  cull_span = 4 cells in world units.
  const uint32_t* __restrict__ coarse;
  float cull_span;
  // x-major copy of the (single-cascade, 128^3) bitfield: bit (z*128 + y)*128 + x of a 32-bit word array (linearize_bitfield_kernel).
  // The same bits under a cheaper address: the Morton spread costs 24 of the ~85 instructions of a marching step.  NULL = off.
  const uint32_t* __restrict__ linear;
};

struct Ray {
  float ox, oy, oz, dx, dy, dz, dx_inv, dy_inv, dz_inv, sx, sy, sz;
};

__device__ __forceinline__ Ray load_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t r) {
  Ray q;
  q.ox = __ldg(rays_o + 3 * r); q.oy = __ldg(rays_o + 3 * r + 1); q.oz = __ldg(rays_o + 3 * r + 2);
  q.dx = __ldg(rays_d + 3 * r); q.dy = __ldg(rays_d + 3 * r + 1); q.dz = __ldg(rays_d + 3 * r + 2);
  q.dx_inv = __fdiv_rn(1.0f, q.dx); q.dy_inv = __fdiv_rn(1.0f, q.dy); q.dz_inv = __fdiv_rn(1.0f, q.dz);
  // 0.5f*signf(d)  (signf = copysignf(1,d), raymarching.cu:7)
  q.sx = copysignf(0.5f, q.dx); q.sy = copysignf(0.5f, q.dy); q.sz = copysignf(0.5f, q.dz);
  return q;
}

// One marching step at parameter t (raymarching.cu:205-232).  Returns true when the cell is occupied
// (then dt is the step to take and x,y,z the sample); otherwise advances t past the empty cell.
// kSimple (cascades == 1 and exp_step_factor == 0, the synthetic-scene case): calc_dt(t) is the constant
// fmaxf(dt_min, fminf(t*0, dt_max)) = dt_min for every finite t, both mip selectors clamp to 0, and
// mip_bound = fminf(2^-1, scale) — the same VALUES the general path computes, hoisted out of the loop.
template <bool kSimple = false, bool kLinear = false>
__device__ __forceinline__ bool march_step(const Ray& q, const MarchParams& p, float& t, float& x, float& y,
                                           float& z, float& dt) {
  x = __fmaf_rn(t, q.dx, q.ox); y = __fmaf_rn(t, q.dy, q.oy); z = __fmaf_rn(t, q.dz, q.oz);
  dt = kSimple ? p.dt0 : calc_dt(t, p.dt);
  const int mip = kSimple ? 0 : max(mip_from_pos(x, y, z, p.cascades), mip_from_dt(dt, p.grid_size, p.cascades));
  const float mip_bound = kSimple ? p.mb0 : fminf(scalbnf(1.0f, mip - 1), p.scale);
  const float mip_bound_inv = kSimple ? p.mb0_inv : __fdiv_rn(1.0f, mip_bound);
  // round down to the containing cell: (int)clamp(0.5f*(x*inv+1)*G, 0, G-1)
  const int nx = (int)fmaxf(0.0f, fminf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(x, mip_bound_inv, 1.0f)), p.grid_f), p.grid_m1));
  const int ny = (int)fmaxf(0.0f, fminf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(y, mip_bound_inv, 1.0f)), p.grid_f), p.grid_m1));
  const int nz = (int)fmaxf(0.0f, fminf(__fmul_rn(__fmul_rn(0.5f, __fmaf_rn(z, mip_bound_inv, 1.0f)), p.grid_f), p.grid_m1));
  bool occ;
  if (kLinear) {                                 // kSimple, grid 128: mip == 0
    const uint32_t lin = ((uint32_t)nz * 128u + (uint32_t)ny) * 128u + (uint32_t)nx;
    occ = (__ldg(p.linear + (lin >> 5)) >> (lin & 31u)) & 1u;
  } else {
    const uint32_t idx = (uint32_t)mip * p.grid3 + morton3D((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
    occ = __ldg(p.bitfield + (idx >> 3)) & (1u << (idx & 7u));
  }
  if (occ) return true;
  // distance to the far faces of this cell along the ray
  const float tx = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)nx, 0.5f), q.sx), p.grid_inv), 2.0f, -1.0f), mip_bound, -x), q.dx_inv);
  const float ty = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)ny, 0.5f), q.sy), p.grid_inv), 2.0f, -1.0f), mip_bound, -y), q.dy_inv);
  const float tz = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmul_rn(__fadd_rn(__fadd_rn((float)nz, 0.5f), q.sz), p.grid_inv), 2.0f, -1.0f), mip_bound, -z), q.dz_inv);
  const float t_target = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
  if (kSimple) { do { t = __fadd_rn(t, p.dt0); } while (t < t_target); }
  else { do { t = __fadd_rn(t, calc_dt(t, p.dt)); } while (t < t_target); }
  return false;
}

// Block-wide exclusive scan of one int per thread (kMarchBlock threads).  Returns the exclusive
// prefix; *total receives the block sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[kMarchBlock / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int s = lane < kMarchBlock / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < kMarchBlock / 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += n;
    }
    if (lane < kMarchBlock / 32) warp_sums[lane] = s;
  }
  __syncthreads();
  const int warp_off = wid == 0 ? 0 : warp_sums[wid - 1];
  *total = warp_sums[kMarchBlock / 32 - 1];
  return warp_off + inc - v;
}

// Empty-ray culling for the single-cascade training march.  77 % of the rays of a Lego-shaped batch cross the box without
// meeting one occupied cell, and each of them pays ~200 trips of the bit-exact stepping loop to find that out.  A ray
// whose segment [t, t2] stays clear of every occupied cell by a margin can skip the loop: its count is 0 whatever the
// roundings of the stepping sequence are, so parity is untouched.
//   coarse bit (bx,by,bz) of a 32^3 lattice = "some occupied fine cell in the 4^3 block or in one of its 26 neighbours"
//   (a 4^3 block is 64 consecutive Morton codes = one aligned 8-byte word of the bitfield);
//   a ray is checked at points at most 4 cells apart (max norm): every point of the segment is within 2 cells of a
//   checked point, so an occupied cell that the march could ever look up lies inside the dilated block (+-4 cells) of a
//   checked point, with 2 cells to spare for the roundings.  ~30 short trips instead of ~200 long ones; the 87 % of the
//   empty rays that stay clear of the object's dilated hull leave after them (measured on the Lego-shaped scene).
__global__ void __launch_bounds__(256) coarse_occupancy_kernel(const uint8_t* __restrict__ bitfield, uint32_t* __restrict__ coarse) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;          // 32^3 threads; lane = bx
  const int bx = i & 31, by = (i >> 5) & 31, bz = i >> 10;
  bool any = false;
  for (int dz = -1; dz <= 1; dz++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        const int x = bx + dx, y = by + dy, z = bz + dz;
        if (x < 0 || y < 0 || z < 0 || x > 31 || y > 31 || z > 31) continue;
        any |= __ldg(reinterpret_cast<const unsigned long long*>(bitfield) + morton3D((uint32_t)x, (uint32_t)y, (uint32_t)z)) != 0ull;
      }
  const unsigned w = __ballot_sync(0xffffffffu, any);
  if ((i & 31u) == 0) coarse[i >> 5] = w;
}

// true when the segment [t, t2] of a ray provably stays clear of every occupied cell (coarse lattice above): the stepping
// loop would then find no sample whatever its roundings.  false = "cannot tell, march it" (also for degenerate directions).
__device__ __forceinline__ bool segment_is_clear(float ox, float oy, float oz, float dx, float dy, float dz, float t, float t2,
                                                 const MarchParams& p) {
  const float step = p.cull_span / fmaxf(fabsf(dx), fmaxf(fabsf(dy), fabsf(dz)));   // 4 cells along the fastest axis
  if (!(step > 0.f) || !(step < 1e30f)) return false;                                // degenerate direction: just march it
  int guard = 0;
  for (float tc = t;; tc += step) {
    if (++guard > 1024) return false;                                                // step below the resolution of t: just march it
    const float tt = fminf(tc, t2);
    const float cx = __fmaf_rn(tt, dx, ox), cy = __fmaf_rn(tt, dy, oy), cz = __fmaf_rn(tt, dz, oz);
    const int bx = (int)fmaxf(0.0f, fminf(16.0f * (cx * p.mb0_inv + 1.0f), 31.0f));
    const int by = (int)fmaxf(0.0f, fminf(16.0f * (cy * p.mb0_inv + 1.0f), 31.0f));
    const int bz = (int)fmaxf(0.0f, fminf(16.0f * (cz * p.mb0_inv + 1.0f), 31.0f));
    if ((__ldg(p.coarse + by + 32 * bz) >> bx) & 1u) return false;                   // near something occupied
    if (tc >= t2) return true;                                                       // clear all the way
  }
}

// Pass 0 (single cascade with a coarse lattice): thread per ray, ~30 short trips.  Rays that cannot produce a sample get
// n_samples = 0 here; the others are appended to `live`, the queue the persistent pass 1 draws from (the order of the
// queue is arbitrary, every output is indexed by the ray).  Doing the check inside pass 1 instead made its warps run
// the check path and the marching path back to back (0.494 -> 0.454 ms only).
__global__ void __launch_bounds__(kMarchBlock) march_cull_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ hits_t,
    const float* __restrict__ noise, MarchParams p, int64_t n_rays, int32_t* __restrict__ n_samples,
    int32_t* __restrict__ live, int* __restrict__ n_live) {
  const int64_t r = (int64_t)blockIdx.x * kMarchBlock + threadIdx.x;
  bool keep = false;
  if (r < n_rays) {
    const float dx = __ldg(rays_d + 3 * r), dy = __ldg(rays_d + 3 * r + 1), dz = __ldg(rays_d + 3 * r + 2);
    float t = __ldg(hits_t + 2 * r);
    const float t2 = __ldg(hits_t + 2 * r + 1);
    if (t >= 0) t = __fmaf_rn(p.dt0, __ldg(noise + r), t);            // the march's own start (raymarching.cu:195-198)
    if (0 <= t && t < t2) {
      const float ox = __ldg(rays_o + 3 * r), oy = __ldg(rays_o + 3 * r + 1), oz = __ldg(rays_o + 3 * r + 2);
      keep = !segment_is_clear(ox, oy, oz, dx, dy, dz, t, t2, p);
    }
    if (!keep) n_samples[r] = 0;
  }
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if (m) {
    const unsigned lane = threadIdx.x & 31u;
    int base = 0;
    const int leader = __ffs(m) - 1;
    if ((int)lane == leader) base = atomicAdd(n_live, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (keep) live[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)r;
  }
}

// Morton-ordered bitfield (the reference's layout, raymarching.cu:35-62) -> x-major words for the single-cascade marcher:
// thread per 32 cells of one x-row.
__global__ void __launch_bounds__(256) linearize_bitfield_kernel(const uint8_t* __restrict__ bitfield, uint32_t* __restrict__ linear) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;          // 128^3 / 32 words
  const uint32_t x0 = (w & 3u) * 32u, y = (w >> 2) & 127u, z = w >> 9;
  const uint32_t myz = morton3D(0u, y, z);
  uint32_t bits = 0;
#pragma unroll 8
  for (uint32_t i = 0; i < 32; i++) {
    const uint32_t idx = morton3D(x0 + i, 0u, 0u) | myz;
    bits |= ((uint32_t)(__ldg(bitfield + (idx >> 3)) >> (idx & 7u)) & 1u) << i;
  }
  linear[w] = bits;
}

// Pass 1, persistent: every lane owns one ray at a time and pulls the next ray index from a global
// counter the moment its ray is finished (warp-aggregated atomicAdd), so the 32 lanes of a warp stay busy
// although ray lengths differ by two orders of magnitude (first ncu capture: 12 of 32 lanes active).
// One loop trip = one marching step of every lane's current ray.
template <bool kSimple, bool kLinear = false>
__global__ void __launch_bounds__(kMarchBlock) march_count_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ hits_t,
    const float* __restrict__ noise, MarchParams p, int max_samples, int64_t n_rays,
    int32_t* __restrict__ n_samples, float2* __restrict__ scratch, int row_len, int* __restrict__ next_ray,
    const int32_t* __restrict__ live = nullptr, const int* __restrict__ n_live = nullptr) {
  const unsigned lane = threadIdx.x & 31u;
  const int64_t n_fetch = live ? (int64_t)__ldg(n_live) : n_rays;      // culled rays never enter the queue (march_cull_kernel)
  bool have = false, exhausted = false;     // exhausted: the counter has run past n_rays, stop asking
  int64_t r = 0;
  Ray q;
  float t = 0.f, t2 = 0.f, x, y, z, dt;
  int N = 0;
  float2* row = nullptr;
  while (true) {
    const unsigned need = __ballot_sync(0xffffffffu, !have && !exhausted);
    if (need) {
      int base = 0;
      const int leader = __ffs(need) - 1;
      if ((int)lane == leader) base = atomicAdd(next_ray, __popc(need));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (!have && !exhausted) {
        r = (int64_t)base + __popc(need & ((1u << lane) - 1u));
        if (r >= n_fetch) exhausted = true;
        else {
          if (live) r = __ldg(live + r);
          have = true;
          q = load_ray(rays_o, rays_d, r);
          float t1 = __ldg(hits_t + 2 * r);
          t2 = __ldg(hits_t + 2 * r + 1);
          if (t1 >= 0) t1 = __fmaf_rn(kSimple ? p.dt0 : calc_dt(t1, p.dt), __ldg(noise + r), t1);  // raymarching.cu:195-198
          t = t1; N = 0; row = scratch + r * (int64_t)row_len;
        }
      }
    }
    if (!__any_sync(0xffffffffu, have)) break;
    if (have) {
      if (0 <= t && t < t2 && N < max_samples) {
        if (march_step<kSimple, kLinear>(q, p, t, x, y, z, dt)) {
          if (N < row_len) row[N] = make_float2(t, dt);
          t = __fadd_rn(t, dt); N++;
        }
      } else {
        n_samples[r] = N;
        have = false;
      }
    }
  }
}

// per-256-ray sums of n_samples (the persistent pass 1 does not visit rays in block order)
__global__ void __launch_bounds__(kMarchBlock) block_sums_kernel(const int32_t* __restrict__ n_samples, int64_t n_rays,
                                                                 int32_t* __restrict__ block_sums) {
  const int64_t r = (int64_t)blockIdx.x * kMarchBlock + threadIdx.x;
  int total;
  block_exclusive_scan(r < n_rays ? n_samples[r] : 0, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// One CTA: exclusive scan of n_blocks block sums (int64 to survive R*max_samples > 2^31).
// counter[0] = total samples, counter[1] = n_rays (raymarching.cu:237-238 leaves the same pair).
__global__ void __launch_bounds__(1024) block_scan_kernel(const int32_t* __restrict__ block_sums, int n_blocks,
                                                          int64_t* __restrict__ block_offsets,
                                                          int32_t* __restrict__ counter, int64_t n_rays,
                                                          int64_t* __restrict__ total_out) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int64_t v = i < n_blocks ? (int64_t)block_sums[i] : 0;
    int64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int64_t s = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t n = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += n;
      }
      warp_tot[lane] = s;
    }
    __syncthreads();
    const int64_t carry = carry_s;
    const int64_t off = carry + (wid == 0 ? 0 : warp_tot[wid - 1]) + inc - v;
    if (i < n_blocks) block_offsets[i] = off;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int64_t tot = carry_s;
    if (counter) { counter[0] = (int32_t)tot; counter[1] = (int32_t)n_rays; }
    if (total_out) *total_out = tot;
  }
}

// Pass 2.  The CTA covers the same ray range [blockIdx.x*256, +256) as pass 1 so that the block sums / offsets line
// up, but its threads walk the block's SAMPLES, not its rays: thread -> output slot -> owning ray by a binary search
// over the 256 block-local start offsets.  Most rays of a batch are empty and the others carry ~130 samples (Lego-shaped
// scene: 77 % / 127), so lanes-per-ray layouts idle 3 lanes out of 4; this one is balanced and every store is a
// contiguous run (8 lanes per ray: 0.29 ms; flat: see profiles).
__global__ void __launch_bounds__(kMarchBlock) march_emit_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ hits_t,
    MarchParams p, int64_t n_rays, const int32_t* __restrict__ n_samples,
    const int64_t* __restrict__ block_offsets, const float2* __restrict__ scratch, int64_t capacity,
    int64_t* __restrict__ rays_a, float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
    float* __restrict__ ts, const int64_t* __restrict__ slot_to_ray, int row_len) {
  // slot_to_ray (optional): position i of n_samples / scratch / rays_a belongs to ray slot_to_ray[i]
  // (test-time wavefront: slots are the currently alive rays); NULL = identity (training).
  __shared__ int s_start[kMarchBlock + 1];       // block-local exclusive prefix of the sample counts
  const int64_t r0 = (int64_t)blockIdx.x * kMarchBlock;
  const int64_t boff = block_offsets[blockIdx.x];
  int total;
  {
    const int64_t r = r0 + threadIdx.x;
    const int N = r < n_rays ? n_samples[r] : 0;
    const int lstart = block_exclusive_scan(N, &total);
    s_start[threadIdx.x] = lstart;
    if (threadIdx.x == 0) s_start[kMarchBlock] = total;
    if (r < n_rays) { rays_a[3 * r] = slot_to_ray ? (N > 0 ? slot_to_ray[r] : 0) : r; rays_a[3 * r + 1] = boff + lstart; rays_a[3 * r + 2] = N; }
  }
  __syncthreads();
  for (int lo = threadIdx.x; lo < total; lo += kMarchBlock) {
    // last ray whose start is <= lo (empty rays share their start with the next non-empty one, which wins)
    int a = 0, b = kMarchBlock;                     // invariant: s_start[a] <= lo < s_start[b]
#pragma unroll
    for (int it = 0; it < 8; it++) { const int m = (a + b) >> 1; if (s_start[m] <= lo) a = m; else b = m; }
    const int s = lo - s_start[a];
    const int64_t o = boff + lo;
    if (s >= row_len || o >= capacity) continue;    // beyond the scratch row (max_samples > row length only): re-marched below
    const int64_t r = r0 + a;
    const int64_t ray = slot_to_ray ? slot_to_ray[r] : r;
    const float ox = __ldg(rays_o + 3 * ray), oy = __ldg(rays_o + 3 * ray + 1), oz = __ldg(rays_o + 3 * ray + 2);
    const float dx = __ldg(rays_d + 3 * ray), dy = __ldg(rays_d + 3 * ray + 1), dz = __ldg(rays_d + 3 * ray + 2);
    const float2 td = scratch[r * (int64_t)row_len + s];
    xyzs[3 * o] = __fmaf_rn(td.x, dx, ox); xyzs[3 * o + 1] = __fmaf_rn(td.x, dy, oy); xyzs[3 * o + 2] = __fmaf_rn(td.x, dz, oz);
    dirs[3 * o] = dx; dirs[3 * o + 1] = dy; dirs[3 * o + 2] = dz;
    ts[o] = td.x; deltas[o] = td.y;
  }
  {
    // long ray (more samples than the scratch row holds): its owner resumes the serial march right after the last
    // recorded sample
    const int64_t r = r0 + threadIdx.x;
    const int N = r < n_rays ? n_samples[r] : 0;
    if (N > row_len) {
      const int64_t ray = slot_to_ray ? slot_to_ray[r] : r;
      const int64_t start = boff + s_start[threadIdx.x];
      const Ray q = load_ray(rays_o, rays_d, ray);
      const float t2 = __ldg(hits_t + 2 * ray + 1);
      const float2 last = scratch[r * (int64_t)row_len + row_len - 1];
      float t = __fadd_rn(last.x, last.y), x, y, z, dt;
      int s = row_len;
      while (t < t2 && s < N) {
        if (march_step(q, p, t, x, y, z, dt)) {
          const int64_t o = start + s;
          if (o < capacity) {
            xyzs[3 * o] = x; xyzs[3 * o + 1] = y; xyzs[3 * o + 2] = z;
            dirs[3 * o] = q.dx; dirs[3 * o + 1] = q.dy; dirs[3 * o + 2] = q.dz;
            ts[o] = t; deltas[o] = dt;
          }
          t = __fadd_rn(t, dt); s++;
        }
      }
    }
  }
}

// raymarching.cu:335-404.  Thread per alive ray; dense (N_alive, N_samples, .) outputs; unused
// slots are written as zeros by the owning thread (the reference pre-zeroes the tensors on the
// host and models/rendering.py:91 relies on dirs==0 to find them).
__global__ void __launch_bounds__(kMarchBlock) march_test_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, float* __restrict__ hits_t,
    const int64_t* __restrict__ alive, MarchParams p, int n_samples_max, int64_t n_alive,
    float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas, float* __restrict__ ts,
    int32_t* __restrict__ n_eff, const int32_t* __restrict__ live = nullptr, const int* __restrict__ n_live = nullptr) {
  int64_t n = (int64_t)blockIdx.x * kMarchBlock + threadIdx.x;
  if (live) {                                   // slots that survived march_test_cull_kernel, compacted: full warps of real work
    if (n >= (int64_t)__ldg(n_live)) return;
    n = __ldg(live + n);
  } else if (n >= n_alive) return;
  const int64_t r = alive[n];
  const Ray q = load_ray(rays_o, rays_d, r);
  float t = hits_t[2 * r];
  const float t2 = hits_t[2 * r + 1];
  float x, y, z, dt, t_mark = t;
  int s = 0;
  const int64_t base = n * n_samples_max;
  while (t < t2 && s < n_samples_max) {
    if (march_step(q, p, t, x, y, z, dt)) {
      const int64_t o = base + s;
      xyzs[3 * o] = x; xyzs[3 * o + 1] = y; xyzs[3 * o + 2] = z;
      dirs[3 * o] = q.dx; dirs[3 * o + 1] = q.dy; dirs[3 * o + 2] = q.dz;
      ts[o] = t; deltas[o] = dt;
      t = __fadd_rn(t, dt);
      t_mark = t;  // raymarching.cu:390: the resume point is the t right after the last ACCEPTED sample
      s++;
    }
  }
  if (s > 0) hits_t[2 * r] = t_mark;
  n_eff[n] = s;
  for (int k = s; k < n_samples_max; k++) {
    const int64_t o = base + k;
    xyzs[3 * o] = 0.f; xyzs[3 * o + 1] = 0.f; xyzs[3 * o + 2] = 0.f;
    dirs[3 * o] = 0.f; dirs[3 * o + 1] = 0.f; dirs[3 * o + 2] = 0.f;
    ts[o] = 0.f; deltas[o] = 0.f;
  }
}

// Pass 0 of the test-time march for single-cascade scenes (most rays of a frame miss the object): a slot whose remaining
// segment [t, t2] is provably empty (segment_is_clear) gets exactly what the stepping loop would give it — N_eff = 0, a zero
// row, an untouched hits_t — without the ~200 loop trips; the others are appended to `live`, which the march kernel then
// walks with full warps (doing the test inside the march kernel leaves the warps as divergent as before: measured, no gain).
__global__ void __launch_bounds__(kMarchBlock) march_test_cull_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ hits_t,
    const int64_t* __restrict__ alive, MarchParams p, int n_samples_max, int64_t n_alive,
    float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas, float* __restrict__ ts,
    int32_t* __restrict__ n_eff, int32_t* __restrict__ live, int* __restrict__ n_live) {
  const int64_t n = (int64_t)blockIdx.x * kMarchBlock + threadIdx.x;
  bool keep = false;
  if (n < n_alive) {
    const int64_t r = alive[n];
    const float t = __ldg(hits_t + 2 * r), t2 = __ldg(hits_t + 2 * r + 1);
    if (t < t2)
      keep = !segment_is_clear(__ldg(rays_o + 3 * r), __ldg(rays_o + 3 * r + 1), __ldg(rays_o + 3 * r + 2), __ldg(rays_d + 3 * r),
                               __ldg(rays_d + 3 * r + 1), __ldg(rays_d + 3 * r + 2), t, t2, p);
    if (!keep) {
      n_eff[n] = 0;
      for (int k = 0; k < n_samples_max; k++) {
        const int64_t o = n * n_samples_max + k;
        xyzs[3 * o] = 0.f; xyzs[3 * o + 1] = 0.f; xyzs[3 * o + 2] = 0.f;
        dirs[3 * o] = 0.f; dirs[3 * o + 1] = 0.f; dirs[3 * o + 2] = 0.f;
        ts[o] = 0.f; deltas[o] = 0.f;
      }
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if (m) {
    const unsigned lane = threadIdx.x & 31u;
    int base = 0;
    const int leader = __ffs(m) - 1;
    if ((int)lane == leader) base = atomicAdd(n_live, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (keep) live[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)n;
  }
}

// Round-0 culling of the wavefront renderer: the frame's rays whose whole segment is provably empty never enter the alive list.
__global__ void __launch_bounds__(kMarchBlock) render_cull_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ hits_t,
    const int64_t* __restrict__ alive_in, int64_t n_alive_in, MarchParams p, int32_t* __restrict__ live, int* __restrict__ n_live) {
  const int64_t i = (int64_t)blockIdx.x * kMarchBlock + threadIdx.x;
  bool keep = false;
  int64_t r = 0;
  if (i < n_alive_in) {
    r = alive_in ? alive_in[i] : i;
    const float t = __ldg(hits_t + 2 * r), t2 = __ldg(hits_t + 2 * r + 1);
    if (t < t2)
      keep = !segment_is_clear(__ldg(rays_o + 3 * r), __ldg(rays_o + 3 * r + 1), __ldg(rays_o + 3 * r + 2), __ldg(rays_d + 3 * r),
                               __ldg(rays_d + 3 * r + 1), __ldg(rays_d + 3 * r + 2), t, t2, p);
  }
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if (m) {
    const unsigned lane = threadIdx.x & 31u;
    int base = 0;
    const int leader = __ffs(m) - 1;
    if ((int)lane == leader) base = atomicAdd(n_live, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (keep) live[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)r;
  }
}

// ------------------------------------------------------------------------------------------------
// Test-time wavefront renderer (replaces the Python loop of models/rendering.py:46-133 + the per-round
// raymarching_test / composite_test_fw pair).  One launch per round, one thread per currently alive ray:
//   1. composite the samples this ray got in the PREVIOUS round (volumerendering.cu:335-373 semantics:
//      resume T = 1 - opacity, stop at T <= T_threshold),
//   2. if the ray is still alive and a next round exists, append it to the next alive list and march its
//      next <= n_next occupied samples (raymarching.cu:335-404 incl. the calc_dt(cascades) argument),
//      recording (t, dt) in the slot's scratch row; march_emit then packs them.
// Rays never wait for each other across rounds, the host reads back two counters per round, and the
// field is only evaluated on samples that exist.
// the normal / semantic streams of the reference's test-time compositor (volumerendering.cu:335-373): per-sample inputs of the
// previous round and the per-ray accumulators they are composited into; all NULL / 0 for fields without those heads
struct RenderAux {
  const float* normals_pred; const float* normals_raw; const float* sems;      // (S,3), (S,3), (S,C)
  float* normal; float* normal_raw; float* sem;                                // (R,3), (R,3), (R,C)
  int classes;
};
constexpr int kMaxClasses = 32;

__global__ void __launch_bounds__(kMarchBlock) render_advance_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, float* __restrict__ hits_t,
    const int64_t* __restrict__ alive_in, int64_t n_alive_in, const int64_t* __restrict__ prev_rays_a,
    const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
    const float* __restrict__ ts, float T_thr, MarchParams p, int n_next, float* __restrict__ opacity,
    float* __restrict__ depth, float* __restrict__ rgb, int64_t* __restrict__ alive_out, int32_t* __restrict__ counters,
    int32_t* __restrict__ n_samples, float2* __restrict__ scratch, const int32_t* __restrict__ live = nullptr,
    const int* __restrict__ n_live = nullptr, RenderAux aux = RenderAux{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0}) {
  const int64_t i = (int64_t)blockIdx.x * kMarchBlock + threadIdx.x;
  // n_live: the live slot count in device memory (n_alive_in is then only the bound the launch is sized for); live: round 0
  // after render_cull_kernel (ray ids, compacted)
  bool alive = i < (n_live ? min((int64_t)__ldg(n_live), n_alive_in) : n_alive_in);
  int64_t r = 0;
  if (alive) {
    r = live ? (int64_t)__ldg(live + i) : (alive_in ? alive_in[i] : i);
    if (prev_rays_a) {
      const int64_t start = prev_rays_a[3 * i + 1];
      const int N = (int)prev_rays_a[3 * i + 2];
      if (N == 0) alive = false;                       // marched to the end of the box without a sample
      else {
        float T = 1.0f - opacity[r];
        float aO = 0.f, aD = 0.f, aR = 0.f, aG = 0.f, aB = 0.f;
        if (aux.normal == nullptr) {
          for (int s = 0; s < N; s++) {
            const int64_t o = start + s;
            const float a = 1.0f - __expf(-__ldg(sigmas + o) * __ldg(deltas + o));
            const float w = a * T;
            aR = fmaf(w, __ldg(rgbs + 3 * o), aR); aG = fmaf(w, __ldg(rgbs + 3 * o + 1), aG); aB = fmaf(w, __ldg(rgbs + 3 * o + 2), aB);
            aD = fmaf(w, __ldg(ts + o), aD);
            aO += w;
            T *= 1.0f - a;
            if (T <= T_thr) { alive = false; break; }
          }
        } else {                        // + normal_pred, normal_raw and the C semantic channels (volumerendering.cu:352-366)
          float nP[3] = {0.f, 0.f, 0.f}, nR[3] = {0.f, 0.f, 0.f}, aS[kMaxClasses];
          const int C = aux.classes;
          for (int c = 0; c < C; c++) aS[c] = 0.f;
          for (int s = 0; s < N; s++) {
            const int64_t o = start + s;
            const float a = 1.0f - __expf(-__ldg(sigmas + o) * __ldg(deltas + o));
            const float w = a * T;
            aR = fmaf(w, __ldg(rgbs + 3 * o), aR); aG = fmaf(w, __ldg(rgbs + 3 * o + 1), aG); aB = fmaf(w, __ldg(rgbs + 3 * o + 2), aB);
            aD = fmaf(w, __ldg(ts + o), aD);
#pragma unroll
            for (int k = 0; k < 3; k++) {
              nP[k] = fmaf(w, __ldg(aux.normals_pred + 3 * o + k), nP[k]);
              nR[k] = fmaf(w, __ldg(aux.normals_raw + 3 * o + k), nR[k]);
            }
            for (int c = 0; c < C; c++) aS[c] = fmaf(w, __ldg(aux.sems + o * C + c), aS[c]);
            aO += w;
            T *= 1.0f - a;
            if (T <= T_thr) { alive = false; break; }
          }
#pragma unroll
          for (int k = 0; k < 3; k++) { aux.normal[3 * r + k] += nP[k]; aux.normal_raw[3 * r + k] += nR[k]; }
          for (int c = 0; c < C; c++) aux.sem[r * C + c] += aS[c];
        }
        opacity[r] += aO; depth[r] += aD;
        rgb[3 * r] += aR; rgb[3 * r + 1] += aG; rgb[3 * r + 2] += aB;
      }
    }
  }
  if (n_next <= 0) return;
  // compact the survivors (warp-aggregated append; order is irrelevant, every ray owns its slot)
  const unsigned m = __ballot_sync(0xffffffffu, alive);
  int base = 0;
  const unsigned lane = threadIdx.x & 31u;
  if (m && (int)lane == __ffs(m) - 1) base = atomicAdd(counters, __popc(m));
  base = __shfl_sync(0xffffffffu, base, m ? __ffs(m) - 1 : 0);
  if (!alive) return;
  const int64_t j = (int64_t)base + __popc(m & ((1u << lane) - 1u));
  alive_out[j] = r;
  const Ray q = load_ray(rays_o, rays_d, r);
  float t = hits_t[2 * r], x, y, z, dt, t_mark = t;
  const float t2 = hits_t[2 * r + 1];
  float2* row = scratch + j * kScratch;
  int s = 0;
  while (t < t2 && s < n_next) {
    if (march_step(q, p, t, x, y, z, dt)) {
      row[s] = make_float2(t, dt);
      t = __fadd_rn(t, dt); t_mark = t; s++;
    }
  }
  if (s > 0) hits_t[2 * r] = t_mark;
  n_samples[j] = s;
}

static MarchParams make_params(const uint8_t* bitfield, int cascades, float scale, float dt_scale, float esf,
                               int grid_size, int max_samples) {
  MarchParams p;
  p.bitfield = bitfield;
  p.cascades = cascades;
  p.grid_size = grid_size;
  p.scale = scale;
  p.dt = make_dt_params(esf, max_samples, grid_size, dt_scale);
  p.grid_f = (float)grid_size;
  p.grid_inv = 1.0f / (float)grid_size;
  p.grid_m1 = (float)grid_size - 1.0f;
  p.grid3 = (uint32_t)grid_size * grid_size * grid_size;
  p.dt0 = fmaxf(p.dt.dt_min, fminf(0.0f, p.dt.dt_max));
  p.mb0 = fminf(0.5f, scale);
  p.mb0_inv = 1.0f / p.mb0;
  p.coarse = nullptr;
  p.linear = nullptr;
  p.cull_span = 8.0f * p.mb0 / (float)grid_size;      // 4 cells of 2*mip_bound/G
  return p;
}

static bool cull_enabled() {
  static const bool on = !(getenv("NGP_MARCH_CULL") && atoi(getenv("NGP_MARCH_CULL")) == 0);
  return on;
}

}  // namespace ngp

using namespace ngp;

// Workspace layout for the training marcher (caller-provided device memory):
//   int32 n_samples[R] | float t_start[R] | int32 block_sums[B] | int64 block_offsets[B] | int64 total |
//   float2 scratch[R][row]      row = kTrainRow (training) / kScratch (test-time wavefront)
constexpr int64_t kLinearBytes = 128 * 128 * 128 / 8;
static int64_t march_ws_bytes(int64_t n_rays, int row) {
  const int64_t B = ceil_div(n_rays, kMarchBlock);
  auto al = [](int64_t x) { return (x + 255) / 256 * 256; };
  return al(n_rays * 4) + al(n_rays * 4) + al(B * 4) + al(B * 8) + 256 + al(n_rays * row * 8) + 4096 + kLinearBytes;   // + the 32^3-bit coarse lattice, the x-major bitfield
}
NGP_API int64_t ngp_raymarching_train_workspace_bytes(int64_t n_rays) { return march_ws_bytes(n_rays, kTrainRow); }
// workspace of ngp_render_advance / ngp_render_emit for n_alive_in slots
NGP_API int64_t ngp_render_workspace_bytes(int64_t n_slots) { return march_ws_bytes(n_slots, kScratch); }

struct MarchWs { int32_t* n_samples; float* t_start; int32_t* block_sums; int64_t* block_offsets; int64_t* total; float2* scratch; };
static uint32_t* coarse_of(const MarchWs& w, int64_t n_rays, int row) {
  return reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(w.scratch) + (n_rays * row * 8 + 255) / 256 * 256);
}
// the x-major copy of the bitfield, behind the coarse lattice
static uint32_t* linear_of(const MarchWs& w, int64_t n_rays, int row) {
  return reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(coarse_of(w, n_rays, row)) + 4096);
}
static MarchWs carve(void* ws, int64_t n_rays) {
  const int64_t B = ceil_div(n_rays, kMarchBlock);
  auto al = [](int64_t x) { return (x + 255) / 256 * 256; };
  char* p = (char*)ws;
  MarchWs w;
  w.n_samples = (int32_t*)p; p += al(n_rays * 4);
  w.t_start = (float*)p; p += al(n_rays * 4);
  w.block_sums = (int32_t*)p; p += al(B * 4);
  w.block_offsets = (int64_t*)p; p += al(B * 8);
  w.total = (int64_t*)p; p += 256;
  w.scratch = (float2*)p;
  return w;
}

// Phase 1+2 of vren.raymarching_train (binding.cpp:60-81 -> raymarching.cu:283-332): count the
// samples of every ray and scan.  After this call `counter` = [total_samples, n_rays] on the device
// (and ws.total as int64); the caller may read it to size the outputs exactly, or skip the
// read-back and pass a capacity bound to ngp_raymarching_train_write.
NGP_API int ngp_raymarching_train_count(const float* rays_o, const float* rays_d, const float* hits_t,
                                        const uint8_t* density_bitfield, int cascades, float scale,
                                        float exp_step_factor, const float* noise, int grid_size, int max_samples,
                                        int64_t n_rays, int32_t* counter, void* workspace, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_rays <= 0) { if (counter) cudaMemsetAsync(counter, 0, 8, s); return 0; }
  const MarchWs w = carve(workspace, n_rays);
  const MarchParams p = make_params(density_bitfield, cascades, scale, scale, exp_step_factor, grid_size, max_samples);
  const int B = (int)ceil_div(n_rays, kMarchBlock);
  int* next_ray = reinterpret_cast<int*>(w.total) + 4;        // w.total: [int64 total | . | next_ray | n_live | ...] (256 bytes)
  int* n_live = next_ray + 1;
  cudaMemsetAsync(next_ray, 0, 2 * sizeof(int), s);
  const bool simple = cascades == 1 && exp_step_factor == 0.0f;
  static const int ctas_per_sm = getenv("NGP_MARCH_CTAS_PER_SM") ? atoi(getenv("NGP_MARCH_CTAS_PER_SM")) : 4;   // swept 1..6 on B200 (r01 call 19): 4 is the minimum
  const int64_t gmax = (int64_t)kSMs * (ctas_per_sm < 1 ? 1 : ctas_per_sm);
  const int G = (int)(ceil_div(n_rays, kMarchBlock) < gmax ? ceil_div(n_rays, kMarchBlock) : gmax);
  const int row_len = max_samples < kTrainRow ? (max_samples < 1 ? 1 : max_samples) : kTrainRow;
  const bool cull_on = cull_enabled();
  if (simple && cull_on && grid_size == 128 && n_rays >= 2048 && ((uintptr_t)density_bitfield & 7u) == 0) {
    static const bool linear_on = !(getenv("NGP_MARCH_LINEAR") && atoi(getenv("NGP_MARCH_LINEAR")) == 0);   // 0: Morton lookups (A/B only)
    MarchParams pc = p, pm = p;
    uint32_t* coarse = coarse_of(w, n_rays, kTrainRow);
    int32_t* live = reinterpret_cast<int32_t*>(w.t_start);      // n_rays ints: the (otherwise unused) t_start area
    coarse_occupancy_kernel<<<128, 256, 0, s>>>(density_bitfield, coarse);
    NGP_LAUNCH_CHECK("ngp_raymarching_train_count/coarse");
    pc.coarse = coarse;
    if (linear_on) {
      uint32_t* lin = linear_of(w, n_rays, kTrainRow);
      linearize_bitfield_kernel<<<128 * 128 * 128 / 32 / 256, 256, 0, s>>>(density_bitfield, lin);
      NGP_LAUNCH_CHECK("ngp_raymarching_train_count/linearize");
      pm.linear = lin;
    }
    march_cull_kernel<<<B, kMarchBlock, 0, s>>>(rays_o, rays_d, hits_t, noise, pc, n_rays, w.n_samples, live, n_live);
    NGP_LAUNCH_CHECK("ngp_raymarching_train_count/cull");
    if (linear_on) march_count_kernel<true, true><<<G, kMarchBlock, 0, s>>>(rays_o, rays_d, hits_t, noise, pm, max_samples, n_rays, w.n_samples, w.scratch, row_len, next_ray, live, n_live);
    else march_count_kernel<true><<<G, kMarchBlock, 0, s>>>(rays_o, rays_d, hits_t, noise, pm, max_samples, n_rays, w.n_samples, w.scratch, row_len, next_ray, live, n_live);
  } else if (simple) march_count_kernel<true><<<G, kMarchBlock, 0, s>>>(rays_o, rays_d, hits_t, noise, p, max_samples, n_rays, w.n_samples, w.scratch, row_len, next_ray);
  else march_count_kernel<false><<<G, kMarchBlock, 0, s>>>(rays_o, rays_d, hits_t, noise, p, max_samples, n_rays, w.n_samples, w.scratch, row_len, next_ray);
  NGP_LAUNCH_CHECK("ngp_raymarching_train_count/count");
  block_sums_kernel<<<B, kMarchBlock, 0, s>>>(w.n_samples, n_rays, w.block_sums);
  NGP_LAUNCH_CHECK("ngp_raymarching_train_count/sums");
  block_scan_kernel<<<1, 1024, 0, s>>>(w.block_sums, B, w.block_offsets, counter, n_rays, w.total);
  NGP_LAUNCH_CHECK("ngp_raymarching_train_count/scan");
  return 0;
}

// Phase 3: write rays_a (R,3) i64 = [ray_idx, start_idx, N_samples] in ray-index order and the
// packed samples.  Rows >= capacity are dropped (never happens when capacity >= counter[0]).
NGP_API int ngp_raymarching_train_write(const float* rays_o, const float* rays_d, const float* hits_t,
                                        const uint8_t* density_bitfield, int cascades, float scale,
                                        float exp_step_factor, int grid_size, int max_samples, int64_t n_rays,
                                        const void* workspace, int64_t capacity, int64_t* rays_a, float* xyzs,
                                        float* dirs, float* deltas, float* ts, void* stream) {
  if (n_rays <= 0) return 0;
  const MarchWs w = carve((void*)workspace, n_rays);
  const MarchParams p = make_params(density_bitfield, cascades, scale, scale, exp_step_factor, grid_size, max_samples);
  const int B = (int)ceil_div(n_rays, kMarchBlock);
  const int row_len = max_samples < kTrainRow ? (max_samples < 1 ? 1 : max_samples) : kTrainRow;
  march_emit_kernel<<<B, kMarchBlock, 0, (cudaStream_t)stream>>>(rays_o, rays_d, hits_t, p, n_rays, w.n_samples,
                                                                 w.block_offsets, w.scratch, capacity,
                                                                 rays_a, xyzs, dirs, deltas, ts, nullptr, row_len);
  NGP_LAUNCH_CHECK("ngp_raymarching_train_write");
  return 0;
}

// Replaces vren.raymarching_test (binding.cpp:84-106 -> raymarching.cu:407-454).  hits_t (R,2) is
// advanced in place.  Reproduces the reference's calc_dt(..., cascades) argument (raymarching.cu:370,399).
NGP_API int ngp_raymarching_test(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_indices,
                                 const uint8_t* density_bitfield, int cascades, float scale, float exp_step_factor,
                                 int grid_size, int max_samples, int n_samples, int64_t n_alive, float* xyzs,
                                 float* dirs, float* deltas, float* ts, int32_t* n_eff_samples, void* stream) {
  if (n_alive <= 0) return 0;
  MarchParams p = make_params(density_bitfield, cascades, scale, (float)cascades, exp_step_factor, grid_size, max_samples);
  const int B = (int)ceil_div(n_alive, kMarchBlock);
  cudaStream_t s = (cudaStream_t)stream;
  // empty-ray culling (single cascade, no exponential stepping: the synthetic-scene case, where most rays of a frame miss the
  // object): the 4 KB coarse lattice lives in a stream-ordered allocation, released right after the launch
  char* tmp = nullptr;             // [4 KB coarse lattice | 16 B counter | n_alive x int32 live list], stream-ordered, released after the launches
  if (cull_enabled() && cascades == 1 && exp_step_factor == 0.0f && grid_size == 128 && n_alive >= 2048 && ((uintptr_t)density_bitfield & 7u) == 0 &&
      cudaMallocAsync((void**)&tmp, 4096 + 16 + (size_t)n_alive * 4, s) == cudaSuccess) {
    uint32_t* coarse = reinterpret_cast<uint32_t*>(tmp);
    int* n_live = reinterpret_cast<int*>(tmp + 4096);
    int32_t* live = reinterpret_cast<int32_t*>(tmp + 4096 + 16);
    coarse_occupancy_kernel<<<128, 256, 0, s>>>(density_bitfield, coarse);
    cudaMemsetAsync(n_live, 0, sizeof(int), s);
    p.coarse = coarse;
    march_test_cull_kernel<<<B, kMarchBlock, 0, s>>>(rays_o, rays_d, hits_t, alive_indices, p, n_samples, n_alive, xyzs, dirs, deltas, ts,
                                                    n_eff_samples, live, n_live);
    NGP_LAUNCH_CHECK("ngp_raymarching_test/cull");
    march_test_kernel<<<B, kMarchBlock, 0, s>>>(rays_o, rays_d, hits_t, alive_indices, p, n_samples, n_alive, xyzs, dirs, deltas, ts,
                                               n_eff_samples, live, n_live);
    NGP_LAUNCH_CHECK("ngp_raymarching_test");
    cudaFreeAsync(tmp, s);
    return 0;
  }
  cudaGetLastError();
  march_test_kernel<<<B, kMarchBlock, 0, s>>>(rays_o, rays_d, hits_t, alive_indices, p, n_samples,
                                             n_alive, xyzs, dirs, deltas, ts, n_eff_samples);
  NGP_LAUNCH_CHECK("ngp_raymarching_test");
  return 0;
}

// ---- test-time wavefront renderer ------------------------------------------------------------------
// Round k of the renderer that replaces models/rendering.py:46-133 (volume_render).  Composites the previous
// round's packed samples of every alive ray (prev_rays_a (n_alive_in,3) = [ray, start, N] per slot, NULL in
// round 0), appends survivors to alive_out (counters[0] = their number) and marches their next <= n_next
// samples into the workspace; ngp_render_emit then packs them.  n_next = 0 -> composite only (last round).
// n_next must be <= 256.  workspace: ngp_render_workspace_bytes(n_alive_in).
static int render_advance_launch(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_in,
                               int64_t n_alive_in, const int32_t* n_alive_dev, const int64_t* prev_rays_a, const float* sigmas, const float* rgbs,
                               const float* deltas, const float* ts, float T_threshold, const uint8_t* density_bitfield,
                               int cascades, float scale, float exp_step_factor, int grid_size, int max_samples,
                               int n_next, float* opacity, float* depth, float* rgb, int64_t* alive_out,
                               int32_t* counters, void* workspace, const float* normals_pred, const float* normals_raw,
                               const float* sems, int classes, float* normal, float* normal_raw, float* sem, void* stream);
NGP_API int ngp_render_advance_full(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_in,
                               int64_t n_alive_in, const int64_t* prev_rays_a, const float* sigmas, const float* rgbs,
                               const float* deltas, const float* ts, float T_threshold, const uint8_t* density_bitfield,
                               int cascades, float scale, float exp_step_factor, int grid_size, int max_samples,
                               int n_next, float* opacity, float* depth, float* rgb, int64_t* alive_out,
                               int32_t* counters, void* workspace, const float* normals_pred, const float* normals_raw,
                               const float* sems, int classes, float* normal, float* normal_raw, float* sem, void* stream) {
  return render_advance_launch(rays_o, rays_d, hits_t, alive_in, n_alive_in, nullptr, prev_rays_a, sigmas, rgbs, deltas, ts, T_threshold,
                               density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples, n_next, opacity, depth, rgb,
                               alive_out, counters, workspace, normals_pred, normals_raw, sems, classes, normal, normal_raw, sem, stream);
}
static int render_advance_launch(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_in,
                               int64_t n_alive_in, const int32_t* n_alive_dev, const int64_t* prev_rays_a, const float* sigmas, const float* rgbs,
                               const float* deltas, const float* ts, float T_threshold, const uint8_t* density_bitfield,
                               int cascades, float scale, float exp_step_factor, int grid_size, int max_samples,
                               int n_next, float* opacity, float* depth, float* rgb, int64_t* alive_out,
                               int32_t* counters, void* workspace, const float* normals_pred, const float* normals_raw,
                               const float* sems, int classes, float* normal, float* normal_raw, float* sem, void* stream) {
  if (n_alive_in <= 0) return 0;
  if (classes < 0 || classes > kMaxClasses) return set_error_msg("ngp_render_advance_full: classes must be in [0, 32]");
  if ((normal == nullptr) != (normal_raw == nullptr) || (classes > 0 && normal != nullptr && sem == nullptr))
    return set_error_msg("ngp_render_advance_full: pass normal, normal_raw (and sem when classes > 0) together or not at all");
  RenderAux aux{normals_pred, normals_raw, sems, normal, normal_raw, sem, normal ? classes : 0};
  if (prev_rays_a == nullptr || normals_pred == nullptr) aux.normal = nullptr;      // round 0: nothing to composite
  if (n_next > kScratch) return set_error_msg("ngp_render_advance: n_next must be <= 256");
  cudaStream_t s = (cudaStream_t)stream;
  const MarchWs w = carve(workspace, n_alive_in);
  MarchParams p = make_params(density_bitfield, cascades, scale, (float)cascades, exp_step_factor, grid_size, max_samples);
  const int32_t* live = nullptr;
  const int* n_live = nullptr;
  if (prev_rays_a == nullptr && n_next > 0 && cull_enabled() && cascades == 1 && exp_step_factor == 0.0f && grid_size == 128 && n_alive_in >= 2048 &&
      ((uintptr_t)density_bitfield & 7u) == 0) {       // round 0 sees every ray of the frame: the provably empty ones never enter the alive list
    uint32_t* coarse = coarse_of(w, n_alive_in, kScratch);
    int32_t* lv = reinterpret_cast<int32_t*>(w.t_start);
    int* nl = reinterpret_cast<int*>(w.total) + 5;
    coarse_occupancy_kernel<<<128, 256, 0, s>>>(density_bitfield, coarse);
    cudaMemsetAsync(nl, 0, sizeof(int), s);
    p.coarse = coarse;
    render_cull_kernel<<<(int)ceil_div(n_alive_in, kMarchBlock), kMarchBlock, 0, s>>>(rays_o, rays_d, hits_t, alive_in, n_alive_in, p, lv, nl);
    NGP_LAUNCH_CHECK("ngp_render_advance/cull");
    live = lv; n_live = nl;
  } else if (n_alive_dev) n_live = reinterpret_cast<const int*>(n_alive_dev);
  cudaMemsetAsync(counters, 0, 2 * sizeof(int32_t), s);
  if (n_next > 0) cudaMemsetAsync(w.n_samples, 0, n_alive_in * sizeof(int32_t), s);
  const int B = (int)ceil_div(n_alive_in, kMarchBlock);
  render_advance_kernel<<<B, kMarchBlock, 0, s>>>(rays_o, rays_d, hits_t, alive_in, n_alive_in, prev_rays_a, sigmas, rgbs,
                                                 deltas, ts, T_threshold, p, n_next, opacity, depth, rgb, alive_out, counters,
                                                 w.n_samples, w.scratch, live, n_live, aux);
  NGP_LAUNCH_CHECK("ngp_render_advance");
  return 0;
}
NGP_API int ngp_render_advance(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_in,
                               int64_t n_alive_in, const int64_t* prev_rays_a, const float* sigmas, const float* rgbs,
                               const float* deltas, const float* ts, float T_threshold, const uint8_t* density_bitfield,
                               int cascades, float scale, float exp_step_factor, int grid_size, int max_samples,
                               int n_next, float* opacity, float* depth, float* rgb, int64_t* alive_out,
                               int32_t* counters, void* workspace, void* stream) {
  return ngp_render_advance_full(rays_o, rays_d, hits_t, alive_in, n_alive_in, prev_rays_a, sigmas, rgbs, deltas, ts, T_threshold,
                                 density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples, n_next, opacity, depth, rgb,
                                 alive_out, counters, workspace, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, stream);
}

// Packs the samples recorded by ngp_render_advance: rays_a (n_slots,3) = [ray, start, N] per slot of alive_out,
// xyzs/dirs (S,3), deltas/ts (S); counters[1] = S.  n_slots = an upper bound of counters[0] (n_alive_in is fine:
// unused slots hold N = 0).  capacity >= n_slots * n_next is always enough.
NGP_API int ngp_render_emit(const float* rays_o, const float* rays_d, const float* hits_t, const int64_t* alive_out,
                            int64_t n_slots, const uint8_t* density_bitfield, int cascades, float scale,
                            float exp_step_factor, int grid_size, int max_samples, const void* workspace,
                            int64_t capacity, int64_t* rays_a, float* xyzs, float* dirs, float* deltas, float* ts,
                            int32_t* counters, void* stream) {
  if (n_slots <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const MarchWs w = carve((void*)workspace, n_slots);
  const MarchParams p = make_params(density_bitfield, cascades, scale, (float)cascades, exp_step_factor, grid_size, max_samples);
  const int B = (int)ceil_div(n_slots, kMarchBlock);
  block_sums_kernel<<<B, kMarchBlock, 0, s>>>(w.n_samples, n_slots, w.block_sums);
  NGP_LAUNCH_CHECK("ngp_render_emit/sums");
  block_scan_kernel<<<1, 1024, 0, s>>>(w.block_sums, B, w.block_offsets, nullptr, n_slots, w.total);
  NGP_LAUNCH_CHECK("ngp_render_emit/scan");
  march_emit_kernel<<<B, kMarchBlock, 0, s>>>(rays_o, rays_d, hits_t, p, n_slots, w.n_samples, w.block_offsets, w.scratch,
                                             capacity, rays_a, xyzs, dirs, deltas, ts, alive_out, kScratch);
  NGP_LAUNCH_CHECK("ngp_render_emit/emit");
  cudaMemcpyAsync(counters + 1, w.total, sizeof(int32_t), cudaMemcpyDeviceToDevice, s);   // low word of the int64 total
  return 0;
}

// ---- one whole round of the test-time renderer for the ngp_pl-shaped field, sized by BOUNDS --------------------------
// advance (composite the previous round, compact, march) -> emit (pack) -> field (hash grid -> bf16 tiles -> density net ->
// colour net) in one call, every launch taking its live element count from DEVICE memory: the alive count of the
// previous round (n_alive_dev, NULL in round 0) and this round's own counters.  The host therefore never has to wait for
// a count before it can enqueue the next round: it passes an upper bound of the alive slots (n_alive_bound, e.g. the
// count read back one round late) and buffers sized for n_alive_bound * n_next samples.  counters (2 x int32, one pair
// per round so that late read-backs stay valid) = [alive slots after this round, samples marched in this round].
// Replaces one trip of the loop of models/rendering.py:75-124 (raymarching_test, forward_test, composite_test_fw and the
// alive-index compaction) for fields of the NGPCompact shape (networks.py): density MLP 32 -> width -> 16 with
// sigma = exp(h0), colour MLP [SH4(d) | h] -> width -> width -> 3, sigmoid.
NGP_API int ngp_render_round_compact(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_in,
                                     int64_t n_alive_bound, const int32_t* n_alive_dev, const int64_t* prev_rays_a,
                                     const float* prev_sigmas, const float* prev_rgbs, const float* prev_deltas, const float* prev_ts,
                                     float T_threshold, const uint8_t* density_bitfield, int cascades, float scale,
                                     float exp_step_factor, int grid_size, int max_samples, int n_next, float* opacity,
                                     float* depth, float* rgb, int64_t* alive_out, int32_t* counters, void* workspace,
                                     int64_t* rays_a, float* xyzs, float* dirs, float* deltas, float* ts, const float* aabb,
                                     const void* table, int table_dtype, int n_levels, int n_features, int log2_hashmap_size,
                                     int base_resolution, float per_level_scale, const float* sigma_params, const float* rgb_params,
                                     int width, int rgb_hidden, void* feat_tiles, float* h, float* sigmas, float* rgbs, void* stream) {
  if (n_alive_bound <= 0) return 0;
  int rc = render_advance_launch(rays_o, rays_d, hits_t, alive_in, n_alive_bound, n_alive_dev, prev_rays_a, prev_sigmas, prev_rgbs,
                                 prev_deltas, prev_ts, T_threshold, density_bitfield, cascades, scale, exp_step_factor, grid_size,
                                 max_samples, n_next, opacity, depth, rgb, alive_out, counters, workspace, nullptr, nullptr, nullptr, 0,
                                 nullptr, nullptr, nullptr, stream);
  if (rc || n_next <= 0) return rc;
  const int64_t cap = n_alive_bound * n_next;
  rc = ngp_render_emit(rays_o, rays_d, hits_t, alive_out, n_alive_bound, density_bitfield, cascades, scale, exp_step_factor, grid_size,
                       max_samples, workspace, cap, rays_a, xyzs, dirs, deltas, ts, counters, stream);
  if (rc) return rc;
  const int32_t* n_pts = counters + 1;
  const int LF = n_levels * n_features;
  rc = hashgrid_fw_tiles_launch(xyzs, aabb, table, table_dtype, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale,
                                cap, n_pts, feat_tiles, stream);
  if (rc) return rc;
  {   // density net on the feature tiles; sigma = exp(h0) from the same epilogue
    const float* sp[1] = {reinterpret_cast<const float*>(feat_tiles)};
    const int sw[1] = {LF}, sk[1] = {2};
    const int64_t ss[1] = {0};
    rc = mlp_fw_launch(1, sp, sw, sk, ss, sigma_params, width, 1, 16, /*ReLU*/ 1, /*None*/ 0, cap, n_pts, h, 16, sigmas, stream);
    if (rc) return rc;
  }
  {   // colour net on [SH4(normalised d) | h]
    const float* sp[2] = {dirs, h};
    const int sw[2] = {16, 16}, sk[2] = {1, 0};
    const int64_t ss[2] = {3, 16};
    rc = mlp_fw_launch(2, sp, sw, sk, ss, rgb_params, width, rgb_hidden, 3, /*ReLU*/ 1, /*Sigmoid*/ 2, cap, n_pts, rgbs, 3, nullptr, stream);
  }
  return rc;
}

// The field part of a round on its own: hash grid -> bf16 tiles -> density net (sigma = exp(h0)) -> colour net on the `cap` sample
// slots of the packed buffers, every launch taking the live count from DEVICE memory (n_dev: the round's counters[1]).  Lets the
// host enqueue the field of round k right behind its emit pass and read the round's two counters (from a side stream) WHILE the
// field runs: the next round is then sized by exact counts and the GPU never waits for the host (rendering.render_wavefront_pipelined).
NGP_API int ngp_field_compact_fw(const float* xyzs, const float* dirs, int64_t cap, const int32_t* n_dev, const float* aabb, const void* table,
                                 int table_dtype, int n_levels, int n_features, int log2_hashmap_size, int base_resolution,
                                 float per_level_scale, const float* sigma_params, const float* rgb_params, int width, int rgb_hidden,
                                 void* feat_tiles, float* h, float* sigmas, float* rgbs, void* stream) {
  if (cap <= 0) return 0;
  const int LF = n_levels * n_features;
  int rc = hashgrid_fw_tiles_launch(xyzs, aabb, table, table_dtype, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale,
                                    cap, n_dev, feat_tiles, stream);
  if (rc) return rc;
  {
    const float* sp[1] = {reinterpret_cast<const float*>(feat_tiles)};
    const int sw[1] = {LF}, sk[1] = {2};
    const int64_t ss[1] = {0};
    rc = mlp_fw_launch(1, sp, sw, sk, ss, sigma_params, width, 1, 16, /*ReLU*/ 1, /*None*/ 0, cap, n_dev, h, 16, sigmas, stream);
    if (rc) return rc;
  }
  const float* sp[2] = {dirs, h};
  const int sw[2] = {16, 16}, sk[2] = {1, 0};
  const int64_t ss[2] = {3, 16};
  return mlp_fw_launch(2, sp, sw, sk, ss, rgb_params, width, rgb_hidden, 3, /*ReLU*/ 1, /*Sigmoid*/ 2, cap, n_dev, rgbs, 3, nullptr, stream);
}
