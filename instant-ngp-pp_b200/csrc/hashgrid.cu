// Multiresolution hash-grid encoding: forward, parameter gradient (atomic scatter), input gradient
// and the double-backward of the input gradient.
//
// Semantics follow tiny-cuda-nn's GridEncoding (type Hash, interpolation Linear) as the reference
// uses it (models/networks.py:40-52, 67-76).  tiny-cuda-nn is NOT part of
// /root/reference, so this half of the path is written from its published algorithm:
//   scale_l = exp2(l*log2(per_level_scale))*base_resolution - 1 ; res_l = ceil(scale_l)+1
//   n_l     = min(next_multiple(res_l^3, 8), 2^log2_T)      (dense while res_l^3 <= n_l)
//   pos     = fma(scale_l, x, 0.5) ; cell = floor(pos) ; w = pos - cell
//   index   = dense ? x + y*res + z*res^2 : x ^ y*2654435761 ^ z*805459861        (mod n_l)
//   y[l*F+f]= sum_corners prod_d (c_d ? w_d : 1-w_d) * table[offset_l + index][f]
//
// B200 mapping: one thread owns (sample, chunk of LC consecutive levels) with LC*F >= 8 floats, so
// the thread issues 8*LC independent vector gathers (ld.global.nc.v2/v4) before it needs any of
// them, and writes one full 32-byte sector of the (N, L*F) output.  blockIdx.y walks level chunks,
// so at any instant a CTA's gathers hit one or two levels (coarse levels stay L1/L2 resident; the
// 43 MB fp32 table of the L16/F2/T2^19 shape fits the 126 MB L2 outright).  Gradients are
// scattered with vector reductions (red.global.add.v2/v4.f32).
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace ngp {

constexpr int kMaxLevels = 32;

struct GridMeta {
  int n_levels;
  int n_features;          // F
  uint32_t offset[kMaxLevels + 1];  // in entries (each entry = F params)
  uint32_t size[kMaxLevels];        // entries in level (hashmap size)
  uint32_t res[kMaxLevels];
  float scale[kMaxLevels];
  uint8_t dense[kMaxLevels];
  // optional world -> unit-cube map fused into every kernel: x01 = (x - lo) / range with IEEE sub and div, i.e.
  // bit-identical to the (x - xyz_min) / (xyz_max - xyz_min) tensor pass of models/networks.py:174,188
  int affine;              // 0 none, 1 (x - lo) / range, 2 (x - lo) * (1/range) with power-of-two ranges (exact)
  float lo[3], range[3], inv_range[3];
  int k0p;                 // L*F rounded up to 16: operand-tile width of the feature-tile / gradient-tile layouts
  // Block order (block_coords below).  0 = level chunk FASTEST: the n_chunks CTAs that touch the same rows of x / y run
  // back to back, so the rows are served from L2 instead of being swept from HBM once per chunk — right while the whole
  // table is L2 resident (44 MB at T=2^19 F=2).  1 = level chunk SLOWEST: all sample blocks of chunk 0, then chunk 1, ...
  // so that only the 1-4 levels of one chunk (32 MB each at T=2^22 F=2) are live in the 126 MB L2 at any time — for
  // tables beyond the L2, where a random entry otherwise costs a DRAM sector fetch (gather) or a fetch + write-back
  // (reduction); the price is re-reading x (12 B) per chunk and, for the scatter, a dL/dy sector shared by two chunks.
  // Measured on the street shape (114 M samples, 363 MB table, tools/hash_order_probe.py): scatter 71.7 -> 39.3 ms,
  // gather 19.7 -> 19.0 ms (its chunk is 4 levels = 128 MB live).  Tried and dropped: one level per lane pair in the
  // scatter (51 ms: twice the row loads), evict-first (ld.global.cs) row loads (41 ms), prefetch.global.L2 of a run's x /
  // dL/dy rows ahead of its serial loop (40.8 ms: no change, although ncu shows long-scoreboard stalls with DRAM at 21 %
  // and L2 at 46 % — the misses that matter are the reductions' own sector fetches, L2 hit rate 54 %).
  int chunk_major;
  uint32_t n_sblocks;      // sample blocks of this launch (chunk_major only)
  const int32_t* n_dev;    // gather only, optional: the sample count lives in device memory (n = min(n, *n_dev)); blocks past it exit
  int chunk0, level_end;   // scatter over a level RANGE [chunk0 * LC, level_end) (ngp_hashgrid_bw_params_tiles_range); default 0, n_levels
};
__device__ __forceinline__ void block_coords(const GridMeta& m, int n_chunks, uint32_t& sblock, int& chunk) {
  if (m.chunk_major) { chunk = (int)(blockIdx.x / m.n_sblocks); sblock = blockIdx.x % m.n_sblocks; }
  else if ((n_chunks & (n_chunks - 1)) == 0) {                 // 16 levels in chunks of 2 / 4 / 8: shift and mask instead of the
    const int sh = __ffs(n_chunks) - 1;                         // integer division (11 % of the gather's stall samples sat on it)
    sblock = blockIdx.x >> sh; chunk = (int)(blockIdx.x & (uint32_t)(n_chunks - 1));
  } else { sblock = blockIdx.x / n_chunks; chunk = (int)(blockIdx.x % n_chunks); }
}
__device__ __forceinline__ void to_unit(const GridMeta& m, float& x, float& y, float& z) {
  if (m.affine == 2) {            // every range is a power of two: the multiplication by 1/range is exact, same bits as the division
    x = __fmul_rn(__fsub_rn(x, m.lo[0]), m.inv_range[0]);
    y = __fmul_rn(__fsub_rn(y, m.lo[1]), m.inv_range[1]);
    z = __fmul_rn(__fsub_rn(z, m.lo[2]), m.inv_range[2]);
  } else if (m.affine) {
    x = __fdiv_rn(__fsub_rn(x, m.lo[0]), m.range[0]);
    y = __fdiv_rn(__fsub_rn(y, m.lo[1]), m.range[1]);
    z = __fdiv_rn(__fsub_rn(z, m.lo[2]), m.range[2]);
  }
}

__device__ __forceinline__ uint32_t grid_index(uint32_t x, uint32_t y, uint32_t z, uint32_t res, uint32_t size,
                                               bool dense) {
  if (dense) {
    uint32_t i = x + y * res + z * res * res;
    if (i >= size) i -= size;  // == i % size for the reachable range (x,y,z <= res)
    return i;
  }
  const uint32_t h = x ^ (y * 2654435761u) ^ (z * 805459861u);
  return (size & (size - 1)) == 0 ? (h & (size - 1)) : (h % size);
}

// The four (y, z) corners of x-neighbour xh of a cell, with the level's addressing mode decided ONCE (the per-corner
// form above re-tests dense / power-of-two for every corner: index arithmetic was most of the gather's instructions
// once the lane-pair layout had lifted the L1 bound, ncu r01c: issue slots 89 % busy).  Same values as grid_index.
__device__ __forceinline__ void corner4(uint32_t x, uint32_t py, uint32_t pz, uint32_t res, uint32_t size, bool dense,
                                        uint32_t* idx) {
  if (dense) {
    const uint32_t r2 = res * res;
    const uint32_t b = x + py * res + pz * r2;
    idx[0] = b; idx[1] = b + res; idx[2] = b + r2; idx[3] = b + res + r2;
#pragma unroll
    for (int p = 0; p < 4; p++) if (idx[p] >= size) idx[p] -= size;   // == % size for the reachable range
  } else {
    const uint32_t hy0 = py * 2654435761u, hy1 = hy0 + 2654435761u;
    const uint32_t hz0 = pz * 805459861u, hz1 = hz0 + 805459861u;
    idx[0] = x ^ hy0 ^ hz0; idx[1] = x ^ hy1 ^ hz0; idx[2] = x ^ hy0 ^ hz1; idx[3] = x ^ hy1 ^ hz1;
    if ((size & (size - 1)) == 0) {
#pragma unroll
      for (int p = 0; p < 4; p++) idx[p] &= size - 1;
    } else {
#pragma unroll
      for (int p = 0; p < 4; p++) idx[p] %= size;
    }
  }
}

template <int F, typename TP> struct Vec;
template <> struct Vec<1, float> { static __device__ __forceinline__ void ld(const float* p, float* o) { o[0] = __ldg(p); } };
template <> struct Vec<2, float> { static __device__ __forceinline__ void ld(const float* p, float* o) { const float2 v = __ldg((const float2*)p); o[0] = v.x; o[1] = v.y; } };
template <> struct Vec<4, float> { static __device__ __forceinline__ void ld(const float* p, float* o) { const float4 v = __ldg((const float4*)p); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; } };
template <> struct Vec<8, float> { static __device__ __forceinline__ void ld(const float* p, float* o) {
  const float4 a = __ldg((const float4*)p), b = __ldg((const float4*)p + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w; } };
template <> struct Vec<1, __half> { static __device__ __forceinline__ void ld(const __half* p, float* o) { o[0] = __half2float(__ldg(p)); } };
template <> struct Vec<2, __half> { static __device__ __forceinline__ void ld(const __half* p, float* o) { const float2 v = __half22float2(__ldg((const __half2*)p)); o[0] = v.x; o[1] = v.y; } };
template <> struct Vec<4, __half> { static __device__ __forceinline__ void ld(const __half* p, float* o) {
  const uint2 r = __ldg((const uint2*)p);
  const float2 a = __half22float2(*(const __half2*)&r.x), b = __half22float2(*(const __half2*)&r.y);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; } };
template <> struct Vec<8, __half> { static __device__ __forceinline__ void ld(const __half* p, float* o) {
  const uint4 r = __ldg((const uint4*)p);
  const float2 a = __half22float2(*(const __half2*)&r.x), b = __half22float2(*(const __half2*)&r.y);
  const float2 c = __half22float2(*(const __half2*)&r.z), d = __half22float2(*(const __half2*)&r.w);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x; o[7] = d.y; } };

template <int F> __device__ __forceinline__ void red_add(float* p, const float* v) {
  if constexpr (F == 1) atomicAdd(p, v[0]);
  else if constexpr (F == 2) atomicAdd((float2*)p, make_float2(v[0], v[1]));
  else if constexpr (F == 4) atomicAdd((float4*)p, make_float4(v[0], v[1], v[2], v[3]));
  else { atomicAdd((float4*)p, make_float4(v[0], v[1], v[2], v[3])); atomicAdd((float4*)p + 1, make_float4(v[4], v[5], v[6], v[7])); }
}

struct Cell { uint32_t px, py, pz; float wx, wy, wz; };
__device__ __forceinline__ Cell locate(float x, float y, float z, float scale) {
  Cell c;
  float fx = fmaf(scale, x, 0.5f), fy = fmaf(scale, y, 0.5f), fz = fmaf(scale, z, 0.5f);
  const float gx = floorf(fx), gy = floorf(fy), gz = floorf(fz);
  c.px = (uint32_t)(int)gx; c.py = (uint32_t)(int)gy; c.pz = (uint32_t)(int)gz;
  c.wx = fx - gx; c.wy = fy - gy; c.wz = fz - gz;
  return c;
}

template <int F> constexpr int levels_per_thread() { return F >= 8 ? 1 : 8 / F; }
// the scatter keeps 4 corner accumulators per level in registers: 16 bytes of dL/dy (one float4) per lane and sample
// give 70 instead of 104 registers and were 6 % faster (tools/hash_sweep.py)
template <int F> constexpr int scatter_levels_per_thread() { return F >= 4 ? 1 : 4 / F; }

// ----------------------------------------------------------------------------------- forward
// Feature tiles (private layout of the fused density path, see mlp.cu "kSegTiles"): per 128-sample tile the
// encoder writes bf16 features directly in the tcgen05 operand layout of the MLP's input tile
//     byte(tile, r, c) = tile*kFeatTileBytes(k0p) + (c/8)*(128*16 + 64) + r*16 + (c%8)*2
// — one 16-byte row chunk per thread (LC*F == 8 columns), consecutive samples contiguous — so the MLP loads a
// tile with ONE bulk copy and no conversion pass, and the fp32 (N, L*F) feature matrix never exists.
__host__ __device__ __forceinline__ uint32_t feat_chunk_stride() { return 128u * 16u + 64u; }
__host__ __device__ __forceinline__ uint32_t feat_tile_bytes(int k0p) { return (uint32_t)(k0p / 8) * feat_chunk_stride(); }

// The gather is bound by the L1 tag stage: ONE 128-byte line per clock per SM, whatever the sectors
// (tools/probes/l1_probe.cu: 32 lanes on 32 lines 285 G loads/s, lane pairs on the same line 562 G, octets 1159 G).
// So the two x-neighbours of a corner pair — adjacent entries for dense levels, h(x) and h(x+1) = entries that differ
// only in their low bits for hashed ones, i.e. the same line 15 times out of 16 — are fetched by the two lanes of a
// lane PAIR in the same instruction: ~4.3 line lookups per (sample, level) instead of 6 (4 when x is even and the pair
// could go out as one 16-byte load, 8 when it is odd).  The pair's partial sums meet through one shuffle per feature.
//
// H2 (double backward of the input gradient, see ngp_hashgrid_bwbw_input): the same gather with the trilinear weight of a
// corner replaced by the coefficient of that corner in g2 . dy/dx,
//     coef = scale_l * (g2_x * s_x * w_y * w_z + g2_y * s_y * w_x * w_z + g2_z * s_z * w_x * w_y),   s_d = +1 / -1 for the far / near corner,
// so that y = d(g2 . dL/dx)/d(dL/dy).
template <int F, typename TP, bool TILES, bool H2 = false>
__global__ void __launch_bounds__(256) hashgrid_fw_kernel(const float* __restrict__ x, const TP* __restrict__ table,
                                                          GridMeta m, int64_t n, float* __restrict__ y,
                                                          const float* __restrict__ g2 = nullptr) {
  constexpr int LC = levels_per_thread<F>();
  const int n_chunks = TILES ? m.k0p / 8 : (m.n_levels + LC - 1) / LC;
  uint32_t sblock; int chunk;
  block_coords(m, n_chunks, sblock, chunk);                                                        // GridMeta::chunk_major
  if (m.n_dev) {                                    // launch sized for a bound: whole blocks past the live count leave at once
    const int64_t nd = (int64_t)__ldg(m.n_dev);
    if (nd < n) n = nd;
    if ((int64_t)sblock * 128 >= ((n + 127) >> 7 << 7)) return;
  }
  const int64_t i = (int64_t)sblock * (blockDim.x >> 1) + (threadIdx.x >> 1);                      // 128 samples per CTA
  const uint32_t xh = threadIdx.x & 1u;                                                            // which x-neighbour
  const int l0 = chunk * LC;
  const bool in_range = i < n;
  const int64_t ii = in_range ? i : n - 1;          // out-of-range lanes still take part in the pair shuffles
  float xx = __ldg(x + 3 * ii), xy = __ldg(x + 3 * ii + 1), xz = __ldg(x + 3 * ii + 2);
  to_unit(m, xx, xy, xz);
  float hx = 0.f, hy = 0.f, hz = 0.f;
  if (H2) { hx = __ldg(g2 + 3 * ii); hy = __ldg(g2 + 3 * ii + 1); hz = __ldg(g2 + 3 * ii + 2); }
  float out[LC * F];
#pragma unroll
  for (int k = 0; k < LC * F; k++) out[k] = 0.f;
#pragma unroll
  for (int li = 0; li < LC; li++) {
    const int l = l0 + li;
    if (l < m.n_levels) {
      const Cell c = locate(xx, xy, xz, m.scale[l]);
      const TP* base = table + (size_t)m.offset[l] * F;
      const uint32_t res = m.res[l], size = m.size[l];
      const bool dense = m.dense[l];
      const float wxs = xh ? c.wx : 1.f - c.wx;
      float v[4][F];
      uint32_t idx[4];
      corner4(c.px + xh, c.py, c.pz, res, size, dense, idx);
#pragma unroll
      for (int p = 0; p < 4; p++) Vec<F, TP>::ld(base + (size_t)idx[p] * F, v[p]);
#pragma unroll
      for (int p = 0; p < 4; p++) {
        const float wy = (p & 1) ? c.wy : 1.f - c.wy, wz = ((p >> 1) & 1) ? c.wz : 1.f - c.wz;
        float w;
        if (H2) w = m.scale[l] * (hx * (xh ? 1.f : -1.f) * wy * wz + hy * ((p & 1) ? 1.f : -1.f) * wxs * wz + hz * (((p >> 1) & 1) ? 1.f : -1.f) * wxs * wy);
        else w = wxs * wy * wz;
#pragma unroll
        for (int f = 0; f < F; f++) out[li * F + f] = fmaf(w, v[p][f], out[li * F + f]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < LC * F; k++) out[k] += __shfl_xor_sync(0xffffffffu, out[k], 1);
  if (xh) return;
  if (TILES) {
    static_assert(LC * F == 8, "one 16-byte bf16 row chunk per (sample, level chunk)");
    // n_chunks covers the PADDED width (k0p/8 chunks); rows past n of the last tile and chunks past the last level
    // come out as zeros (out[] is 0 there / forced here): the MLP multiplies whole 128-row tiles
    if (i >= ((n + 127) >> 7 << 7)) return;
    uint8_t* dstt = reinterpret_cast<uint8_t*>(y) + (i >> 7) * (int64_t)feat_tile_bytes(m.k0p) + (uint32_t)(l0 / LC) * feat_chunk_stride() + (uint32_t)(i & 127) * 16u;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (in_range) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(out[0], out[1]), h1 = __floats2bfloat162_rn(out[2], out[3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(out[4], out[5]), h3 = __floats2bfloat162_rn(out[6], out[7]);
      q.x = *reinterpret_cast<uint32_t*>(&h0); q.y = *reinterpret_cast<uint32_t*>(&h1);
      q.z = *reinterpret_cast<uint32_t*>(&h2); q.w = *reinterpret_cast<uint32_t*>(&h3);
    }
    *reinterpret_cast<uint4*>(dstt) = q;
    return;
  }
  if (!in_range) return;
  const int LF = m.n_levels * F;
  float* dst = y + i * LF + (int64_t)l0 * F;
  if (l0 + LC <= m.n_levels && (LF % 4) == 0 && ((l0 * F) % 4) == 0) {
#pragma unroll
    for (int k = 0; k < LC * F; k += 4) *(float4*)(dst + k) = make_float4(out[k], out[k + 1], out[k + 2], out[k + 3]);
  } else {
#pragma unroll
    for (int k = 0; k < LC * F; k++) if (l0 * F + k < LF) dst[k] = out[k];
  }
}

// ----------------------------------------------------------------------------------- bw (params)
// dL/dtable[corner] += w_corner * dL/dy  (fp32 accumulation regardless of the table's storage type).
//
// The scatter is bound by L2 atomic throughput (one red per corner per level per sample = 128 per
// sample for L=16), so the kernel is organised to issue FEWER reds, not faster ones:
//  * a thread owns SPT consecutive samples (consecutive samples of a ray are packed together, and at a
//    level whose cells are wider than the marching step they fall into the same cell): contributions
//    to the same cell are merged in registers and flushed once per run;
//  * the two x-neighbours of a corner pair are issued by the two lanes of a lane pair in the same
//    instruction, so that they are ONE sector request whenever they share a sector (below);
//  * exactly-zero upstream rows (samples past early termination) are skipped.
constexpr int kSPT = 24;   // 8 -> 16: one forced flush per run, -8 % (tools/hash_sweep.py)

// Lane pairs again (see hashgrid_fw_kernel): lane xh of a pair accumulates the four corners with x = px + xh.  The L2
// reduction rate is bound by 32-byte SECTOR requests (~220 G/s, tools/probes/l2_red_probe.cu), and two lanes of one
// instruction that hit the same sector are one request: the x-neighbours share a sector whenever x is even (adjacent
// entries, what the former 16-byte pair red exploited) AND when x = 1 mod 4 for hashed levels (h(x+1) = h(x) ^ 3), i.e.
// 1.25 requests per corner pair instead of 1.5; halving the per-thread accumulators also lifts the occupancy.
template <int F> struct CellAcc {
  uint32_t px, py, pz;
  bool has;
  float a[4][F];
};

template <int F>
__device__ __forceinline__ void flush_cell(const CellAcc<F>& c, uint32_t xh, float* __restrict__ base, uint32_t res, uint32_t size, bool dense) {
  uint32_t idx[4];
  corner4(c.px + xh, c.py, c.pz, res, size, dense, idx);
#pragma unroll
  for (int p = 0; p < 4; p++) red_add<F>(base + (size_t)idx[p] * F, c.a[p]);
}

// DYT: dL/dy arrives in "gradient tiles" (written by the MLP backward, mlp.cu): fp32, per 128-sample tile
//     float(tile, r, c) at tile*(128*k0p) + (c/8)*(128*8) + r*8 + (c%8)
// i.e. the 8 columns of a level chunk are one 32-byte sector per sample and consecutive samples are contiguous —
// exactly what a (run of samples, level chunk) lane pair reads, and what a row-per-thread MLP epilogue writes coalesced.
// H2: the table gradient of the double backward (dtable += d(g2 . dL/dx)/dtable): the same run-merged scatter with the
// corner weight replaced by the coefficient `coef` of hashgrid_fw_kernel<.., H2>; dy is then the FIRST-order upstream.
// PF: software pipelining of the run loop — the rows (x, dL/dy[, g2]) of sample j+1 are requested before sample j is
// processed.  In chunk-major order they come from DRAM on every chunk sweep and 69 % of the stall samples sat on their
// first use (ncu r01e, street shape, profiles/r01e_ncu_street_scatter_order1.txt + per-line profile); costs 7 registers.
template <int F, int LC, bool DYT, bool H2 = false, bool PF = false>
__global__ void __launch_bounds__(128) hashgrid_bw_params_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                 GridMeta m, int64_t n, float* __restrict__ dtable, int spt,
                                                                 const float* __restrict__ g2 = nullptr) {
  const int n_chunks = (m.level_end + LC - 1) / LC - m.chunk0;
  uint32_t sblock; int chunk;
  block_coords(m, n_chunks, sblock, chunk);                 // GridMeta::chunk_major
  const int64_t s0 = ((int64_t)sblock * (blockDim.x >> 1) + (threadIdx.x >> 1)) * spt;
  const uint32_t xh = threadIdx.x & 1u;
  if (s0 >= n) return;
  const int l0 = (chunk + m.chunk0) * LC;
  const int LF = m.n_levels * F;
  CellAcc<F> acc[LC];
#pragma unroll
  for (int li = 0; li < LC; li++) acc[li].has = false;
  auto dy_row = [&](int64_t i) -> const float* {
    return DYT ? dy + (i >> 7) * (int64_t)(128 * m.k0p) + (int64_t)((l0 * F) >> 3) * (128 * 8) + (i & 127) * 8 + ((l0 * F) & 7)
               : dy + i * LF + (int64_t)l0 * F;
  };
  struct Row { float x, y, z, hx, hy, hz, g[LC * F]; };
  const bool vec4 = l0 + LC <= m.n_levels && (LF & 3) == 0 && ((l0 * F) & 3) == 0;
  auto load_row = [&](int64_t i, Row& r) {
    r.x = __ldg(x + 3 * i); r.y = __ldg(x + 3 * i + 1); r.z = __ldg(x + 3 * i + 2);
    r.hx = r.hy = r.hz = 0.f;
    if (H2) { r.hx = __ldg(g2 + 3 * i); r.hy = __ldg(g2 + 3 * i + 1); r.hz = __ldg(g2 + 3 * i + 2); }
    const float* src = dy_row(i);
    if (vec4) {
#pragma unroll
      for (int k = 0; k < LC * F; k += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + k));
        r.g[k] = v.x; r.g[k + 1] = v.y; r.g[k + 2] = v.z; r.g[k + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < LC * F; k++) r.g[k] = (l0 * F + k < LF) ? __ldg(src + k) : 0.f;
    }
  };
  Row nxt;
  if (PF) load_row(s0, nxt);
#pragma unroll 1
  for (int j = 0; j < spt; j++) {
    const int64_t i = s0 + j;
    if (i >= n) break;
    Row cur;
    if (PF) { cur = nxt; if (j + 1 < spt && i + 1 < n) load_row(i + 1, nxt); }
    else load_row(i, cur);
    float xx = cur.x, xy = cur.y, xz = cur.z;
    to_unit(m, xx, xy, xz);
    const float hx = cur.hx, hy = cur.hy, hz = cur.hz;
    const float* g = cur.g;
#pragma unroll
    for (int li = 0; li < LC; li++) {
      const int l = l0 + li;
      bool any = false;
#pragma unroll
      for (int f = 0; f < F; f++) any |= g[li * F + f] != 0.f;
      if (H2) any = any && (hx != 0.f || hy != 0.f || hz != 0.f);
      if (l < m.level_end && any) {
        const Cell c = locate(xx, xy, xz, m.scale[l]);
        CellAcc<F>& A = acc[li];
        if (!A.has || A.px != c.px || A.py != c.py || A.pz != c.pz) {
          if (A.has) flush_cell<F>(A, xh, dtable + (size_t)m.offset[l] * F, m.res[l], m.size[l], m.dense[l]);
          A.px = c.px; A.py = c.py; A.pz = c.pz; A.has = true;
#pragma unroll
          for (int k = 0; k < 4; k++)
#pragma unroll
            for (int f = 0; f < F; f++) A.a[k][f] = 0.f;
        }
        const float wxs = xh ? c.wx : 1.f - c.wx;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const float wy = (k & 1) ? c.wy : 1.f - c.wy, wz = ((k >> 1) & 1) ? c.wz : 1.f - c.wz;
          float w;
          if (H2) w = m.scale[l] * (hx * (xh ? 1.f : -1.f) * wy * wz + hy * ((k & 1) ? 1.f : -1.f) * wxs * wz + hz * (((k >> 1) & 1) ? 1.f : -1.f) * wxs * wy);
          else w = wxs * wy * wz;
#pragma unroll
          for (int f = 0; f < F; f++) A.a[k][f] = fmaf(w, g[li * F + f], A.a[k][f]);
        }
      }
    }
  }
#pragma unroll
  for (int li = 0; li < LC; li++) {
    const int l = l0 + li;
    if (l < m.level_end && acc[li].has)
      flush_cell<F>(acc[li], xh, dtable + (size_t)m.offset[l] * F, m.res[l], m.size[l], m.dense[l]);
  }
}

// F = 8 (the reference's grids, networks.py:40-52,67-76): an entry is 32 bytes = one whole sector, so the x-neighbours of
// a corner pair can never share a sector and every corner costs TWO 16-byte reductions = two L2 sector requests in the
// kernel above.  Here the lanes of a pair split the FEATURES instead: lane h owns features 4h..4h+3 of all eight
// corners, the pair's two float4 reductions to one entry are the same instruction and the same sector = ONE request:
// 8 requests per (cell, level) instead of 16.  Same run merging, same zero-row skip, same H2 weights.
// MODE 0: dtable += sum w * dy.  MODE 1 (H2): dtable += sum coef(g2) * dy.  MODE 2 (dual): both terms of the table
// gradient of a field whose output AND input gradient are used (density + normals, networks.py:186-196) in one pass:
// dtable += sum w * dy + coef(g2) * dy2 — the two scatters hit the same corners, so the second one's reductions are free.
template <bool DYT, int MODE>
__global__ void __launch_bounds__(128) hashgrid_bw_params_f8_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                    GridMeta m, int64_t n, float* __restrict__ dtable, int spt,
                                                                    const float* __restrict__ g2 = nullptr,
                                                                    const float* __restrict__ dy2 = nullptr) {
  constexpr int F = 8;
  constexpr bool H2 = MODE != 0;
  uint32_t sblock; int l;
  block_coords(m, m.level_end - m.chunk0, sblock, l);       // one level per lane pair
  l += m.chunk0;
  const int64_t s0 = ((int64_t)sblock * (blockDim.x >> 1) + (threadIdx.x >> 1)) * spt;
  const uint32_t fh = (threadIdx.x & 1u) * 4u;              // this lane's feature half
  if (s0 >= n) return;
  const int LF = m.n_levels * F;
  const float scale = m.scale[l];
  const uint32_t res = m.res[l], size = m.size[l];
  const bool dense = m.dense[l];
  float* base = dtable + (size_t)m.offset[l] * F + fh;
  uint32_t cx = 0, cy = 0, cz = 0;
  bool has = false;
  float a[8][4];                                            // corner (x bit 2, z bit 1, y bit 0) x feature of this half
  auto flush = [&]() {
#pragma unroll
    for (int xb = 0; xb < 2; xb++) {
      uint32_t idx[4];
      corner4(cx + xb, cy, cz, res, size, dense, idx);
#pragma unroll
      for (int p = 0; p < 4; p++) red_add<4>(base + (size_t)idx[p] * F, a[xb * 4 + p]);
    }
  };
#pragma unroll 1
  for (int j = 0; j < spt; j++) {
    const int64_t i = s0 + j;
    if (i >= n) break;
    float xx = __ldg(x + 3 * i), xy = __ldg(x + 3 * i + 1), xz = __ldg(x + 3 * i + 2);
    to_unit(m, xx, xy, xz);
    float hx = 0.f, hy = 0.f, hz = 0.f;
    if (H2) { hx = __ldg(g2 + 3 * i); hy = __ldg(g2 + 3 * i + 1); hz = __ldg(g2 + 3 * i + 2); }
    const int64_t off = DYT ? (i >> 7) * (int64_t)(128 * m.k0p) + (int64_t)l * (128 * 8) + (i & 127) * 8 + fh
                            : i * LF + (int64_t)l * F + fh;
    const float4 g = __ldg(reinterpret_cast<const float4*>(dy + off));
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);            // MODE 2: the rows weighted by coef(g2)
    if (MODE == 2) q = __ldg(reinterpret_cast<const float4*>(dy2 + off));
    const bool hnz = hx != 0.f || hy != 0.f || hz != 0.f;
    const bool gnz = g.x != 0.f || g.y != 0.f || g.z != 0.f || g.w != 0.f;
    const bool qnz = q.x != 0.f || q.y != 0.f || q.z != 0.f || q.w != 0.f;
    const bool any = MODE == 0 ? gnz : MODE == 1 ? (gnz && hnz) : (gnz || (qnz && hnz));
    if (!any) continue;
    const Cell c = locate(xx, xy, xz, scale);
    if (!has || cx != c.px || cy != c.py || cz != c.pz) {
      if (has) flush();
      cx = c.px; cy = c.py; cz = c.pz; has = true;
#pragma unroll
      for (int k = 0; k < 8; k++) { a[k][0] = 0.f; a[k][1] = 0.f; a[k][2] = 0.f; a[k][3] = 0.f; }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int xb = k >> 2, yb = k & 1, zb = (k >> 1) & 1;
      const float wx = xb ? c.wx : 1.f - c.wx, wy = yb ? c.wy : 1.f - c.wy, wz = zb ? c.wz : 1.f - c.wz;
      const float w = wx * wy * wz;
      float coef = 0.f;
      if (H2) coef = scale * (hx * (xb ? 1.f : -1.f) * wy * wz + hy * (yb ? 1.f : -1.f) * wx * wz + hz * (zb ? 1.f : -1.f) * wx * wy);
      if (MODE == 2) {
        a[k][0] = fmaf(coef, q.x, fmaf(w, g.x, a[k][0])); a[k][1] = fmaf(coef, q.y, fmaf(w, g.y, a[k][1]));
        a[k][2] = fmaf(coef, q.z, fmaf(w, g.z, a[k][2])); a[k][3] = fmaf(coef, q.w, fmaf(w, g.w, a[k][3]));
      } else {
        const float u = MODE == 1 ? coef : w;
        a[k][0] = fmaf(u, g.x, a[k][0]); a[k][1] = fmaf(u, g.y, a[k][1]);
        a[k][2] = fmaf(u, g.z, a[k][2]); a[k][3] = fmaf(u, g.w, a[k][3]);
      }
    }
  }
  if (has) flush();
}

// ----------------------------------------------------------------------------------- bw (input)
// dL/dx_d = sum_l scale_l * sum_f dL/dy_{l,f} * sum_{corners of the other two dims} w_other * (v[d=1]-v[d=0])
// Lane pairs as in hashgrid_fw_kernel: lane xh fetches the four corners with x = px + xh (one L1 line lookup for the two
// x-neighbours of a corner pair instead of two), forms d_p = dL/dy . v_p for them, and contributes
//     gx += scale * s_x * sum_p w_y w_z d_p          (s_x = +1 for xh = 1, -1 for xh = 0)
//     gy += scale * w_x * sum_z w_z (d[y=1] - d[y=0]) ,  gz likewise;
// the two lanes' partial sums over ALL levels meet through three shuffles at the end.  The first version (thread per
// sample, eight 32-byte gathers per level for F = 8) took 10.5 ms on 14 M samples against 6.3 ms for the forward gather.
template <int F, typename TP>
__global__ void __launch_bounds__(256) hashgrid_bw_input_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                const TP* __restrict__ table, GridMeta m, int64_t n,
                                                                float* __restrict__ dx) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 1) + (threadIdx.x >> 1);
  const uint32_t xh = threadIdx.x & 1u;
  const bool in_range = i < n;
  const int64_t ii = in_range ? i : n - 1;          // out-of-range lanes still take part in the pair shuffles
  float xx = __ldg(x + 3 * ii), xy = __ldg(x + 3 * ii + 1), xz = __ldg(x + 3 * ii + 2);
  to_unit(m, xx, xy, xz);
  const int LF = m.n_levels * F;
  const float sx = xh ? 1.f : -1.f;
  float gx = 0.f, gy = 0.f, gz = 0.f;
  for (int l = 0; l < m.n_levels; l++) {
    const Cell c = locate(xx, xy, xz, m.scale[l]);
    const TP* base = table + (size_t)m.offset[l] * F;
    float g[F];
    const float* gp = dy + ii * LF + l * F;
    if constexpr (F % 4 == 0) {
      if ((LF & 3) == 0) {
#pragma unroll
        for (int f = 0; f < F; f += 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(gp + f)); g[f] = v.x; g[f + 1] = v.y; g[f + 2] = v.z; g[f + 3] = v.w; }
      } else {
#pragma unroll
        for (int f = 0; f < F; f++) g[f] = __ldg(gp + f);
      }
    } else {
#pragma unroll
      for (int f = 0; f < F; f++) g[f] = __ldg(gp + f);
    }
    uint32_t idx[4];
    corner4(c.px + xh, c.py, c.pz, m.res[l], m.size[l], m.dense[l], idx);
    float v[4][F], d[4];      // p: bit0 = y, bit1 = z
#pragma unroll
    for (int p = 0; p < 4; p++) Vec<F, TP>::ld(base + (size_t)idx[p] * F, v[p]);
#pragma unroll
    for (int p = 0; p < 4; p++) {
      float a = 0.f;
#pragma unroll
      for (int f = 0; f < F; f++) a = fmaf(g[f], v[p][f], a);
      d[p] = a;
    }
    const float ay = 1.f - c.wy, az = 1.f - c.wz, wxs = xh ? c.wx : 1.f - c.wx;
    const float s = m.scale[l];
    gx += s * sx * (ay * az * d[0] + c.wy * az * d[1] + ay * c.wz * d[2] + c.wy * c.wz * d[3]);
    gy += s * wxs * (az * (d[1] - d[0]) + c.wz * (d[3] - d[2]));
    gz += s * wxs * (ay * (d[2] - d[0]) + c.wy * (d[3] - d[1]));
  }
  gx += __shfl_xor_sync(0xffffffffu, gx, 1);
  gy += __shfl_xor_sync(0xffffffffu, gy, 1);
  gz += __shfl_xor_sync(0xffffffffu, gz, 1);
  if (xh == 0 && in_range) { dx[3 * i] = gx; dx[3 * i + 1] = gy; dx[3 * i + 2] = gz; }
}

// ----------------------------------------------------------------------------------- double bw
// Given g2 = dL/d(dL/dx) (N,3) and the first-order upstream gy = dL/dy (N, L*F):
//   dL/dtable[corner c] += gy * scale * sum_d g2_d * sign_d(c) * prod_{d' != d} w_{d'}(c)     -> hashgrid_bw_params_kernel<.., H2>
//   dL/d(gy)_{l,f}       = sum_d g2_d * dy_{l,f}/dx_d                                           -> hashgrid_fw_kernel<.., H2>
// (second derivatives of the trilinear kernel w.r.t. x itself are not propagated: x is a leaf produced by the marcher
// under no_grad in the reference, models/rendering.py:207-212).  The first version was one thread per (sample, level)
// issuing its 8 corner reductions unmerged: 119 ms for 14 M samples of the street shape (F=8), 35 % of the step of the
// reference-literal field — consecutive samples of a ray sit in the same coarse cell, i.e. 32-way same-address
// reductions; the run-merged lane-pair kernels do the same work in the time of an ordinary gather + scatter.

static int fill_meta(GridMeta& m, int n_levels, int F, int log2_T, int base_res, float per_level_scale, const float* aabb = nullptr) {
  if (n_levels < 1 || n_levels > kMaxLevels) return -1;
  if (!(F == 1 || F == 2 || F == 4 || F == 8)) return -1;
  m.n_levels = n_levels; m.n_features = F;
  m.k0p = (n_levels * F + 15) / 16 * 16;
  m.affine = aabb != nullptr;
  bool pow2 = aabb != nullptr;
  for (int d = 0; d < 3; d++) {
    m.lo[d] = aabb ? aabb[d] : 0.f; m.range[d] = aabb ? aabb[3 + d] : 1.f;
    int e; const float fr = frexpf(m.range[d], &e);
    pow2 = pow2 && fr == 0.5f && e > -100 && e < 100;      // normal power of two: x * (1/range) == x / range bit for bit
    m.inv_range[d] = 1.0f / m.range[d];
  }
  if (pow2) m.affine = 2;
  const float log2_pls = log2f(per_level_scale);
  uint32_t off = 0;
  for (int l = 0; l < n_levels; l++) {
    const float scale = exp2f(l * log2_pls) * base_res - 1.0f;   // tcnn grid_scale
    const uint32_t res = (uint32_t)ceilf(scale) + 1;             // tcnn grid_resolution
    const uint64_t dense_n = (uint64_t)res * res * res;
    const uint64_t cap = 1ull << log2_T;
    uint64_t sz = (dense_n + 7) / 8 * 8;
    if (sz > cap) sz = cap;
    m.offset[l] = off; m.size[l] = (uint32_t)sz; m.res[l] = res; m.scale[l] = scale;
    m.dense[l] = dense_n <= sz ? 1 : 0;
    off += (uint32_t)sz;
  }
  m.offset[n_levels] = off;
  m.chunk_major = 0; m.n_sblocks = 1; m.chunk0 = 0; m.level_end = n_levels; m.n_dev = nullptr;
  return 0;
}

// Block order of the gather / scatter launches (GridMeta::chunk_major): chunk-major once the table no longer fits the
// L2 with room for the streamed rows.  NGP_HASH_ORDER=0|1 overrides (tools/hash_order_probe.py).
constexpr size_t kL2ResidentTableBytes = 72u << 20;
// software-pipelined scatter (PF): where the rows come from DRAM, i.e. in chunk-major order; NGP_HASH_PF=0|1 overrides
static bool use_pf(const GridMeta& m) {
  const char* e = getenv("NGP_HASH_PF");
  return e ? atoi(e) != 0 : m.chunk_major != 0;
}
static void set_block_order(GridMeta& m, size_t entry_bytes, int64_t n_sblocks) {
  const char* e = getenv("NGP_HASH_ORDER");
  m.chunk_major = e ? atoi(e) != 0 : (size_t)m.offset[m.n_levels] * entry_bytes > kL2ResidentTableBytes;
  m.n_sblocks = (uint32_t)n_sblocks;
}

}  // namespace ngp

using namespace ngp;

// Level table of a tcnn-shaped grid.  Fills offsets (L+1, in entries), sizes, resolutions, scales,
// dense flags (each L) and returns the total number of parameters (entries*F), or -1 on bad config.
// Replaces the constructor-side bookkeeping of tcnn.Encoding({"otype":"Grid"/"HashGrid"}) used at
// models/networks.py:40-52,67-76.
NGP_API int64_t ngp_hashgrid_layout(int n_levels, int n_features, int log2_hashmap_size, int base_resolution,
                                    float per_level_scale, uint32_t* offsets, uint32_t* sizes, uint32_t* resolutions,
                                    float* scales, uint8_t* dense) {
  GridMeta m;
  if (fill_meta(m, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale)) {
    set_error_msg("ngp_hashgrid_layout: need 1<=n_levels<=32 and n_features in {1,2,4,8}");
    return -1;
  }
  for (int l = 0; l < n_levels; l++) {
    if (offsets) offsets[l] = m.offset[l];
    if (sizes) sizes[l] = m.size[l];
    if (resolutions) resolutions[l] = m.res[l];
    if (scales) scales[l] = m.scale[l];
    if (dense) dense[l] = m.dense[l];
  }
  if (offsets) offsets[n_levels] = m.offset[n_levels];
  return (int64_t)m.offset[n_levels] * n_features;
}

#define NGP_F_DISPATCH(F_, ...)                                 \
  switch (F_) {                                                 \
    case 1: { constexpr int F = 1; __VA_ARGS__; } break;        \
    case 2: { constexpr int F = 2; __VA_ARGS__; } break;        \
    case 4: { constexpr int F = 4; __VA_ARGS__; } break;        \
    default: { constexpr int F = 8; __VA_ARGS__; } break;       \
  }

// y (N, L*F) f32 = encode(x (N,3) f32 in [0,1]).  table_dtype: 0 = f32, 1 = f16.
// aabb (all four entry points; HOST pointer, may be NULL): {lo_x, lo_y, lo_z, range_x, range_y, range_z} — the kernels
// then take world-space x and map it with (x - lo) / range themselves (networks.py:174 fused; bit-identical).
// Replaces tcnn.Encoding.forward for the Grid encoding (called from models/networks.py:177,182).
NGP_API int ngp_hashgrid_fw(const float* x, const float* aabb, const void* table, int table_dtype, int n_levels, int n_features,
                            int log2_hashmap_size, int base_resolution, float per_level_scale, int64_t n, float* y,
                            void* stream) {
  if (n <= 0) return 0;
  GridMeta m;
  if (fill_meta(m, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale, aabb))
    return set_error_msg("ngp_hashgrid_fw: bad grid config");
  cudaStream_t st = (cudaStream_t)stream;
  NGP_F_DISPATCH(n_features, {
    constexpr int LC = levels_per_thread<F>();
    const unsigned grid = (unsigned)(ceil_div(n, 128) * ceil_div(n_levels, LC));     // lane pair per (sample, level chunk)
    set_block_order(m, (size_t)F * (table_dtype == 0 ? 4 : 2), ceil_div(n, 128));
    if (table_dtype == 0) hashgrid_fw_kernel<F, float, false><<<grid, 256, 0, st>>>(x, (const float*)table, m, n, y);
    else hashgrid_fw_kernel<F, __half, false><<<grid, 256, 0, st>>>(x, (const __half*)table, m, n, y);
  });
  NGP_LAUNCH_CHECK("ngp_hashgrid_fw");
  return 0;
}

// dtable (n_params) f32 += scatter(dL/dy).  The caller zeroes dtable (or accumulates on purpose).
NGP_API int ngp_hashgrid_bw_params(const float* x, const float* aabb, const float* dL_dy, int n_levels, int n_features,
                                   int log2_hashmap_size, int base_resolution, float per_level_scale, int64_t n,
                                   float* dtable, void* stream) {
  if (n <= 0) return 0;
  GridMeta m;
  if (fill_meta(m, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale, aabb))
    return set_error_msg("ngp_hashgrid_bw_params: bad grid config");
  NGP_F_DISPATCH(n_features, {
    constexpr int LC = scatter_levels_per_thread<F>();
    const char* e = getenv("NGP_HASH_SPT");
    const int spt = e ? atoi(e) : kSPT;
    const unsigned grid = (unsigned)(ceil_div(ceil_div(n, spt), 64) * ceil_div(n_levels, LC));
    set_block_order(m, (size_t)F * 4, ceil_div(ceil_div(n, spt), 64));
    if (F == 8 && (n_levels * F) % 4 == 0) hashgrid_bw_params_f8_kernel<false, 0><<<grid, 128, 0, (cudaStream_t)stream>>>(x, dL_dy, m, n, dtable, spt);
    else if (use_pf(m)) hashgrid_bw_params_kernel<F, LC, false, false, true><<<grid, 128, 0, (cudaStream_t)stream>>>(x, dL_dy, m, n, dtable, spt);
    else hashgrid_bw_params_kernel<F, LC, false><<<grid, 128, 0, (cudaStream_t)stream>>>(x, dL_dy, m, n, dtable, spt);
  });
  NGP_LAUNCH_CHECK("ngp_hashgrid_bw_params");
  return 0;
}

// dx (N,3) f32 = (dy/dx)^T dL/dy
NGP_API int ngp_hashgrid_bw_input(const float* x, const float* aabb, const float* dL_dy, const void* table, int table_dtype, int n_levels,
                                  int n_features, int log2_hashmap_size, int base_resolution, float per_level_scale,
                                  int64_t n, float* dL_dx, void* stream) {
  if (n <= 0) return 0;
  GridMeta m;
  if (fill_meta(m, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale, aabb))
    return set_error_msg("ngp_hashgrid_bw_input: bad grid config");
  cudaStream_t st = (cudaStream_t)stream;
  NGP_F_DISPATCH(n_features, {
    const unsigned grid = (unsigned)ceil_div(n, 128);        // lane pair per sample
    if (table_dtype == 0) hashgrid_bw_input_kernel<F, float><<<grid, 256, 0, st>>>(x, dL_dy, (const float*)table, m, n, dL_dx);
    else hashgrid_bw_input_kernel<F, __half><<<grid, 256, 0, st>>>(x, dL_dy, (const __half*)table, m, n, dL_dx);
  });
  NGP_LAUNCH_CHECK("ngp_hashgrid_bw_input");
  return 0;
}

// Double backward of ngp_hashgrid_bw_input: g2 = dL/d(dL_dx) (N,3).
//   dtable += d(g2 . dL_dx)/dtable   (skipped when dtable == NULL)
//   d_dL_dy = d(g2 . dL_dx)/d(dL_dy) (skipped when d_dL_dy == NULL; dL_dy may then be NULL too)
NGP_API int ngp_hashgrid_bwbw_input(const float* x, const float* aabb, const float* g2, const float* dL_dy, const void* table,
                                    int table_dtype, int n_levels, int n_features, int log2_hashmap_size,
                                    int base_resolution, float per_level_scale, int64_t n, float* dtable,
                                    float* d_dL_dy, void* stream) {
  if (n <= 0) return 0;
  GridMeta m;
  if (fill_meta(m, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale, aabb))
    return set_error_msg("ngp_hashgrid_bwbw_input: bad grid config");
  cudaStream_t st = (cudaStream_t)stream;
  NGP_F_DISPATCH(n_features, {
    if (d_dL_dy) {
      constexpr int LC = levels_per_thread<F>();
      const unsigned grid = (unsigned)(ceil_div(n, 128) * ceil_div(n_levels, LC));
      set_block_order(m, (size_t)F * (table_dtype == 0 ? 4 : 2), ceil_div(n, 128));
      if (table_dtype == 0) hashgrid_fw_kernel<F, float, false, true><<<grid, 256, 0, st>>>(x, (const float*)table, m, n, d_dL_dy, g2);
      else hashgrid_fw_kernel<F, __half, false, true><<<grid, 256, 0, st>>>(x, (const __half*)table, m, n, d_dL_dy, g2);
    }
    if (dtable && dL_dy) {
      constexpr int LC = scatter_levels_per_thread<F>();
      const unsigned grid = (unsigned)(ceil_div(ceil_div(n, kSPT), 64) * ceil_div(n_levels, LC));
      set_block_order(m, (size_t)F * 4, ceil_div(ceil_div(n, kSPT), 64));
      if (F == 8) hashgrid_bw_params_f8_kernel<false, 1><<<grid, 128, 0, st>>>(x, dL_dy, m, n, dtable, kSPT, g2);
      else hashgrid_bw_params_kernel<F, LC, false, true><<<grid, 128, 0, st>>>(x, dL_dy, m, n, dtable, kSPT, g2);
    }
  });
  NGP_LAUNCH_CHECK("ngp_hashgrid_bwbw_input");
  return 0;
}

// dtable += scatter(w, dL_dy) + scatter(coef(g2), dL_dy_first): BOTH terms of the table gradient of a field evaluated with
// its input gradient (density + normals, models/networks.py:186-196), where dL_dy is the upstream of the encoding itself,
// g2 (N,3) the upstream of the input gradient and dL_dy_first (N, L*F) the first-order upstream the input gradient was
// taken with.  One pass for F = 8 (the two scatters hit the same corners); otherwise the two launches of
// ngp_hashgrid_bw_params + ngp_hashgrid_bwbw_input back to back.  g2 == NULL: the first term only.
NGP_API int ngp_hashgrid_bw_params_dual(const float* x, const float* aabb, const float* dL_dy, const float* g2,
                                        const float* dL_dy_first, int n_levels, int n_features, int log2_hashmap_size,
                                        int base_resolution, float per_level_scale, int64_t n, float* dtable, void* stream) {
  if (n <= 0) return 0;
  if (g2 == nullptr || dL_dy_first == nullptr || n_features != 8) {
    int rc = ngp_hashgrid_bw_params(x, aabb, dL_dy, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale, n, dtable, stream);
    if (rc || g2 == nullptr || dL_dy_first == nullptr) return rc;
    return ngp_hashgrid_bwbw_input(x, aabb, g2, dL_dy_first, nullptr, 0, n_levels, n_features, log2_hashmap_size, base_resolution,
                                   per_level_scale, n, dtable, nullptr, stream);
  }
  GridMeta m;
  if (fill_meta(m, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale, aabb))
    return set_error_msg("ngp_hashgrid_bw_params_dual: bad grid config");
  const unsigned grid = (unsigned)(ceil_div(ceil_div(n, kSPT), 64) * n_levels);
  set_block_order(m, (size_t)8 * 4, ceil_div(ceil_div(n, kSPT), 64));
  hashgrid_bw_params_f8_kernel<false, 2><<<grid, 128, 0, (cudaStream_t)stream>>>(x, dL_dy, m, n, dtable, kSPT, g2, dL_dy_first);
  NGP_LAUNCH_CHECK("ngp_hashgrid_bw_params_dual");
  return 0;
}

// ---- feature-tile / gradient-tile variants (the fused density path of the ngp_pl-shaped field) ---------------------
// Bytes of one 128-sample bf16 feature tile for a grid of n_levels*n_features columns (padded to 16).
NGP_API int64_t ngp_feature_tile_bytes(int n_levels, int n_features) {
  return (int64_t)feat_tile_bytes((n_levels * n_features + 15) / 16 * 16);
}
// y_tiles (ceil(N/128) * ngp_feature_tile_bytes) = bf16(encode(x)) in the MLP's operand-tile layout (see the kernel).
int ngp::hashgrid_fw_tiles_launch(const float* x, const float* aabb, const void* table, int table_dtype, int n_levels,
                                  int n_features, int log2_hashmap_size, int base_resolution, float per_level_scale,
                                  int64_t n, const int32_t* n_dev, void* y_tiles, void* stream) {
  if (n <= 0) return 0;
  GridMeta m;
  if (fill_meta(m, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale, aabb))
    return set_error_msg("ngp_hashgrid_fw_tiles: bad grid config");
  m.n_dev = n_dev;
  cudaStream_t st = (cudaStream_t)stream;
  NGP_F_DISPATCH(n_features, {
    const unsigned grid = (unsigned)(ceil_div(n, 128) * (m.k0p / 8));      // one lane pair per (sample, 16-byte chunk), padded width
    set_block_order(m, (size_t)F * (table_dtype == 0 ? 4 : 2), ceil_div(n, 128));
    if (table_dtype == 0) hashgrid_fw_kernel<F, float, true><<<grid, 256, 0, st>>>(x, (const float*)table, m, n, (float*)y_tiles);
    else hashgrid_fw_kernel<F, __half, true><<<grid, 256, 0, st>>>(x, (const __half*)table, m, n, (float*)y_tiles);
  });
  NGP_LAUNCH_CHECK("ngp_hashgrid_fw_tiles");
  return 0;
}
NGP_API int ngp_hashgrid_fw_tiles(const float* x, const float* aabb, const void* table, int table_dtype, int n_levels,
                                  int n_features, int log2_hashmap_size, int base_resolution, float per_level_scale,
                                  int64_t n, void* y_tiles, void* stream) {
  return ngp::hashgrid_fw_tiles_launch(x, aabb, table, table_dtype, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale,
                                       n, nullptr, y_tiles, stream);
}
// dtable += scatter(dL/dy) with dL/dy in gradient tiles (ceil(N/128) * 128 * k0p floats, see the kernel), levels
// [level_begin, level_end) only.  Launching the scatter as a few level ranges lets the caller hand each finished slice of the
// table gradient (levels are contiguous in the flat table) to the gradient all-reduce while the next range is still being
// scattered (ngp_b200/trainer.py).  level_begin must be a multiple of the kernel's levels-per-lane-pair (4/F, 1 for F >= 4).
NGP_API int ngp_hashgrid_bw_params_tiles_range(const float* x, const float* aabb, const float* dy_tiles, int n_levels,
                                               int n_features, int log2_hashmap_size, int base_resolution,
                                               float per_level_scale, int64_t n, float* dtable, int level_begin, int level_end,
                                               void* stream) {
  if (n <= 0) return 0;
  GridMeta m;
  if (fill_meta(m, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale, aabb))
    return set_error_msg("ngp_hashgrid_bw_params_tiles: bad grid config");
  if (level_begin < 0 || level_end > n_levels || level_begin >= level_end)
    return set_error_msg("ngp_hashgrid_bw_params_tiles_range: need 0 <= level_begin < level_end <= n_levels");
  NGP_F_DISPATCH(n_features, {
    constexpr int LC = scatter_levels_per_thread<F>();
    if (level_begin % LC) return set_error_msg("ngp_hashgrid_bw_params_tiles_range: level_begin must be a multiple of 4/F");
    const int n_chunks = (int)ceil_div(level_end - level_begin, LC);
    const unsigned grid = (unsigned)(ceil_div(ceil_div(n, kSPT), 64) * n_chunks);
    set_block_order(m, (size_t)F * 4, ceil_div(ceil_div(n, kSPT), 64));
    m.chunk0 = level_begin / LC; m.level_end = level_end;
    if (F == 8) hashgrid_bw_params_f8_kernel<true, 0><<<grid, 128, 0, (cudaStream_t)stream>>>(x, dy_tiles, m, n, dtable, kSPT);
    else if (use_pf(m)) hashgrid_bw_params_kernel<F, LC, true, false, true><<<grid, 128, 0, (cudaStream_t)stream>>>(x, dy_tiles, m, n, dtable, kSPT);
    else hashgrid_bw_params_kernel<F, LC, true><<<grid, 128, 0, (cudaStream_t)stream>>>(x, dy_tiles, m, n, dtable, kSPT);
  });
  NGP_LAUNCH_CHECK("ngp_hashgrid_bw_params_tiles_range");
  return 0;
}
NGP_API int ngp_hashgrid_bw_params_tiles(const float* x, const float* aabb, const float* dy_tiles, int n_levels,
                                         int n_features, int log2_hashmap_size, int base_resolution,
                                         float per_level_scale, int64_t n, float* dtable, void* stream) {
  return ngp_hashgrid_bw_params_tiles_range(x, aabb, dy_tiles, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale, n,
                                            dtable, 0, n_levels, stream);
}
