// Operand-tile layout and tcgen05 issue helpers shared by the fused MLP kernels (mlp.cu) and the density net
// (density_net.cu).  See tc05.cuh for the descriptor conventions.
#pragma once
#include "common.cuh"
#include "tc05.cuh"

namespace ngp {
using namespace tc05;

constexpr int kTile = 128;          // samples (rows) per tile = TMEM lanes

// Operand tiles: layout of tc05.cuh with the chunk stride padded by 64 B (keeps 8-byte staging stores
// at the 2-wavefront minimum): byte_offset(r,c) = (c/8)*CS(rows) + r*16 + (c%8)*2.
__host__ __device__ __forceinline__ uint32_t chunk_stride(uint32_t rows) { return rows * 16u + 64u; }
__device__ __forceinline__ uint32_t toff(uint32_t rows, uint32_t r, uint32_t c) {
  return (c >> 3) * chunk_stride(rows) + r * 16u + (c & 7u) * 2u;
}
__device__ __forceinline__ void st_chunk(uint8_t* tile, uint32_t rows, uint32_t r, uint32_t c, const float* v) {
  uint4 q;
  q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]); q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(tile + toff(rows, r, c)) = q;
}
__device__ __forceinline__ void st_chunk_relu(uint8_t* tile, uint32_t rows, uint32_t r, uint32_t c, const float* v) {
  uint4 q;
  q.x = pack_bf16_relu(v[0], v[1]); q.y = pack_bf16_relu(v[2], v[3]); q.z = pack_bf16_relu(v[4], v[5]); q.w = pack_bf16_relu(v[6], v[7]);
  *reinterpret_cast<uint4*>(tile + toff(rows, r, c)) = q;
}
// ReLU backward on packed operands: the chunk holds H = relu(Z) as bf16 (so H > 0 <=> halfword != 0); it is
// overwritten IN PLACE by bf16(dH) where H > 0 and 0 elsewhere — no unpacking of H, no per-element select.
__device__ __forceinline__ uint32_t relu_mask2(uint32_t h2) {
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&h2), z);
}
__device__ __forceinline__ void relu_bw_chunk(uint8_t* tile, uint32_t rows, uint32_t r, uint32_t c, const float* g) {
  uint4* p = reinterpret_cast<uint4*>(tile + toff(rows, r, c));
  const uint4 h = *p;
  uint4 q;
  q.x = pack_bf16(g[0], g[1]) & relu_mask2(h.x); q.y = pack_bf16(g[2], g[3]) & relu_mask2(h.y);
  q.z = pack_bf16(g[4], g[5]) & relu_mask2(h.z); q.w = pack_bf16(g[6], g[7]) & relu_mask2(h.w);
  *p = q;
}
__device__ __forceinline__ void st_quad(uint8_t* tile, uint32_t rows, uint32_t r, uint32_t c, float4 v) {
  uint2 q;
  q.x = pack_bf16(v.x, v.y); q.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(tile + toff(rows, r, c)) = q;
}
__device__ __forceinline__ void ld_chunk(const uint8_t* tile, uint32_t rows, uint32_t r, uint32_t c, float* v) {
  const uint4 q = *reinterpret_cast<const uint4*>(tile + toff(rows, r, c));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; i++) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void st_elem(uint8_t* tile, uint32_t rows, uint32_t r, uint32_t c, float v) {
  *reinterpret_cast<__nv_bfloat16*>(tile + toff(rows, r, c)) = __float2bfloat16_rn(v);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- MMA issue helpers (single thread) ------------------------------------------------------------
// Descriptors are built ONCE per operand and stepped along K by adding to the 14-bit start-address field: shared
// memory ends below 2^18 bytes, so (address >> 4) never carries out of the field and the step is one 32-bit add.
__device__ __forceinline__ uint64_t desc_step(uint64_t d, uint32_t bytes) {
  const uint32_t lo = (uint32_t)d + (bytes >> 4);
  return (d & 0xFFFFFFFF00000000ull) | lo;
}
// D[128 x N] = A[128 x K] (K-major tile, 128 rows) * B[N x K]^T (K-major tile with b_rows rows)
__device__ __forceinline__ void issue_fwd(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t b_rows, int N, int K) {
  const uint32_t id = idesc_bf16(kTile, N, 0, 0);
  const uint64_t da = smem_desc(a_addr, chunk_stride(kTile), 128);
  const uint64_t db = smem_desc(b_addr, chunk_stride(b_rows), 128);
  for (int k = 0; k < K; k += 16)
    mma_bf16(tmem_d, desc_step(da, (k >> 3) * chunk_stride(kTile)), desc_step(db, (k >> 3) * chunk_stride(b_rows)), id, k > 0);
}
// D[128 x N] = dZ[128 x K] (K-major) * W[K x N] where W's tile is [w_rows(K) x N cols]: MN-major B.
__device__ __forceinline__ void issue_dgrad(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, uint32_t w_rows, int N, int K) {
  const uint32_t id = idesc_bf16(kTile, N, 0, 1);
  const uint64_t da = smem_desc(a_addr, chunk_stride(kTile), 128);
  const uint64_t db = smem_desc(w_addr, 128, chunk_stride(w_rows));   // LBO = next 8 rows(k), SBO = next 8 cols(n)
  for (int k = 0; k < K; k += 16)
    mma_bf16(tmem_d, desc_step(da, (k >> 3) * chunk_stride(kTile)), desc_step(db, k * 16), id, k > 0);
}
// D[M x N] += P[128 x M]^T * Q[128 x N]   (both 128-row sample tiles read MN-major, K = samples); the
// accumulator is zero-initialised once per CTA, every slot accumulates into it.
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_d, uint32_t p_addr, uint32_t q_addr, int M, int N) {
  const uint32_t id = idesc_bf16(M, N, 1, 1);
  const uint64_t da = smem_desc(p_addr, 128, chunk_stride(kTile));
  const uint64_t db = smem_desc(q_addr, 128, chunk_stride(kTile));
#pragma unroll
  for (int k = 0; k < kTile; k += 16)
    mma_bf16(tmem_d, desc_step(da, k * 16), desc_step(db, k * 16), id, true);
}


// Row-major fp32 output from a thread-per-row register layout: lane r of a warp holds COLS consecutive columns of row r (what a
// 32x32b TMEM load gives).  Stored directly, every instruction touches 32 different lines (16 B each) — the L1 serves one line
// per clock.  Through a padded per-warp scratch (conflict-free 16-byte phases both ways) each store instruction covers
// 32*4/COLS whole rows of COLS*4 contiguous bytes: 4 (COLS = 32) or 8 (COLS = 16) lines instead of 32.
template <int COLS>
__device__ __forceinline__ void store_rows_coalesced(float* scratch, const float* v, float* gbase, int64_t gstride, int rows_valid, uint32_t lane) {
  constexpr int LD = COLS + 4, LPR = COLS / 4, RPI = 32 / LPR;
#pragma unroll
  for (int j = 0; j < COLS / 4; j++)
    *reinterpret_cast<float4*>(scratch + lane * LD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int rsub = lane / LPR, c4 = (lane % LPR) * 4;
#pragma unroll
  for (int it = 0; it < 32 / RPI; it++) {
    const int row = it * RPI + rsub;
    const float4 x = *reinterpret_cast<const float4*>(scratch + row * LD + c4);
    if (row < rows_valid) __stcs(reinterpret_cast<float4*>(gbase + row * gstride + c4), x);
  }
  __syncwarp();
}

// Wide fp32 rows -> bf16 operand tile, warp-cooperatively: consecutive lanes read consecutive float4 of a row (w/128 lines per
// load instruction per row) instead of one row per thread (32 lines per instruction: the L1 serves one line per clock).  Warp wq of
// the 4 warps that own a 128-row tile stages rows [32 wq, 32 wq + 32).  w: power of two >= 32 floats; col % 4 == 0; 16-byte aligned rows.
__device__ __forceinline__ void stage_rows_coop(const float* __restrict__ seg, int64_t stride, int w, int col, int64_t row0, int64_t n,
                                                uint32_t wq, uint32_t lane, uint8_t* Xs) {
  const int q = w >> 2;                                           // float4 per row
  if (q <= 32) {
    const int sh = __ffs(q) - 1, rpi = 32 >> sh;                  // rows per instruction
    const uint32_t rsub = lane >> sh, f4 = lane & (uint32_t)(q - 1);
    constexpr int B = 8;                                          // rows (instructions) in flight per lane: the loads of a batch overlap
    for (int it = 0; it < 32 / rpi; it += B) {
      float4 v[B];
#pragma unroll
      for (int j = 0; j < B; j++) {
        const uint32_t r = 32u * wq + (uint32_t)(it + j) * rpi + rsub;
        const int64_t g = row0 + r;
        v[j] = ((it + j) * rpi < 32 && g < n) ? __ldg(reinterpret_cast<const float4*>(seg + g * stride) + f4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < B; j++) {
        const uint32_t r = 32u * wq + (uint32_t)(it + j) * rpi + rsub;
        if ((it + j) * rpi < 32) {
          uint2 qq; qq.x = pack_bf16(v[j].x, v[j].y); qq.y = pack_bf16(v[j].z, v[j].w);
          *reinterpret_cast<uint2*>(Xs + toff(kTile, r, col + 4 * f4)) = qq;
        }
      }
    }
  } else {
    for (int it = 0; it < 32; it++) {
      const uint32_t r = 32u * wq + it;
      const int64_t g = row0 + r;
      for (int p0 = 0; p0 < q; p0 += 32) {
        const float4 v = g < n ? __ldg(reinterpret_cast<const float4*>(seg + g * stride) + p0 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint2 qq; qq.x = pack_bf16(v.x, v.y); qq.y = pack_bf16(v.z, v.w);
        *reinterpret_cast<uint2*>(Xs + toff(kTile, r, col + 4 * (p0 + lane))) = qq;
      }
    }
  }
}

}  // namespace ngp
