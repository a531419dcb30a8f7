// Fused small-MLP forward / backward on the 5th-gen tensor cores (tcgen05.mma, accumulators in
// TMEM), one 128-sample tile at a time, activations never leaving shared memory / TMEM between
// layers.
//
// Replaces tcnn.Network({"otype":"CutlassMLP"|"FullyFusedMLP", ...}) as the reference uses it
// (models/networks.py:89-162: rgb_net, norm_pred_header, semantic_header, skybox_rgb_net,
// tonemapper_net_*; bias-free, ReLU hidden, None/Sigmoid output) — the reference runs one CUTLASS
// GEMM launch per layer with every activation round-tripping HBM.  tiny-cuda-nn is not in
// /root/reference; semantics follow SURVEY.md Appendix B.
//
// Parameter layout (flat fp32, row-major (out,in) per layer, layers concatenated):
//   W_0 [width x k0] | W_1.. [width x width] (n_hidden-1 of them) | W_out [n_out_pad16 x width]
//
// Kernel shape: CTA = 256 threads = 8 warps; thread t works on tile row (sample) t%128 and on column
// half t/128 in the TMEM->register epilogues (warp w can read TMEM lanes 32*(w%4)..+31).  Thread 0
// issues every tcgen05.mma and commits to one mbarrier; the next tile's inputs arrive by bulk copy
// (UBLKCP) while the current one is computed; the CTA loops over tiles with stride gridDim.x, and the
// 2-3 co-resident CTAs per SM overlap each other's MMA / epilogue phases.
//   forward : X -> [MMA -> act -> bf16 smem]* -> MMA -> act_out -> global
//   backward: recompute the forward chain (hidden activations stay in smem), then per layer, top
//             down:  wgrad (accumulated in TMEM for the CTA's whole lifetime, flushed once with
//             fp32 atomics) and dgrad, both reading the SAME smem tiles as MN-major operands.
// Operand tiles use the no-swizzle layout documented in tc05.cuh.
#include "common.cuh"
#include "tc05.cuh"
#include <stdio.h>
#include <string.h>

namespace ngp {
using namespace tc05;

constexpr int kTile = 128;          // samples (rows) per tile = TMEM lanes
constexpr int kThreads = 128;       // 4 warps: warp w reads TMEM lanes 32*w..  (256 = two column halves: measured no faster, r01 call 16)
constexpr int kHalves = kThreads / kTile;
constexpr int kMaxSeg = 3;
constexpr int kMaxHidden = 6;

enum Act { kActNone = 0, kActReLU = 1, kActSigmoid = 2, kActExp = 3 };
enum SegKind { kSegPlain = 0, kSegSH4 = 1 };

struct MlpCfg {
  int n_seg;
  int seg_w[kMaxSeg];
  int seg_kind[kMaxSeg];
  int64_t seg_stride[kMaxSeg];   // in floats
  int k0, k0p;                   // input width / padded to 16
  int w, wp;                     // hidden width / padded to 64 or 128
  int nh;                        // hidden layers >= 1
  int no, nop;                   // output width / padded to 16
  int act_h, act_o;
  // shared-memory byte offsets
  uint32_t off_w[kMaxHidden + 1];  // weight tiles: layer 0..nh-1, then output layer at [nh]
  uint32_t off_x, off_h[kMaxHidden], off_dz;
  // raw fp32 landing zone of the next tile (bulk copies): per segment, then dL/dy (backward)
  uint32_t off_raw[kMaxSeg], raw_bytes[kMaxSeg];
  int bulk;                        // set by the host when every segment is contiguous + 16-byte aligned
  uint32_t smem_bytes;
  // TMEM column offsets
  uint32_t tm_cols;                  // allocation (power of two)
  uint32_t tm_wg[kMaxHidden + 1];    // wgrad accumulators (backward only)
  // parameter offsets (floats)
  int64_t p_off[kMaxHidden + 1];
};

struct SegPtrs { const float* p[kMaxSeg]; };
struct SegGrads { float* p[kMaxSeg]; int64_t stride[kMaxSeg]; };

__device__ __forceinline__ float act_apply(int a, float z) {
  switch (a) {
    case kActReLU: return fmaxf(z, 0.f);
    case kActSigmoid: return 1.f / (1.f + __expf(-z));
    case kActExp: return __expf(z);
    default: return z;
  }
}
// derivative expressed through the activation's OUTPUT y (what is kept in shared memory)
__device__ __forceinline__ float act_grad_from_out(int a, float y) {
  switch (a) {
    case kActReLU: return y > 0.f ? 1.f : 0.f;
    case kActSigmoid: return y * (1.f - y);
    case kActExp: return y;
    default: return 1.f;
  }
}

// Array forms: ONE dispatch on the (launch-constant) activation id, then straight-line code — the scalar
// forms above inside an unrolled loop compile to a jump table per element (BRX + range checks were 15 %
// of all issued instructions in the first ncu capture of mlp_bw_kernel).
template <int N> __device__ __forceinline__ void act_apply_arr(int a, float* v) {
  if (a == kActReLU) {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = fmaxf(v[i], 0.f);
  } else if (a == kActSigmoid) {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = 1.f / (1.f + __expf(-v[i]));
  } else if (a == kActExp) {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = __expf(v[i]);
  }
}
// v[i] *= act'(.) given the activation OUTPUT y[i]
template <int N> __device__ __forceinline__ void act_grad_mul_arr(int a, float* v, const float* y) {
  if (a == kActReLU) {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = y[i] > 0.f ? v[i] : 0.f;
  } else if (a == kActSigmoid) {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] *= y[i] * (1.f - y[i]);
  } else if (a == kActExp) {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] *= y[i];
  }
}

__device__ __forceinline__ void sh4(float x, float y, float z, float* o) {
  const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  o[0] = 0.28209479177387814f;
  o[1] = -0.48860251190291987f * y; o[2] = 0.48860251190291987f * z; o[3] = -0.48860251190291987f * x;
  o[4] = 1.0925484305920792f * xy; o[5] = -1.0925484305920792f * yz;
  o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = -1.0925484305920792f * xz; o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2); o[10] = 2.8906114426405538f * xy * z;
  o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2); o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2); o[14] = 1.4453057213202769f * z * (x2 - y2);
  o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

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

// All 128 threads stage rows [row0, row0+128) of every input segment as bf16 into the X tile.
// Plain segments are read as a flat run of float4 units (consecutive threads -> consecutive 16-byte
// units of the same row, then the next row: fully coalesced when stride == width), converted and
// written with 8-byte stores; SH segments are evaluated by the row's owner thread.
__device__ __forceinline__ void stage_input(const MlpCfg& c, const SegPtrs& in, int64_t row0, int64_t n, uint32_t t,
                                            uint8_t* Xs) {
  int col = 0;
  for (int s = 0; s < c.n_seg; s++) {
    const int w = c.seg_w[s];
    if (c.seg_kind[s] == kSegSH4) {
      if (t >= kTile) { col += w; continue; }            // one thread per row evaluates the harmonics
      const int64_t row = row0 + t;
      float o[16];
      if (row < n) {
        const float* d = in.p[s] + row * c.seg_stride[s];
        const float dx = __ldg(d), dy = __ldg(d + 1), dz = __ldg(d + 2);
        const float inv = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-6f);  // F.normalize(eps=1e-6), networks.py:221
        sh4(dx * inv, dy * inv, dz * inv, o);
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) o[i] = 0.f;
      }
      if ((col & 7) == 0) { st_chunk(Xs, kTile, t, col, o); st_chunk(Xs, kTile, t, col + 8, o + 8); }
      else { for (int i = 0; i < 16; i++) st_elem(Xs, kTile, t, col + i, o[i]); }
    } else {
      const float* base = in.p[s];
      const int64_t stride = c.seg_stride[s];
      const bool vec = ((col & 3) == 0) && ((w & 3) == 0) && ((stride & 3) == 0) && ((((uintptr_t)base) & 15) == 0);
      if (vec) {
        const int upr = w >> 2;                       // float4 units per row
        for (int u = t; u < kTile * upr; u += kThreads) {
          const int r = u / upr, c4 = u - r * upr;
          const int64_t row = row0 + r;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < n) v = __ldg(reinterpret_cast<const float4*>(base + row * stride) + c4);
          st_quad(Xs, kTile, r, col + c4 * 4, v);
        }
      } else {
        for (int u = t; u < kTile * w; u += kThreads) {
          const int r = u / w, cc = u - r * w;
          const int64_t row = row0 + r;
          st_elem(Xs, kTile, r, col + cc, row < n ? __ldg(base + row * stride + cc) : 0.f);
        }
      }
    }
    col += w;
  }
  if (col < c.k0p) {
    const int padw = c.k0p - col;
    for (int u = t; u < kTile * padw; u += kThreads) st_elem(Xs, kTile, u / padw, col + u % padw, 0.f);
  }
}

// ---- next-tile prefetch: one thread arms `full` and fires one bulk copy per segment (+ dL/dy) ------
__device__ __forceinline__ bool tile_is_bulk(const MlpCfg& c, int64_t tile, int64_t n) {
  return c.bulk && (tile + 1) * kTile <= n;
}
__device__ __forceinline__ void issue_prefetch(const MlpCfg& c, const SegPtrs& in, int64_t tile, uint8_t* smem,
                                               uint64_t* full) {
  const int64_t row0 = tile * kTile;
  uint32_t total = 0;
  for (int s = 0; s < c.n_seg; s++) total += c.raw_bytes[s];
  mbar_expect_tx(full, total);
  for (int s = 0; s < c.n_seg; s++)
    bulk_g2s(smem + c.off_raw[s], in.p[s] + row0 * c.seg_stride[s], c.raw_bytes[s], full);
}
// Raw fp32 landing zone -> bf16 operand tile.  Thread r converts row r; in step i it reads float4
// #((i + r) mod upr) of its row, so the 32 lanes of a warp hit distinct bank groups even though rows
// are w*4 bytes apart (a straight per-row walk would be a 32-way conflict), and no index division is
// needed.
__device__ __forceinline__ void convert_raw(const MlpCfg& c, const uint8_t* smem, uint32_t t, uint8_t* Xs) {
  if (t >= kTile) return;
  int col = 0;
  for (int s = 0; s < c.n_seg; s++) {
    const int w = c.seg_w[s];
    const float* raw = reinterpret_cast<const float*>(smem + c.off_raw[s]);
    if (c.seg_kind[s] == kSegSH4) {
      const float dx = raw[3 * t], dy = raw[3 * t + 1], dz = raw[3 * t + 2];
      const float inv = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-6f);
      float o[16];
      sh4(dx * inv, dy * inv, dz * inv, o);
      if ((col & 7) == 0) { st_chunk(Xs, kTile, t, col, o); st_chunk(Xs, kTile, t, col + 8, o + 8); }
      else { for (int i = 0; i < 16; i++) st_elem(Xs, kTile, t, col + i, o[i]); }
    } else if (((col | w) & 3) == 0) {
      const int upr = w >> 2;
      const float4* rowp = reinterpret_cast<const float4*>(raw) + (size_t)t * upr;
      int c4 = (int)(t % (uint32_t)upr);
#pragma unroll 4
      for (int i = 0; i < upr; i++) {
        st_quad(Xs, kTile, t, col + c4 * 4, rowp[c4]);
        c4 = (c4 + 1 == upr) ? 0 : c4 + 1;
      }
    } else {
      for (int cc = 0; cc < w; cc++) st_elem(Xs, kTile, t, col + cc, raw[t * w + cc]);
    }
    col += w;
  }
}
// All threads: zero the padding columns [k0, k0p) of the X tile once (segments never touch them).
__device__ __forceinline__ void zero_pad_cols(const MlpCfg& c, uint32_t t, uint8_t* Xs) {
  if (c.k0 < c.k0p) {
    const int padw = c.k0p - c.k0;
    for (int u = t; u < kTile * padw; u += kThreads) st_elem(Xs, kTile, u / padw, c.k0 + u % padw, 0.f);
  }
}

// All threads: convert the fp32 parameter vector to bf16 operand tiles (zero padded).
__device__ __forceinline__ void stage_weights(const MlpCfg& c, const float* __restrict__ params, uint8_t* smem) {
  for (int l = 0; l <= c.nh; l++) {
    const int rows_t = (l == c.nh) ? c.nop : c.wp;           // tile rows
    const int cols_t = (l == 0) ? c.k0p : c.wp;              // tile cols
    const int rows = (l == c.nh) ? c.nop : c.w;              // rows present in params
    const int cols = (l == 0) ? c.k0 : c.w;
    const float* W = params + c.p_off[l];
    uint8_t* tile = smem + c.off_w[l];
    for (int i = threadIdx.x; i < rows_t * cols_t; i += blockDim.x) {
      const int r = i / cols_t, cc = i % cols_t;
      const float v = (r < rows && cc < cols) ? __ldg(W + (int64_t)r * cols + cc) : 0.f;
      st_elem(tile, rows_t, r, cc, v);
    }
  }
}

// ---- MMA issue helpers (single thread) ------------------------------------------------------------
// D[128 x N] = A[128 x K] (K-major tile, 128 rows) * B[N x K]^T (K-major tile with b_rows rows)
__device__ __forceinline__ void issue_fwd(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t b_rows, int N, int K) {
  const uint32_t id = idesc_bf16(kTile, N, 0, 0);
  for (int k = 0; k < K; k += 16) {
    const uint64_t da = smem_desc(a_addr + (k >> 3) * chunk_stride(kTile), chunk_stride(kTile), 128);
    const uint64_t db = smem_desc(b_addr + (k >> 3) * chunk_stride(b_rows), chunk_stride(b_rows), 128);
    mma_bf16(tmem_d, da, db, id, k > 0);
  }
}
// D[128 x N] = dZ[128 x K] (K-major) * W[K x N] where W's tile is [w_rows(K) x N cols]: MN-major B.
__device__ __forceinline__ void issue_dgrad(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, uint32_t w_rows, int N, int K) {
  const uint32_t id = idesc_bf16(kTile, N, 0, 1);
  for (int k = 0; k < K; k += 16) {
    const uint64_t da = smem_desc(a_addr + (k >> 3) * chunk_stride(kTile), chunk_stride(kTile), 128);
    const uint64_t db = smem_desc(w_addr + k * 16, 128, chunk_stride(w_rows));   // LBO = next 8 rows(k), SBO = next 8 cols(n)
    mma_bf16(tmem_d, da, db, id, k > 0);
  }
}
// D[M x N] (+)= P[128 x M]^T * Q[128 x N]   (both 128-row sample tiles read MN-major, K = samples)
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_d, uint32_t p_addr, uint32_t q_addr, int M, int N, bool accumulate) {
  const uint32_t id = idesc_bf16(M, N, 1, 1);
  for (int k = 0; k < kTile; k += 16) {
    const uint64_t da = smem_desc(p_addr + k * 16, 128, chunk_stride(kTile));
    const uint64_t db = smem_desc(q_addr + k * 16, 128, chunk_stride(kTile));
    mma_bf16(tmem_d, da, db, id, accumulate || k > 0);
  }
}

struct CtaCtx {
  uint64_t* bar; uint64_t* full; uint32_t* tmem_slot; uint32_t tmem; uint32_t phase; uint32_t fphase;
};

__device__ __forceinline__ void cta_setup(CtaCtx& x, uint8_t* smem, uint32_t tm_cols) {
  x.bar = reinterpret_cast<uint64_t*>(smem);
  x.full = reinterpret_cast<uint64_t*>(smem + 8);
  x.tmem_slot = reinterpret_cast<uint32_t*>(smem + 16);
  x.phase = 0; x.fphase = 0;
  if (threadIdx.x == 0) { mbar_init(x.bar, 1); mbar_init(x.full, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { __syncwarp(); tmem_alloc(x.tmem_slot, tm_cols); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  x.tmem = *x.tmem_slot;
}
__device__ __forceinline__ void cta_teardown(CtaCtx& x, uint32_t tm_cols) {
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(x.tmem, tm_cols);
}
// every thread: wait until all MMAs committed so far have finished.
// NOTE: mbarrier.try_wait parks the WARP; the single-thread issue blocks are therefore followed by
// __syncwarp() so that lane 0 never has work left when its sibling lanes start waiting.
__device__ __forceinline__ void wait_mma(CtaCtx& x) {
  mbar_wait(x.bar, x.phase);
  x.phase ^= 1;
  fence_after_sync();
}
// every thread: smem written by this thread is ready for the tensor core, TMEM reads are done
__device__ __forceinline__ void publish() {
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
}

// =============================================================================== forward kernel
// Thread geometry shared by both kernels.
struct Lane {
  uint32_t t, row, half, trow;      // thread id, tile row, column half (0/1), TMEM address of this warp's lanes
};
__device__ __forceinline__ Lane make_lane(uint32_t tmem) {
  Lane L;
  L.t = threadIdx.x; L.row = L.t & (kTile - 1); L.half = L.t >> 7;
  L.trow = tmem + ((((L.t >> 5) & 3u) * 32u) << 16);
  return L;
}
// hidden-layer epilogue: this thread's share of the wp accumulator columns -> act -> bf16 tile row.
// Columns are fetched 64 at a time (two x32 TMEM loads in flight, ONE wait).
__device__ __forceinline__ void epilogue_hidden(const MlpCfg& c, const Lane& L, uint8_t* H) {
  const int cw = c.wp / kHalves, cb = (int)L.half * cw;
  for (int c0 = cb; c0 < cb + cw; c0 += 64) {
    uint32_t r0[32], r1[32];
    const bool two = c0 + 32 < cb + cw;
    tmem_ld32_nowait(L.trow + c0, r0);
    if (two) tmem_ld32_nowait(L.trow + c0 + 32, r1);
    tmem_wait_ld();
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r0[i]);
    act_apply_arr<32>(c.act_h, v);
#pragma unroll
    for (int q = 0; q < 4; q++) st_chunk(H, kTile, L.row, c0 + 8 * q, v + 8 * q);
    if (two) {
#pragma unroll
      for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r1[i]);
      act_apply_arr<32>(c.act_h, v);
#pragma unroll
      for (int q = 0; q < 4; q++) st_chunk(H, kTile, L.row, c0 + 32 + 8 * q, v + 8 * q);
    }
  }
}
// 16 consecutive columns [c0, c0+16) of one row -> global, clipped to [0, ncols); 16-byte stores when the
// destination allows it (each thread owns a contiguous 64-byte run of its row)
__device__ __forceinline__ void store_row16(float* dst_row, int c0, int ncols, bool vec, const float* v) {
  if (vec && c0 + 16 <= ncols) {
#pragma unroll
    for (int q = 0; q < 4; q++)
      *reinterpret_cast<float4*>(dst_row + c0 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; i++) if (c0 + i < ncols) dst_row[c0 + i] = v[i];
  }
}

// aux_exp (optional): aux_exp[row] = exp(z_out[row][0]) — the density head of the ngp_pl-shaped field
// (sigma = TruncExp(h[:,0])) produced by the same epilogue instead of a strided select + exp pass.
__global__ void __launch_bounds__(kThreads) mlp_fw_kernel(MlpCfg c, SegPtrs in, const float* __restrict__ params, int64_t n,
                                                          float* __restrict__ out, int64_t out_stride,
                                                          float* __restrict__ aux_exp) {
  extern __shared__ __align__(128) uint8_t smem[];
  CtaCtx cx;
  cta_setup(cx, smem, c.tm_cols);
  stage_weights(c, params, smem);
  const Lane L = make_lane(cx.tmem);
  const uint32_t t = L.t;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t tacc = cx.tmem;                               // accumulator columns [0, max(wp,nop))
  uint8_t* Xs = smem + c.off_x;
  uint8_t* Hs = smem + c.off_h[0];
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const bool out_vec = ((out_stride & 3) == 0) && ((((uintptr_t)out) & 15) == 0);

  zero_pad_cols(c, t, Xs);
  if (t == 0 && (int64_t)blockIdx.x < n_tiles && tile_is_bulk(c, blockIdx.x, n)) issue_prefetch(c, in, blockIdx.x, smem, cx.full);
  __syncwarp();
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTile;
    const int64_t row = row0 + L.row;
    if (tile_is_bulk(c, tile, n)) { mbar_wait(cx.full, cx.fphase); cx.fphase ^= 1; convert_raw(c, smem, t, Xs); }
    else stage_input(c, in, row0, n, t, Xs);
    publish();
    {   // the landing zone is free again: fetch the next tile while this one is computed
      const int64_t nxt = tile + gridDim.x;
      if (t == 0 && nxt < n_tiles && tile_is_bulk(c, nxt, n)) issue_prefetch(c, in, nxt, smem, cx.full);
      __syncwarp();
    }
    for (int l = 0; l <= c.nh; l++) {
      const bool last = l == c.nh;
      const int N = last ? c.nop : c.wp;
      const int K = l == 0 ? c.k0p : c.wp;
      if (t == 0) {
        fence_after_sync();
        issue_fwd(tacc, sbase + (l == 0 ? c.off_x : c.off_h[0]), sbase + c.off_w[l], (uint32_t)N, N, K);
        mma_commit(cx.bar);
      }
      __syncwarp();
      wait_mma(cx);
      if (!last) {
        epilogue_hidden(c, L, Hs);
        publish();
      } else {
        for (int c0 = 16 * (int)L.half; c0 < c.nop; c0 += 16 * kHalves) {
          float v[16];
          tmem_ld16(L.trow + c0, v);
          if (row < n) {
            if (aux_exp && c0 == 0) aux_exp[row] = __expf(v[0]);
            act_apply_arr<16>(c.act_o, v);
            store_row16(out + row * out_stride, c0, c.no, out_vec, v);
          }
        }
        fence_before_sync();
        __syncthreads();                 // all TMEM reads of this tile are done before the next tile's first MMA
      }
    }
  }
  cta_teardown(cx, c.tm_cols);
}

// =============================================================================== backward kernel
__global__ void __launch_bounds__(kThreads) mlp_bw_kernel(MlpCfg c, SegPtrs in, const float* __restrict__ params, int64_t n,
                                                          const float* __restrict__ dout, int64_t dout_stride,
                                                          float* __restrict__ dparams, SegGrads dseg,
                                                          const float* __restrict__ d_aux_exp) {
  extern __shared__ __align__(128) uint8_t smem[];
  CtaCtx cx;
  cta_setup(cx, smem, c.tm_cols);
  stage_weights(c, params, smem);
  const Lane L = make_lane(cx.tmem);
  const uint32_t t = L.t;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t tacc = cx.tmem;
  uint8_t* Xs = smem + c.off_x;
  uint8_t* dZ = smem + c.off_dz;
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  bool have_wgrad = false;
  bool want_dx = false;
  for (int s = 0; s < c.n_seg; s++) want_dx |= dseg.p[s] != nullptr;
  const bool dout_vec = ((dout_stride & 3) == 0) && ((((uintptr_t)dout) & 15) == 0);

  zero_pad_cols(c, t, Xs);
  if (t == 0 && (int64_t)blockIdx.x < n_tiles && tile_is_bulk(c, blockIdx.x, n)) issue_prefetch(c, in, blockIdx.x, smem, cx.full);
  __syncwarp();
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTile;
    const int64_t row = row0 + L.row;
    const bool valid = row < n;
    if (tile_is_bulk(c, tile, n)) { mbar_wait(cx.full, cx.fphase); cx.fphase ^= 1; convert_raw(c, smem, t, Xs); }
    else stage_input(c, in, row0, n, t, Xs);
    publish();
    {
      const int64_t nxt = tile + gridDim.x;
      if (t == 0 && nxt < n_tiles && tile_is_bulk(c, nxt, n)) issue_prefetch(c, in, nxt, smem, cx.full);
      __syncwarp();
    }
    // this thread's first 16 columns of dL/dy are requested NOW and consumed after the forward recompute,
    // so their HBM latency hides behind the MMA phases
    float dreg[16];
    const int d0 = 16 * (int)L.half;
    {
      const float* drow = dout + row * dout_stride;
      if (valid && dout_vec && d0 + 16 <= c.no) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(drow + d0) + q);
          dreg[4 * q] = g4.x; dreg[4 * q + 1] = g4.y; dreg[4 * q + 2] = g4.z; dreg[4 * q + 3] = g4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) dreg[i] = (valid && d0 + i < c.no) ? __ldg(drow + d0 + i) : 0.f;
      }
    }
    const float daux = (d_aux_exp && valid && L.half == 0) ? __ldg(d_aux_exp + row) : 0.f;
    // ---- recompute the forward chain; H_{l+1} kept in smem (post-activation, bf16)
    for (int l = 0; l < c.nh; l++) {
      if (t == 0) {
        fence_after_sync();
        issue_fwd(tacc, sbase + (l == 0 ? c.off_x : c.off_h[l - 1]), sbase + c.off_w[l], (uint32_t)c.wp, c.wp, l == 0 ? c.k0p : c.wp);
        mma_commit(cx.bar);
      }
      __syncwarp();
      wait_mma(cx);
      epilogue_hidden(c, L, smem + c.off_h[l]);
      publish();
    }
    // ---- output layer pre-activation -> dZ_out = dL/dy * act_o'(z)
    if (t == 0) {
      fence_after_sync();
      issue_fwd(tacc, sbase + c.off_h[c.nh - 1], sbase + c.off_w[c.nh], (uint32_t)c.nop, c.nop, c.wp);
      mma_commit(cx.bar);
    }
    __syncwarp();
    wait_mma(cx);
    for (int c0 = d0; c0 < c.nop; c0 += 16 * kHalves) {
      float v[16];
      tmem_ld16(L.trow + c0, v);
      if (c0 != d0) {                      // n_out > 16*kHalves: later column groups are read at use
#pragma unroll
        for (int i = 0; i < 16; i++) dreg[i] = (valid && c0 + i < c.no) ? __ldg(dout + row * dout_stride + c0 + i) : 0.f;
      }
      const float z0 = v[0];
      float g[16];
#pragma unroll
      for (int i = 0; i < 16; i++) g[i] = dreg[i];
      act_apply_arr<16>(c.act_o, v);                 // v = y
      act_grad_mul_arr<16>(c.act_o, g, v);           // g = dL/dy * act'(z)
      // TruncExp backward of the density head: + dL/dsigma * exp(clamp(z0, -7, 7))  (custom_functions.py:211)
      if (c0 == 0 && d_aux_exp) g[0] = fmaf(daux, __expf(fminf(fmaxf(z0, -7.f), 7.f)), g[0]);
#pragma unroll
      for (int i = 0; i < 16; i++) v[i] = (c0 + i < c.no) ? g[i] : 0.f;
      st_chunk(dZ, kTile, L.row, c0, v);
      st_chunk(dZ, kTile, L.row, c0 + 8, v + 8);
    }
    publish();
    // ---- top-down: wgrad + dgrad per layer.  dZ_out lives in its own (narrow) tile; every hidden dZ_l is
    // written IN PLACE over H_{l} (the thread that reads a 16-byte chunk for the activation mask is the one
    // that overwrites it), so no wp-wide gradient tile is needed.
    for (int l = c.nh; l >= 0; l--) {
      const bool is_out = l == c.nh;
      const int Nz = is_out ? c.nop : c.wp;                       // width of dZ_l
      const int Kin = l == 0 ? c.k0p : c.wp;                      // width of the layer's input
      const uint32_t ain = sbase + (l == 0 ? c.off_x : c.off_h[l - 1]);
      const uint32_t dza = sbase + (is_out ? c.off_dz : c.off_h[l]);
      const bool need_dgrad = l > 0 || want_dx;
      if (t == 0) {
        fence_after_sync();
        if (is_out) issue_wgrad(cx.tmem + c.tm_wg[l], ain, dza, c.wp, c.nop, have_wgrad);   // D^T[in x out]
        else issue_wgrad(cx.tmem + c.tm_wg[l], dza, ain, c.wp, Kin, have_wgrad);            // D[out x in]
        if (need_dgrad) issue_dgrad(tacc, dza, sbase + c.off_w[l], (uint32_t)Nz, Kin, Nz);
        mma_commit(cx.bar);
      }
      __syncwarp();
      wait_mma(cx);
      if (l > 0) {
        uint8_t* H = smem + c.off_h[l - 1];
        const int cw = c.wp / kHalves, cb = (int)L.half * cw;
        for (int c0 = cb; c0 < cb + cw; c0 += 32) {
          float v[32], h[32];
          tmem_ld32(L.trow + c0, v);
#pragma unroll
          for (int q = 0; q < 4; q++) ld_chunk(H, kTile, L.row, c0 + 8 * q, h + 8 * q);
          act_grad_mul_arr<32>(c.act_h, v, h);
#pragma unroll
          for (int q = 0; q < 4; q++) st_chunk(H, kTile, L.row, c0 + 8 * q, v + 8 * q);
        }
        publish();
      } else {
        if (want_dx) {
          // input gradient straight from TMEM to global: each thread owns 64-byte runs of its row
          for (int c0 = 16 * (int)L.half; c0 < c.k0; c0 += 16 * kHalves) {
            float v[16];
            tmem_ld16(L.trow + c0, v);
            if (valid) {
              int col = 0;
              for (int s = 0; s < c.n_seg; s++) {
                const int w = c.seg_w[s];
                float* dst = dseg.p[s];
                if (dst && col < c0 + 16 && col + w > c0) {
                  float* drow = dst + row * dseg.stride[s];
                  const bool vec = ((dseg.stride[s] & 3) == 0) && ((col & 3) == 0) && ((w & 3) == 0) && ((((uintptr_t)dst) & 15) == 0);
#pragma unroll
                  for (int q = 0; q < 4; q++) {
                    const int cc = c0 + 4 * q;
                    if (vec && cc >= col && cc + 4 <= col + w) {
                      *reinterpret_cast<float4*>(drow + (cc - col)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    } else {
#pragma unroll
                      for (int i = 0; i < 4; i++) if (cc + i >= col && cc + i < col + w) drow[cc + i - col] = v[4 * q + i];
                    }
                  }
                }
                col += w;
              }
            }
          }
        }
        publish();
      }
    }
    have_wgrad = true;
  }

  // ---- flush the TMEM-resident weight gradients with fp32 atomics
  if (have_wgrad) {
    fence_after_sync();
    // accumulator row m of an M=128 tile sits in TMEM lane m; of an M=64 tile in lane (m%16)+32*(m/16)
    const bool m128 = c.wp == 128;
    const uint32_t lane = t & 31, quad = (t >> 5) & 3;
    const int m = m128 ? (int)L.row : (lane < 16 ? (int)(quad * 16 + lane) : -1);
    for (int l = 0; l <= c.nh; l++) {
      const bool is_out = l == c.nh;
      const int ncols = is_out ? c.nop : (l == 0 ? c.k0p : c.wp);
      const int in_true = l == 0 ? c.k0 : c.w;
      for (int c0 = 16 * (int)L.half; c0 < ncols; c0 += 16 * kHalves) {
        float v[16];
        tmem_ld16(L.trow + c.tm_wg[l] + c0, v);
        if (m >= 0 && m < c.w) {
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const int cc = c0 + i;
            if (is_out) { if (cc < c.no) atomicAdd(dparams + c.p_off[l] + (int64_t)cc * c.w + m, v[i]); }   // D^T[in=m][out=cc]
            else if (cc < in_true) atomicAdd(dparams + c.p_off[l] + (int64_t)m * in_true + cc, v[i]);       // D[out=m][in=cc]
          }
        }
      }
    }
  }
  cta_teardown(cx, c.tm_cols);
}

static uint32_t next_pow2_cols(uint32_t x) { uint32_t p = 32; while (p < x) p <<= 1; return p; }

// Fills the derived fields of cfg; returns 0 or a negative error.
static int finalize_cfg(MlpCfg& c, bool backward) {
  if (c.n_seg < 1 || c.n_seg > kMaxSeg) return -1;
  c.k0 = 0;
  for (int s = 0; s < c.n_seg; s++) {
    if (c.seg_kind[s] == kSegSH4 && c.seg_w[s] != 16) return -2;
    c.k0 += c.seg_w[s];
  }
  c.k0p = (c.k0 + 15) / 16 * 16;
  if (c.w < 1 || c.w > 128 || c.nh < 1 || c.nh > kMaxHidden || c.no < 1 || c.no > 128 || c.k0p > 256) return -3;
  c.wp = c.w <= 64 ? 64 : 128;
  c.nop = (c.no + 15) / 16 * 16;
  auto al = [](uint32_t x) { return (x + 127u) / 128u * 128u; };
  auto tile_bytes = [](int rows, int cols) { return (uint32_t)(cols / 8) * chunk_stride((uint32_t)rows); };
  uint32_t off = 128;  // mbarrier + tmem slot
  int64_t poff = 0;
  for (int l = 0; l <= c.nh; l++) {
    const int rows_t = l == c.nh ? c.nop : c.wp, cols_t = l == 0 ? c.k0p : c.wp;
    c.off_w[l] = off; off += al(tile_bytes(rows_t, cols_t));
    c.p_off[l] = poff;
    poff += (int64_t)(l == c.nh ? c.nop : c.w) * (l == 0 ? c.k0 : c.w);
  }
  c.off_x = off; off += al(tile_bytes(kTile, c.k0p));
  const int n_h = backward ? c.nh : 1;
  for (int l = 0; l < n_h; l++) { c.off_h[l] = off; off += al(tile_bytes(kTile, c.wp)); }
  for (int l = n_h; l < kMaxHidden; l++) c.off_h[l] = c.off_h[0];
  c.off_dz = off;
  if (backward) off += al(tile_bytes(kTile, c.nop));      // dZ_out only; hidden dZ_l overwrite H_l in place
  // raw landing zone (optional: dropped when it does not fit)
  if (c.bulk) {
    uint32_t o2 = off;
    for (int s = 0; s < c.n_seg; s++) {
      c.raw_bytes[s] = (uint32_t)(kTile * (c.seg_kind[s] == kSegSH4 ? 3 : c.seg_w[s]) * 4);
      c.off_raw[s] = o2; o2 += al(c.raw_bytes[s]);
    }
    if (o2 <= 227 * 1024) off = o2; else c.bulk = 0;
  }
  c.smem_bytes = off;
  uint32_t cols = (uint32_t)c.wp;
  if ((uint32_t)c.nop > cols) cols = c.nop;
  if (backward && (uint32_t)c.k0p > cols) cols = c.k0p;
  if (backward) {
    for (int l = 0; l <= c.nh; l++) {
      c.tm_wg[l] = cols;
      cols += l == c.nh ? c.nop : (l == 0 ? c.k0p : c.wp);
    }
  }
  if (cols > 512) return -4;
  c.tm_cols = next_pow2_cols(cols);
  if (c.smem_bytes > 227 * 1024) return -5;
  return 0;
}

static int build_cfg(MlpCfg& c, int n_seg, const float* const* seg_ptr, const int* seg_w, const int* seg_kind,
                     const int64_t* seg_stride, int width, int n_hidden, int n_out, int act_hidden, int act_out,
                     bool backward, const float* dout, int64_t dout_stride) {
  memset(&c, 0, sizeof(c));
  c.n_seg = n_seg;
  for (int s = 0; s < n_seg && s < kMaxSeg; s++) { c.seg_w[s] = seg_w[s]; c.seg_kind[s] = seg_kind[s]; c.seg_stride[s] = seg_stride[s]; }
  c.w = width; c.nh = n_hidden; c.no = n_out; c.act_h = act_hidden; c.act_o = act_out;
  c.bulk = 1;
  for (int s = 0; s < n_seg && s < kMaxSeg; s++) {
    const int raw_w = seg_kind[s] == kSegSH4 ? 3 : seg_w[s];
    if (seg_stride[s] != raw_w || (((uintptr_t)seg_ptr[s]) & 15) != 0) c.bulk = 0;
  }
  (void)dout; (void)dout_stride;
  return finalize_cfg(c, backward);
}

// Persistent grid: SMs x co-resident CTAs.  Residency is limited by shared memory (227 KB/SM, +1 KB
// driver reservation per CTA), TMEM columns (512/SM) and 16 warps' worth of registers; it is computed
// here directly (the kernels' per-tile latency chain is hidden ONLY by co-resident CTAs, so a
// too-small answer is a 3x slowdown, not a detail).
static int launch_grid(const void* fn, const MlpCfg& c, int64_t n) {
  int occ = (int)((227 * 1024) / (c.smem_bytes + 1024));
  const int by_tmem = 512 / (int)c.tm_cols;
  if (occ > by_tmem) occ = by_tmem;
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, fn) == cudaSuccess && fa.numRegs > 0) {
    const int regs = (fa.numRegs + 7) / 8 * 8;
    const int by_regs = 65536 / (regs * kThreads);
    if (occ > by_regs) occ = by_regs;
  }
  if (occ > 8) occ = 8;
  if (occ < 1) occ = 1;
  int64_t g = (int64_t)kSMs * occ;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (g > tiles) g = tiles;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace ngp

using namespace ngp;

// Number of fp32 parameters of the MLP (layout in the header of this file).
NGP_API int64_t ngp_mlp_param_count(int n_input, int width, int n_hidden, int n_out) {
  if (n_input < 1 || width < 1 || n_hidden < 1 || n_out < 1) return -1;
  const int64_t nop = (n_out + 15) / 16 * 16;
  return (int64_t)width * n_input + (int64_t)(n_hidden - 1) * width * width + nop * width;
}

// out (N, n_out) = MLP(cat(segments)).  Segment kinds: 0 = fp32 rows of seg_width floats at
// seg_ptr + row*seg_stride; 1 = degree-4 SH of the normalised (N,3) direction at seg_ptr (width 16).
// Activations: 0 none, 1 ReLU, 2 sigmoid, 3 exp.  aux_exp_out (optional, N floats) = exp(out[:,0]) taken
// before the output activation (the ngp_pl density head).
NGP_API int ngp_mlp_fw(int n_seg, const float* const* seg_ptr, const int* seg_width, const int* seg_kind,
                       const int64_t* seg_stride, const float* params, int width, int n_hidden, int n_out,
                       int act_hidden, int act_out, int64_t n, float* out, int64_t out_stride, float* aux_exp_out,
                       void* stream) {
  if (n <= 0) return 0;
  MlpCfg c;
  const int rc = build_cfg(c, n_seg, seg_ptr, seg_width, seg_kind, seg_stride, width, n_hidden, n_out, act_hidden, act_out, false, nullptr, 0);
  if (rc) { char b[128]; snprintf(b, sizeof b, "ngp_mlp_fw: unsupported MLP shape (code %d)", rc); return set_error_msg(b); }
  SegPtrs in; for (int s = 0; s < kMaxSeg; s++) in.p[s] = s < n_seg ? seg_ptr[s] : nullptr;
  cudaError_t e = cudaFuncSetAttribute(mlp_fw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem_bytes);
  if (e != cudaSuccess) return set_error(e, "ngp_mlp_fw/attr");
  const int grid = launch_grid((const void*)mlp_fw_kernel, c, n);
  mlp_fw_kernel<<<grid, kThreads, c.smem_bytes, (cudaStream_t)stream>>>(c, in, params, n, out, out_stride, aux_exp_out);
  NGP_LAUNCH_CHECK("ngp_mlp_fw");
  return 0;
}

// dparams (+=, fp32 atomics; caller zeroes) and optional per-segment input gradients
// dseg_ptr[s] (N, seg_width[s]) with row stride dseg_stride[s] (NULL = not needed; SH segments
// never receive one).  dL_dout is (N, n_out) with row stride dout_stride.
NGP_API int ngp_mlp_bw(int n_seg, const float* const* seg_ptr, const int* seg_width, const int* seg_kind,
                       const int64_t* seg_stride, const float* params, int width, int n_hidden, int n_out,
                       int act_hidden, int act_out, int64_t n, const float* dL_dout, int64_t dout_stride,
                       float* dparams, float* const* dseg_ptr, const int64_t* dseg_stride, const float* dL_daux_exp,
                       void* stream) {
  if (n <= 0) return 0;
  MlpCfg c;
  const int rc = build_cfg(c, n_seg, seg_ptr, seg_width, seg_kind, seg_stride, width, n_hidden, n_out, act_hidden, act_out, true, dL_dout, dout_stride);
  if (rc) { char b[128]; snprintf(b, sizeof b, "ngp_mlp_bw: unsupported MLP shape (code %d)", rc); return set_error_msg(b); }
  SegPtrs in; SegGrads dg;
  for (int s = 0; s < kMaxSeg; s++) {
    in.p[s] = s < n_seg ? seg_ptr[s] : nullptr;
    dg.p[s] = (s < n_seg && dseg_ptr && seg_kind[s] == kSegPlain) ? dseg_ptr[s] : nullptr;
    dg.stride[s] = (s < n_seg && dseg_stride) ? dseg_stride[s] : 0;
  }
  cudaError_t e = cudaFuncSetAttribute(mlp_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem_bytes);
  if (e != cudaSuccess) return set_error(e, "ngp_mlp_bw/attr");
  const int grid = launch_grid((const void*)mlp_bw_kernel, c, n);
  mlp_bw_kernel<<<grid, kThreads, c.smem_bytes, (cudaStream_t)stream>>>(c, in, params, n, dL_dout, dout_stride, dparams, dg, dL_daux_exp);
  NGP_LAUNCH_CHECK("ngp_mlp_bw");
  return 0;
}
