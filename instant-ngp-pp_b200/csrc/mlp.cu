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
// Kernel shape (v6, "slots"): ONE persistent CTA per SM holding SLOTS independent 128-sample tiles in
// flight.  Slot s = warps 4s..4s+3 (128 threads, thread t owns tile row t, warp w reads TMEM lanes
// 32*(w%4)..+31); it has its own operand tiles in shared memory, its own accumulator columns in TMEM, its
// own mbarrier and its own named barrier (bar.sync s+1, 128), and thread 0 of the slot issues that slot's
// tcgen05.mma.  The per-tile chain  MMA -> TMEM read -> activation -> bf16 smem -> MMA  is latency bound
// (ncu r01: IPC 1.1, tensor pipe 2 % busy), so throughput = tiles in flight per SM: the slots share ONE copy
// of the weight tiles and ONE set of weight-gradient accumulators in TMEM (all slots' wgrad MMAs accumulate
// into the same columns; the tensor pipe executes MMAs in order), which is what lets 4-8 tiles fit where the
// one-tile-per-CTA design was capped at 2-4 by TMEM columns (power-of-two allocations) and shared memory.
//   forward : X -> [MMA -> act -> bf16 smem]* -> MMA -> act_out -> global
//   backward: recompute the forward chain (hidden activations stay in smem), then per layer, top
//             down:  wgrad (accumulated in TMEM for the CTA's whole lifetime, flushed once with
//             fp32 atomics) and dgrad, both reading the SAME smem tiles as MN-major operands.
// Operand tiles use the no-swizzle layout documented in tc05.cuh.
#include "common.cuh"
#include "tc05.cuh"
#include "mlp_tiles.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace ngp {
using namespace tc05;

constexpr int kSlotThreads = 128;   // 4 warps per slot: warp w reads TMEM lanes 32*(w%4)..
constexpr int kMaxSlots = 8;
constexpr int kMaxSeg = 3;
constexpr int kMaxHidden = 6;

enum Act { kActNone = 0, kActReLU = 1, kActSigmoid = 2, kActExp = 3 };
enum SegKind { kSegPlain = 0, kSegSH4 = 1, kSegTiles = 2 };   // 2: bf16 feature tiles already in operand layout (hashgrid.cu)

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
  int slots;                     // tiles in flight per CTA
  int coop_stage;                // wide fp32 input rows staged warp-cooperatively (whole rows per load instruction)
  const int32_t* n_dev;          // forward only, optional: the row count lives in device memory (n = min(n, *n_dev)) — test-time wavefront rounds
  // shared memory: [0,64) MMA mbarriers, [64,128) landing-zone mbarriers, [128,132) TMEM base, weights (shared by all slots), then one region per slot
  uint32_t off_w[kMaxHidden + 1];  // weight tiles: layer 0..nh-1, then output layer at [nh]   (absolute)
  uint32_t slot_base, slot_bytes;
  uint32_t off_x, off_h[kMaxHidden], off_dz;   // relative to the slot's region
  // raw fp32 landing zone of the slot's NEXT tile (one bulk copy per segment), relative to the slot's region
  uint32_t off_raw[kMaxSeg], raw_bytes[kMaxSeg], raw_total;
  int bulk;                        // every segment contiguous + 16-byte aligned, and the zone fits
  int x_tiles;                     // the single input segment is a feature-tile array: X is double buffered and filled by ONE bulk copy
  uint32_t x_bytes;                // bytes of one X operand tile
  uint32_t smem_bytes;
  // TMEM columns: slot s accumulates in [s*acc_cols, (s+1)*acc_cols); wgrad accumulators (backward) follow
  uint32_t acc_cols, tm_cols;
  uint32_t tm_wg[kMaxHidden + 1], wg_base, wg_cols;
  // parameter offsets (floats)
  int64_t p_off[kMaxHidden + 1];
};

struct SegPtrs { const float* p[kMaxSeg]; };
struct SegGrads { float* p[kMaxSeg]; int64_t stride[kMaxSeg]; int tiles; };   // tiles: p[0] is a gradient-tile array (hashgrid.cu DYT)

__device__ __forceinline__ float act_apply(int a, float z) {
  switch (a) {
    case kActReLU: return fmaxf(z, 0.f);
    case kActSigmoid: return __fdividef(1.f, 1.f + __expf(-z));
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
    for (int i = 0; i < N; i++) v[i] = __fdividef(1.f, 1.f + __expf(-v[i]));
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

// first `cnt` (<= N, warp-uniform) elements only: the padded output columns never pay for a sigmoid / exp
template <int N> __device__ __forceinline__ void act_apply_cnt(int a, float* v, int cnt) {
  if (a == kActNone) return;
  if (cnt >= N) { act_apply_arr<N>(a, v); return; }
#pragma unroll
  for (int i = 0; i < N; i++) if (i < cnt) v[i] = act_apply(a, v[i]);
}
template <int N> __device__ __forceinline__ void act_grad_mul_cnt(int a, float* v, const float* y, int cnt) {
  if (a == kActNone) return;
  if (cnt >= N) { act_grad_mul_arr<N>(a, v, y); return; }
#pragma unroll
  for (int i = 0; i < N; i++) if (i < cnt) v[i] *= act_grad_from_out(a, y[i]);
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


// ---- slot geometry ---------------------------------------------------------------------------------
struct Slot {
  uint32_t id, t;           // slot index in the CTA; thread index in the slot (= tile row)
  bool lead;                // this warp is the slot's first warp: one elected lane of it issues the slot's MMAs / bulk copies
  uint32_t tacc;            // TMEM address of the slot's accumulator (lane 0)
  uint32_t trow;            // same columns, this warp's lane quadrant
  uint32_t twg;             // TMEM address (lane 0) of the CTA-wide wgrad accumulators
  uint64_t* bar; uint32_t phase;
  uint64_t* full; uint32_t fphase;  // landing zone of the next tile (bulk copies)
  uint8_t* base; uint32_t sbase;   // the slot's shared-memory region: generic pointer / shared-window address
};
__device__ __forceinline__ void slot_sync(const Slot& S) { asm volatile("bar.sync %0, 128;" ::"r"(S.id + 1) : "memory"); }
// every thread of the slot: smem written by this thread is ready for the tensor core, its TMEM reads are done
__device__ __forceinline__ void publish(const Slot& S) {
  fence_async_smem();
  fence_before_sync();
  slot_sync(S);
}
// every thread of the slot: wait until the MMAs committed last by the slot's issuer have finished.
// NOTE: mbarrier.try_wait parks the WARP; single-thread issue blocks are therefore followed by __syncwarp()
// so that lane 0 never has work left when its sibling lanes start waiting.
__device__ __forceinline__ void wait_mma(Slot& S) {
  mbar_wait(S.bar, S.phase);
  S.phase ^= 1;
  fence_after_sync();
}

// Shape of the MLP as the kernels see it: compile-time constants for the hot shapes of the field (every loop
// over layers / columns / segments unrolls, no index arithmetic on runtime widths), copied from the launch
// configuration for everything else.
struct Dims {
  int n_seg, seg_w[kMaxSeg], seg_kind[kMaxSeg];
  int k0, k0p, w, wp, nh, no, nop, act_h, act_o;
};
struct GenericShape { static constexpr bool kStatic = false; };
template <int NSEG, int S0K, int S0W, int S1K, int S1W, int W, int NH, int NO, int AH, int AO>
struct StaticShape {
  static constexpr bool kStatic = true;
  static constexpr int n_seg = NSEG, s0k = S0K, s0w = S0W, s1k = S1K, s1w = S1W, w = W, nh = NH, no = NO, ah = AH, ao = AO;
};
// the two networks of the ngp_pl-shaped field (networks.py NGPCompact): density 32 -> 64 -> 16, colour [SH4(d) | h16] -> 64 -> 64 -> 3
using SigmaShape = StaticShape<1, kSegPlain, 32, 0, 0, 64, 1, 16, kActReLU, kActNone>;
using RgbShape = StaticShape<2, kSegSH4, 16, kSegPlain, 16, 64, 2, 3, kActReLU, kActSigmoid>;
using SigmaTilesShape = StaticShape<1, kSegTiles, 32, 0, 0, 64, 1, 16, kActReLU, kActNone>;   // density net fed by feature tiles

template <typename SH> __device__ __forceinline__ Dims make_dims(const MlpCfg& c) {
  Dims d;
  if constexpr (SH::kStatic) {
    d.n_seg = SH::n_seg;
    d.seg_w[0] = SH::s0w; d.seg_kind[0] = SH::s0k; d.seg_w[1] = SH::s1w; d.seg_kind[1] = SH::s1k; d.seg_w[2] = 0; d.seg_kind[2] = 0;
    d.k0 = SH::s0w + SH::s1w; d.k0p = (d.k0 + 15) / 16 * 16;
    d.w = SH::w; d.wp = SH::w <= 64 ? 64 : 128; d.nh = SH::nh; d.no = SH::no; d.nop = (SH::no + 15) / 16 * 16;
    d.act_h = SH::ah; d.act_o = SH::ao;
  } else {
    d.n_seg = c.n_seg;
#pragma unroll
    for (int s = 0; s < kMaxSeg; s++) { d.seg_w[s] = c.seg_w[s]; d.seg_kind[s] = c.seg_kind[s]; }
    d.k0 = c.k0; d.k0p = c.k0p; d.w = c.w; d.wp = c.wp; d.nh = c.nh; d.no = c.no; d.nop = c.nop; d.act_h = c.act_h; d.act_o = c.act_o;
  }
  return d;
}

// CTA prologue: mbarriers, TMEM allocation, weights -> bf16 operand tiles; returns the TMEM base.
__device__ __forceinline__ uint32_t cta_setup(const MlpCfg& c, const Dims& d, const float* __restrict__ params, uint8_t* smem) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 128);
  if (threadIdx.x == 0) {
    for (int s = 0; s < c.slots; s++) { mbar_init(bars + s, 1); mbar_init(bars + 8 + s, 1); }
    mbar_fence_init();
  }
  if (threadIdx.x < 32) { __syncwarp(); tmem_alloc(tslot, c.tm_cols); }
  // fp32 parameter vector -> bf16 operand tiles (zero padded), all threads
  for (int l = 0; l <= d.nh; l++) {
    const int rows_t = (l == d.nh) ? d.nop : d.wp;           // tile rows
    const int cols_t = (l == 0) ? d.k0p : d.wp;              // tile cols
    const int rows = (l == d.nh) ? d.nop : d.w;              // rows present in params
    const int cols = (l == 0) ? d.k0 : d.w;
    const float* W = params + c.p_off[l];
    uint8_t* tile = smem + c.off_w[l];
    for (int i = threadIdx.x; i < rows_t * cols_t; i += blockDim.x) {
      const int r = i / cols_t, cc = i - r * cols_t;
      const float v = (r < rows && cc < cols) ? __ldg(W + (int64_t)r * cols + cc) : 0.f;
      st_elem(tile, rows_t, r, cc, v);
    }
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  return *tslot;
}
__device__ __forceinline__ Slot make_slot(const MlpCfg& c, const Dims& d, uint8_t* smem, uint32_t tmem) {
  Slot S;
  // Everything but S.t is the same for the 32 lanes of a warp.  Routing the warp index (and the TMEM base read from
  // shared memory) through a lane-0 shuffle tells the compiler so: slot addresses, TMEM addresses and with them the
  // MMA descriptors then live in UNIFORM registers, and the single-thread issue blocks lose the per-MMA
  // R2UR.BROADCAST / ELECT waterfall they had when these were per-lane values (18 -> 4 instructions per UTCHMMA,
  // profiles/sass/mlp_*: the issue block is serial time of the slot's tile chain).
  const uint32_t warp_u = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  tmem = __shfl_sync(0xffffffffu, tmem, 0);
  S.id = warp_u >> 2; S.t = threadIdx.x & 127u; S.lead = (warp_u & 3u) == 0u;
  S.tacc = tmem + S.id * c.acc_cols;
  S.trow = S.tacc + (((warp_u & 3u) * 32u) << 16);
  S.twg = tmem + c.wg_base;
  S.bar = reinterpret_cast<uint64_t*>(smem) + S.id; S.phase = 0;
  S.full = reinterpret_cast<uint64_t*>(smem) + 8 + S.id; S.fphase = 0;
  S.base = smem + c.slot_base + S.id * c.slot_bytes;
  S.sbase = smem_u32(S.base);
  return S;
}
__device__ __forceinline__ void cta_teardown(const MlpCfg& c, const Dims& d, uint32_t tmem) {
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, c.tm_cols);
}

// Thread t of a slot stages ITS row of every input segment as bf16 into the slot's X tile (16-byte row chunks:
// a warp writes 512 contiguous bytes per store) — rows are 48..640 contiguous bytes in global memory, so every
// sector a thread touches is fully used; SH segments are evaluated in registers (no dir_encoder pass, no cat).
__device__ __forceinline__ void stage_row(const MlpCfg& c, const Dims& d, const SegPtrs& in, int64_t row, bool valid, uint32_t t, uint8_t* Xs, int64_t n) {
  int col = 0;
  _Pragma("unroll") for (int s = 0; s < kMaxSeg; s++) { if (s >= d.n_seg) break;
    const int w = d.seg_w[s];
    const float* src = in.p[s] + row * c.seg_stride[s];
    if (d.seg_kind[s] == kSegSH4) {
      float o[16];
      if (valid) {
        const float dx = __ldg(src), dy = __ldg(src + 1), dz = __ldg(src + 2);
        const float inv = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-6f);  // F.normalize(eps=1e-6), networks.py:221
        sh4(dx * inv, dy * inv, dz * inv, o);
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) o[i] = 0.f;
      }
      if ((col & 7) == 0) { st_chunk(Xs, kTile, t, col, o); st_chunk(Xs, kTile, t, col + 8, o + 8); }
      else { for (int i = 0; i < 16; i++) st_elem(Xs, kTile, t, col + i, o[i]); }
    } else {
      const bool vec = ((w & 3) == 0) && ((c.seg_stride[s] & 3) == 0) && ((((uintptr_t)in.p[s]) & 15) == 0);
      if (c.coop_stage && vec && w >= 32 && (w & (w - 1)) == 0 && (col & 3) == 0) {
        // wide rows (the (S,128) feature matrix of the reference's heads): whole rows per load instruction, not one row per thread
        stage_rows_coop(in.p[s], c.seg_stride[s], w, col, row - t, n, t >> 5, t & 31u, Xs);
      } else if (vec && ((col | w) & 7) == 0) {
#pragma unroll 4
        for (int c8 = 0; c8 < w; c8 += 8) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
          if (valid) { a = __ldg(reinterpret_cast<const float4*>(src + c8)); b = __ldg(reinterpret_cast<const float4*>(src + c8 + 4)); }
          const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
          st_chunk(Xs, kTile, t, col + c8, v);
        }
      } else if (vec && (col & 3) == 0) {
        for (int c4 = 0; c4 < w; c4 += 4)
          st_quad(Xs, kTile, t, col + c4, valid ? __ldg(reinterpret_cast<const float4*>(src + c4)) : make_float4(0.f, 0.f, 0.f, 0.f));
      } else {
        for (int cc = 0; cc < w; cc++) st_elem(Xs, kTile, t, col + cc, valid ? __ldg(src + cc) : 0.f);
      }
    }
    col += w;
  }
}
// ---- next-tile prefetch through the TMA engine: the slot's thread 0 arms `full` and fires one bulk copy per
// segment into the slot's raw fp32 landing zone; the copy overlaps the whole MMA/epilogue chain of the
// current tile and costs no registers and no LSU instructions.
__device__ __forceinline__ bool tile_is_bulk(const MlpCfg& c, const Dims& d, int64_t tile, int64_t n) {
  return c.bulk && (tile + 1) * kTile <= n;
}
// feature tiles: tile k of a slot lands in X buffer (k & 1) — already bf16, already in operand layout
__device__ __forceinline__ void issue_tile_prefetch(const MlpCfg& c, const SegPtrs& in, int64_t tile, int buf, const Slot& S) {
  mbar_expect_tx(S.full, c.x_bytes);
  bulk_g2s(S.base + c.off_x + buf * c.x_bytes, reinterpret_cast<const uint8_t*>(in.p[0]) + tile * (int64_t)c.x_bytes, c.x_bytes, S.full);
}
__device__ __forceinline__ void issue_prefetch(const MlpCfg& c, const Dims& d, const SegPtrs& in, int64_t tile, const Slot& S) {
  const int64_t row0 = tile * kTile;
  mbar_expect_tx(S.full, c.raw_total);
  _Pragma("unroll") for (int s = 0; s < kMaxSeg; s++)
    if (s < d.n_seg) bulk_g2s(S.base + c.off_raw[s], in.p[s] + row0 * c.seg_stride[s], c.raw_bytes[s], S.full);
}
// Raw fp32 landing zone -> bf16 operand tile.  Thread r converts row r; in step i it reads 16-byte unit
// #((i + r) mod upr) of its row, so the 32 lanes of a warp hit distinct bank groups even though rows
// are w*4 bytes apart (a straight per-row walk would be a 32-way conflict), and no index division is needed.
__device__ __forceinline__ void convert_raw(const MlpCfg& c, const Dims& d, const Slot& S, uint8_t* Xs) {
  const uint32_t t = S.t;
  int col = 0;
  _Pragma("unroll") for (int s = 0; s < kMaxSeg; s++) { if (s >= d.n_seg) break;
    const int w = d.seg_w[s];
    const float* raw = reinterpret_cast<const float*>(S.base + c.off_raw[s]);
    if (d.seg_kind[s] == kSegSH4) {
      const float dx = raw[3 * t], dy = raw[3 * t + 1], dz = raw[3 * t + 2];
      const float inv = 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-6f);
      float o[16];
      sh4(dx * inv, dy * inv, dz * inv, o);
      if ((col & 7) == 0) { st_chunk(Xs, kTile, t, col, o); st_chunk(Xs, kTile, t, col + 8, o + 8); }
      else { for (int i = 0; i < 16; i++) st_elem(Xs, kTile, t, col + i, o[i]); }
    } else if (((col | w) & 7) == 0) {
      const int upr = w >> 3;                                 // 32-byte (8-float) units per row
      const float4* rowp = reinterpret_cast<const float4*>(raw) + (size_t)t * (upr * 2);
      int u = (int)(t % (uint32_t)upr);
#pragma unroll 4
      for (int i = 0; i < upr; i++) {
        const float4 a = rowp[2 * u], b = rowp[2 * u + 1];
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        st_chunk(Xs, kTile, t, col + 8 * u, v);
        u = (u + 1 == upr) ? 0 : u + 1;
      }
    } else if (((col | w) & 3) == 0) {
      const int upr = w >> 2;
      const float4* rowp = reinterpret_cast<const float4*>(raw) + (size_t)t * upr;
      int c4 = (int)(t % (uint32_t)upr);
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
// the row this thread will stage for the slot's NEXT tile: pull it into L2 now (no registers, no smem)
__device__ __forceinline__ void prefetch_row(const MlpCfg& c, const Dims& d, const SegPtrs& in, int64_t row) {
  _Pragma("unroll") for (int s = 0; s < kMaxSeg; s++) { if (s >= d.n_seg) break;
    const char* p = reinterpret_cast<const char*>(in.p[s] + row * c.seg_stride[s]);
    const int bytes = (d.seg_kind[s] == kSegSH4 ? 3 : d.seg_w[s]) * 4;
    for (int b = 0; b < bytes; b += 128) prefetch_l2(p + b);
  }
}
// zero the padding columns [k0, k0p) of the slot's X tile once (segments never touch them)
__device__ __forceinline__ void zero_pad_cols(const MlpCfg& c, const Dims& d, uint32_t t, uint8_t* Xs) {
  for (int cc = d.k0; cc < d.k0p; cc++) st_elem(Xs, kTile, t, cc, 0.f);
}

// hidden-layer epilogue: the thread's row of the wp accumulator columns -> act -> bf16 tile row, CH columns
// per TMEM load (CH = 32: two loads in flight per wait when registers allow; 16 for the 1024-thread CTAs)
template <int W>
__device__ __forceinline__ void act_store(const MlpCfg& c, const Dims& d, const Slot& S, uint8_t* H, int c0, float* v) {
  if (d.act_h == kActReLU) {
#pragma unroll
    for (int q = 0; q < W / 8; q++) st_chunk_relu(H, kTile, S.t, c0 + 8 * q, v + 8 * q);
  } else {
    act_apply_arr<W>(d.act_h, v);
#pragma unroll
    for (int q = 0; q < W / 8; q++) st_chunk(H, kTile, S.t, c0 + 8 * q, v + 8 * q);
  }
}
// CH = 64: two x32 TMEM loads in flight per wait (needs ~100 registers); 32 / 16: one load per wait
template <int CH>
__device__ __forceinline__ void epilogue_hidden(const MlpCfg& c, const Dims& d, const Slot& S, uint8_t* H) {
  if (CH == 64) {
    for (int c0 = 0; c0 < d.wp; c0 += 64) {          // wp is 64 or 128
      uint32_t r0[32], r1[32];
      tmem_ld32_nowait(S.trow + c0, r0);
      tmem_ld32_nowait(S.trow + c0 + 32, r1);
      tmem_wait_ld();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r0[i]);
      act_store<32>(c, d, S, H, c0, v);
#pragma unroll
      for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r1[i]);
      act_store<32>(c, d, S, H, c0 + 32, v);
    }
  } else if (CH == 32) {
    for (int c0 = 0; c0 < d.wp; c0 += 32) {
      float v[32];
      tmem_ld32(S.trow + c0, v);
      act_store<32>(c, d, S, H, c0, v);
    }
  } else {
    for (int c0 = 0; c0 < d.wp; c0 += 16) {
      float v[16];
      tmem_ld16(S.trow + c0, v);
      act_store<16>(c, d, S, H, c0, v);
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

// =============================================================================== forward kernel
// aux_exp (optional): aux_exp[row] = exp(z_out[row][0]) — the density head of the ngp_pl-shaped field
// (sigma = TruncExp(h[:,0])) produced by the same epilogue instead of a strided select + exp pass.
template <int SLOTS, typename SH>
__global__ void __launch_bounds__(kSlotThreads * SLOTS) mlp_fw_kernel(MlpCfg c, SegPtrs in, const float* __restrict__ params, int64_t n,
                                                                      float* __restrict__ out, int64_t out_stride,
                                                                      float* __restrict__ aux_exp) {
  extern __shared__ __align__(128) uint8_t smem[];
  if (c.n_dev) { const int64_t nd = (int64_t)__ldg(c.n_dev); if (nd < n) n = nd; }
  const Dims d = make_dims<SH>(c);
  const uint32_t tmem = cta_setup(c, d, params, smem);
  Slot S = make_slot(c, d, smem, tmem);
  const uint32_t t = S.t;
  uint8_t* Xs = S.base + c.off_x;
  uint8_t* Hs = S.base + c.off_h[0];
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int64_t tstride = (int64_t)gridDim.x * SLOTS;
  const bool out_vec = ((out_stride & 3) == 0) && ((((uintptr_t)out) & 15) == 0);
  constexpr int CH = SLOTS > 6 ? 16 : (SLOTS > 4 ? 32 : 64);

  const bool xt = d.seg_kind[0] == kSegTiles;
  if (!xt) zero_pad_cols(c, d, t, Xs);
  {
    const int64_t first = (int64_t)blockIdx.x * SLOTS + S.id;
    if (S.lead && first < n_tiles && elect_one()) {
      if (xt) issue_tile_prefetch(c, in, first, 0, S);
      else if (tile_is_bulk(c, d, first, n)) issue_prefetch(c, d, in, first, S);
    }
    __syncwarp();
  }
  int buf = 0;
  for (int64_t tile = (int64_t)blockIdx.x * SLOTS + S.id; tile < n_tiles; tile += tstride, buf ^= 1) {
    const int64_t row = tile * kTile + t;
    const uint32_t xoff = c.off_x + (xt ? buf * c.x_bytes : 0u);
    if (xt) { mbar_wait(S.full, S.fphase); S.fphase ^= 1; }
    else if (tile_is_bulk(c, d, tile, n)) { mbar_wait(S.full, S.fphase); S.fphase ^= 1; convert_raw(c, d, S, Xs); }
    else stage_row(c, d, in, row, row < n, t, Xs, n);
    publish(S);
    {   // the landing zone is free again: fetch the slot's next tile while this one is computed
      const int64_t nxt = tile + tstride;
      if (nxt < n_tiles) {
        if (xt) { if (S.lead && elect_one()) issue_tile_prefetch(c, in, nxt, buf ^ 1, S); }
        else if (tile_is_bulk(c, d, nxt, n)) { if (S.lead && elect_one()) issue_prefetch(c, d, in, nxt, S); }
        else if (row + tstride * kTile < n) prefetch_row(c, d, in, row + tstride * kTile);
      }
      __syncwarp();
    }
    for (int l = 0; l <= d.nh; l++) {
      const bool last = l == d.nh;
      const int N = last ? d.nop : d.wp;
      const int K = l == 0 ? d.k0p : d.wp;
      if (S.lead && elect_one()) {
        fence_after_sync();
        issue_fwd(S.tacc, S.sbase + (l == 0 ? xoff : c.off_h[0]), smem_u32(smem) + c.off_w[l], (uint32_t)N, N, K);
        mma_commit(S.bar);
      }
      __syncwarp();
      wait_mma(S);
      if (!last) {
        epilogue_hidden<CH>(c, d, S, Hs);
        publish(S);
      } else {
        if (d.no <= 4) {                               // colour / normal / density heads: 1..4 real columns
          float v[4];
          tmem_ld4(S.trow, v);
          if (row < n) {
            if (aux_exp) aux_exp[row] = __expf(v[0]);
            float* o = out + row * out_stride;
            if (d.act_o == kActSigmoid) {
#pragma unroll
              for (int i = 0; i < 4; i++) v[i] = __fdividef(1.f, 1.f + __expf(-v[i]));
            } else if (d.act_o == kActExp) {
#pragma unroll
              for (int i = 0; i < 4; i++) v[i] = __expf(v[i]);
            } else if (d.act_o == kActReLU) {
#pragma unroll
              for (int i = 0; i < 4; i++) v[i] = fmaxf(v[i], 0.f);
            }
            o[0] = v[0];
            if (d.no > 1) o[1] = v[1];
            if (d.no > 2) o[2] = v[2];
            if (d.no > 3) o[3] = v[3];
          }
        } else {
          for (int c0 = 0; c0 < d.nop; c0 += 16) {
            float v[16];
            tmem_ld16(S.trow + c0, v);
            if (row < n) {
              if (aux_exp && c0 == 0) aux_exp[row] = __expf(v[0]);
              act_apply_cnt<16>(d.act_o, v, d.no - c0);
              store_row16(out + row * out_stride, c0, d.no, out_vec, v);
            }
          }
        }
        // the next tile's publish() (fence + slot barrier) orders these TMEM reads before its first MMA
      }
    }
  }
  cta_teardown(c, d, tmem);
}

// =============================================================================== backward kernel
template <int SLOTS, typename SH>
__global__ void __launch_bounds__(kSlotThreads * SLOTS) mlp_bw_kernel(MlpCfg c, SegPtrs in, const float* __restrict__ params, int64_t n,
                                                                      const float* __restrict__ dout, int64_t dout_stride,
                                                                      float* __restrict__ dparams, SegGrads dseg,
                                                                      const float* __restrict__ d_aux_exp,
                                                                      const float* __restrict__ yout, int64_t yout_stride) {
  extern __shared__ __align__(128) uint8_t smem[];
  const Dims d = make_dims<SH>(c);
  const uint32_t tmem = cta_setup(c, d, params, smem);
  Slot S = make_slot(c, d, smem, tmem);
  const uint32_t t = S.t;
  const uint32_t wbase = smem_u32(smem);
  uint8_t* Xs = S.base + c.off_x;
  uint8_t* dZ = S.base + c.off_dz;
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  const int64_t tstride = (int64_t)gridDim.x * SLOTS;
  bool want_dx = false;
  _Pragma("unroll") for (int s = 0; s < kMaxSeg; s++) if (s < d.n_seg) want_dx |= dseg.p[s] != nullptr;
  const bool dout_vec = ((dout_stride & 3) == 0) && ((((uintptr_t)dout) & 15) == 0);
  const uint32_t warp = threadIdx.x >> 5, quad = warp & 3u, part = warp >> 2;
  // The forward's OUTPUT (saved by autograd anyway) replaces the recompute of the output layer: act_o' is a function
  // of y, and the density head's z0 is y0 when act_o is None — one MMA -> TMEM -> epilogue round trip less per tile.
  const bool skip_out = yout != nullptr && (d.act_o == kActNone || (d.no <= 4 && d_aux_exp == nullptr));

  // zero the CTA-wide weight-gradient accumulators (every slot accumulates into them from its first tile on)
  for (uint32_t c0 = 16 * part; c0 < c.wg_cols; c0 += 16 * SLOTS) tmem_st16_zero(S.twg + ((quad * 32u) << 16) + c0);
  tmem_wait_st();
  if (d.seg_kind[0] != kSegTiles) zero_pad_cols(c, d, t, Xs);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const bool xt = d.seg_kind[0] == kSegTiles;
  {
    const int64_t first = (int64_t)blockIdx.x * SLOTS + S.id;
    if (S.lead && first < n_tiles && elect_one()) {
      if (xt) issue_tile_prefetch(c, in, first, 0, S);
      else if (tile_is_bulk(c, d, first, n)) issue_prefetch(c, d, in, first, S);
    }
    __syncwarp();
  }
  int buf = 0;
  for (int64_t tile = (int64_t)blockIdx.x * SLOTS + S.id; tile < n_tiles; tile += tstride, buf ^= 1) {
    const int64_t row = tile * kTile + t;
    const bool valid = row < n;
    const uint32_t xoff = c.off_x + (xt ? buf * c.x_bytes : 0u);
    if (xt) { mbar_wait(S.full, S.fphase); S.fphase ^= 1; }
    else if (tile_is_bulk(c, d, tile, n)) { mbar_wait(S.full, S.fphase); S.fphase ^= 1; convert_raw(c, d, S, Xs); }
    else stage_row(c, d, in, row, valid, t, Xs, n);
    publish(S);
    {
      const int64_t nxt = tile + tstride;
      if (S.lead && nxt < n_tiles && elect_one()) {
        if (xt) issue_tile_prefetch(c, in, nxt, buf ^ 1, S);
        else if (tile_is_bulk(c, d, nxt, n)) issue_prefetch(c, d, in, nxt, S);
      }
      __syncwarp();
    }
    // the first 16 columns of this row's dL/dy are requested NOW and consumed after the forward recompute,
    // so their HBM latency hides behind the MMA phases
    float dreg[16];
    if (d.no <= 4) {
      const float* drow = dout + row * dout_stride;
#pragma unroll
      for (int i = 0; i < 4; i++) dreg[i] = (valid && i < d.no) ? __ldg(drow + i) : 0.f;
    } else {
      const float* drow = dout + row * dout_stride;
      if (valid && dout_vec && 16 <= d.no) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(drow) + q);
          dreg[4 * q] = g4.x; dreg[4 * q + 1] = g4.y; dreg[4 * q + 2] = g4.z; dreg[4 * q + 3] = g4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) dreg[i] = (valid && i < d.no) ? __ldg(drow + i) : 0.f;
      }
    }
    const float daux = (d_aux_exp && valid) ? __ldg(d_aux_exp + row) : 0.f;
    float yreg[4] = {0.f, 0.f, 0.f, 0.f};
    if (skip_out && valid) {
      const float* yr = yout + row * yout_stride;
      if (d.no <= 4) {
#pragma unroll
        for (int i = 0; i < 4; i++) if (i < d.no) yreg[i] = __ldg(yr + i);
      } else yreg[0] = __ldg(yr);
    }
    {
      const int64_t nrow = row + tstride * kTile;
      if (nrow < n) {
        if (!xt && !tile_is_bulk(c, d, tile + tstride, n)) prefetch_row(c, d, in, nrow);
        for (int b = 0; b < d.no * 4; b += 128) prefetch_l2(reinterpret_cast<const char*>(dout + nrow * dout_stride) + b);
      }
    }
    // ---- recompute the forward chain; H_{l+1} kept in smem (post-activation, bf16)
    for (int l = 0; l < d.nh; l++) {
      if (S.lead && elect_one()) {
        fence_after_sync();
        issue_fwd(S.tacc, S.sbase + (l == 0 ? xoff : c.off_h[l - 1]), wbase + c.off_w[l], (uint32_t)d.wp, d.wp, l == 0 ? d.k0p : d.wp);
        mma_commit(S.bar);
      }
      __syncwarp();
      wait_mma(S);
      epilogue_hidden<(SLOTS > 4 ? 32 : 64)>(c, d, S, S.base + c.off_h[l]);
      if (skip_out && l == d.nh - 1) {
        // dZ_out = dL/dy * act_o'(y) from the saved output, written together with the last hidden tile
        if (d.no <= 4) {
          float g[16];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float y = yreg[i];
            float dy = 1.f;
            if (d.act_o == kActSigmoid) dy = y * (1.f - y);
            else if (d.act_o == kActExp) dy = y;
            else if (d.act_o == kActReLU) dy = y > 0.f ? 1.f : 0.f;
            g[i] = i < d.no ? dreg[i] * dy : 0.f;
          }
          if (d_aux_exp) g[0] = fmaf(daux, __expf(fminf(fmaxf(yreg[0], -7.f), 7.f)), g[0]);
#pragma unroll
          for (int i = 4; i < 16; i++) g[i] = 0.f;
          st_chunk(dZ, kTile, t, 0, g);
          st_chunk(dZ, kTile, t, 8, g + 8);
        } else {
          for (int c0 = 0; c0 < d.nop; c0 += 16) {
            float g[16];
#pragma unroll
            for (int i = 0; i < 16; i++) g[i] = c0 == 0 ? dreg[i] : ((valid && c0 + i < d.no) ? __ldg(dout + row * dout_stride + c0 + i) : 0.f);
            if (c0 == 0 && d_aux_exp) g[0] = fmaf(daux, __expf(fminf(fmaxf(yreg[0], -7.f), 7.f)), g[0]);
#pragma unroll
            for (int i = 0; i < 16; i++) if (c0 + i >= d.no) g[i] = 0.f;
            st_chunk(dZ, kTile, t, c0, g);
            st_chunk(dZ, kTile, t, c0 + 8, g + 8);
          }
        }
      }
      publish(S);
    }
    // ---- output layer pre-activation -> dZ_out = dL/dy * act_o'(z)
    if (!skip_out) {
    if (S.lead && elect_one()) {
      fence_after_sync();
      issue_fwd(S.tacc, S.sbase + c.off_h[d.nh - 1], wbase + c.off_w[d.nh], (uint32_t)d.nop, d.nop, d.wp);
      mma_commit(S.bar);
    }
    __syncwarp();
    wait_mma(S);
    if (d.no <= 4) {
      float v[4], g[16];
      tmem_ld4(S.trow, v);
      const float z0 = v[0];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float y = v[i], dy = 1.f;
        if (d.act_o == kActSigmoid) { y = __fdividef(1.f, 1.f + __expf(-y)); dy = y * (1.f - y); }
        else if (d.act_o == kActExp) { y = __expf(y); dy = y; }
        else if (d.act_o == kActReLU) { dy = y > 0.f ? 1.f : 0.f; }
        g[i] = i < d.no ? dreg[i] * dy : 0.f;
      }
      if (d_aux_exp) g[0] = fmaf(daux, __expf(fminf(fmaxf(z0, -7.f), 7.f)), g[0]);
#pragma unroll
      for (int i = 4; i < 16; i++) g[i] = 0.f;
      st_chunk(dZ, kTile, t, 0, g);
      st_chunk(dZ, kTile, t, 8, g + 8);
    } else
    for (int c0 = 0; c0 < d.nop; c0 += 16) {
      float v[16];
      tmem_ld16(S.trow + c0, v);
      if (c0 != 0) {                       // n_out > 16: later column groups are read at use
#pragma unroll
        for (int i = 0; i < 16; i++) dreg[i] = (valid && c0 + i < d.no) ? __ldg(dout + row * dout_stride + c0 + i) : 0.f;
      }
      const float z0 = v[0];
      float g[16];
#pragma unroll
      for (int i = 0; i < 16; i++) g[i] = dreg[i];
      act_apply_cnt<16>(d.act_o, v, d.no - c0);      // v = y
      act_grad_mul_cnt<16>(d.act_o, g, v, d.no - c0);   // g = dL/dy * act'(z)
      // TruncExp backward of the density head: + dL/dsigma * exp(clamp(z0, -7, 7))  (custom_functions.py:211)
      if (c0 == 0 && d_aux_exp) g[0] = fmaf(daux, __expf(fminf(fmaxf(z0, -7.f), 7.f)), g[0]);
#pragma unroll
      for (int i = 0; i < 16; i++) v[i] = (c0 + i < d.no) ? g[i] : 0.f;
      st_chunk(dZ, kTile, t, c0, v);
      st_chunk(dZ, kTile, t, c0 + 8, v + 8);
    }
    publish(S);
    }
    // ---- top-down: wgrad + dgrad per layer.  dZ_out lives in its own (narrow) tile; every hidden dZ_l is
    // written IN PLACE over H_{l} (the thread that reads a 16-byte chunk for the activation mask is the one
    // that overwrites it), so no wp-wide gradient tile is needed.
    for (int l = d.nh; l >= 0; l--) {
      const bool is_out = l == d.nh;
      const int Nz = is_out ? d.nop : d.wp;                       // width of dZ_l
      const int Kin = l == 0 ? d.k0p : d.wp;                      // width of the layer's input
      const uint32_t ain = S.sbase + (l == 0 ? xoff : c.off_h[l - 1]);
      const uint32_t dza = S.sbase + (is_out ? c.off_dz : c.off_h[l]);
      const bool need_dgrad = l > 0 || want_dx;
      if (S.lead && elect_one()) {
        fence_after_sync();
        if (is_out) issue_wgrad(S.twg + c.tm_wg[l], ain, dza, d.wp, d.nop);   // D^T[in x out]
        else issue_wgrad(S.twg + c.tm_wg[l], dza, ain, d.wp, Kin);            // D[out x in]
        if (need_dgrad) issue_dgrad(S.tacc, dza, wbase + c.off_w[l], (uint32_t)Nz, Kin, Nz);
        mma_commit(S.bar);
      }
      __syncwarp();
      wait_mma(S);
      if (l > 0) {
        uint8_t* H = S.base + c.off_h[l - 1];
        if (d.act_h == kActReLU) {
          for (int c0 = 0; c0 < d.wp; c0 += 32) {
            float v[32];
            tmem_ld32(S.trow + c0, v);
#pragma unroll
            for (int q = 0; q < 4; q++) relu_bw_chunk(H, kTile, t, c0 + 8 * q, v + 8 * q);
          }
        } else {
          for (int c0 = 0; c0 < d.wp; c0 += 32) {
            float v[32], h[32];
            tmem_ld32(S.trow + c0, v);
#pragma unroll
            for (int q = 0; q < 4; q++) ld_chunk(H, kTile, t, c0 + 8 * q, h + 8 * q);
            act_grad_mul_arr<32>(d.act_h, v, h);
#pragma unroll
            for (int q = 0; q < 4; q++) st_chunk(H, kTile, t, c0 + 8 * q, v + 8 * q);
          }
        }
        publish(S);
      } else if (want_dx && dseg.tiles) {
        // gradient tiles: row r of chunk q is 32 bytes at tile*(128*k0p) + q*(128*8) + r*8 floats — consecutive rows
        // contiguous, so a warp's two 16-byte stores per chunk cover 1 KB of whole sectors
        float* tbase = dseg.p[0] + tile * (int64_t)(kTile * d.k0p) + t * 8;
        for (int c0 = 0; c0 < d.k0p; c0 += 16) {
          float v[16];
          tmem_ld16(S.trow + c0, v);
#pragma unroll
          for (int q = 0; q < 2; q++) {
            float4* dst = reinterpret_cast<float4*>(tbase + ((c0 >> 3) + q) * (kTile * 8));
            dst[0] = make_float4(v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3]);
            dst[1] = make_float4(v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]);
          }
        }
      } else if (want_dx) {
        // input gradient straight from TMEM to global: each thread owns 64-byte runs of its row
        for (int c0 = 0; c0 < d.k0; c0 += 16) {
          float v[16];
          tmem_ld16(S.trow + c0, v);
          if (valid) {
            int col = 0;
            _Pragma("unroll") for (int s = 0; s < kMaxSeg; s++) { if (s >= d.n_seg) break;
              const int w = d.seg_w[s];
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
      // the next tile's publish() orders the l == 0 TMEM reads (and the wgrad reads of X) before its first MMA
    }
  }

  // ---- flush the TMEM-resident weight gradients with fp32 atomics (all warps: quadrant x column stripe)
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if ((int64_t)blockIdx.x * SLOTS < n_tiles) {
    // accumulator row m of an M=128 tile sits in TMEM lane m; of an M=64 tile in lane (m%16)+32*(m/16)
    const bool m128 = d.wp == 128;
    const uint32_t lane = threadIdx.x & 31;
    const int m = m128 ? (int)(quad * 32 + lane) : (lane < 16 ? (int)(quad * 16 + lane) : -1);
    for (int l = 0; l <= d.nh; l++) {
      const bool is_out = l == d.nh;
      const int ncols = is_out ? d.nop : (l == 0 ? d.k0p : d.wp);
      const int in_true = l == 0 ? d.k0 : d.w;
      for (int c0 = 16 * (int)part; c0 < ncols; c0 += 16 * SLOTS) {
        float v[16];
        tmem_ld16(S.twg + ((quad * 32u) << 16) + c.tm_wg[l] + c0, v);
        if (m >= 0 && m < d.w) {
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const int cc = c0 + i;
            if (is_out) { if (cc < d.no) atomicAdd(dparams + c.p_off[l] + (int64_t)cc * d.w + m, v[i]); }   // D^T[in=m][out=cc]
            else if (cc < in_true) atomicAdd(dparams + c.p_off[l] + (int64_t)m * in_true + cc, v[i]);       // D[out=m][in=cc]
          }
        }
      }
    }
  }
  cta_teardown(c, d, tmem);
}

// Fills the derived fields of cfg (everything but .slots / .tm_cols / .smem_bytes, which pick_slots sets);
// returns 0 or a negative error.
static int finalize_cfg(MlpCfg& c, bool backward) {
  if (c.n_seg < 1 || c.n_seg > kMaxSeg) return -1;
  c.k0 = 0;
  c.x_tiles = 0;
  for (int s = 0; s < c.n_seg; s++) {
    if (c.seg_kind[s] == kSegSH4 && c.seg_w[s] != 16) return -2;
    if (c.seg_kind[s] == kSegTiles) { if (c.n_seg != 1) return -6; c.x_tiles = 1; c.bulk = 0; }
    c.k0 += c.seg_w[s];
  }
  c.k0p = (c.k0 + 15) / 16 * 16;
  if (c.w < 1 || c.w > 128 || c.nh < 1 || c.nh > kMaxHidden || c.no < 1 || c.no > 128 || c.k0p > 256) return -3;
  c.wp = c.w <= 64 ? 64 : 128;
  c.nop = (c.no + 15) / 16 * 16;
  auto al = [](uint32_t x) { return (x + 127u) / 128u * 128u; };
  auto tile_bytes = [](int rows, int cols) { return (uint32_t)(cols / 8) * chunk_stride((uint32_t)rows); };
  uint32_t off = 256;  // mbarriers + tmem slot
  int64_t poff = 0;
  for (int l = 0; l <= c.nh; l++) {
    const int rows_t = l == c.nh ? c.nop : c.wp, cols_t = l == 0 ? c.k0p : c.wp;
    c.off_w[l] = off; off += al(tile_bytes(rows_t, cols_t));
    c.p_off[l] = poff;
    poff += (int64_t)(l == c.nh ? c.nop : c.w) * (l == 0 ? c.k0 : c.w);
  }
  c.slot_base = off;
  uint32_t rel = 0;
  c.x_bytes = tile_bytes(kTile, c.k0p);                   // == ngp_feature_tile_bytes: a multiple of 128
  c.off_x = rel; rel += al(c.x_bytes) * (c.x_tiles ? 2u : 1u);   // feature tiles: double buffered, filled by bulk copy
  const int n_h = backward ? c.nh : 1;
  for (int l = 0; l < n_h; l++) { c.off_h[l] = rel; rel += al(tile_bytes(kTile, c.wp)); }
  for (int l = n_h; l < kMaxHidden; l++) c.off_h[l] = c.off_h[0];
  c.off_dz = rel;
  if (backward) rel += al(tile_bytes(kTile, c.nop));      // dZ_out only; hidden dZ_l overwrite H_l in place
  c.slot_bytes = rel;                                     // without the landing zone; pick_slots may add it
  c.raw_total = 0;
  for (int s = 0; s < c.n_seg; s++) {
    c.raw_bytes[s] = (uint32_t)(kTile * (c.seg_kind[s] == kSegSH4 ? 3 : c.seg_w[s]) * 4);
    c.off_raw[s] = rel + c.raw_total; c.raw_total += al(c.raw_bytes[s]);
  }
  uint32_t acc = (uint32_t)c.wp;
  if ((uint32_t)c.nop > acc) acc = c.nop;
  if (backward && (uint32_t)c.k0p > acc) acc = c.k0p;
  c.acc_cols = acc;
  c.wg_cols = 0;
  if (backward) {
    for (int l = 0; l <= c.nh; l++) {
      c.tm_wg[l] = c.wg_cols;
      c.wg_cols += l == c.nh ? c.nop : (l == 0 ? c.k0p : c.wp);
    }
  }
  if (c.slot_base + c.slot_bytes > 227 * 1024) return -5;
  if (c.acc_cols + c.wg_cols > 512) return -4;
  return 0;
}

static uint32_t next_pow2_cols(uint32_t x) { uint32_t p = 32; while (p < x) p <<= 1; return p; }

static int env_int(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}

// Tiles in flight per CTA: as many as shared memory (227 KB), TMEM (512 columns) and the 1024-thread CTA
// limit allow, from the compiled set; NGP_MLP_SLOTS_FW / NGP_MLP_SLOTS_BW override (tuning only).
static int best_compiled(int cap, const int* compiled, int n_compiled) {
  int s = 0;
  for (int i = 0; i < n_compiled; i++) if (compiled[i] <= cap && compiled[i] > s) s = compiled[i];
  return s;
}
static void pick_slots(MlpCfg& c, bool backward, const int* compiled, int n_compiled) {
  int cap = kMaxSlots;
  const int by_tmem = (int)((512u - c.wg_cols) / c.acc_cols);
  if (cap > by_tmem) cap = by_tmem;
  const int forced = env_int(backward ? "NGP_MLP_SLOTS_BW" : "NGP_MLP_SLOTS_FW");
  if (forced > 0 && forced < cap) cap = forced;
  const uint32_t room = 227u * 1024u - c.slot_base;
  const int plain = best_compiled(cap < (int)(room / c.slot_bytes) ? cap : (int)(room / c.slot_bytes), compiled, n_compiled);
  const int with_raw = best_compiled(cap < (int)(room / (c.slot_bytes + c.raw_total)) ? cap : (int)(room / (c.slot_bytes + c.raw_total)), compiled, n_compiled);
  // The landing zone takes the input rows off the LSU/L1 path (thread-per-row loads cost one L1 tag lookup per
  // lane: at >= 128 B of fp32 per row the density net was L1-bound, ncu r01 l1tex 73 %), extra slots buy tiles in
  // flight; measured on B200 (tools/mlp_sweep.py): wide rows want the zone, narrow rows (colour net: 76 B) the
  // slots.  NGP_MLP_BULK=0/1 forces the choice (tuning only).
  const int env_bulk = getenv("NGP_MLP_BULK") ? env_int("NGP_MLP_BULK") : -1;
  uint32_t row_bytes = 0;
  for (int s = 0; s < c.n_seg; s++) row_bytes += c.raw_bytes[s] / kTile;
  bool use_raw = c.bulk && with_raw >= 2 && row_bytes >= 128;
  if (env_bulk == 0) use_raw = false;
  if (env_bulk == 1 && c.bulk && with_raw >= 1) use_raw = true;
  c.bulk = use_raw ? 1 : 0;
  if (use_raw) c.slot_bytes += c.raw_total;
  const int s = use_raw ? with_raw : (plain > 0 ? plain : 1);
  c.slots = s;
  c.wg_base = (uint32_t)s * c.acc_cols;
  c.tm_cols = next_pow2_cols(c.wg_base + c.wg_cols);
  c.smem_bytes = c.slot_base + (uint32_t)s * c.slot_bytes;
}

static int build_cfg(MlpCfg& c, int n_seg, const float* const* seg_ptr, const int* seg_w, const int* seg_kind, const int64_t* seg_stride, int width,
                     int n_hidden, int n_out, int act_hidden, int act_out, bool backward) {
  memset(&c, 0, sizeof(c));
  c.n_seg = n_seg;
  for (int s = 0; s < n_seg && s < kMaxSeg; s++) { c.seg_w[s] = seg_w[s]; c.seg_kind[s] = seg_kind[s]; c.seg_stride[s] = seg_stride[s]; }
  c.w = width; c.nh = n_hidden; c.no = n_out; c.act_h = act_hidden; c.act_o = act_out;
  c.bulk = 1;
  for (int s = 0; s < n_seg && s < kMaxSeg; s++) {
    const int raw_w = seg_kind[s] == kSegSH4 ? 3 : seg_w[s];
    if (seg_stride[s] != raw_w || (((uintptr_t)seg_ptr[s]) & 15) != 0) c.bulk = 0;
  }
  return finalize_cfg(c, backward);
}

// Persistent grid: SMs x co-resident CTAs (normally 1: a CTA's slots take the whole SM's TMEM / most of its
// shared memory; small forced slot counts leave room for several CTAs).
static int launch_grid(const void* fn, const MlpCfg& c, int64_t n) {
  int occ = (int)((227 * 1024) / (c.smem_bytes + 1024));
  const int by_tmem = 512 / (int)c.tm_cols;
  if (occ > by_tmem) occ = by_tmem;
  const int threads = kSlotThreads * c.slots;
  if (occ > 2048 / threads) occ = 2048 / threads;
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, fn) == cudaSuccess && fa.numRegs > 0) {
    const int regs = (fa.numRegs + 7) / 8 * 8;
    const int by_regs = 65536 / (regs * threads);
    if (occ > by_regs) occ = by_regs;
  }
  if (occ < 1) occ = 1;
  int64_t g = (int64_t)kSMs * occ;
  const int64_t groups = ((n + kTile - 1) / kTile + c.slots - 1) / c.slots;
  if (g > groups) g = groups;
  return (int)(g < 1 ? 1 : g);
}

static const int kFwSlots[] = {1, 2, 3, 4, 5, 6, 8};
static const int kBwSlots[] = {1, 2, 3, 4, 5, 6};

}  // namespace ngp

using namespace ngp;

// Number of fp32 parameters of the MLP (layout in the header of this file).
NGP_API int64_t ngp_mlp_param_count(int n_input, int width, int n_hidden, int n_out) {
  if (n_input < 1 || width < 1 || n_hidden < 1 || n_out < 1) return -1;
  const int64_t nop = (n_out + 15) / 16 * 16;
  return (int64_t)width * n_input + (int64_t)(n_hidden - 1) * width * width + nop * width;
}

#define NGP_MLP_DISPATCH(SLOTS_, KERNEL_, SH_, ...)                                                   \
  case SLOTS_: {                                                                                      \
    const void* fn = (const void*)KERNEL_<SLOTS_, SH_>;                                               \
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem_bytes); \
    if (e != cudaSuccess) return set_error(e, #KERNEL_ "/attr");                                      \
    const int grid = launch_grid(fn, c, n);                                                           \
    KERNEL_<SLOTS_, SH_><<<grid, kSlotThreads * SLOTS_, c.smem_bytes, (cudaStream_t)stream>>>(__VA_ARGS__); \
    launched = true;                                                                                  \
  } break;

// 0 generic, 1 SigmaShape, 2 RgbShape, 3 SigmaTilesShape
static int match_shape(const MlpCfg& c) {
  if (env_int("NGP_MLP_GENERIC") > 0) return 0;
  if (c.w != 64 || c.act_h != kActReLU) return 0;
  if (c.n_seg == 1 && c.seg_kind[0] == kSegPlain && c.seg_w[0] == 32 && c.nh == 1 && c.no == 16 && c.act_o == kActNone) return 1;
  if (c.n_seg == 1 && c.seg_kind[0] == kSegTiles && c.seg_w[0] == 32 && c.nh == 1 && c.no == 16 && c.act_o == kActNone) return 3;
  if (c.n_seg == 2 && c.seg_kind[0] == kSegSH4 && c.seg_kind[1] == kSegPlain && c.seg_w[1] == 16 && c.nh == 2 && c.no == 3 &&
      c.act_o == kActSigmoid) return 2;
  return 0;
}

// out (N, n_out) = MLP(cat(segments)).  Segment kinds: 0 = fp32 rows of seg_width floats at
// seg_ptr + row*seg_stride; 1 = degree-4 SH of the normalised (N,3) direction at seg_ptr (width 16); 2 = bf16 feature
// tiles written by ngp_hashgrid_fw_tiles (must be the only segment; its input gradient comes back as gradient tiles).
// Activations: 0 none, 1 ReLU, 2 sigmoid, 3 exp.  aux_exp_out (optional, N floats) = exp(out[:,0]) taken
// before the output activation (the ngp_pl density head).
// n_dev (optional, device): the live row count, n then being an upper bound the launch is sized for (ngp_render_round_compact)
int ngp::mlp_fw_launch(int n_seg, const float* const* seg_ptr, const int* seg_width, const int* seg_kind,
                       const int64_t* seg_stride, const float* params, int width, int n_hidden, int n_out,
                       int act_hidden, int act_out, int64_t n, const int32_t* n_dev, float* out, int64_t out_stride, float* aux_exp_out,
                       void* stream) {
  if (n <= 0) return 0;
  MlpCfg c;
  const int rc = build_cfg(c, n_seg, seg_ptr, seg_width, seg_kind, seg_stride, width, n_hidden, n_out, act_hidden, act_out, false);
  if (rc) { char b[128]; snprintf(b, sizeof b, "ngp_mlp_fw: unsupported MLP shape (code %d)", rc); return set_error_msg(b); }
  pick_slots(c, false, kFwSlots, (int)(sizeof(kFwSlots) / sizeof(int)));
  c.n_dev = n_dev;
  c.coop_stage = getenv("NGP_MLP_STAGE_COOP_FW") ? env_int("NGP_MLP_STAGE_COOP_FW") : 1;
  SegPtrs in; for (int s = 0; s < kMaxSeg; s++) in.p[s] = s < n_seg ? seg_ptr[s] : nullptr;
  bool launched = false;
  const int shape = match_shape(c);
#define FW_ARGS c, in, params, n, out, out_stride, aux_exp_out
#define FW_STATIC(SH_) switch (c.slots) {                \
    NGP_MLP_DISPATCH(4, mlp_fw_kernel, SH_, FW_ARGS)     \
    NGP_MLP_DISPATCH(5, mlp_fw_kernel, SH_, FW_ARGS)     \
    NGP_MLP_DISPATCH(6, mlp_fw_kernel, SH_, FW_ARGS)     \
    NGP_MLP_DISPATCH(8, mlp_fw_kernel, SH_, FW_ARGS)     \
    default: break; }
  if (shape == 1) { FW_STATIC(SigmaShape) } else if (shape == 2) { FW_STATIC(RgbShape) } else if (shape == 3) { FW_STATIC(SigmaTilesShape) }
  if (!launched) switch (c.slots) {
    NGP_MLP_DISPATCH(1, mlp_fw_kernel, GenericShape, FW_ARGS)
    NGP_MLP_DISPATCH(2, mlp_fw_kernel, GenericShape, FW_ARGS)
    NGP_MLP_DISPATCH(3, mlp_fw_kernel, GenericShape, FW_ARGS)
    NGP_MLP_DISPATCH(4, mlp_fw_kernel, GenericShape, FW_ARGS)
    NGP_MLP_DISPATCH(5, mlp_fw_kernel, GenericShape, FW_ARGS)
    NGP_MLP_DISPATCH(6, mlp_fw_kernel, GenericShape, FW_ARGS)
    NGP_MLP_DISPATCH(8, mlp_fw_kernel, GenericShape, FW_ARGS)
    default: return set_error_msg("ngp_mlp_fw: no kernel for this slot count");
  }
  NGP_LAUNCH_CHECK("ngp_mlp_fw");
  return 0;
}
NGP_API int ngp_mlp_fw(int n_seg, const float* const* seg_ptr, const int* seg_width, const int* seg_kind,
                       const int64_t* seg_stride, const float* params, int width, int n_hidden, int n_out,
                       int act_hidden, int act_out, int64_t n, float* out, int64_t out_stride, float* aux_exp_out,
                       void* stream) {
  return ngp::mlp_fw_launch(n_seg, seg_ptr, seg_width, seg_kind, seg_stride, params, width, n_hidden, n_out, act_hidden, act_out, n, nullptr,
                            out, out_stride, aux_exp_out, stream);
}

// dparams (+=, fp32 atomics; caller zeroes) and optional per-segment input gradients
// dseg_ptr[s] (N, seg_width[s]) with row stride dseg_stride[s] (NULL = not needed; SH segments
// never receive one).  dL_dout is (N, n_out) with row stride dout_stride.  saved_out (optional): the forward's
// output (N, n_out), row stride saved_out_stride — lets the kernel skip recomputing the output layer.
NGP_API int ngp_mlp_bw(int n_seg, const float* const* seg_ptr, const int* seg_width, const int* seg_kind,
                       const int64_t* seg_stride, const float* params, int width, int n_hidden, int n_out,
                       int act_hidden, int act_out, int64_t n, const float* dL_dout, int64_t dout_stride,
                       float* dparams, float* const* dseg_ptr, const int64_t* dseg_stride, const float* dL_daux_exp,
                       const float* saved_out, int64_t saved_out_stride, void* stream) {
  if (n <= 0) return 0;
  MlpCfg c;
  const int rc = build_cfg(c, n_seg, seg_ptr, seg_width, seg_kind, seg_stride, width, n_hidden, n_out, act_hidden, act_out, true);
  if (rc) { char b[128]; snprintf(b, sizeof b, "ngp_mlp_bw: unsupported MLP shape (code %d)", rc); return set_error_msg(b); }
  pick_slots(c, true, kBwSlots, (int)(sizeof(kBwSlots) / sizeof(int)));
  // Wide-row staging stays per thread in the backward kernel.  Measured on the reference heads at 14 M samples (tools/step_profile_ngp.py,
  // profiles/r02d_mlp_wide_rows.txt): cooperative staging gains 0.75 ms per forward launch, but with the backward's 2 slots of 128
  // threads a thread's 32 independent row loads hide more latency than warp-wide batches do (+0.1 ms), and routing the input
  // gradient through a transposed whole-row store cost +0.8 ms against the direct 64-byte runs (tried, removed).
  c.coop_stage = getenv("NGP_MLP_STAGE_COOP_BW") ? env_int("NGP_MLP_STAGE_COOP_BW") : 0;
  SegPtrs in; SegGrads dg;
  for (int s = 0; s < kMaxSeg; s++) {
    in.p[s] = s < n_seg ? seg_ptr[s] : nullptr;
    dg.p[s] = (s < n_seg && dseg_ptr && seg_kind[s] != kSegSH4) ? dseg_ptr[s] : nullptr;
    dg.stride[s] = (s < n_seg && dseg_stride) ? dseg_stride[s] : 0;
  }
  // a feature-tile input gets its gradient back as gradient tiles (hashgrid.cu DYT layout): ceil(N/128)*128*k0p floats
  dg.tiles = (n_seg == 1 && seg_kind[0] == kSegTiles) ? 1 : 0;
  bool launched = false;
  const int shape = match_shape(c);
#define BW_ARGS c, in, params, n, dL_dout, dout_stride, dparams, dg, dL_daux_exp, saved_out, saved_out_stride
#define BW_STATIC(SH_) switch (c.slots) {                \
    NGP_MLP_DISPATCH(2, mlp_bw_kernel, SH_, BW_ARGS)     \
    NGP_MLP_DISPATCH(3, mlp_bw_kernel, SH_, BW_ARGS)     \
    NGP_MLP_DISPATCH(4, mlp_bw_kernel, SH_, BW_ARGS)     \
    NGP_MLP_DISPATCH(5, mlp_bw_kernel, SH_, BW_ARGS)     \
    NGP_MLP_DISPATCH(6, mlp_bw_kernel, SH_, BW_ARGS)     \
    default: break; }
  if (shape == 1) { BW_STATIC(SigmaShape) } else if (shape == 2) { BW_STATIC(RgbShape) } else if (shape == 3) { BW_STATIC(SigmaTilesShape) }
  if (!launched) switch (c.slots) {
    NGP_MLP_DISPATCH(1, mlp_bw_kernel, GenericShape, BW_ARGS)
    NGP_MLP_DISPATCH(2, mlp_bw_kernel, GenericShape, BW_ARGS)
    NGP_MLP_DISPATCH(3, mlp_bw_kernel, GenericShape, BW_ARGS)
    NGP_MLP_DISPATCH(4, mlp_bw_kernel, GenericShape, BW_ARGS)
    NGP_MLP_DISPATCH(5, mlp_bw_kernel, GenericShape, BW_ARGS)
    NGP_MLP_DISPATCH(6, mlp_bw_kernel, GenericShape, BW_ARGS)
    default: return set_error_msg("ngp_mlp_bw: no kernel for this slot count");
  }
  NGP_LAUNCH_CHECK("ngp_mlp_bw");
  return 0;
}
