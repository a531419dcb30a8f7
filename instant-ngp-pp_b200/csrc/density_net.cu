// The reference's density net on the 5th-gen tensor cores: ONE kernel per direction for
//     z1 = W1 e + b1 ; a1 = softplus(z1) ; z2 = w2 . a1 + b2 ; sigma = softplus(z2)          (xyz_net + sigma_act)
//     g_e = d sigma / d e = W1^T (s2 * s1 * w2),  s1 = sigmoid(z1), s2 = sigmoid(z2)          (autograd normals)
// and for the backward of BOTH outputs (sigma and g_e: the double backward the reference obtains from
// torch.autograd.grad(sigmas, x, create_graph=True) + loss.backward(), models/networks.py:54-59,172-196).
//
// The reference runs this net as torch ops: 2 Linear layers + ~30 elementwise kernels on (S,128) tensors per step.  Round 1
// of this library fused the elementwise stages (density_head.cu) and left six (S,128)x(128,128) GEMMs on cuBLAS TF32 with
// z1 / t / dz1 / v round-tripping HBM as fp32 matrices (2 KB/sample saved for backward, 20 % of the playground-shaped step).
// Here the 128x128 weight matrix lives in shared memory as ONE bf16 operand tile that serves every product (K-major for
// e W1^T and dg W1^T, MN-major for u W1 and dz1 W1), the accumulators live in TMEM, and the activations never leave the SM:
//
//   forward  (256 threads, up to 3 CTAs / SM, one 128-sample tile at a time per CTA):
//     e rows (coalesced fp32) -> bf16 tile X -> MMA Z1 = X W1^T -> epilogue: a1, s1; z2 row sums; u = s1*w2 written over X
//     -> MMA G = U W1 -> epilogue: g_e = s2 * G  (the row scale s2 commutes with the product, so t = s2*u is never built).
//   backward (512 threads, 1 CTA / SM): with upstream (dsigma, dg) and the forward's saved s2, g_e:
//     dz2  = (1 - s2) * (dg . g_e) + dsigma * s2          [ds2*s2*(1-s2) with ds2 = (dg W1^T).(s1*w2) = (dg . g_e)/s2: no pass
//                                                           over the columns is needed before the main one, and no division]
//     MMAs Z1 = X W1^T, V = DG W1^T  -> epilogue per element:  uv = s1*v ;  dz1 = w2*(uv*s2*(1-s1) + dz2*s1) ;  t = s2*s1*w2
//     -> MMA DE = dZ1 W1 (-> global), and into CTA-lifetime TMEM accumulators:  dW1 += dZ1^T X + T^T DG ,  db1 += dZ1^T 1
//     (a tile of ones: the bias gradient is one more small MMA instead of 128 shuffle reductions per tile);
//     dw2 = sum_rows (s2*uv + dz2*a1) is kept in registers (a thread's columns are the same for every tile) and reduced once.
// bf16 operands, fp32 accumulation — the precision class of every other head of the field (mlp.cu); the TF32 / fp32 torch
// path stays selectable on the host side (networks.py NGP.density_net) and is what the parity tests compare against.
//
// Bounds at S samples: forward reads 512 B + writes 516 B per sample (HBM 2.3 ms at 14 M samples), 3 MUFU per hidden
// unit; backward reads 3 x 512 B and writes 512 B.  Row-major fp32 matrices at the boundary (what the hash-grid kernels on
// either side exchange); the per-thread row stores of g_e / de are the L1-tag-bound part (32 lines per store instruction).
#include "common.cuh"
#include "tc05.cuh"
#include "mlp_tiles.cuh"
#include <stdio.h>

namespace ngp {

constexpr int kDn = 128;                                        // input width == hidden width (networks.py:54-59 at L*F = 128)
constexpr uint32_t kDnTile = 16u * (128u * 16u + 64u);          // one 128 x 128 bf16 operand tile (mlp_tiles.cuh layout): 33 792 B
constexpr uint32_t kOffB1 = 128, kOffW2 = 640, kOffScr = 1152;  // fp32 b1[128], w2[128], scratch[512]
constexpr uint32_t kOffW1 = 3200;                               // W1 tile
constexpr uint32_t kOffT0 = kOffW1 + kDnTile;                   // first sample tile
constexpr uint32_t kFwScratch = 8u * 32u * 36u * 4u;             // g_e transpose scratch (8 warps x 32 rows x (32+4) floats): aliases the X tile
constexpr uint32_t kFwSmem = kOffT0 + kFwScratch;               // 73 856 B: 3 CTAs per SM
constexpr uint32_t kOffX = kOffT0, kOffDG = kOffX + kDnTile, kOffT = kOffDG + kDnTile, kOffDZ = kOffT + kDnTile;
constexpr uint32_t kOffOnes = kOffDZ + kDnTile;                 // 128 x 16 tile of ones
constexpr uint32_t kOffBwScr = kOffOnes + 2u * (128u * 16u + 64u);
constexpr uint32_t kBwSmem = kOffBwScr + 16u * 32u * 20u * 4u;  // de transpose scratch (16 warps x 32 rows x (16+4) floats): 217 344 B
// TMEM columns of the backward kernel
constexpr uint32_t kColZ1 = 0, kColV = 128, kColDE = 0, kColDW1 = 256, kColDB1 = 384;

__device__ __forceinline__ void st_quad_bf16(uint8_t* tile, uint32_t r, uint32_t c, float4 v) {
  uint2 q;
  q.x = pack_bf16(v.x, v.y); q.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(tile + toff(kTile, r, c)) = q;
}

// CTA prologue shared by both kernels: barriers, TMEM, b1 / w2 / W1 -> shared memory
__device__ __forceinline__ uint32_t dn_setup(uint8_t* smem, const float* __restrict__ W1, const float* __restrict__ b1,
                                             const float* __restrict__ w2, uint32_t tm_cols, uint32_t warp) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 32);
  if (threadIdx.x == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); mbar_fence_init(); }
  if (warp == 0) { __syncwarp(); tmem_alloc(tslot, tm_cols); }
  float* sb1 = reinterpret_cast<float*>(smem + kOffB1);
  float* sw2 = reinterpret_cast<float*>(smem + kOffW2);
  for (int i = threadIdx.x; i < kDn; i += blockDim.x) { sb1[i] = __ldg(b1 + i); sw2[i] = __ldg(w2 + i); }
  uint8_t* Wt = smem + kOffW1;
  for (int i = threadIdx.x; i < kDn * 16; i += blockDim.x) {           // (row, 8-column chunk) units of the (out, in) matrix
    const int r = i >> 4, c8 = (i & 15) * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(W1 + r * kDn + c8)), b = __ldg(reinterpret_cast<const float4*>(W1 + r * kDn + c8) + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    st_chunk(Wt, kTile, r, c8, v);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  return *tslot;
}

// =============================================================================================== forward
__global__ void __launch_bounds__(256, 3) density_net_fw_kernel(const float* __restrict__ e, const float* __restrict__ W1,
                                                             const float* __restrict__ b1, const float* __restrict__ w2,
                                                             const float* __restrict__ b2p, int64_t n, float* __restrict__ sigma,
                                                             float* __restrict__ s2_out, float* __restrict__ ge) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t lane = threadIdx.x & 31u, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t tmem = __shfl_sync(0xffffffffu, dn_setup(smem, W1, b1, w2, 128, warp), 0);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  const float* sb1 = reinterpret_cast<const float*>(smem + kOffB1);
  const float* sw2 = reinterpret_cast<const float*>(smem + kOffW2);
  float* scr = reinterpret_cast<float*>(smem + kOffScr);
  uint8_t* Xt = smem + kOffT0;
  const uint32_t xaddr = smem_u32(Xt), waddr = smem_u32(smem + kOffW1);
  const uint32_t q = warp & 3u, h = warp >> 2;                   // TMEM lane quadrant; column half
  const uint32_t trow = tmem + ((q * 32u) << 16);
  const uint32_t r_t = q * 32u + lane;                           // this thread's row of the tile in the epilogues
  const float b2 = __ldg(b2p);
  uint32_t phase = 0;
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTile;
    // ---- stage X: a warp reads whole 512-byte rows (4 lines per load instruction), 8 rows in flight
#pragma unroll
    for (int rr = 0; rr < 16; rr += 8) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int64_t g = row0 + warp + 8 * (rr + j);
        v[j] = g < n ? __ldcs(reinterpret_cast<const float4*>(e + g * kDn) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 8; j++) st_quad_bf16(Xt, warp + 8 * (rr + j), lane * 4, v[j]);
    }
    {   // this CTA's next tile -> L2 while the current one is computed
      const int64_t nrow0 = row0 + (int64_t)gridDim.x * kTile;
#pragma unroll
      for (int k2 = 0; k2 < 16; k2 += 8) {
        const int64_t g = nrow0 + warp + 8 * (k2 + (int)(lane >> 2));
        if (g < n) prefetch_l2(e + g * kDn + (lane & 3u) * 32);
      }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      issue_fwd(tmem, xaddr, waddr, kTile, kDn, kDn);            // Z1 = X W1^T
      mma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, phase); phase ^= 1;
    fence_after_sync();
    // ---- epilogue A: softplus / sigmoid of the thread's 64 columns, z2 partial sum, u = s1 * w2 over X
    float dot = 0.f;
#pragma unroll
    for (int g = 0; g < 2; g++) {
      const uint32_t c0 = 64u * h + 32u * g;
      float z[32];
      tmem_ld32(trow + c0, z);
#pragma unroll
      for (int k = 0; k < 32; k += 8) {
        float u[8], w[8], b[8];
        *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(sw2 + c0 + k); *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(sw2 + c0 + k + 4);
        *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(sb1 + c0 + k); *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(sb1 + c0 + k + 4);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const SpSg a = softplus_sigmoid(z[k + i] + b[i]);
          dot = fmaf(a.sp, w[i], dot);
          u[i] = a.sg * w[i];
        }
        if (ge) st_chunk(Xt, kTile, r_t, c0 + k, u);
      }
    }
    scr[h * kTile + r_t] = dot;
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (ge && warp == 0 && elect_one()) {
      fence_after_sync();
      issue_dgrad(tmem, xaddr, waddr, kDn, kDn, kDn);            // G = U W1
      mma_commit(bar);
    }
    __syncwarp();
    const SpSg o = softplus_sigmoid(scr[r_t] + scr[kTile + r_t] + b2);
    const int64_t grow = row0 + r_t;
    if (h == 0 && grow < n) { sigma[grow] = o.sp; s2_out[grow] = o.sg; }
    if (ge) {
      mbar_wait(bar, phase); phase ^= 1;
      fence_after_sync();
      // G is complete, so the X / U tile is dead: it is the transpose scratch of the coalesced g_e store
      float* scratch = reinterpret_cast<float*>(Xt) + warp * (32 * 36);
      const int64_t wrow0 = row0 + q * 32;                       // first row of this warp's quadrant
      const int rows_valid = (int)(n - wrow0 < 32 ? (n - wrow0 < 0 ? 0 : n - wrow0) : 32);
#pragma unroll
      for (int g = 0; g < 2; g++) {
        const uint32_t c0 = 64u * h + 32u * g;
        float v[32];
        tmem_ld32(trow + c0, v);
#pragma unroll
        for (int k = 0; k < 32; k++) v[k] *= o.sg;
        store_rows_coalesced<32>(scratch, v, ge + wrow0 * kDn + c0, kDn, rows_valid, lane);
      }
      __syncthreads();                                           // the scratch is the next tile's X
    }
    // the next tile's first barrier orders these TMEM / scratch reads before its MMA / scratch writes
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// =============================================================================================== backward
__global__ void __launch_bounds__(512) density_net_bw_kernel(const float* __restrict__ e, const float* __restrict__ dge,
                                                             const float* __restrict__ ge, const float* __restrict__ dsigma,
                                                             const float* __restrict__ s2, const float* __restrict__ W1,
                                                             const float* __restrict__ b1, const float* __restrict__ w2, int64_t n,
                                                             float* __restrict__ de, float* __restrict__ dW1, float* __restrict__ db1,
                                                             float* __restrict__ dw2, float* __restrict__ db2) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t lane = threadIdx.x & 31u, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t tmem = __shfl_sync(0xffffffffu, dn_setup(smem, W1, b1, w2, 512, warp), 0);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* bar2 = bar + 1;
  const float* sb1 = reinterpret_cast<const float*>(smem + kOffB1);
  const float* sw2 = reinterpret_cast<const float*>(smem + kOffW2);
  float* rowdot = reinterpret_cast<float*>(smem + kOffScr);
  uint8_t *Xt = smem + kOffX, *DGt = smem + kOffDG, *Tt = smem + kOffT, *DZt = smem + kOffDZ, *Ones = smem + kOffOnes;
  const uint32_t xaddr = smem_u32(Xt), dgaddr = smem_u32(DGt), taddr = smem_u32(Tt), dzaddr = smem_u32(DZt), oaddr = smem_u32(Ones);
  const uint32_t waddr = smem_u32(smem + kOffW1);
  const uint32_t q = warp & 3u, h = warp >> 2;                   // TMEM lane quadrant; column quarter (32 columns)
  const uint32_t trow = tmem + ((q * 32u) << 16);
  const uint32_t r_t = q * 32u + lane;
  const bool has_v = dge != nullptr;

  // ones tile; zero the CTA-lifetime accumulators (dW1: 128 columns, db1: 16)
  for (uint32_t i = threadIdx.x; i < 2u * (128u * 16u + 64u) / 4u; i += blockDim.x) reinterpret_cast<uint32_t*>(Ones)[i] = 0x3F803F80u;
  if (h == 0) for (uint32_t c0 = kColDW1; c0 < kColDB1 + 16; c0 += 16) tmem_st16_zero(trow + c0);
  tmem_wait_st();
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  float aw[32];                                                  // dw2 partial sums of this thread's 32 columns over all its rows
#pragma unroll
  for (int i = 0; i < 32; i++) aw[i] = 0.f;
  float adz2 = 0.f;
  uint32_t phase = 0, phase2 = 0;
  bool pending = false;                                          // wgrad MMAs of the previous tile still reading the sample tiles
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTile;
    const int64_t grow = row0 + r_t;
    const bool valid = grow < n;
    const float sg = valid ? __ldg(s2 + grow) : 0.f;
    const float dsig = (valid && dsigma) ? __ldg(dsigma + grow) : 0.f;
    if (pending) { mbar_wait(bar2, phase2); phase2 ^= 1; }
    // ---- stage X (and DG, and the row dot dg . g_e): a warp per row, 8 rows per warp, 4 in flight
#pragma unroll
    for (int rr = 0; rr < 8; rr += 4) {
      float4 xe[4], xd[4], xg[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int64_t g = row0 + warp + 16 * (rr + j);
        const bool ok = g < n;
        xe[j] = ok ? __ldcs(reinterpret_cast<const float4*>(e + g * kDn) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        xd[j] = (ok && has_v) ? __ldcs(reinterpret_cast<const float4*>(dge + g * kDn) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        xg[j] = (ok && has_v) ? __ldcs(reinterpret_cast<const float4*>(ge + g * kDn) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t r = warp + 16 * (rr + j);
        st_quad_bf16(Xt, r, lane * 4, xe[j]);
        if (has_v) {
          st_quad_bf16(DGt, r, lane * 4, xd[j]);
          const float d = warp_sum(xd[j].x * xg[j].x + xd[j].y * xg[j].y + xd[j].z * xg[j].z + xd[j].w * xg[j].w);
          if (lane == 0) rowdot[r] = d;
        }
      }
    }
    {   // next tile's rows -> L2 while this one is computed (this CTA is the only one on the SM)
      const int64_t nrow = row0 + (int64_t)gridDim.x * kTile + warp * 8 + (lane >> 2);
      if (nrow < n) {
        prefetch_l2(e + nrow * kDn + (lane & 3) * 32);
        if (has_v) { prefetch_l2(dge + nrow * kDn + (lane & 3) * 32); prefetch_l2(ge + nrow * kDn + (lane & 3) * 32); }
      }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      issue_fwd(tmem + kColZ1, xaddr, waddr, kTile, kDn, kDn);                 // Z1 = X W1^T
      if (has_v) issue_fwd(tmem + kColV, dgaddr, waddr, kTile, kDn, kDn);      // V  = DG W1^T
      mma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, phase); phase ^= 1;
    fence_after_sync();
    // ---- epilogue: dz1 and t of the thread's 32 columns
    const float dz2 = valid ? (has_v ? (1.f - sg) * rowdot[r_t] : 0.f) + dsig * sg : 0.f;
    if (h == 0) adz2 += dz2;
#pragma unroll
    for (int g = 0; g < 2; g++) {
      const uint32_t c0 = 32u * h + 16u * g;
      float z[16], v[16];
      tmem_ld16(trow + kColZ1 + c0, z);
      if (has_v) tmem_ld16(trow + kColV + c0, v);
#pragma unroll
      for (int i8 = 0; i8 < 16; i8 += 8) {
        float dz[8], tt[8];
#pragma unroll
        for (int i4 = 0; i4 < 8; i4 += 4) {            // one LDS.128 per four columns (the MIO queue is shared with the MUFU ops)
          const float4 w4 = *reinterpret_cast<const float4*>(sw2 + c0 + i8 + i4), b4 = *reinterpret_cast<const float4*>(sb1 + c0 + i8 + i4);
          const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int i = i8 + i4 + j;
            const float w = wv[j];
            const SpSg a = softplus_sigmoid(z[i] + bv[j]);
            const float uv = has_v ? a.sg * v[i] : 0.f;
            dz[i4 + j] = valid ? w * (uv * sg * (1.f - a.sg) + dz2 * a.sg) : 0.f;
            tt[i4 + j] = sg * a.sg * w;                          // sg = 0 on rows past the end
            aw[16 * g + i] += valid ? fmaf(dz2, a.sp, sg * uv) : 0.f;
          }
        }
        st_chunk(DZt, kTile, r_t, c0 + i8, dz);
        if (has_v) st_chunk(Tt, kTile, r_t, c0 + i8, tt);
      }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      issue_dgrad(tmem + kColDE, dzaddr, waddr, kDn, kDn, kDn);                // DE = dZ1 W1   (Z1 columns: consumed above)
      mma_commit(bar);
      issue_wgrad(tmem + kColDW1, dzaddr, xaddr, kDn, kDn);                    // dW1 += dZ1^T X
      if (has_v) issue_wgrad(tmem + kColDW1, taddr, dgaddr, kDn, kDn);         //      + T^T DG
      issue_wgrad(tmem + kColDB1, dzaddr, oaddr, kDn, 16);                     // db1 += dZ1^T 1
      mma_commit(bar2);
    }
    __syncwarp();
    pending = true;
    mbar_wait(bar, phase); phase ^= 1;
    fence_after_sync();
    if (de) {
      float* scratch = reinterpret_cast<float*>(smem + kOffBwScr) + warp * (32 * 20);
      const int64_t wrow0 = row0 + q * 32;
      const int rows_valid = (int)(n - wrow0 < 32 ? (n - wrow0 < 0 ? 0 : n - wrow0) : 32);
#pragma unroll
      for (int g = 0; g < 2; g++) {
        const uint32_t c0 = 32u * h + 16u * g;
        float v[16];
        tmem_ld16(trow + kColDE + c0, v);
        store_rows_coalesced<16>(scratch, v, de + wrow0 * kDn + c0, kDn, rows_valid, lane);
      }
    }
  }
  if (pending) { mbar_wait(bar2, phase2); phase2 ^= 1; }
  fence_after_sync();
  // ---- flush: TMEM-resident dW1 / db1, register-resident dw2 / db2
  if (pending) {
    const int m = (int)r_t;                                      // accumulator row = output unit
#pragma unroll
    for (int g = 0; g < 2; g++) {
      const uint32_t c0 = 32u * h + 16u * g;
      float v[16];
      tmem_ld16(trow + kColDW1 + c0, v);
#pragma unroll
      for (int i = 0; i < 16; i++) atomicAdd(dW1 + m * kDn + c0 + i, v[i]);
    }
    if (h == 0) {
      float v[16];
      tmem_ld16(trow + kColDB1, v);
      atomicAdd(db1 + m, v[0]);
    }
#pragma unroll
    for (int i = 0; i < 32; i++) {
      const float s = warp_sum(aw[i]);
      if (lane == 0) atomicAdd(dw2 + 32 * h + i, s);
    }
    if (h == 0) {
      const float s = warp_sum(adz2);
      if (lane == 0) atomicAdd(db2, s);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace ngp

using namespace ngp;

static int dn_check(const char* who, int n_in, int width, const void* a, const void* b, const void* c, const void* d) {
  if (n_in != kDn || width != kDn) { char m[160]; snprintf(m, sizeof m, "%s: the tensor-core density net is built for 128 -> 128 -> 1 (got %d -> %d)", who, n_in, width); return set_error_msg(m); }
  if ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c) | ((uintptr_t)d)) & 15) { char m[160]; snprintf(m, sizeof m, "%s: row-major matrices must be 16-byte aligned", who); return set_error_msg(m); }
  return 0;
}

// sigma (N) = Softplus(Linear(128,1)(Softplus(Linear(128,128)(e))))  and, when g_e != NULL, g_e (N,128) = d sigma / d e
// (models/networks.py:54-59 xyz_net, :181 sigma_act, :186-196 autograd normals up to the encoder).  e (N,128) row-major fp32,
// W1 (128,128) row-major (out,in), b1 (128), w2 (128), b2 (1).  s2 (N) = sigmoid of the output pre-activation: saved for the backward.
NGP_API int ngp_density_net_fw(const float* e, const float* W1, const float* b1, const float* w2, const float* b2, int64_t n, int n_in,
                               int width, float* sigma, float* s2, float* g_e, void* stream) {
  if (n <= 0) return 0;
  if (int rc = dn_check("ngp_density_net_fw", n_in, width, e, W1, g_e, nullptr)) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t er = cudaFuncSetAttribute(density_net_fw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwSmem);
    if (er != cudaSuccess) return set_error(er, "ngp_density_net_fw/attr");
    attr = true;
  }
  const int64_t tiles = (n + kTile - 1) / kTile;
  const int grid = (int)(tiles < (int64_t)kSMs * 3 ? tiles : (int64_t)kSMs * 3);
  density_net_fw_kernel<<<grid, 256, kFwSmem, (cudaStream_t)stream>>>(e, W1, b1, w2, b2, n, sigma, s2, g_e);
  NGP_LAUNCH_CHECK("ngp_density_net_fw");
  return 0;
}

// Backward of both outputs.  Upstream: dsigma (N) | NULL, d_ge (N,128) | NULL (then g_e, the forward's output, is required).
// Out: de (N,128) | NULL = dL/de; += into dW1 (128,128), db1 (128), dw2 (128), db2 (1) (caller zeroes).
NGP_API int ngp_density_net_bw(const float* e, const float* d_ge, const float* g_e, const float* dsigma, const float* s2, const float* W1,
                               const float* b1, const float* w2, int64_t n, int n_in, int width, float* de, float* dW1, float* db1,
                               float* dw2, float* db2, void* stream) {
  if (n <= 0) return 0;
  if (int rc = dn_check("ngp_density_net_bw", n_in, width, e, d_ge, g_e, de)) return rc;
  if (d_ge && !g_e) return set_error_msg("ngp_density_net_bw: d_ge needs the forward's g_e");
  static bool attr = false;
  if (!attr) {
    cudaError_t er = cudaFuncSetAttribute(density_net_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwSmem);
    if (er != cudaSuccess) return set_error(er, "ngp_density_net_bw/attr");
    attr = true;
  }
  const int64_t tiles = (n + kTile - 1) / kTile;
  const int grid = (int)(tiles < (int64_t)kSMs ? tiles : (int64_t)kSMs);
  density_net_bw_kernel<<<grid, 512, kBwSmem, (cudaStream_t)stream>>>(e, d_ge, g_e, dsigma, s2, W1, b1, w2, n, de, dW1, db1, dw2, db2);
  NGP_LAUNCH_CHECK("ngp_density_net_bw");
  return 0;
}
