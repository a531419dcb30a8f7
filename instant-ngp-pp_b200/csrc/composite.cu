// Front-to-back compositing over packed rays: training forward / backward, alpha-only, and the
// incremental test-time compositor.
// Replaces reference models/csrc/volumerendering.cu:5-63 (composite_alpha_fw_cu), :65-164
// (composite_train_fw_cu), :167-311 (composite_train_bw_cu), :314-423 (composite_test_fw_cu).
//
// Reference: one thread per ray, serial loop, (8+C) per-ray accumulators read-modify-written in
// GLOBAL memory at every sample, every output pre-zeroed by the host, an extra elementwise kernel
// plus an in-place per-thread sequential scan in the backward.
//
// Here: a group of G lanes owns a ray and consumes its packed samples G at a time (coalesced
// loads); transmittance is a multiplicative segmented scan, running sums are additive scans, the
// per-ray accumulators live in registers and are written once.  Every output element is written
// by the kernel (zeros after early termination), so no host-side fill is needed.
#include "scan.cuh"
#include <stdlib.h>

namespace ngp {

constexpr int kCP = 16;  // semantic classes handled per pass (register accumulators)
constexpr int kFwDepth = 4, kBwDepth = 2;   // chunks of G samples whose inputs are in flight together (forward / backward compositor)

// ------------------------------------------------------------------------------------------ fw
// do_main: also produce opacity/depth/rgb/normal_pred/ws/total_samples (first class chunk only).
template <int G, bool kTiled>
__global__ void __launch_bounds__(256) composite_train_fw_kernel(
    const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ normals,
    const float* __restrict__ sems, const float* __restrict__ deltas, const float* __restrict__ ts,
    const int64_t* __restrict__ rays_a, float T_thr, int classes, int c0, int nc, bool do_main, int64_t n_rays,
    int64_t* __restrict__ total_samples, float* __restrict__ opacity, float* __restrict__ depth,
    float* __restrict__ rgb, float* __restrict__ normal_pred, float* __restrict__ sem, float* __restrict__ ws, int ray_tile) {
  // an empty ray (most rays of a batch are): zeros, no reductions
  auto write_empty = [&](const Seg& sg) {
    const int64_t r = sg.ray;
    if (do_main) {
      total_samples[r] = 0;
      opacity[r] = 0.f; depth[r] = 0.f;
      rgb[3 * r] = 0.f; rgb[3 * r + 1] = 0.f; rgb[3 * r + 2] = 0.f;
      if (normal_pred) { normal_pred[3 * r] = 0.f; normal_pred[3 * r + 1] = 0.f; normal_pred[3 * r + 2] = 0.f; }
    }
    for (int i = 0; i < nc; i++) sem[r * classes + c0 + i] = 0.f;
  };
  for_each_ray<G, kTiled>(rays_a, n_rays, ray_tile, [&](const Seg& sg, const int j) {
  if (G == 32 && sg.n <= 0) {          // one warp per ray (kTiled == false): the warp's ray is empty
    if (j == 0 && sg.ray >= 0) write_empty(sg);
    return;
  }

  float aO = 0.f, aD = 0.f, aR = 0.f, aG = 0.f, aB = 0.f, aNx = 0.f, aNy = 0.f, aNz = 0.f;
  float aS[kCP];
#pragma unroll
  for (int i = 0; i < kCP; i++) aS[i] = 0.f;

  // kFwDepth chunks of G samples per trip: every input of the trip is requested before the first scan, so a ray of up to
  // kFwDepth * G samples (the mean non-empty ray of a Lego-shaped batch has 127) pays ONE memory round trip, not one (or, with
  // the weights' consumers loading where they are used, two) per chunk — the kernel is a chain of dependent round trips per ray,
  // far from the HBM rate (tools/composite_sweep.py).  A ray that has already terminated requests nothing: its remaining samples
  // only get ws = 0.
  TState st;
  for (int base = 0; warp_any(base < sg.n); base += kFwDepth * G) {
    float a[kFwDepth], t[kFwDepth], c_r[kFwDepth], c_g[kFwDepth], c_b[kFwDepth];
#pragma unroll
    for (int m = 0; m < kFwDepth; m++) {
      const int k = base + m * G + j;
      a[m] = 0.f; t[m] = 0.f; c_r[m] = 0.f; c_g[m] = 0.f; c_b[m] = 0.f;
      if (k < sg.n && !st.done) {
        const int64_t s = sg.start + k;
        const float sig = __ldg(sigmas + s), dl = __ldg(deltas + s);
        if (do_main) { t[m] = __ldg(ts + s); c_r[m] = __ldg(rgbs + 3 * s); c_g[m] = __ldg(rgbs + 3 * s + 1); c_b[m] = __ldg(rgbs + 3 * s + 2); }
        a[m] = sample_alpha(sig, dl);
      }
    }
#pragma unroll
    for (int m = 0; m < kFwDepth; m++) {
      const int cb = base + m * G;
      if (m > 0 && !warp_any(cb < sg.n)) break;
      const bool valid = cb + j < sg.n;
      const int64_t s = sg.start + cb + j;
      float n_x = 0.f, n_y = 0.f, n_z = 0.f;
      if (normals && do_main && valid && !st.done) { n_x = __ldg(normals + 3 * s); n_y = __ldg(normals + 3 * s + 1); n_z = __ldg(normals + 3 * s + 2); }
      float Tb, Ta; bool active;
      chunk_transmittance<G>(st, a[m], valid, j, cb, T_thr, Tb, Ta, active);
      const float w = active ? a[m] * Tb : 0.f;
      if (valid && do_main) ws[s] = w;
      if (active) {
        if (do_main) {
          aO += w;
          aD = fmaf(w, t[m], aD);
          aR = fmaf(w, c_r[m], aR); aG = fmaf(w, c_g[m], aG); aB = fmaf(w, c_b[m], aB);
          if (normals) { aNx = fmaf(w, n_x, aNx); aNy = fmaf(w, n_y, aNy); aNz = fmaf(w, n_z, aNz); }
        }
        if (nc > 0) {
          const float* sp = sems + s * classes + c0;
#pragma unroll
          for (int i = 0; i < kCP; i++) if (i < nc) aS[i] = fmaf(w, __ldg(sp + i), aS[i]);
        }
      }
    }
  }
  if (!st.done) st.n_done = sg.n;

  if (do_main) {
    aO = group_sum<G>(aO); aD = group_sum<G>(aD);
    aR = group_sum<G>(aR); aG = group_sum<G>(aG); aB = group_sum<G>(aB);
    aNx = group_sum<G>(aNx); aNy = group_sum<G>(aNy); aNz = group_sum<G>(aNz);
  }
  if (nc > 0) {
#pragma unroll
    for (int i = 0; i < kCP; i++) if (i < nc) aS[i] = group_sum<G>(aS[i]);
  }

  if (j == 0 && sg.ray >= 0) {
    const int64_t r = sg.ray;
    if (do_main) {
      total_samples[r] = st.n_done;
      opacity[r] = aO; depth[r] = aD;
      rgb[3 * r] = aR; rgb[3 * r + 1] = aG; rgb[3 * r + 2] = aB;
      if (normal_pred) { normal_pred[3 * r] = aNx; normal_pred[3 * r + 1] = aNy; normal_pred[3 * r + 2] = aNz; }
    }
    if (nc > 0) {
#pragma unroll
      for (int i = 0; i < kCP; i++) if (i < nc) sem[r * classes + c0 + i] = aS[i];
    }
  }
  }, write_empty);
}

template <int G, bool kTiled>
__global__ void __launch_bounds__(256) composite_alpha_fw_kernel(
    const float* __restrict__ sigmas, const float* __restrict__ deltas, const int64_t* __restrict__ rays_a,
    float T_thr, int64_t n_rays, float* __restrict__ alphas, float* __restrict__ ws, int ray_tile) {
  for_each_ray<G, kTiled>(rays_a, n_rays, ray_tile, [&](const Seg& sg, const int j) {
  TState st;
  for (int base = 0; warp_any(base < sg.n); base += G) {
    const bool valid = base + j < sg.n;
    const int64_t s = sg.start + base + j;
    float a = 0.f;
    if (valid) a = sample_alpha(__ldg(sigmas + s), __ldg(deltas + s));
    float Tb, Ta; bool active;
    chunk_transmittance<G>(st, a, valid, j, base, T_thr, Tb, Ta, active);
    if (valid) { alphas[s] = active ? a : 0.f; ws[s] = active ? a * Tb : 0.f; }
  }
  });
}

// ------------------------------------------------------------------------------------------ bw
// volumerendering.cu:212-245.  dsigma has NO normal / semantic terms (reference behaviour).
template <int G, bool kTiled>
__global__ void __launch_bounds__(256) composite_train_bw_kernel(
    const float* __restrict__ dL_dopacity, const float* __restrict__ dL_ddepth, const float* __restrict__ dL_drgb,
    const float* __restrict__ dL_dnormal, const float* __restrict__ dL_dsem, const float* __restrict__ dL_dws,
    const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ ws_saved,
    const float* __restrict__ deltas, const float* __restrict__ ts, const int64_t* __restrict__ rays_a,
    const float* __restrict__ opacity, const float* __restrict__ depth, const float* __restrict__ rgb,
    float T_thr, int classes, int64_t n_rays, float* __restrict__ dL_dsigmas, float* __restrict__ dL_drgbs,
    float* __restrict__ dL_dnormals, float* __restrict__ dL_dsems, int ray_tile) {
  for_each_ray<G, kTiled>(rays_a, n_rays, ray_tile, [&](const Seg& sg, const int j) {
  if (G == 32 && sg.n <= 0) return;          // nothing to write for an empty ray
  const int64_t r = sg.ray < 0 ? 0 : sg.ray;

  float gO = 0.f, gD = 0.f, gR = 0.f, gG = 0.f, gB = 0.f, gNx = 0.f, gNy = 0.f, gNz = 0.f;
  float O = 0.f, D = 0.f, R = 0.f, Gc = 0.f, B = 0.f;
  if (sg.n > 0) {
    gO = __ldg(dL_dopacity + r); gD = __ldg(dL_ddepth + r);
    gR = __ldg(dL_drgb + 3 * r); gG = __ldg(dL_drgb + 3 * r + 1); gB = __ldg(dL_drgb + 3 * r + 2);
    if (dL_dnormal) { gNx = __ldg(dL_dnormal + 3 * r); gNy = __ldg(dL_dnormal + 3 * r + 1); gNz = __ldg(dL_dnormal + 3 * r + 2); }
    O = __ldg(opacity + r); D = __ldg(depth + r);
    R = __ldg(rgb + 3 * r); Gc = __ldg(rgb + 3 * r + 1); B = __ldg(rgb + 3 * r + 2);
  }
  // pass A: sum over the whole segment of dL_dws*ws (volumerendering.cu:206-210, 277)
  float wsum = 0.f;
  if (dL_dws) {                                   // plain strided loop (no shuffles inside): the compiler batches its loads
#pragma unroll 4
    for (int k = j; k < sg.n; k += G) wsum = fmaf(__ldg(dL_dws + sg.start + k), __ldg(ws_saved + sg.start + k), wsum);
  }
  wsum = group_sum<G>(wsum);

  // pass B (kBwDepth chunks per trip, see the forward kernel)
  TState st;
  float cr = 0.f, cg = 0.f, cb = 0.f, cd = 0.f, cw = 0.f;  // running sums entering the chunk
  for (int base = 0; warp_any(base < sg.n); base += kBwDepth * G) {
    float a_[kBwDepth], dl_[kBwDepth], t_[kBwDepth], c0_[kBwDepth], c1_[kBwDepth], c2_[kBwDepth], gw_[kBwDepth], wsv_[kBwDepth];
#pragma unroll
    for (int m = 0; m < kBwDepth; m++) {
      const int k = base + m * G + j;
      a_[m] = 0.f; dl_[m] = 0.f; t_[m] = 0.f; c0_[m] = 0.f; c1_[m] = 0.f; c2_[m] = 0.f; gw_[m] = 0.f; wsv_[m] = 0.f;
      if (k < sg.n && !st.done) {                 // behind the terminating sample every gradient is 0 (w = 0, ws = 0): nothing to read
        const int64_t s = sg.start + k;
        dl_[m] = __ldg(deltas + s);
        const float sig = __ldg(sigmas + s);
        t_[m] = __ldg(ts + s);
        c0_[m] = __ldg(rgbs + 3 * s); c1_[m] = __ldg(rgbs + 3 * s + 1); c2_[m] = __ldg(rgbs + 3 * s + 2);
        if (dL_dws) { gw_[m] = __ldg(dL_dws + s); wsv_[m] = __ldg(ws_saved + s); }
        a_[m] = sample_alpha(sig, dl_[m]);
      }
    }
#pragma unroll
    for (int m = 0; m < kBwDepth; m++) {
      const int cbase = base + m * G;
      if (m > 0 && !warp_any(cbase < sg.n)) break;
      const bool valid = cbase + j < sg.n;
      const int64_t s = sg.start + cbase + j;
      const float a = a_[m], dl = dl_[m], t = t_[m], c0 = c0_[m], c1 = c1_[m], c2 = c2_[m], gw = gw_[m], wsv = wsv_[m];
      float Tb, Ta; bool active;
      chunk_transmittance<G>(st, a, valid, j, cbase, T_thr, Tb, Ta, active);
      const float w = active ? a * Tb : 0.f;
      // inclusive running sums (r,g,b,d of the reference and the scanned dL_dws*ws)
      const float pr = cr + group_incl_sum<G>(w * c0, j);
      const float pg = cg + group_incl_sum<G>(w * c1, j);
      const float pb = cb + group_incl_sum<G>(w * c2, j);
      const float pd = cd + group_incl_sum<G>(w * t, j);
      const float pw = cw + group_incl_sum<G>(gw * wsv, j);
      cr = group_bcast<G>(pr, G - 1); cg = group_bcast<G>(pg, G - 1); cb = group_bcast<G>(pb, G - 1);
      cd = group_bcast<G>(pd, G - 1); cw = group_bcast<G>(pw, G - 1);
      if (valid) {
        float ds = 0.f;
        if (active) {
          ds = dl * (gR * (c0 * Ta - (R - pr)) + gG * (c1 * Ta - (Gc - pg)) + gB * (c2 * Ta - (B - pb)) +
                     gO * (1.0f - O) + gD * (t * Ta - (D - pd)) + Ta * gw - (wsum - pw));
        }
        dL_dsigmas[s] = ds;
        dL_drgbs[3 * s] = gR * w; dL_drgbs[3 * s + 1] = gG * w; dL_drgbs[3 * s + 2] = gB * w;
        if (dL_dnormals) { dL_dnormals[3 * s] = gNx * w; dL_dnormals[3 * s + 1] = gNy * w; dL_dnormals[3 * s + 2] = gNz * w; }
        float* dsp = dL_dsems + s * classes;
        const float* gsp = dL_dsem + r * classes;
        for (int i = 0; i < classes; i++) dsp[i] = __ldg(gsp + i) * w;
      }
    }
  }
  });
}

// ------------------------------------------------------------------------------------------ test
// volumerendering.cu:314-374.  Thread per alive ray; per-ray state is read once, accumulated in
// registers over <= N_samples steps and written back once.
__global__ void __launch_bounds__(256) composite_test_fw_kernel(
    const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ normals,
    const float* __restrict__ normals_raw, const float* __restrict__ sems, const float* __restrict__ deltas,
    const float* __restrict__ ts, int64_t* __restrict__ alive, float T_thr, int classes,
    const int32_t* __restrict__ n_eff, int n_samples, int64_t n_alive, float* __restrict__ opacity,
    float* __restrict__ depth, float* __restrict__ rgb, float* __restrict__ normal, float* __restrict__ normal_raw,
    float* __restrict__ sem) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_alive) return;
  const int ne = n_eff[n];
  if (ne == 0) { alive[n] = -1; return; }
  const int64_t r = alive[n];
  const int64_t base = n * n_samples;
  const float T0 = 1.0f - opacity[r];
  float T = T0;
  float aO = 0.f, aD = 0.f, aR = 0.f, aG = 0.f, aB = 0.f;
  float aNx = 0.f, aNy = 0.f, aNz = 0.f, aRx = 0.f, aRy = 0.f, aRz = 0.f;
  float aS[kCP];
#pragma unroll
  for (int i = 0; i < kCP; i++) aS[i] = 0.f;
  const int nc = min(classes, kCP);
  bool dead = false;
  int s = 0;
  while (s < ne) {
    const int64_t o = base + s;
    const float a = sample_alpha(__ldg(sigmas + o), __ldg(deltas + o));
    const float w = a * T;
    aR = fmaf(w, __ldg(rgbs + 3 * o), aR); aG = fmaf(w, __ldg(rgbs + 3 * o + 1), aG); aB = fmaf(w, __ldg(rgbs + 3 * o + 2), aB);
    aD = fmaf(w, __ldg(ts + o), aD);
    aO += w;
    aNx = fmaf(w, __ldg(normals + 3 * o), aNx); aNy = fmaf(w, __ldg(normals + 3 * o + 1), aNy); aNz = fmaf(w, __ldg(normals + 3 * o + 2), aNz);
    aRx = fmaf(w, __ldg(normals_raw + 3 * o), aRx); aRy = fmaf(w, __ldg(normals_raw + 3 * o + 1), aRy); aRz = fmaf(w, __ldg(normals_raw + 3 * o + 2), aRz);
    const float* sp = sems + o * classes;
#pragma unroll
    for (int i = 0; i < kCP; i++) if (i < nc) aS[i] = fmaf(w, __ldg(sp + i), aS[i]);
    T *= 1.0f - a;
    if (T <= T_thr) { dead = true; break; }
    s++;
  }
  const int s_end = dead ? s + 1 : s;  // samples that contributed
  // classes beyond the register budget: replay the (cheap) alpha chain per extra chunk
  for (int c0 = kCP; c0 < classes; c0 += kCP) {
    float Tx = T0;
    float xS[kCP];
#pragma unroll
    for (int i = 0; i < kCP; i++) xS[i] = 0.f;
    const int ncx = min(classes - c0, kCP);
    for (int k = 0; k < s_end; k++) {
      const int64_t o = base + k;
      const float a = sample_alpha(__ldg(sigmas + o), __ldg(deltas + o));
      const float w = a * Tx;
      const float* sp = sems + o * classes + c0;
#pragma unroll
      for (int i = 0; i < kCP; i++) if (i < ncx) xS[i] = fmaf(w, __ldg(sp + i), xS[i]);
      Tx *= 1.0f - a;
    }
#pragma unroll
    for (int i = 0; i < kCP; i++) if (i < ncx) sem[r * classes + c0 + i] += xS[i];
  }
  opacity[r] += aO; depth[r] += aD;
  rgb[3 * r] += aR; rgb[3 * r + 1] += aG; rgb[3 * r + 2] += aB;
  normal[3 * r] += aNx; normal[3 * r + 1] += aNy; normal[3 * r + 2] += aNz;
  normal_raw[3 * r] += aRx; normal_raw[3 * r + 1] += aRy; normal_raw[3 * r + 2] += aRz;
#pragma unroll
  for (int i = 0; i < kCP; i++) if (i < nc) sem[r * classes + i] += aS[i];
  if (dead) alive[n] = -1;
}

// lanes per ray from the mean segment length (host-known: N / N_rays)
int pick_group(int64_t n_samples, int64_t n_rays) {
  if (const char* e = getenv("NGP_COMPOSITE_G")) { const int g = atoi(e); if (g == 4 || g == 8 || g == 16 || g == 32) return g; }   // tuning only
  // The mean is over ALL rays, and most rays of a batch are empty (Lego-shaped scene: 77 % of 2^18 random rays carry no
  // sample, the others 127 on average: tools/composite_sweep.py) — so the thresholds sit well below the group size.
  const double avg = n_rays > 0 ? (double)n_samples / (double)n_rays : 0.0;
  if (avg > 12.0) return 32;
  if (avg > 6.0) return 16;
  if (avg > 2.5) return 8;
  return 4;
}

// Launch shape of the group kernels (compositors, distortion / Ref-NeRF losses).  With G == 32 a warp walks kRayTile rays
// (scan.cuh for_each_ray); NGP_COMPOSITE_TILED=0 goes back to one launch slot per ray (1..8: rays per warp), NGP_COMPOSITE_BLOCK (32..256) sets the CTA size
// (tuning only; with one warp per ray a CTA's warp slots come free only when its longest ray is done, so small CTAs won there).
int group_tile(int64_t n_samples, int64_t n_rays) {
  if (const char* e = getenv("NGP_COMPOSITE_TILED")) { const int t = atoi(e); return t < 0 ? 0 : (t > kRayTile ? kRayTile : t); }   // 0 = off, 1..8 = rays per warp
  // ~256 samples per warp: 8 rays where most are empty (Lego-shaped batch: 34 per ray over all rays), one ray per warp where every
  // ray is long (street-shaped batch: 430 per ray — walking 8 such rays one after the other left the GPU with an eighth of the
  // warps it needs: 89 -> 142 ms per step, measured)
  const double avg = n_rays > 0 ? (double)n_samples / (double)n_rays : 0.0;
  const int t = avg > 0.0 ? (int)(256.0 / avg + 0.5) : kRayTile;
  return t < 1 ? 1 : (t > kRayTile ? kRayTile : t);
}
int group_block(bool tiled) {
  if (const char* e = getenv("NGP_COMPOSITE_BLOCK")) { const int b = atoi(e); if (b == 32 || b == 64 || b == 128 || b == 256) return b; }
  return tiled ? 128 : 64;
}

}  // namespace ngp

using namespace ngp;

#define NGP_GROUP_DISPATCH(G_, ...)                                  \
  switch (G_) {                                                      \
    case 32: { constexpr int G = 32; __VA_ARGS__; } break;           \
    case 16: { constexpr int G = 16; __VA_ARGS__; } break;           \
    case 8:  { constexpr int G = 8;  __VA_ARGS__; } break;           \
    default: { constexpr int G = 4;  __VA_ARGS__; } break;           \
  }

// Replaces vren.composite_train_fw (binding.cpp:121-145 -> volumerendering.cu:118-164).
// Every output element is written (ws / per-ray rows), no pre-zeroing needed provided rays_a
// lists every ray once and its segments tile [0, n_samples) — which is what the marcher emits.
NGP_API int ngp_composite_train_fw(const float* sigmas, const float* rgbs, const float* normals_pred,
                                   const float* sems, const float* deltas, const float* ts, const int64_t* rays_a,
                                   float T_threshold, int classes, int64_t n_samples, int64_t n_rays,
                                   int64_t* total_samples, float* opacity, float* depth, float* rgb,
                                   float* normal_pred, float* sem, float* ws, void* stream) {
  if (n_rays <= 0) return 0;
  const int Gsel = pick_group(n_samples, n_rays);
  cudaStream_t st = (cudaStream_t)stream;
  int c0 = 0;
  do {
    const int nc = classes - c0 < kCP ? (classes - c0 < 0 ? 0 : classes - c0) : kCP;
    const bool do_main = c0 == 0;
    NGP_GROUP_DISPATCH(Gsel, {
      NGP_GROUP_LAUNCH(composite_train_fw_kernel, st,
          sigmas, rgbs, normals_pred, sems, deltas, ts, rays_a, T_threshold, classes, c0, nc, do_main, n_rays,
          total_samples, opacity, depth, rgb, normal_pred, sem, ws);
    });
    NGP_LAUNCH_CHECK("ngp_composite_train_fw");
    c0 += kCP;
  } while (c0 < classes);
  return 0;
}

// Replaces vren.composite_alpha_fw (binding.cpp:109-118 -> volumerendering.cu:37-63).
NGP_API int ngp_composite_alpha_fw(const float* sigmas, const float* deltas, const int64_t* rays_a, float T_threshold,
                                   int64_t n_samples, int64_t n_rays, float* alphas, float* ws, void* stream) {
  if (n_rays <= 0) return 0;
  const int Gsel = pick_group(n_samples, n_rays);
  NGP_GROUP_DISPATCH(Gsel, {
    NGP_GROUP_LAUNCH(composite_alpha_fw_kernel, (cudaStream_t)stream, sigmas, deltas, rays_a, T_threshold, n_rays, alphas, ws);
  });
  NGP_LAUNCH_CHECK("ngp_composite_alpha_fw");
  return 0;
}

// Replaces vren.composite_train_bw (binding.cpp:148-188 -> volumerendering.cu:249-311).
// normals_pred / normal_pred of the reference signature are not read by its kernel's maths and are
// therefore not part of this entry point.
NGP_API int ngp_composite_train_bw(const float* dL_dopacity, const float* dL_ddepth, const float* dL_drgb,
                                   const float* dL_dnormal_pred, const float* dL_dsem, const float* dL_dws,
                                   const float* sigmas, const float* rgbs, const float* ws, const float* deltas,
                                   const float* ts, const int64_t* rays_a, const float* opacity, const float* depth,
                                   const float* rgb, float T_threshold, int classes, int64_t n_samples,
                                   int64_t n_rays, float* dL_dsigmas, float* dL_drgbs, float* dL_dnormals_pred,
                                   float* dL_dsems, void* stream) {
  if (n_rays <= 0) return 0;
  const int Gsel = pick_group(n_samples, n_rays);
  NGP_GROUP_DISPATCH(Gsel, {
    NGP_GROUP_LAUNCH(composite_train_bw_kernel, (cudaStream_t)stream,
        dL_dopacity, dL_ddepth, dL_drgb, dL_dnormal_pred, dL_dsem, dL_dws, sigmas, rgbs, ws, deltas, ts, rays_a,
        opacity, depth, rgb, T_threshold, classes, n_rays, dL_dsigmas, dL_drgbs, dL_dnormals_pred, dL_dsems);
  });
  NGP_LAUNCH_CHECK("ngp_composite_train_bw");
  return 0;
}

// Replaces vren.composite_test_fw (binding.cpp:242-284 -> volumerendering.cu:376-423).
// In place: alive_indices (-1 = finished), opacity/depth/rgb/normal/normal_raw/sem.
NGP_API int ngp_composite_test_fw(const float* sigmas, const float* rgbs, const float* normals,
                                  const float* normals_raw, const float* sems, const float* deltas, const float* ts,
                                  int64_t* alive_indices, float T_threshold, int classes, const int32_t* n_eff_samples,
                                  int n_samples, int64_t n_alive, float* opacity, float* depth, float* rgb,
                                  float* normal, float* normal_raw, float* sem, void* stream) {
  if (n_alive <= 0) return 0;
  const int64_t blocks = ceil_div(n_alive, 256);
  composite_test_fw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      sigmas, rgbs, normals, normals_raw, sems, deltas, ts, alive_indices, T_threshold, classes, n_eff_samples,
      n_samples, n_alive, opacity, depth, rgb, normal, normal_raw, sem);
  NGP_LAUNCH_CHECK("ngp_composite_test_fw");
  return 0;
}

// ------------------------------------------------------------------------------------------- per-ray -> per-sample
// Per-ray tensors (appearance embeddings, ...) repeated for every sample of their ray: models/rendering.py:217-219
//     kwargs[k] = torch.repeat_interleave(v[rays_a[:, 0]], rays_a[:, 2], 0)
// = an index, a repeats -> indices kernel and an index_select whose gather kernel runs one CTA per output row (3.6 ms
// for 14 M rows of 8 floats, twice per step, profiles/r01e_step_profile_playground_after.txt); the backward is an
// index_add.  Here: one warp per ray in both directions; a ray's rows are contiguous in the packed sample order, so
// the stores / loads of a warp are contiguous runs.  W <= 32 floats per row.
namespace ngp {
__global__ void __launch_bounds__(256) expand_per_ray_kernel(const float* __restrict__ v, const int64_t* __restrict__ rays_a,
                                                             int64_t n_rays, int W, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int sub = lane / W, c = lane - sub * W, per = 32 / W;      // `per` rows per warp pass; lanes >= per*W idle
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rays; r += n_warps) {
    const int64_t ray = __ldg(rays_a + 3 * r), start = __ldg(rays_a + 3 * r + 1), N = __ldg(rays_a + 3 * r + 2);
    if (sub >= per || N <= 0) continue;
    const float val = __ldg(v + ray * W + c);
    for (int64_t k = sub; k < N; k += per) out[(start + k) * W + c] = val;
  }
}
__global__ void __launch_bounds__(256) reduce_per_ray_kernel(const float* __restrict__ dout, const int64_t* __restrict__ rays_a,
                                                             int64_t n_rays, int W, float* __restrict__ dv) {
  const int lane = threadIdx.x & 31;
  const int sub = lane / W, c = lane - sub * W, per = 32 / W;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rays; r += n_warps) {
    const int64_t ray = __ldg(rays_a + 3 * r), start = __ldg(rays_a + 3 * r + 1), N = __ldg(rays_a + 3 * r + 2);
    if (sub >= per || N <= 0) continue;
    float acc = 0.f;
    for (int64_t k = sub; k < N; k += per) acc += __ldg(dout + (start + k) * W + c);
    atomicAdd(dv + ray * W + c, acc);          // <= 32/W partial sums per (ray, column); dv is zeroed by the caller
  }
}
}  // namespace ngp

// out (S,W)[start_r + k] = v (., W)[ray_r] for every ray r of rays_a (R,3) i64 [ray_idx, start_idx, N_samples] and k < N_r
// (models/rendering.py:217-219).  W <= 32.
NGP_API int ngp_expand_per_ray(const float* v, const int64_t* rays_a, int64_t n_rays, int width, float* out, void* stream) {
  if (n_rays <= 0) return 0;
  if (width < 1 || width > 32) return set_error_msg("ngp_expand_per_ray: width must be in [1, 32]");
  const int64_t blocks = ceil_div(n_rays, 8);
  expand_per_ray_kernel<<<(unsigned)(blocks < (int64_t)kSMs * 16 ? blocks : (int64_t)kSMs * 16), 256, 0, (cudaStream_t)stream>>>(v, rays_a, n_rays, width, out);
  NGP_LAUNCH_CHECK("ngp_expand_per_ray");
  return 0;
}
// Backward of ngp_expand_per_ray: dv (., W)[ray_r] += sum_k dout (S,W)[start_r + k]; the caller zeroes dv.
NGP_API int ngp_reduce_per_ray(const float* dout, const int64_t* rays_a, int64_t n_rays, int width, float* dv, void* stream) {
  if (n_rays <= 0) return 0;
  if (width < 1 || width > 32) return set_error_msg("ngp_reduce_per_ray: width must be in [1, 32]");
  const int64_t blocks = ceil_div(n_rays, 8);
  reduce_per_ray_kernel<<<(unsigned)(blocks < (int64_t)kSMs * 16 ? blocks : (int64_t)kSMs * 16), 256, 0, (cudaStream_t)stream>>>(dout, rays_a, n_rays, width, dv);
  NGP_LAUNCH_CHECK("ngp_reduce_per_ray");
  return 0;
}
