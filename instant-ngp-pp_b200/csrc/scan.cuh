// Sub-warp segmented-scan building blocks shared by the compositing / loss kernels.
//
// A "group" is G consecutive lanes (G in {4,8,16,32}) that own ONE ray and walk its packed
// samples G at a time; 32/G rays share a warp.  All shuffles are full-mask with width=G, and
// every loop that contains them runs a warp-uniform trip count (see `warp_any`).
#pragma once
#include "common.cuh"
#include <type_traits>

namespace ngp {

constexpr unsigned kFull = 0xffffffffu;

template <int G> __device__ __forceinline__ float group_incl_sum(float v, int j) {
#pragma unroll
  for (int o = 1; o < G; o <<= 1) {
    const float n = __shfl_up_sync(kFull, v, o, G);
    if (j >= o) v += n;
  }
  return v;
}
template <int G> __device__ __forceinline__ float group_incl_prod(float v, int j) {
#pragma unroll
  for (int o = 1; o < G; o <<= 1) {
    const float n = __shfl_up_sync(kFull, v, o, G);
    if (j >= o) v *= n;
  }
  return v;
}
template <int G> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o, G);
  return v;
}
// value held by lane `src` (0..G-1) of my group
template <int G> __device__ __forceinline__ float group_bcast(float v, int src) {
  return __shfl_sync(kFull, v, src, G);
}
// index (0..G-1) of the first lane of my group whose predicate is set, or G when none is.
template <int G> __device__ __forceinline__ int group_first(bool pred) {
  const unsigned b = __ballot_sync(kFull, pred);
  const int lane = threadIdx.x & 31;
  const unsigned gm = (G == 32) ? b : ((b >> (lane & ~(G - 1))) & ((1u << G) - 1u));
  return gm ? (__ffs(gm) - 1) : G;
}
__device__ __forceinline__ bool warp_any(bool p) { return __any_sync(kFull, p); }

// alpha of one sample exactly as the reference spells it: 1.0f - __expf(-sigma*delta)
// (volumerendering.cu:24,94,214,349; ref_loss.cu:26,106).
__device__ __forceinline__ float sample_alpha(float sigma, float delta) { return 1.0f - __expf(-sigma * delta); }

// Per-ray segment descriptor from a rays_a row (int64 x3: ray_idx, start_idx, N_samples).
struct Seg { int64_t ray; int64_t start; int n; };
__device__ __forceinline__ Seg load_seg(const int64_t* __restrict__ rays_a, int64_t row, int64_t n_rows) {
  Seg s;
  if (row < n_rows) {
    s.ray = __ldg(rays_a + 3 * row); s.start = __ldg(rays_a + 3 * row + 1); s.n = (int)__ldg(rays_a + 3 * row + 2);
  } else { s.ray = -1; s.start = 0; s.n = 0; }
  return s;
}

// Which rays a thread works on.
//   kTiled == false: group g of the grid (G consecutive threads) owns rays_a row g — one launch slot per ray.
//   kTiled == true (G == 32 only): a warp owns `tile` (1..kRayTile, chosen by the host from the mean samples per ray so that a
//   warp gets ~256 samples: group_tile) consecutive rows.  Lanes < tile fetch the rows' descriptors in one coalesced request and
//   the warp then walks them one after the other.  Most rays of a batch are empty (Lego-shaped scene: 77 %),
//   and a warp launched for an empty ray is pure launch overhead: with one warp per ray the kernels below were bound by the rate
//   at which CTAs start, not by memory or issue slots (tools/composite_sweep.py, profiles/r02h_composite_sweep.txt).
// body(sg, j) is called once per row with j = the thread's lane in the group; it must write the row's outputs also for n == 0.
// empty(sg) (optional, kTiled only): called by ONE lane per empty row (n == 0) — all empty rows of the tile in one pass, side by
// side — and body() is then not called for them; without it body() sees every row.
constexpr int kRayTile = 8;
struct NoEmptyHandler { __device__ __forceinline__ void operator()(const Seg&) const {} };
template <int G, bool kTiled, typename Body, typename Empty = NoEmptyHandler>
__device__ __forceinline__ void for_each_ray(const int64_t* __restrict__ rays_a, int64_t n_rays, int tile, Body body, Empty empty = Empty()) {
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if constexpr (G == 32 && kTiled) {
    constexpr bool kHasEmpty = !std::is_same<Empty, NoEmptyHandler>::value;
    const int lane = (int)(threadIdx.x & 31u);
    const int64_t row0 = (gtid >> 5) * tile;
    const Seg mine = load_seg(rays_a, lane < tile ? row0 + lane : n_rays, n_rays);
    if (kHasEmpty && mine.ray >= 0 && mine.n <= 0) empty(mine);
#pragma unroll 1
    for (int k = 0; k < tile; k++) {
      Seg sg;
      sg.ray = __shfl_sync(kFull, mine.ray, k); sg.start = __shfl_sync(kFull, mine.start, k); sg.n = __shfl_sync(kFull, mine.n, k);
      if (sg.ray < 0) break;                       // past the last row (warp-uniform)
      if (kHasEmpty && sg.n <= 0) continue;
      body(sg, lane);
    }
  } else {
    body(load_seg(rays_a, gtid / G, n_rays), (int)(gtid % G));
  }
}
// threads a launch needs for n_rays rows
__host__ __forceinline__ int64_t group_threads(int64_t n_rays, int G, int tile) {
  return (G == 32 && tile > 0) ? ceil_div(n_rays, tile) * 32 : n_rays * G;
}

int group_tile(int64_t n_samples, int64_t n_rays);   // composite.cu: rays per warp at G == 32, 0 = one launch slot per ray
int group_block(bool tiled);    // composite.cu
// inside NGP_GROUP_DISPATCH (G is the compile-time group size): kernel<G, tiled><<<grid, block, 0, stream>>>(args)
#define NGP_GROUP_LAUNCH(KERNEL_, STREAM_, ...)                                                            \
  do {                                                                                                     \
    const int tile_ = G == 32 ? ngp::group_tile(n_samples, n_rays) : 0;                                    \
    const int bs_ = ngp::group_block(tile_ > 0);                                                           \
    const unsigned blocks_ = (unsigned)ceil_div(ngp::group_threads(n_rays, G, tile_), bs_);                \
    if (tile_ > 0) KERNEL_<G, (G == 32)><<<blocks_, bs_, 0, STREAM_>>>(__VA_ARGS__, tile_);                \
    else KERNEL_<G, false><<<blocks_, bs_, 0, STREAM_>>>(__VA_ARGS__, 0);                                  \
  } while (0)

// Transmittance bookkeeping for one chunk of G samples of one ray.
//   in : a (alpha of my sample, 0 for lanes past the end), valid, T_carry (T before the chunk), done
//   out: T_before / T_after for my sample, active (sample contributes: before or AT the
//        terminating sample, volumerendering.cu:109-112), and the updated group state.
struct TState {
  float T_carry = 1.0f;   // transmittance entering the next chunk
  bool done = false;      // ray already hit T <= T_threshold
  int n_done = 0;         // reference's `samples` counter (volumerendering.cu:90,112)
};
template <int G>
__device__ __forceinline__ void chunk_transmittance(TState& st, float a, bool valid, int j, int base, float T_thr,
                                                    float& T_before, float& T_after, bool& active) {
  const float om = valid ? 1.0f - a : 1.0f;
  const float P = group_incl_prod<G>(om, j);
  float Pex = __shfl_up_sync(kFull, P, 1, G);
  if (j == 0) Pex = 1.0f;
  T_before = st.T_carry * Pex;
  T_after = T_before * om;
  const bool stop = valid && !st.done && (T_after <= T_thr);
  const int f = group_first<G>(stop);
  active = valid && !st.done && j <= f;
  const float T_last = group_bcast<G>(T_after, G - 1);
  if (!st.done) {
    if (f < G) { st.done = true; st.n_done = base + f; }
    else st.T_carry = T_last;
  }
}

}  // namespace ngp
