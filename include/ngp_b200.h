/* ngp_b200.h — C ABI of libngp_b200.so: the B200 (sm_100a) hot path of instant-ngp-pp.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a CUDA stream
 * (`void* stream` = cudaStream_t, NULL = legacy default stream), launches asynchronously on that
 * stream and returns 0 on success or a non-zero status (ngp_last_error() has the text).
 * No torch types cross this boundary; nothing here allocates device memory; there is no CPU
 * fallback.  All float tensors are fp32, contiguous, row-major unless a stride is given.
 *
 * Each function names the reference interface it replaces (file:line under the reference tree
 * zhihao-lin/instant-ngp-pp).  `vren.X` = the pybind function registered at
 * models/csrc/binding.cpp:323-342.
 */
#ifndef NGP_B200_H_
#define NGP_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ library */
int ngp_abi_version(void);
const char* ngp_last_error(void);
int ngp_check_device(void); /* 0 iff the current device is compute capability 10.x */
/* diagnostic (no reference counterpart): issues `iters` rounds of lane-pair red.global.add.v2.f32 into random 32-byte
 * sectors of `table` (the hash-gradient scatter's access pattern); returns the number of L2 sector requests issued.
 * bench.py times it to measure, in the same run, the L2 reduction request rate that bounds hashgrid_bw_params. */
int64_t ngp_probe_l2_reduction(float* table, int64_t table_bytes, int iters, void* stream);

/* ------------------------------------------------------------------ a1: intersections
 * vren.ray_aabb_intersect  binding.cpp:4-16  -> intersection.cu:59-100
 * hit_cnt (R) i32, hits_t (R,max_hits,2) f32 (-1 = empty), hits_voxel_idx (R,max_hits) i64 (-1). */
int ngp_ray_aabb_intersect(const float* rays_o, const float* rays_d, const float* centers, const float* half_sizes,
                           int64_t n_rays, int64_t n_voxels, int max_hits, int32_t* hit_cnt, float* hits_t,
                           int64_t* hits_voxel_idx, void* stream);
/* vren.ray_sphere_intersect  binding.cpp:19-31 -> intersection.cu:156-197 */
int ngp_ray_sphere_intersect(const float* rays_o, const float* rays_d, const float* centers, const float* radii,
                             int64_t n_rays, int64_t n_spheres, int max_hits, int32_t* hit_cnt, float* hits_t,
                             int64_t* hits_sphere_idx, void* stream);

/* ray generation: get_rays  datasets/ray_utils.py:49-72 with the gathers of train.py:136-137 fused:
 * rays_d = poses[img_idx][:, :3] . directions[pix_idx], rays_o = poses[img_idx][:, 3].  directions (P,3), poses (V,3,4);
 * img_idx NULL = poses[0] for every ray, pix_idx NULL = directions[i]. */
/* hits_t[(t1 >= 0) & (t1 < near), 0, 0] = near with t1 = hits_t[:, 0, 0]   models/rendering.py:29-30; in place; row_stride = floats
 * between consecutive rays' t1 (2 * max_hits) */
int ngp_near_clamp(float* hits_t, int64_t n_rays, int64_t row_stride, float near_distance, void* stream);
int ngp_get_rays(const float* directions, const float* poses, const int64_t* img_idx, const int64_t* pix_idx,
                 int64_t n_rays, float* rays_o, float* rays_d, void* stream);

/* ------------------------------------------------------------------ a9: occupancy grid
 * vren.morton3D          binding.cpp:46-50 -> raymarching.cu:72-88    coords (N,3) i32 -> (N) i32
 * vren.morton3D_invert   binding.cpp:53-57 -> raymarching.cu:103-119
 * vren.packbits          binding.cpp:35-43 -> raymarching.cu:143-161  dtype 0=f32 1=f16 2=f64 */
int ngp_morton3D(const int32_t* coords, int64_t n, int32_t* indices, void* stream);
int ngp_morton3D_invert(const int32_t* indices, int64_t n, int32_t* coords, void* stream);
int ngp_packbits(const void* density_grid, int dtype, int64_t n_bytes, float density_threshold,
                 uint8_t* density_bitfield, void* stream);
/* threshold in device memory (networks.py:404-408, thr = min(mean, density_threshold), without the .item() sync) */
int ngp_packbits_dthr(const float* density_grid, int64_t n_bytes, const float* density_threshold_dev,
                      uint8_t* density_bitfield, void* stream);

/* ------------------------------------------------------------------ f2: occupancy update as one kernel chain
 * NGP.sample_uniform_and_occupied_cells + NGP.update_density_grid   models/networks.py:294-333,379-408
 * (per cascade: randint, morton3D, nonzero, randint, fancy index, morton3D_invert, cat, rand_like, index_put, then
 * where / maximum over the grid, a masked mean with .item(), packbits).  Split at the one step that belongs to the
 * model — density(xyzs):
 *   sample : mask(grid > thr) -> popcount scan -> one thread per draw: M uniform cells + M occupied cells per cascade
 *            (warmup: every cell), jittered world positions          -> indices (Cc*n) i32 (-1 = void), xyzs (Cc*n,3)
 *   update : decay -> atomicMax scatter of the densities -> mean of the positive cells -> thr = min(mean, threshold)
 *            -> packbits, all on the stream; grid and bitfield are updated in place.
 * workspace: ngp_occupancy_workspace_bytes(cascades) bytes, 16-byte aligned, zero-filled once by the caller. */
int64_t ngp_occupancy_workspace_bytes(int cascades);
int ngp_occupancy_sample(const float* density_grid, int cascades, float scale, float density_threshold, int64_t M, int warmup,
                         uint64_t seed, void* workspace, int32_t* indices, float* xyzs, void* stream);
int ngp_occupancy_update(float* density_grid, int cascades, const int32_t* indices, const float* density, int64_t n_per_cascade,
                         float decay, const float* count_grid, float density_threshold, void* workspace,
                         uint8_t* density_bitfield, void* stream);

/* ------------------------------------------------------------------ a2: training ray marcher
 * vren.raymarching_train  binding.cpp:60-81 -> raymarching.cu:283-332, split so the caller can
 * size the outputs exactly (or not sync at all):
 *   count : per-ray sample counts + scan; counter (2) i32 = [total_samples, n_rays]
 *   write : rays_a (R,3) i64 [ray_idx,start_idx,N] in ray-index order, xyzs/dirs (S,3), deltas/ts (S)
 * workspace: ngp_raymarching_train_workspace_bytes(n_rays) bytes of device memory shared by both. */
int64_t ngp_raymarching_train_workspace_bytes(int64_t n_rays);
int64_t ngp_render_workspace_bytes(int64_t n_slots);   /* workspace of ngp_render_advance / ngp_render_emit */
int ngp_raymarching_train_count(const float* rays_o, const float* rays_d, const float* hits_t,
                                const uint8_t* density_bitfield, int cascades, float scale, float exp_step_factor,
                                const float* noise, int grid_size, int max_samples, int64_t n_rays, int32_t* counter,
                                void* workspace, void* stream);
int ngp_raymarching_train_write(const float* rays_o, const float* rays_d, const float* hits_t,
                                const uint8_t* density_bitfield, int cascades, float scale, float exp_step_factor,
                                int grid_size, int max_samples, int64_t n_rays, const void* workspace,
                                int64_t capacity, int64_t* rays_a, float* xyzs, float* dirs, float* deltas, float* ts,
                                void* stream);

/* ------------------------------------------------------------------ a3: test-time marcher
 * vren.raymarching_test  binding.cpp:84-106 -> raymarching.cu:407-454.  hits_t (R,2) advanced in
 * place; outputs dense (n_alive, n_samples, .), unused slots zero; n_eff_samples (n_alive) i32. */
int ngp_raymarching_test(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_indices,
                         const uint8_t* density_bitfield, int cascades, float scale, float exp_step_factor,
                         int grid_size, int max_samples, int n_samples, int64_t n_alive, float* xyzs, float* dirs,
                         float* deltas, float* ts, int32_t* n_eff_samples, void* stream);

/* ------------------------------------------------------------------ f3: test-time wavefront renderer
 * models/rendering.py:46-133 (volume_render loop) = per round raymarching_test + composite_test_fw + Python
 * compaction.  One advance launch per round composites the previous round and marches the next one. */
int ngp_render_advance(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_in,
                       int64_t n_alive_in, const int64_t* prev_rays_a, const float* sigmas, const float* rgbs,
                       const float* deltas, const float* ts, float T_threshold, const uint8_t* density_bitfield,
                       int cascades, float scale, float exp_step_factor, int grid_size, int max_samples, int n_next,
                       float* opacity, float* depth, float* rgb, int64_t* alive_out, int32_t* counters, void* workspace,
                       void* stream);
/* the same round with the normal / semantic streams of the reference's test-time compositor (volumerendering.cu:335-373):
 * normals_pred, normals_raw (S,3) and sems (S,classes) of the previous round's samples are composited into normal, normal_raw
 * (R,3) and sem (R,classes) next to rgb / depth / opacity.  Pass NULL / 0 for all seven to get ngp_render_advance. */
int ngp_render_advance_full(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_in, int64_t n_alive_in,
                            const int64_t* prev_rays_a, const float* sigmas, const float* rgbs, const float* deltas,
                            const float* ts, float T_threshold, const uint8_t* density_bitfield, int cascades, float scale,
                            float exp_step_factor, int grid_size, int max_samples, int n_next, float* opacity, float* depth,
                            float* rgb, int64_t* alive_out, int32_t* counters, void* workspace, const float* normals_pred,
                            const float* normals_raw, const float* sems, int classes, float* normal, float* normal_raw,
                            float* sem, void* stream);
/* One whole round for the ngp_pl-shaped field (density MLP L*F -> width -> 16 with sigma = exp(h0); colour MLP [SH4(d) | h] ->
 * width (x rgb_hidden) -> 3 sigmoid): advance + emit + hash grid -> bf16 tiles -> both MLPs, every launch reading its live
 * element count from device memory (n_alive_dev = the previous round's counters, NULL in round 0; counters[1] of this round),
 * so the host enqueues round after round from BOUNDS (n_alive_bound slots, buffers of n_alive_bound * n_next samples) and
 * reads the counters back late.  One trip of the loop of models/rendering.py:75-124 without a host round trip.
 * counters: 2 x int32 of THIS round = [alive slots after it, samples marched in it].  feat_tiles: workspace of
 * ceil(cap/128) * ngp_feature_tile_bytes bytes; h (cap,16), sigmas (cap), rgbs (cap,3). */
/* the field part of such a round alone (hash grid -> tiles -> both MLPs, live count from device memory n_dev): enqueued right behind
 * ngp_render_emit so that the host can fetch the round's counters while it runs (rendering.render_wavefront_pipelined) */
int ngp_field_compact_fw(const float* xyzs, const float* dirs, int64_t cap, const int32_t* n_dev, const float* aabb, const void* table,
                         int table_dtype, int n_levels, int n_features, int log2_hashmap_size, int base_resolution,
                         float per_level_scale, const float* sigma_params, const float* rgb_params, int width, int rgb_hidden,
                         void* feat_tiles, float* h, float* sigmas, float* rgbs, void* stream);
int ngp_render_round_compact(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive_in,
                             int64_t n_alive_bound, const int32_t* n_alive_dev, const int64_t* prev_rays_a,
                             const float* prev_sigmas, const float* prev_rgbs, const float* prev_deltas, const float* prev_ts,
                             float T_threshold, const uint8_t* density_bitfield, int cascades, float scale, float exp_step_factor,
                             int grid_size, int max_samples, int n_next, float* opacity, float* depth, float* rgb,
                             int64_t* alive_out, int32_t* counters, void* workspace, int64_t* rays_a, float* xyzs, float* dirs,
                             float* deltas, float* ts, const float* aabb, const void* table, int table_dtype, int n_levels,
                             int n_features, int log2_hashmap_size, int base_resolution, float per_level_scale,
                             const float* sigma_params, const float* rgb_params, int width, int rgb_hidden, void* feat_tiles,
                             float* h, float* sigmas, float* rgbs, void* stream);
int ngp_render_emit(const float* rays_o, const float* rays_d, const float* hits_t, const int64_t* alive_out,
                    int64_t n_slots, const uint8_t* density_bitfield, int cascades, float scale, float exp_step_factor,
                    int grid_size, int max_samples, const void* workspace, int64_t capacity, int64_t* rays_a,
                    float* xyzs, float* dirs, float* deltas, float* ts, int32_t* counters, void* stream);

/* ------------------------------------------------------------------ a4: training compositor
 * vren.composite_train_fw  binding.cpp:121-145 -> volumerendering.cu:118-164
 * vren.composite_train_bw  binding.cpp:148-188 -> volumerendering.cu:249-311
 * vren.composite_alpha_fw  binding.cpp:109-118 -> volumerendering.cu:37-63 */
int ngp_composite_train_fw(const float* sigmas, const float* rgbs, const float* normals_pred, const float* sems,
                           const float* deltas, const float* ts, const int64_t* rays_a, float T_threshold, int classes,
                           int64_t n_samples, int64_t n_rays, int64_t* total_samples, float* opacity, float* depth,
                           float* rgb, float* normal_pred, float* sem, float* ws, void* stream);
int ngp_composite_train_bw(const float* dL_dopacity, const float* dL_ddepth, const float* dL_drgb,
                           const float* dL_dnormal_pred, const float* dL_dsem, const float* dL_dws, const float* sigmas,
                           const float* rgbs, const float* ws, const float* deltas, const float* ts,
                           const int64_t* rays_a, const float* opacity, const float* depth, const float* rgb,
                           float T_threshold, int classes, int64_t n_samples, int64_t n_rays, float* dL_dsigmas,
                           float* dL_drgbs, float* dL_dnormals_pred, float* dL_dsems, void* stream);
int ngp_composite_alpha_fw(const float* sigmas, const float* deltas, const int64_t* rays_a, float T_threshold,
                           int64_t n_samples, int64_t n_rays, float* alphas, float* ws, void* stream);

/* ------------------------------------------------------------------ a5: test-time compositor
 * vren.composite_test_fw  binding.cpp:242-284 -> volumerendering.cu:376-423 (in place) */
int ngp_composite_test_fw(const float* sigmas, const float* rgbs, const float* normals, const float* normals_raw,
                          const float* sems, const float* deltas, const float* ts, int64_t* alive_indices,
                          float T_threshold, int classes, const int32_t* n_eff_samples, int n_samples, int64_t n_alive,
                          float* opacity, float* depth, float* rgb, float* normal, float* normal_raw, float* sem,
                          void* stream);

/* ------------------------------------------------------------------ a6: Ref-NeRF normal losses
 * vren.composite_refloss_fw  binding.cpp:191-208 -> ref_loss.cu:41-73
 * vren.composite_refloss_bw  binding.cpp:211-239 -> ref_loss.cu:133-175 */
int ngp_composite_refloss_fw(const float* sigmas, const float* normals_diff, const float* normals_ori,
                             const float* deltas, const int64_t* rays_a, float T_threshold, int64_t n_samples,
                             int64_t n_rays, float* loss_o, float* loss_p, void* stream);
int ngp_composite_refloss_bw(const float* dL_dloss_o, const float* dL_dloss_p, const float* sigmas,
                             const float* normals_diff, const float* normals_ori, const float* deltas,
                             const int64_t* rays_a, const float* loss_o, const float* loss_p, float T_threshold,
                             int64_t n_samples, int64_t n_rays, float* dL_dsigmas, float* dL_dnormals_diff,
                             float* dL_dnormals_ori, void* stream);

/* ------------------------------------------------------------------ a7: distortion loss
 * vren.distortion_loss_fw  binding.cpp:287-298 -> losses.cu:62-107
 * vren.distortion_loss_bw  binding.cpp:301-320 -> losses.cu:143-173 */
int ngp_distortion_loss_fw(const float* ws, const float* deltas, const float* ts, const int64_t* rays_a,
                           int64_t n_samples, int64_t n_rays, float* loss, float* ws_inclusive_scan,
                           float* wts_inclusive_scan, void* stream);
int ngp_distortion_loss_bw(const float* dL_dloss, const float* ws_inclusive_scan, const float* wts_inclusive_scan,
                           const float* ws, const float* deltas, const float* ts, const int64_t* rays_a,
                           int64_t n_samples, int64_t n_rays, float* dL_dws, void* stream);

/* ------------------------------------------------------------------ a10: hash-grid encoding
 * tcnn.Encoding(3, {"otype":"Grid"|"HashGrid", ...})  models/networks.py:40-52, 67-76
 * (tiny-cuda-nn is an un-vendored dependency of the reference; semantics per SURVEY.md App. B).
 * table_dtype: 0 = f32, 1 = f16.  Gradients are always fp32.
 * aabb: HOST pointer to {lo[3], range[3]} or NULL.  When given, x is world-space and every kernel applies
 * (x - lo) / range itself — the (x - xyz_min) / (xyz_max - xyz_min) pass of models/networks.py:174,188 fused
 * (IEEE sub + div: bit-identical to the tensor ops).  dL_dx is always w.r.t. the unit-cube coordinate. */
int64_t ngp_hashgrid_layout(int n_levels, int n_features, int log2_hashmap_size, int base_resolution,
                            float per_level_scale, uint32_t* offsets /*L+1*/, uint32_t* sizes, uint32_t* resolutions,
                            float* scales, uint8_t* dense);
int ngp_hashgrid_fw(const float* x, const float* aabb, const void* table, int table_dtype, int n_levels, int n_features,
                    int log2_hashmap_size, int base_resolution, float per_level_scale, int64_t n, float* y,
                    void* stream);
int ngp_hashgrid_bw_params(const float* x, const float* aabb, const float* dL_dy, int n_levels, int n_features, int log2_hashmap_size,
                           int base_resolution, float per_level_scale, int64_t n, float* dtable, void* stream);
int ngp_hashgrid_bw_input(const float* x, const float* aabb, const float* dL_dy, const void* table, int table_dtype, int n_levels,
                          int n_features, int log2_hashmap_size, int base_resolution, float per_level_scale, int64_t n,
                          float* dL_dx, void* stream);
int ngp_hashgrid_bwbw_input(const float* x, const float* aabb, const float* g2, const float* dL_dy, const void* table, int table_dtype,
                            int n_levels, int n_features, int log2_hashmap_size, int base_resolution,
                            float per_level_scale, int64_t n, float* dtable, float* d_dL_dy, void* stream);
/* both terms of the table gradient of a field used with its input gradient (density + normals, networks.py:186-196):
 * dtable += scatter(w, dL_dy) + scatter(coef(g2), dL_dy_first) — one pass for F = 8.  g2 == NULL: first term only. */
int ngp_hashgrid_bw_params_dual(const float* x, const float* aabb, const float* dL_dy, const float* g2,
                                const float* dL_dy_first, int n_levels, int n_features, int log2_hashmap_size,
                                int base_resolution, float per_level_scale, int64_t n, float* dtable, void* stream);
/* Fused density path of the ngp_pl-shaped field (private layouts, no reference counterpart): the encoder writes bf16
 * features directly as 128-sample tcgen05 operand tiles (ngp_feature_tile_bytes each) which ngp_mlp_fw/bw load with one
 * bulk copy (segment kind 2), and ngp_mlp_bw returns dL/dy as fp32 "gradient tiles" (ceil(N/128)*128*k0p floats,
 * level-chunk major) which the scatter consumes — the fp32 (N, L*F) feature / gradient matrices never exist. */
int64_t ngp_feature_tile_bytes(int n_levels, int n_features);
int ngp_hashgrid_fw_tiles(const float* x, const float* aabb, const void* table, int table_dtype, int n_levels,
                          int n_features, int log2_hashmap_size, int base_resolution, float per_level_scale, int64_t n,
                          void* y_tiles, void* stream);
int ngp_hashgrid_bw_params_tiles(const float* x, const float* aabb, const float* dy_tiles, int n_levels, int n_features,
                                 int log2_hashmap_size, int base_resolution, float per_level_scale, int64_t n,
                                 float* dtable, void* stream);
/* the same scatter restricted to levels [level_begin, level_end) (level_begin a multiple of 4/F): a level range is one
 * contiguous slice of the flat table gradient, so the caller can start the data-parallel all-reduce of a finished slice
 * (Lightning DDP's bucketed all-reduce, train.py:431) while the next range is being scattered. */
int ngp_hashgrid_bw_params_tiles_range(const float* x, const float* aabb, const float* dy_tiles, int n_levels, int n_features,
                                       int log2_hashmap_size, int base_resolution, float per_level_scale, int64_t n,
                                       float* dtable, int level_begin, int level_end, void* stream);

/* ------------------------------------------------------------------ a11: SH direction encoding
 * tcnn.Encoding(3, {"otype":"SphericalHarmonics","degree":4|3})  models/networks.py:78-85,128-135 */
int ngp_sh_fw(const float* v, int degree, int64_t n, float* out, void* stream);

/* ------------------------------------------------------------------ a12: fused MLP (tcgen05/TMEM)
 * tcnn.Network(n_in, n_out, {"otype":"CutlassMLP", ...})  models/networks.py:89-162
 * Input = concatenation of up to 3 segments (kind 0: fp32 rows; kind 1: SH4 of normalised dirs; kind 2: bf16 feature
 * tiles from ngp_hashgrid_fw_tiles, single segment only, gradient returned as gradient tiles).
 * Activations: 0 none, 1 ReLU, 2 sigmoid, 3 exp.  aux_exp_out / dL_daux_exp (optional): the density head
 * sigma = TruncExp(out[:,0]) of the ngp_pl-shaped field (custom_functions.py:200-211) fused into the epilogues. */
int64_t ngp_mlp_param_count(int n_input, int width, int n_hidden, int n_out);
int ngp_mlp_fw(int n_seg, const float* const* seg_ptr, const int* seg_width, const int* seg_kind,
               const int64_t* seg_stride, const float* params, int width, int n_hidden, int n_out, int act_hidden,
               int act_out, int64_t n, float* out, int64_t out_stride, float* aux_exp_out, void* stream);
int ngp_mlp_bw(int n_seg, const float* const* seg_ptr, const int* seg_width, const int* seg_kind,
               const int64_t* seg_stride, const float* params, int width, int n_hidden, int n_out, int act_hidden,
               int act_out, int64_t n, const float* dL_dout, int64_t dout_stride, float* dparams,
               float* const* dseg_ptr, const int64_t* dseg_stride, const float* dL_daux_exp,
               const float* saved_out /* optional: forward output, skips the output-layer recompute */,
               int64_t saved_out_stride, void* stream);

/* ------------------------------------------------------------------ a12: the density net on tcgen05 (csrc/density_net.cu)
 * xyz_net = Sequential(Linear(128,128), Softplus, Linear(128,1)) + sigma_act = Softplus   models/networks.py:54-59,172-181
 * and the autograd normals through it (torch.autograd.grad(create_graph=True) + loss.backward())   networks.py:186-196.
 *   fw : e (N,128) -> sigma (N), s2 (N) = sigmoid(pre-activation of sigma), g_e (N,128) | NULL = d sigma / d e
 *   bw : upstream dsigma (N) | NULL, d_ge (N,128) | NULL (+ the forward's g_e) -> de (N,128) | NULL,
 *        += dW1 (128,128), db1 (128), dw2 (128), db2 (1)
 * W1 row-major (out,in) = xyz_net[0].weight, b1 = xyz_net[0].bias, w2 = xyz_net[2].weight (128), b2 = xyz_net[2].bias.
 * bf16 tensor-core operands, fp32 accumulation.  n_in and width must be 128. */
int ngp_density_net_fw(const float* e, const float* W1, const float* b1, const float* w2, const float* b2, int64_t n, int n_in,
                       int width, float* sigma, float* s2, float* g_e, void* stream);
int ngp_density_net_bw(const float* e, const float* d_ge, const float* g_e, const float* dsigma, const float* s2, const float* W1,
                       const float* b1, const float* w2, int64_t n, int n_in, int width, float* de, float* dW1, float* db1,
                       float* dw2, float* db2, void* stream);

/* ------------------------------------------------------------------ a14: photometric + opacity terms of NeRFLoss
 * d['rgb'] = (results['rgb'] - target['rgb'])**2 ; d['opacity'] = lambda_opa * (-o * log(o)), o = opacity + 1e-10   losses.py:89-96
 * and loss = sum(mean) of train.py:310: the scalar and both gradients from one pass (loss[0] +=, caller zeroes). */
int ngp_basic_loss(const float* rgb, const float* target, const float* opacity, int64_t n_rays, float lambda_opa, float* loss,
                   float* drgb, float* dopacity, void* stream);

/* ------------------------------------------------------------------ a13: the field's two normal outputs
 * normals_raw = -F.normalize(grads, eps=1e-6), normals_pred = -F.normalize(norm_pred_header(feat), eps=1e-6)   models/networks.py:209,222-223
 * as one kernel per direction; scale_* = the unit-cube -> world factor of the analytic gradient (1 for none).
 * inv (N): 1 / max(||x * scale||, eps), negated where the norm was clamped (the Jacobian there is -I/eps). */
int ngp_neg_normalize_fw(const float* x, float scale_x, float scale_y, float scale_z, float eps, int64_t n, float* y, float* inv,
                         void* stream);
int ngp_neg_normalize_bw(const float* gy, const float* y, const float* inv, float scale_x, float scale_y, float scale_z, int64_t n,
                         float* gx, void* stream);
/* the per-sample inputs of RefLoss   models/rendering.py:243-246:  normals_diff = (normals_raw - normals_pred)^2,
 * normals_ori = clamp(sum(normals_raw * normalize(dirs)), min=0)^2, one kernel per direction */
int ngp_refloss_prep_fw(const float* normals_raw, const float* normals_pred, const float* dirs, int64_t n, float* normals_diff,
                        float* normals_ori, void* stream);
int ngp_refloss_prep_bw(const float* normals_raw, const float* normals_pred, const float* dirs, const float* g_diff, const float* g_ori,
                        int64_t n, float* g_raw, float* g_pred, void* stream);

/* ------------------------------------------------------------------ f1: fused dense Adam + grad-norm clip
 * torch.optim.Adam + gradient_clip_val=50   train.py:244-251, 435   */
int ngp_grad_sumsq(const float* g, int64_t n, float* accum, void* stream);
int ngp_clip_coef(const float* sumsq, float max_norm, float* coef, void* stream);
int ngp_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, int step, const float* grad_scale, void* stream);
/* lazy variant: entries with an exactly-zero gradient keep p / exp_avg / exp_avg_sq (tiny-cuda-nn's rule for encoding
 * parameters; 4 B instead of 28 B for every hash-table entry the batch did not touch).  Opt-in: not torch.optim.Adam's rule. */
int ngp_adam_step_lazy(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                       float beta1, float beta2, float eps, int step, const float* grad_scale, void* stream);

/* ------------------------------------------------------------------ a14: per-ray tensors repeated per sample
 * torch.repeat_interleave(v[rays_a[:, 0]], rays_a[:, 2], 0)  models/rendering.py:217-219 (appearance embeddings) and its
 * backward.  v (., W) f32, rays_a (R,3) i64 [ray_idx, start_idx, N_samples], out / dout (S,W); W <= 32;
 * reduce accumulates (+=) into dv (caller zeroes). */
int ngp_expand_per_ray(const float* v, const int64_t* rays_a, int64_t n_rays, int width, float* out, void* stream);
int ngp_reduce_per_ray(const float* dout, const int64_t* rays_a, int64_t n_rays, int width, float* dv, void* stream);

/* ------------------------------------------------------------------ a12/a13: density net with analytic normals
 * The elementwise stages of the reference's torch density net  models/networks.py:54-59 (xyz_net = Linear -> Softplus ->
 * Linear(.,1), sigma_act Softplus) together with d sigma / d(encoding) of networks.py:186-196 and the backward of both
 * (what torch.autograd.grad(create_graph=True) + loss.backward() evaluate there).  The GEMMs with W1 stay with the caller.
 *   fw: z1 (N,W) = e W1^T + b1, w2 (W), b2 (1)  ->  sigma (N), s2 (N) = sigmoid(z2), t (N,W) = s2*sigmoid(z1)*w2  (g_e = t W1)
 *   bw: v (N,W) = dL/dg_e W1^T | NULL, dsigma (N) | NULL  ->  dz1 (N,W), dz2 (N); dw2 (W), db1 (W) accumulated (+=).
 * W in {128, 256, 384, 512}. */
int ngp_density_head_fw(const float* z1, const float* w2, const float* b2, int64_t n, int width, float* sigma,
                        float* s2, float* t, void* stream);
int ngp_density_head_bw(const float* z1, const float* v, const float* s2, const float* dsigma, const float* w2,
                        int64_t n, int width, float* dz1, float* dz2, float* dw2, float* db1, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NGP_B200_H_ */
