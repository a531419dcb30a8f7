"""`vren` — the reference's 15-function extension surface (models/csrc/binding.cpp:323-342),
re-implemented over the libngp_b200 C ABI.

Same names, argument order, return structure and in-place behaviour as the reference's pybind
module, so `import vren` in models/custom_functions.py:2, models/rendering.py:7,
models/networks.py:6 and losses.py:4 resolves to this file (see instant-ngp-pp_b200/vren.py).
Differences a caller can observe, all deliberate:
  * outputs are allocated with torch.empty and fully written by the kernels (the reference
    zero-fills N_rays*max_samples rows per marching call, raymarching.cu:298-305);
  * raymarching_train returns exactly total_samples rows (the reference returns
    N_rays*max_samples rows that Python immediately slices to counter[0],
    custom_functions.py:93-98) in deterministic ray-index order (reference: atomic order);
  * kernels run on torch's CURRENT stream (reference: legacy default stream).
"""
import torch

from . import _lib
from ._lib import lib, ptr, check, stream

_I32, _I64, _F32 = torch.int32, torch.int64, torch.float32


def _f32(t, name):
    if t.dtype != _F32:
        raise RuntimeError(f"ngp_b200.vren: {name} must be float32 (got {t.dtype})")
    return t


def ray_aabb_intersect(rays_o, rays_d, centers, half_sizes, max_hits):
    """binding.cpp:4-16.  -> [hit_cnt (R) i32, hits_t (R,max_hits,2) f32, hits_voxel_idx (R,max_hits) i64]"""
    _lib.require_device()
    R, V = rays_o.shape[0], centers.shape[0]
    dev = rays_o.device
    hit_cnt = torch.empty(R, dtype=_I32, device=dev)
    hits_t = torch.empty(R, max_hits, 2, dtype=_F32, device=dev)
    hits_idx = torch.empty(R, max_hits, dtype=_I64, device=dev)
    check(lib.ngp_ray_aabb_intersect(ptr(_f32(rays_o, "rays_o")), ptr(_f32(rays_d, "rays_d")), ptr(centers),
                                     ptr(half_sizes), R, V, int(max_hits), ptr(hit_cnt), ptr(hits_t), ptr(hits_idx),
                                     stream()), "ray_aabb_intersect")
    return [hit_cnt, hits_t, hits_idx]


def ray_sphere_intersect(rays_o, rays_d, centers, radii, max_hits):
    """binding.cpp:19-31."""
    _lib.require_device()
    R, V = rays_o.shape[0], centers.shape[0]
    dev = rays_o.device
    hit_cnt = torch.empty(R, dtype=_I32, device=dev)
    hits_t = torch.empty(R, max_hits, 2, dtype=_F32, device=dev)
    hits_idx = torch.empty(R, max_hits, dtype=_I64, device=dev)
    check(lib.ngp_ray_sphere_intersect(ptr(_f32(rays_o, "rays_o")), ptr(_f32(rays_d, "rays_d")), ptr(centers),
                                       ptr(radii), R, V, int(max_hits), ptr(hit_cnt), ptr(hits_t), ptr(hits_idx),
                                       stream()), "ray_sphere_intersect")
    return [hit_cnt, hits_t, hits_idx]


def morton3D(coords):
    """binding.cpp:46-50.  coords (N,3) int32 -> indices (N) int32"""
    _lib.require_device()
    if coords.dtype != _I32:
        raise RuntimeError("ngp_b200.vren.morton3D: coords must be int32")
    out = torch.empty(coords.shape[0], dtype=_I32, device=coords.device)
    check(lib.ngp_morton3D(ptr(coords), coords.shape[0], ptr(out), stream()), "morton3D")
    return out


def morton3D_invert(indices):
    """binding.cpp:53-57.  indices (N) int32 -> coords (N,3) int32"""
    _lib.require_device()
    if indices.dtype != _I32:
        raise RuntimeError("ngp_b200.vren.morton3D_invert: indices must be int32")
    out = torch.empty(indices.shape[0], 3, dtype=_I32, device=indices.device)
    check(lib.ngp_morton3D_invert(ptr(indices), indices.shape[0], ptr(out), stream()), "morton3D_invert")
    return out


_PACK_DT = {torch.float32: 0, torch.float16: 1, torch.float64: 2}


def packbits(density_grid, density_threshold, density_bitfield):
    """binding.cpp:35-43.  In place on density_bitfield (N bytes); returns None."""
    _lib.require_device()
    if density_grid.dtype not in _PACK_DT:
        raise RuntimeError("ngp_b200.vren.packbits: density_grid must be float32/float16/float64")
    if density_bitfield.dtype != torch.uint8:
        raise RuntimeError("ngp_b200.vren.packbits: density_bitfield must be uint8")
    check(lib.ngp_packbits(ptr(density_grid), _PACK_DT[density_grid.dtype], density_bitfield.shape[0],
                           float(density_threshold), ptr(density_bitfield), stream()), "packbits")


def packbits_dthr(density_grid, density_threshold, density_bitfield):
    """packbits with the threshold as a 0-dim float32 CUDA tensor (no host read-back)."""
    _lib.require_device()
    if density_grid.dtype != torch.float32 or density_threshold.dtype != torch.float32 or density_bitfield.dtype != torch.uint8:
        raise RuntimeError("ngp_b200.vren.packbits_dthr: float32 grid / float32 threshold / uint8 bitfield")
    check(lib.ngp_packbits_dthr(ptr(density_grid), density_bitfield.shape[0], ptr(density_threshold), ptr(density_bitfield),
                                stream()), "packbits_dthr")


class MarchPlan:
    """Result of the count+scan phase of the training marcher (device resident)."""
    __slots__ = ("workspace", "counter", "n_rays")

    def __init__(self, workspace, counter, n_rays):
        self.workspace, self.counter, self.n_rays = workspace, counter, n_rays


def raymarching_train_count(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise,
                            grid_size, max_samples):
    """Phase 1+2 of raymarching_train: per-ray counts + scan, no host sync."""
    _lib.require_device()
    R = rays_o.shape[0]
    dev = rays_o.device
    ws = torch.empty(int(lib.ngp_raymarching_train_workspace_bytes(R)), dtype=torch.uint8, device=dev)
    counter = torch.empty(2, dtype=_I32, device=dev)
    check(lib.ngp_raymarching_train_count(ptr(_f32(rays_o, "rays_o")), ptr(_f32(rays_d, "rays_d")),
                                          ptr(_f32(hits_t, "hits_t")), ptr(density_bitfield), int(cascades),
                                          float(scale), float(exp_step_factor), ptr(_f32(noise, "noise")),
                                          int(grid_size), int(max_samples), R, ptr(counter), ptr(ws), stream()),
          "raymarching_train/count+cull" if (int(cascades) == 1 and float(exp_step_factor) == 0.0 and int(grid_size) == 128
                                             and R >= 2048) else "raymarching_train/count")
    return MarchPlan(ws, counter, R)


def raymarching_train_write(plan, rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor,
                            grid_size, max_samples, capacity):
    """Phase 3: emit rays_a and `capacity` rows of packed samples (rows past the true total are untouched)."""
    R = plan.n_rays
    dev = rays_o.device
    rays_a = torch.empty(R, 3, dtype=_I64, device=dev)
    xyzs = torch.empty(capacity, 3, dtype=_F32, device=dev)
    dirs = torch.empty(capacity, 3, dtype=_F32, device=dev)
    deltas = torch.empty(capacity, dtype=_F32, device=dev)
    ts = torch.empty(capacity, dtype=_F32, device=dev)
    check(lib.ngp_raymarching_train_write(ptr(rays_o), ptr(rays_d), ptr(hits_t), ptr(density_bitfield), int(cascades),
                                          float(scale), float(exp_step_factor), int(grid_size), int(max_samples), R,
                                          ptr(plan.workspace), int(capacity), ptr(rays_a), ptr(xyzs), ptr(dirs),
                                          ptr(deltas), ptr(ts), stream()), "raymarching_train/write")
    return rays_a, xyzs, dirs, deltas, ts


def raymarching_train(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise, grid_size,
                      max_samples):
    """binding.cpp:60-81.  -> [rays_a (R,3) i64, xyzs (S,3), dirs (S,3), deltas (S), ts (S), counter (2) i32]

    One host read-back of counter[0] sizes the outputs exactly (the reference performs the same
    read-back one line later, custom_functions.py:93).
    """
    plan = raymarching_train_count(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise,
                                   grid_size, max_samples)
    total = int(plan.counter[0].item())
    rays_a, xyzs, dirs, deltas, ts = raymarching_train_write(plan, rays_o, rays_d, hits_t, density_bitfield, cascades,
                                                             scale, exp_step_factor, grid_size, max_samples, total)
    return [rays_a, xyzs, dirs, deltas, ts, plan.counter]


def raymarching_test(rays_o, rays_d, hits_t, alive_indices, density_bitfield, cascades, scale, exp_step_factor,
                     grid_size, max_samples, N_samples):
    """binding.cpp:84-106.  hits_t (R,2) is advanced in place.
    -> [xyzs (A,N,3), dirs (A,N,3), deltas (A,N), ts (A,N), N_eff_samples (A) i32]"""
    _lib.require_device()
    A = alive_indices.shape[0]
    dev = rays_o.device
    N = int(N_samples)
    xyzs = torch.empty(A, N, 3, dtype=_F32, device=dev)
    dirs = torch.empty(A, N, 3, dtype=_F32, device=dev)
    deltas = torch.empty(A, N, dtype=_F32, device=dev)
    ts = torch.empty(A, N, dtype=_F32, device=dev)
    n_eff = torch.empty(A, dtype=_I32, device=dev)
    check(lib.ngp_raymarching_test(ptr(_f32(rays_o, "rays_o")), ptr(_f32(rays_d, "rays_d")), ptr(_f32(hits_t, "hits_t")),
                                   ptr(alive_indices), ptr(density_bitfield), int(cascades), float(scale),
                                   float(exp_step_factor), int(grid_size), int(max_samples), N, A, ptr(xyzs), ptr(dirs),
                                   ptr(deltas), ptr(ts), ptr(n_eff), stream()), "raymarching_test")
    return [xyzs, dirs, deltas, ts, n_eff]


def composite_alpha_fw(sigmas, deltas, rays_a, T_threshold):
    """binding.cpp:109-118.  -> [alphas (S), ws (S)]"""
    _lib.require_device()
    S, R = sigmas.shape[0], rays_a.shape[0]
    alphas = torch.zeros_like(sigmas)
    ws = torch.zeros_like(sigmas)
    check(lib.ngp_composite_alpha_fw(ptr(_f32(sigmas, "sigmas")), ptr(deltas), ptr(rays_a), float(T_threshold), S, R,
                                     ptr(alphas), ptr(ws), stream()), "composite_alpha_fw")
    return [alphas, ws]


def composite_train_fw(sigmas, rgbs, normals_pred, sems, deltas, ts, rays_a, T_threshold, classes):
    """binding.cpp:121-145.  -> [total_samples (R) i64, opacity (R), depth (R), rgb (R,3),
    normal_pred (R,3), sem (R,C), ws (S)]"""
    _lib.require_device()
    S, R = sigmas.shape[0], rays_a.shape[0]
    dev = sigmas.device
    C = int(classes)
    # per-ray rows are tiny: zero them so a rays_a that does not list every ray still matches the
    # reference's torch::zeros outputs; ws is fully written by the kernel.
    # ... from ONE zero fill: [total i64 (R) | opacity | depth | rgb (R,3) | normal (R,3) | sem (R,C)] carved out of a flat buffer
    flat = torch.zeros(R * (10 + C), dtype=_F32, device=dev)
    total = flat[:2 * R].view(_I64)
    opacity, depth = flat[2 * R:3 * R], flat[3 * R:4 * R]
    rgb, normal = flat[4 * R:7 * R].view(R, 3), flat[7 * R:10 * R].view(R, 3)
    sem = flat[10 * R:].view(R, C)
    ws = torch.empty(S, dtype=_F32, device=dev)
    check(lib.ngp_composite_train_fw(ptr(_f32(sigmas, "sigmas")), ptr(rgbs), ptr(normals_pred), ptr(sems), ptr(deltas),
                                     ptr(ts), ptr(rays_a), float(T_threshold), C, S, R, ptr(total), ptr(opacity),
                                     ptr(depth), ptr(rgb), ptr(normal), ptr(sem), ptr(ws), stream()),
          "composite_train_fw")
    return [total, opacity, depth, rgb, normal, sem, ws]


def composite_train_bw(dL_dopacity, dL_ddepth, dL_drgb, dL_dnormal_pred, dL_dsem, dL_dws, sigmas, rgbs, normals_pred,
                       ws, deltas, ts, rays_a, opacity, depth, rgb, normal_pred, T_threshold, classes):
    """binding.cpp:148-188.  -> [dL_dsigmas (S), dL_drgbs (S,3), dL_dnormals_pred (S,3), dL_dsems (S,C)]"""
    _lib.require_device()
    S, R = sigmas.shape[0], rays_a.shape[0]
    dev = sigmas.device
    C = int(classes)
    d_sig = torch.empty(S, dtype=_F32, device=dev)
    d_rgb = torch.empty(S, 3, dtype=_F32, device=dev)
    d_nrm = torch.empty(S, 3, dtype=_F32, device=dev)
    d_sem = torch.empty(S, C, dtype=_F32, device=dev)
    check(lib.ngp_composite_train_bw(ptr(dL_dopacity), ptr(dL_ddepth), ptr(dL_drgb), ptr(dL_dnormal_pred), ptr(dL_dsem),
                                     ptr(dL_dws), ptr(_f32(sigmas, "sigmas")), ptr(rgbs), ptr(ws), ptr(deltas), ptr(ts),
                                     ptr(rays_a), ptr(opacity), ptr(depth), ptr(rgb), float(T_threshold), C, S, R,
                                     ptr(d_sig), ptr(d_rgb), ptr(d_nrm), ptr(d_sem), stream()), "composite_train_bw")
    return [d_sig, d_rgb, d_nrm, d_sem]


def composite_refloss_fw(sigmas, normals_diff, normals_ori, deltas, ts, rays_a, T_threshold):
    """binding.cpp:191-208.  -> [loss_o (R), loss_p (R,3)]"""
    _lib.require_device()
    S, R = sigmas.shape[0], rays_a.shape[0]
    dev = sigmas.device
    loss_o = torch.zeros(R, dtype=_F32, device=dev)
    loss_p = torch.zeros(R, 3, dtype=_F32, device=dev)
    check(lib.ngp_composite_refloss_fw(ptr(_f32(sigmas, "sigmas")), ptr(normals_diff), ptr(normals_ori), ptr(deltas),
                                       ptr(rays_a), float(T_threshold), S, R, ptr(loss_o), ptr(loss_p), stream()),
          "composite_refloss_fw")
    return [loss_o, loss_p]


def composite_refloss_bw(dL_dloss_o, dL_dloss_p, sigmas, normals_diff, normals_ori, deltas, ts, rays_a, loss_o, loss_p,
                         T_threshold):
    """binding.cpp:211-239.  -> [dL_dsigmas (S), dL_dnormals_diff (S,3), dL_dnormals_ori (S)]"""
    _lib.require_device()
    S, R = sigmas.shape[0], rays_a.shape[0]
    dev = sigmas.device
    d_sig = torch.empty(S, dtype=_F32, device=dev)
    d_diff = torch.empty(S, 3, dtype=_F32, device=dev)
    d_ori = torch.empty(S, dtype=_F32, device=dev)
    check(lib.ngp_composite_refloss_bw(ptr(dL_dloss_o), ptr(dL_dloss_p), ptr(_f32(sigmas, "sigmas")), ptr(normals_diff),
                                       ptr(normals_ori), ptr(deltas), ptr(rays_a), ptr(loss_o), ptr(loss_p),
                                       float(T_threshold), S, R, ptr(d_sig), ptr(d_diff), ptr(d_ori), stream()),
          "composite_refloss_bw")
    return [d_sig, d_diff, d_ori]


def composite_test_fw(sigmas, rgbs, normals, normals_raw, sems, deltas, ts, hits_t, alive_indices, T_threshold,
                      classes, N_eff_samples, opacity, depth, rgb, normal, normal_raw, sem):
    """binding.cpp:242-284.  In place on alive_indices / opacity / depth / rgb / normal / normal_raw / sem."""
    _lib.require_device()
    A = alive_indices.shape[0]
    N = sigmas.shape[1] if sigmas.dim() == 2 else 1
    check(lib.ngp_composite_test_fw(ptr(_f32(sigmas, "sigmas")), ptr(rgbs), ptr(normals), ptr(normals_raw), ptr(sems),
                                    ptr(deltas), ptr(ts), ptr(alive_indices), float(T_threshold), int(classes),
                                    ptr(N_eff_samples), N, A, ptr(opacity), ptr(depth), ptr(rgb), ptr(normal),
                                    ptr(normal_raw), ptr(sem), stream()), "composite_test_fw")


def distortion_loss_fw(ws, deltas, ts, rays_a):
    """binding.cpp:287-298.  -> [loss (R), ws_inclusive_scan (S), wts_inclusive_scan (S)]"""
    _lib.require_device()
    S, R = ws.shape[0], rays_a.shape[0]
    dev = ws.device
    loss = torch.zeros(R, dtype=_F32, device=dev)
    ws_inc = torch.empty(S, dtype=_F32, device=dev)
    wts_inc = torch.empty(S, dtype=_F32, device=dev)
    check(lib.ngp_distortion_loss_fw(ptr(_f32(ws, "ws")), ptr(deltas), ptr(ts), ptr(rays_a), S, R, ptr(loss),
                                     ptr(ws_inc), ptr(wts_inc), stream()), "distortion_loss_fw")
    return [loss, ws_inc, wts_inc]


def distortion_loss_bw(dL_dloss, ws_inclusive_scan, wts_inclusive_scan, ws, deltas, ts, rays_a):
    """binding.cpp:301-320.  -> dL_dws (S)"""
    _lib.require_device()
    S, R = ws.shape[0], rays_a.shape[0]
    d_ws = torch.empty(S, dtype=_F32, device=ws.device)
    check(lib.ngp_distortion_loss_bw(ptr(_f32(dL_dloss, "dL_dloss")), ptr(ws_inclusive_scan), ptr(wts_inclusive_scan),
                                     ptr(ws), ptr(deltas), ptr(ts), ptr(rays_a), S, R, ptr(d_ws), stream()),
          "distortion_loss_bw")
    return d_ws
