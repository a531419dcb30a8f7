"""Autograd boundary of the hot path — same class names and forward/backward arities as the
reference's models/custom_functions.py:9-244, implemented over ngp_b200.vren.

    RayAABBIntersector  custom_functions.py:9-30      RaySphereIntersector  :33-53
    RayMarcher          custom_functions.py:56-114    VolumeRenderer        :117-163
    RefLoss             custom_functions.py:165-198   TruncExp / ReLU / TruncTanh  :200-244
    DistortionLoss      losses.py:32-58
"""
import torch

from . import vren


class RayAABBIntersector(torch.autograd.Function):
    """(rays_o, rays_d, center, half_size, max_hits) -> hit_cnt, hits_t (R,max_hits,2), hits_voxel_idx."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, center, half_size, max_hits):
        out = vren.ray_aabb_intersect(rays_o, rays_d, center, half_size, max_hits)
        ctx.mark_non_differentiable(*out)
        return tuple(out)


class RaySphereIntersector(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, center, radii, max_hits):
        out = vren.ray_sphere_intersect(rays_o, rays_d, center, radii, max_hits)
        ctx.mark_non_differentiable(*out)
        return tuple(out)


class RayMarcher(torch.autograd.Function):
    """Occupancy-skipping sample generation.

    forward(rays_o, rays_d, hits_t (R,2), density_bitfield, cascades, scale, exp_step_factor,
            grid_size, max_samples) -> rays_a (R,3), xyzs (S,3), dirs (S,3), deltas (S), ts (S),
            total_samples (0-dim int32 tensor)
    The start-jitter noise is drawn here with torch.rand_like, as the reference does
    (custom_functions.py:85), so the RNG stream advances identically.
    """

    @staticmethod
    def forward(ctx, rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, grid_size,
                max_samples):
        noise = torch.rand_like(rays_o[:, 0])
        rays_a, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(
            rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise, grid_size, max_samples)
        ctx.save_for_backward(rays_a, ts)
        ctx.n_rays = rays_o.shape[0]
        ctx.mark_non_differentiable(rays_a, deltas, ts)
        return rays_a, xyzs, dirs, deltas, ts, counter[0]

    @staticmethod
    def backward(ctx, g_rays_a, g_xyzs, g_dirs, g_deltas, g_ts, g_total):
        # xyz = o + t*d, dirs = d  =>  dL/do = sum_s dL/dxyz_s ; dL/dd = sum_s (t_s*dL/dxyz_s + dL/ddirs_s)
        # (custom_functions.py:104-114; unreachable in the reference because the marcher runs under no_grad)
        rays_a, ts = ctx.saved_tensors
        ray_of_sample = torch.repeat_interleave(rays_a[:, 0], rays_a[:, 2])
        g_o = torch.zeros(ctx.n_rays, 3, device=ts.device, dtype=ts.dtype).index_add_(0, ray_of_sample, g_xyzs)
        g_d = torch.zeros_like(g_o).index_add_(0, ray_of_sample, g_xyzs * ts[:, None] + g_dirs)
        return g_o, g_d, None, None, None, None, None, None, None


class VolumeRenderer(torch.autograd.Function):
    """Front-to-back compositing of packed samples (training).

    forward(sigmas (S), rgbs (S,3), normals_pred (S,3), sems (S,C), deltas (S), ts (S), rays_a (R,3),
            T_threshold, classes) -> total_samples (scalar), opacity (R), depth (R), rgb (R,3),
            normal_pred (R,3), sem (R,C), ws (S)
    """

    @staticmethod
    def forward(ctx, sigmas, rgbs, normals_pred, sems, deltas, ts, rays_a, T_threshold, classes):
        total, opacity, depth, rgb, normal_pred, sem, ws = vren.composite_train_fw(
            sigmas, rgbs, normals_pred, sems, deltas, ts, rays_a, T_threshold, classes)
        ctx.save_for_backward(sigmas, rgbs, normals_pred, deltas, ts, rays_a, opacity, depth, rgb, normal_pred, ws)
        ctx.T_threshold, ctx.classes = T_threshold, classes
        return total.sum(), opacity, depth, rgb, normal_pred, sem, ws

    @staticmethod
    def backward(ctx, g_total, g_opacity, g_depth, g_rgb, g_normal_pred, g_sem, g_ws):
        sigmas, rgbs, normals_pred, deltas, ts, rays_a, opacity, depth, rgb, normal_pred, ws = ctx.saved_tensors
        d_sig, d_rgb, d_nrm, d_sem = vren.composite_train_bw(
            g_opacity.contiguous(), g_depth.contiguous(), g_rgb.contiguous(), g_normal_pred.contiguous(),
            g_sem.contiguous(), g_ws.contiguous(), sigmas, rgbs, normals_pred, ws, deltas, ts, rays_a, opacity, depth,
            rgb, normal_pred, ctx.T_threshold, ctx.classes)
        return d_sig, d_rgb, d_nrm, d_sem, None, None, None, None, None


class VolumeRendererLite(torch.autograd.Function):
    """VolumeRenderer for fields without normal / semantic heads (the ngp_pl-shaped BASELINE model):
    same compositor kernels with the normal and semantic streams switched off (NULL pointers).
    forward(sigmas, rgbs, deltas, ts, rays_a, T_threshold) -> total_samples, opacity, depth, rgb, ws"""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, ts, rays_a, T_threshold):
        from . import _lib
        from ._lib import lib, ptr, check, stream
        _lib.require_device()
        S, R = sigmas.shape[0], rays_a.shape[0]
        dev = sigmas.device
        total = torch.empty(R, dtype=torch.int64, device=dev)
        opacity = torch.empty(R, dtype=torch.float32, device=dev)
        depth = torch.empty(R, dtype=torch.float32, device=dev)
        rgb = torch.empty(R, 3, dtype=torch.float32, device=dev)
        ws = torch.empty(S, dtype=torch.float32, device=dev)
        check(lib.ngp_composite_train_fw(ptr(sigmas), ptr(rgbs), None, None, ptr(deltas), ptr(ts), ptr(rays_a),
                                         float(T_threshold), 0, S, R, ptr(total), ptr(opacity), ptr(depth), ptr(rgb), None,
                                         None, ptr(ws), stream()), "composite_train_fw")
        ctx.save_for_backward(sigmas, rgbs, deltas, ts, rays_a, opacity, depth, rgb, ws)
        ctx.T_threshold = T_threshold
        return total.sum(), opacity, depth, rgb, ws

    @staticmethod
    def backward(ctx, g_total, g_opacity, g_depth, g_rgb, g_ws):
        from ._lib import lib, ptr, check, stream
        sigmas, rgbs, deltas, ts, rays_a, opacity, depth, rgb, ws = ctx.saved_tensors
        S, R = sigmas.shape[0], rays_a.shape[0]
        d_sig = torch.empty_like(sigmas)
        d_rgb = torch.empty_like(rgbs)
        # bound to locals: a temporary .contiguous() copy would be freed (and its block reused by the next copy) before the launch
        g_opacity, g_depth, g_rgb, g_ws = g_opacity.contiguous(), g_depth.contiguous(), g_rgb.contiguous(), g_ws.contiguous()
        check(lib.ngp_composite_train_bw(ptr(g_opacity), ptr(g_depth), ptr(g_rgb), None,
                                         None, ptr(g_ws), ptr(sigmas), ptr(rgbs), ptr(ws), ptr(deltas), ptr(ts),
                                         ptr(rays_a), ptr(opacity), ptr(depth), ptr(rgb), float(ctx.T_threshold), 0, S, R,
                                         ptr(d_sig), ptr(d_rgb), None, None, stream()), "composite_train_bw")
        return d_sig, d_rgb, None, None, None, None


class RefLoss(torch.autograd.Function):
    """Composited Ref-NeRF normal losses: forward(sigmas, normals_diff (S,3), normals_ori (S), deltas,
    ts, rays_a, T_threshold) -> loss_o (R), loss_p (R,3).  No gradient reaches sigmas
    (custom_functions.py:198)."""

    @staticmethod
    def forward(ctx, sigmas, normals_diff, normals_ori, deltas, ts, rays_a, T_threshold):
        loss_o, loss_p = vren.composite_refloss_fw(sigmas, normals_diff, normals_ori, deltas, ts, rays_a, T_threshold)
        ctx.save_for_backward(sigmas, normals_diff, normals_ori, deltas, ts, rays_a, loss_o, loss_p)
        ctx.T_threshold = T_threshold
        return loss_o, loss_p

    @staticmethod
    def backward(ctx, g_loss_o, g_loss_p):
        sigmas, normals_diff, normals_ori, deltas, ts, rays_a, loss_o, loss_p = ctx.saved_tensors
        _, d_diff, d_ori = vren.composite_refloss_bw(g_loss_o.contiguous(), g_loss_p.contiguous(), sigmas, normals_diff,
                                                     normals_ori, deltas, ts, rays_a, loss_o, loss_p, ctx.T_threshold)
        return None, d_diff, d_ori, None, None, None, None


class DistortionLoss(torch.autograd.Function):
    """mip-NeRF-360 distortion loss per ray: forward(ws, deltas, ts, rays_a) -> loss (R)  (losses.py:32-58)."""

    @staticmethod
    def forward(ctx, ws, deltas, ts, rays_a):
        loss, ws_inc, wts_inc = vren.distortion_loss_fw(ws, deltas, ts, rays_a)
        ctx.save_for_backward(ws_inc, wts_inc, ws, deltas, ts, rays_a)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        ws_inc, wts_inc, ws, deltas, ts, rays_a = ctx.saved_tensors
        return vren.distortion_loss_bw(g_loss.contiguous(), ws_inc, wts_inc, ws, deltas, ts, rays_a), None, None, None


class NegNormalize(torch.autograd.Function):
    """-F.normalize(x * scale, p=2, dim=-1, eps) on (N,3) rows as one kernel per direction (csrc/normals.cu) — the field's
    normals_raw / normals_pred (models/networks.py:209,222-223).  scale: None or three python floats."""

    @staticmethod
    def forward(ctx, x, scale=None, eps=1e-6):
        from . import _lib
        from ._lib import lib, ptr, check, stream
        _lib.require_device()
        x = x.contiguous().float()
        n = x.shape[0]
        sc = tuple(float(v) for v in scale) if scale is not None else (1.0, 1.0, 1.0)
        y = torch.empty_like(x); inv = torch.empty(n, device=x.device)
        check(lib.ngp_neg_normalize_fw(ptr(x), *sc, float(eps), n, ptr(y), ptr(inv), stream()), "neg_normalize_fw")
        ctx.sc = sc
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        from ._lib import lib, ptr, check, stream
        y, inv = ctx.saved_tensors
        gx = torch.empty_like(y)
        gy = gy.contiguous().float()
        check(lib.ngp_neg_normalize_bw(ptr(gy), ptr(y), ptr(inv), *ctx.sc, y.shape[0], ptr(gx), stream()), "neg_normalize_bw")
        return gx, None, None


class RefLossPrep(torch.autograd.Function):
    """normals_diff = (normals_raw - normals_pred)**2 and normals_ori = clamp(sum(normals_raw * normalize(dirs)), min=0)**2
    (models/rendering.py:243-246) as one kernel per direction (csrc/normals.cu)."""

    @staticmethod
    def forward(ctx, normals_raw, normals_pred, dirs):
        from . import _lib
        from ._lib import lib, ptr, check, stream
        _lib.require_device()
        nr, npd, d = normals_raw.contiguous().float(), normals_pred.contiguous().float(), dirs.contiguous().float()
        n = nr.shape[0]
        diff = torch.empty_like(nr); ori = torch.empty(n, device=nr.device)
        check(lib.ngp_refloss_prep_fw(ptr(nr), ptr(npd), ptr(d), n, ptr(diff), ptr(ori), stream()), "refloss_prep_fw")
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(nr, npd, d)
        return diff, ori

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_diff, g_ori):
        from ._lib import lib, ptr, check, stream
        nr, npd, d = ctx.saved_tensors
        need_r, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_raw = torch.empty_like(nr) if need_r else None
        g_pred = torch.empty_like(nr) if need_p else None
        # bound to locals: two temporary .contiguous() copies inside one call expression can end up in the same freed block
        g_diff = g_diff.contiguous() if g_diff is not None else None
        g_ori = g_ori.contiguous() if g_ori is not None else None
        check(lib.ngp_refloss_prep_bw(ptr(nr), ptr(npd), ptr(d), ptr(g_diff), ptr(g_ori), nr.shape[0], ptr(g_raw), ptr(g_pred), stream()),
              "refloss_prep_bw")
        return g_raw, g_pred, None


class TruncExp(torch.autograd.Function):
    """exp with the backward evaluated at clamp(x, -7, 7) (custom_functions.py:200-211)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-7, 7))


class TruncTanh(torch.autograd.Function):
    """tanh with the backward evaluated at clamp(x, -15, 15) (custom_functions.py:231-244)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.tanh(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * (1 - torch.tanh(x.clamp(-15, 15)) ** 2)


class ReLU(torch.autograd.Function):
    """The reference's leaky-gradient ReLU: zero-side gradient is the constant 1e-6, not g*0
    (custom_functions.py:213-229)."""

    @staticmethod
    def forward(ctx, x):
        mask = x > 0
        ctx.save_for_backward(mask)
        return torch.where(mask, x, torch.zeros_like(x))

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return torch.where(mask, g, torch.full_like(g, 1e-6))


class ExpandPerRay(torch.autograd.Function):
    """Per-ray rows repeated for every sample of their ray — `torch.repeat_interleave(v[rays_a[:, 0]], rays_a[:, 2], 0)`
    of models/rendering.py:217-219 (the appearance embedding) — as one kernel per direction (ngp_expand_per_ray /
    ngp_reduce_per_ray).  v (R_all, W <= 32) float32, rays_a (R,3) int64, n_samples = rays_a[:, 2].sum()."""

    @staticmethod
    def forward(ctx, v, rays_a, n_samples):
        from . import _lib
        from ._lib import lib, ptr, check, stream
        _lib.require_device()
        v = v.contiguous()
        out = torch.empty(n_samples, v.shape[1], dtype=torch.float32, device=v.device)
        check(lib.ngp_expand_per_ray(ptr(v), ptr(rays_a), rays_a.shape[0], v.shape[1], ptr(out), stream()), "expand_per_ray")
        ctx.save_for_backward(rays_a)
        ctx.shape = v.shape
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        from ._lib import lib, ptr, check, stream
        (rays_a,) = ctx.saved_tensors
        dv = torch.zeros(ctx.shape, dtype=torch.float32, device=dout.device)
        check(lib.ngp_reduce_per_ray(ptr(dout.contiguous()), ptr(rays_a), rays_a.shape[0], ctx.shape[1], ptr(dv), stream()),
              "reduce_per_ray")
        return dv, None, None
