"""Renderer orchestration: AABB -> march -> field -> composite, training and test time.

Same entry point and result keys as the reference's models/rendering.py:13-251
(`render(model, rays_o, rays_d, **kwargs)` with kwargs test_time, exp_step_factor, T_threshold,
num_classes, use_skybox, random_bg, embedding_a, max_samples, to_cpu, to_numpy).
"""
import torch
import torch.nn.functional as F

from . import vren
from .custom_functions import ExpandPerRay, RayAABBIntersector, RayMarcher, RefLoss, RefLossPrep, VolumeRenderer, VolumeRendererLite

MAX_SAMPLES = 1024      # models/rendering.py:9
NEAR_DISTANCE = 0.01    # models/rendering.py:10


def render(model, rays_o, rays_d, **kwargs):
    rays_o, rays_d = rays_o.contiguous(), rays_d.contiguous()
    _, hits_t, _ = RayAABBIntersector.apply(rays_o, rays_d, model.center, model.half_size, 1)
    # rays that start inside the box begin at the near plane (rendering.py:29-30)
    if hits_t.is_cuda and hits_t.is_contiguous() and hits_t.dtype == torch.float32:
        from ._lib import lib, ptr, check, stream
        check(lib.ngp_near_clamp(ptr(hits_t), hits_t.shape[0], hits_t.stride(0), NEAR_DISTANCE, stream()), "near_clamp")
    else:
        t1 = hits_t[:, 0, 0]
        hits_t[:, 0, 0] = torch.where((t1 >= 0) & (t1 < NEAR_DISTANCE), torch.full_like(t1, NEAR_DISTANCE), t1)

    fn = _render_rays_test if kwargs.get("test_time", False) else _render_rays_train
    results = fn(model, rays_o, rays_d, hits_t, **kwargs)
    if kwargs.get("to_cpu", False):
        for k, v in results.items():
            if torch.is_tensor(v):
                v = v.cpu()
                if kwargs.get("to_numpy", False):
                    v = v.numpy()
            results[k] = v
    return results


def volume_render(model, rays_o, rays_d, hits_t, opacity, depth, rgb, normal_pred, normal_raw, sem, **kwargs):
    """Adaptive test-time march/compose loop (rendering.py:46-133): while rays are alive, march each
    alive ray to its next N_samples occupied samples, evaluate the field there, composite in place,
    drop converged rays.

    kwargs['sample_schedule']:
      'reference' — the reference's chunk sizes: max(min(N_rays//N_alive, 64), min_samples) per round,
                    i.e. ONE sample per ray per round while most rays are alive (dozens to hundreds of
                    rounds, each with several host syncs);
      'geometric' (default) — 4, 8, 16, 32, 64, then 128 samples per round: <= 13 rounds to reach MAX_SAMPLES and one
                    host read-back per round.  The marcher resumes every ray exactly where it stopped
                    and the compositor is a per-ray sequential recurrence, so the composited result
                    does not depend on the chunking (up to the association of per-round partial sums).
    """
    N_rays = len(rays_o)
    device = rays_o.device
    esf = kwargs.get("exp_step_factor", 0.)
    classes = kwargs.get("num_classes", 7)
    T_thr = kwargs.get("T_threshold", 1e-4)
    max_samples = kwargs.get("max_samples", MAX_SAMPLES)
    min_samples = 1 if esf == 0 else 4
    schedule = kwargs.get("sample_schedule", "geometric")
    alive = torch.arange(N_rays, device=device)
    samples = 0
    rnd = 0
    total_samples = torch.zeros((), dtype=torch.int64, device=device)

    while samples < max_samples:
        N_alive = len(alive)
        if N_alive == 0:
            break
        if schedule == "reference":
            N_samples = max(min(N_rays // N_alive, 64), min_samples)
        else:
            N_samples = min(4 << rnd, 128, max_samples - samples)
        rnd += 1
        samples += N_samples
        xyzs, dirs, deltas, ts, N_eff = vren.raymarching_test(
            rays_o, rays_d, hits_t, alive, model.density_bitfield, model.cascades, model.scale, esf,
            model.grid_size, MAX_SAMPLES, N_samples)
        total_samples += N_eff.sum()
        # valid slots are the first N_eff of each row (the reference finds them as dirs != 0)
        valid = torch.arange(N_samples, device=device)[None, :] < N_eff[:, None]
        flat = valid.reshape(-1)
        idx = torch.nonzero(flat)[:, 0]
        if idx.numel() == 0:
            break
        kw = dict(kwargs)
        if isinstance(kw.get("embedding_a", None), torch.Tensor) and kw["embedding_a"].shape[0] == N_rays:
            ray_of_slot = alive[:, None].expand(-1, N_samples).reshape(-1)[idx]
            kw["embedding_a"] = kw["embedding_a"][ray_of_slot]
        _s, _c, _np, _nr, _sem = model.forward_test(xyzs.reshape(-1, 3)[idx], dirs.reshape(-1, 3)[idx], **kw)
        n_slots = N_alive * N_samples
        sigmas = torch.zeros(n_slots, device=device).index_copy_(0, idx, _s.detach().float())
        rgbs = torch.zeros(n_slots, 3, device=device).index_copy_(0, idx, _c.detach().float())
        n_pred = torch.zeros(n_slots, 3, device=device).index_copy_(0, idx, _np.detach().float())
        n_raw = torch.zeros(n_slots, 3, device=device).index_copy_(0, idx, _nr.detach().float())
        sems = torch.zeros(n_slots, classes, device=device)
        if classes > 0 and _sem.shape[-1] == classes:
            sems.index_copy_(0, idx, _sem.detach().float())
        vren.composite_test_fw(
            sigmas.view(N_alive, N_samples), rgbs.view(N_alive, N_samples, 3), n_pred.view(N_alive, N_samples, 3),
            n_raw.view(N_alive, N_samples, 3), sems.view(N_alive, N_samples, classes), deltas, ts, hits_t, alive,
            T_thr, classes, N_eff, opacity, depth, rgb, normal_pred, normal_raw, sem)
        alive = alive[alive >= 0]

    rgb_bg = model.forward_skybox(rays_d) if kwargs.get("use_skybox", False) else torch.zeros(3, device=device)
    rgb += rgb_bg * (1 - opacity)[:, None]
    return total_samples


@torch.no_grad()
def render_wavefront(model, rays_o, rays_d, hits_t, opacity, depth, rgb, normal_pred=None, normal_raw=None, sem=None, **kwargs):
    """Test-time renderer: the reference's march -> evaluate -> composite -> compact loop (rendering.py:46-133) as ONE fused
    advance kernel per round (composite the previous round incl. the normal / semantic streams, compact the survivors, march
    the next round), a packing pass, and the field evaluation on exactly the samples that exist.  Rounds take 4, 8, 16, ...
    samples; round 0 drops the frame's provably empty rays before marching (single-cascade scenes); the host reads two
    counters per round.  Same samples and the same per-ray compositing recurrence as the reference loop.
    normal_pred / normal_raw (R,3) and sem (R,C): accumulators of fields with normal / semantic heads (all three or none)."""
    from . import _lib
    from ._lib import lib, ptr, check, stream
    _lib.require_device()
    N_rays, dev = len(rays_o), rays_o.device
    esf = float(kwargs.get("exp_step_factor", 0.))
    T_thr = float(kwargs.get("T_threshold", 1e-4))
    max_samples = int(kwargs.get("max_samples", MAX_SAMPLES))
    full = normal_pred is not None
    classes = int(sem.shape[1]) if (full and sem is not None) else 0
    geo = (model.cascades, float(model.scale), esf, model.grid_size, MAX_SAMPLES)
    bitfield = ptr(model.density_bitfield)
    alive_in, n_alive = None, N_rays
    prev = None                        # (rays_a, sigmas, rgbs, deltas, ts, normals_pred, normals_raw, sems) of the previous round
    counters = torch.zeros(2, dtype=torch.int32, device=dev)
    ws = torch.empty(int(lib.ngp_render_workspace_bytes(N_rays)), dtype=torch.uint8, device=dev)     # sized for round 0, reused
    alive_bufs = [torch.empty(N_rays, dtype=torch.int64, device=dev) for _ in range(2)]
    per_ray_emb = isinstance(kwargs.get("embedding_a", None), torch.Tensor) and kwargs["embedding_a"].shape[0] == N_rays
    total = 0
    samples, rnd = 0, 0
    st = stream()
    while n_alive > 0:
        n_next = min(4 << rnd, 128, max_samples - samples) if samples < max_samples else 0
        alive_out = alive_bufs[rnd & 1]
        rnd += 1; samples += n_next
        pr = prev if prev is not None else (None,) * 8
        check(lib.ngp_render_advance_full(ptr(rays_o), ptr(rays_d), ptr(hits_t), ptr(alive_in), n_alive, ptr(pr[0]), ptr(pr[1]),
                                          ptr(pr[2]), ptr(pr[3]), ptr(pr[4]), T_thr, bitfield, *geo, n_next,
                                          ptr(opacity), ptr(depth), ptr(rgb), ptr(alive_out), ptr(counters), ptr(ws),
                                          ptr(pr[5]), ptr(pr[6]), ptr(pr[7]), classes, ptr(normal_pred), ptr(normal_raw), ptr(sem), st),
              "render_advance")
        if n_next == 0:
            break
        cap = n_alive * n_next
        rays_a = torch.empty(n_alive, 3, dtype=torch.int64, device=dev)
        xyzs = torch.empty(cap, 3, device=dev); dirs = torch.empty(cap, 3, device=dev)
        deltas = torch.empty(cap, device=dev); ts = torch.empty(cap, device=dev)
        check(lib.ngp_render_emit(ptr(rays_o), ptr(rays_d), ptr(hits_t), ptr(alive_out), n_alive, bitfield,
                                  *geo, ptr(ws), cap, ptr(rays_a), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(ts), ptr(counters),
                                  st), "render_emit")
        n_alive_out, n_pts = (int(v) for v in counters.tolist())          # the round's one host read-back
        total += n_pts
        if n_alive_out == 0 or n_pts == 0:
            break
        kw = kwargs
        if per_ray_emb:                 # per-ray tensors follow their rays' samples (rendering.py:217-219 at test time)
            kw = dict(kwargs)
            ra = rays_a[:n_alive_out]
            kw["embedding_a"] = torch.repeat_interleave(kwargs["embedding_a"][ra[:, 0]], ra[:, 2], 0, output_size=n_pts)
        out = model.forward_test(xyzs[:n_pts], dirs[:n_pts], **kw)
        sig, col = out[0].contiguous(), out[1].contiguous()
        extra = (out[2].float().contiguous(), out[3].float().contiguous(), out[4].float().contiguous()) if full else (None, None, None)
        prev = (rays_a, sig, col, deltas, ts) + extra
        alive_in, n_alive = alive_out, n_alive_out
    return total


def _round_schedule(max_samples):
    """samples per round: 4, 8, 16, 32, 64, then 128 until max_samples is reached"""
    out, tot, k = [], 0, 0
    while tot < max_samples:
        n = min(4 << k, 128, max_samples - tot)
        out.append(n); tot += n; k += 1
    return out


@torch.no_grad()
def render_wavefront_compact(model, rays_o, rays_d, hits_t, opacity, depth, rgb, **kwargs):
    """render_wavefront for the ngp_pl-shaped field (networks.NGPCompact) with the round loop resident on the device: every
    round is ONE C-ABI call (ngp_render_round_compact: advance + pack + hash grid + both MLPs) whose launches read their live
    counts from device memory, so rounds are enqueued back to back from BOUNDS of the alive count and the host never waits
    for the round it has just enqueued: it blocks once after round 0 (where the frame's empty rays leave: the bound drops 4-5x)
    and from then on reads each round's counters one round late, while the next round is already running.
    The reference's loop (models/rendering.py:75-124) synchronises three times per round."""
    from . import _lib, tcnn
    from ._lib import lib, ptr, check, stream
    _lib.require_device()
    N_rays, dev = len(rays_o), rays_o.device
    esf = float(kwargs.get("exp_step_factor", 0.))
    T_thr = float(kwargs.get("T_threshold", 1e-4))
    sched = _round_schedule(int(kwargs.get("max_samples", MAX_SAMPLES)))
    geo = (model.cascades, float(model.scale), esf, model.grid_size, MAX_SAMPLES)
    enc, snet, cnet = model.xyz_encoder, model.sigma_net, model.rgb_net
    g = enc.grid
    LF = g.n_levels * g.n_features
    tile_bytes = int(lib.ngp_feature_tile_bytes(g.n_levels, g.n_features))
    aabb = tcnn._aabb_arg(model.aabb())
    table, sp, cp = enc.params.detach(), snet.params.detach(), cnet.params.detach()
    field = (aabb, ptr(table), 0, *g.args(), ptr(sp), ptr(cp), snet.mlp.width, cnet.mlp.n_hidden)
    bitfield = ptr(model.density_bitfield)
    R = len(sched)
    counters = torch.zeros(R + 1, 2, dtype=torch.int32, device=dev)                 # one pair per round: late read-backs stay valid
    host = torch.zeros(R + 1, 2, dtype=torch.int32).pin_memory()
    events = [None] * (R + 1)
    ws = torch.empty(int(lib.ngp_render_workspace_bytes(N_rays)), dtype=torch.uint8, device=dev)
    alive_bufs = [torch.empty(N_rays, dtype=torch.int64, device=dev) for _ in range(2)]
    st = stream()
    cbase = counters.data_ptr()
    alive_in, bound, prev = None, N_rays, (None,) * 5
    known = -1                                   # last round whose counters have reached the host
    last = -1
    for k, n_next in enumerate(sched):
        if bound == 0:
            break
        cap = bound * n_next
        alive_out = alive_bufs[k & 1]
        rays_a = torch.empty(bound, 3, dtype=torch.int64, device=dev)
        xyzs = torch.empty(cap, 3, device=dev); dirs = torch.empty(cap, 3, device=dev)
        deltas = torch.empty(cap, device=dev); ts = torch.empty(cap, device=dev)
        tiles = torch.empty((cap + 127) // 128 * tile_bytes, dtype=torch.uint8, device=dev)
        h = torch.empty(cap, 16, device=dev); sig = torch.empty(cap, device=dev); col = torch.empty(cap, 3, device=dev)
        check(lib.ngp_render_round_compact(ptr(rays_o), ptr(rays_d), ptr(hits_t), ptr(alive_in), bound, (cbase + 8 * (k - 1)) if k else None,
                                           ptr(prev[0]), ptr(prev[1]), ptr(prev[2]), ptr(prev[3]), ptr(prev[4]), T_thr, bitfield, *geo, n_next,
                                           ptr(opacity), ptr(depth), ptr(rgb), ptr(alive_out), cbase + 8 * k, ptr(ws),
                                           ptr(rays_a), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(ts), *field, ptr(tiles), ptr(h), ptr(sig), ptr(col), st),
              "render_round_compact")
        host[k].copy_(counters[k], non_blocking=True)
        events[k] = torch.cuda.Event(); events[k].record()
        prev, alive_in, last = (rays_a, sig, col, deltas, ts), alive_out, k
        # the bound of the NEXT round: the newest alive count the host may know without stalling the device — round k itself
        # after round 0 (one blocking wait, the frame's empty rays have just left), round k - 1 afterwards
        want = k if k == 0 else k - 1
        if want > known:
            events[want].synchronize(); known = want
        bound = min(bound, int(host[known, 0]))
    # composite the last marched round (n_next = 0: no march, no field)
    if last >= 0 and bound > 0:
        check(lib.ngp_render_round_compact(ptr(rays_o), ptr(rays_d), ptr(hits_t), ptr(alive_in), bound, cbase + 8 * last,
                                           ptr(prev[0]), ptr(prev[1]), ptr(prev[2]), ptr(prev[3]), ptr(prev[4]), T_thr, bitfield, *geo, 0,
                                           ptr(opacity), ptr(depth), ptr(rgb), ptr(alive_bufs[(last + 1) & 1]), cbase + 8 * R, ptr(ws),
                                           None, None, None, None, None, *field, None, None, None, None, st), "render_round_compact")
    return int(counters[:R, 1].sum())


_PIPE = {}          # per device: (copy stream, pinned counter buffer)


@torch.no_grad()
def render_wavefront_pipelined(model, rays_o, rays_d, hits_t, opacity, depth, rgb, **kwargs):
    """render_wavefront for the ngp_pl-shaped field with the host read-back taken off the critical path: a round's field evaluation
    (ngp_field_compact_fw: hash grid + both MLPs, live sample count read from device memory) is enqueued right behind its emit pass,
    and the round's two counters travel to the host on a side stream WHILE the field runs — so every round is still sized by EXACT
    counts (unlike the bound-sized render_wavefront_compact) but the GPU no longer idles between emit and field while the host waits,
    allocates and launches.  Same kernels on the same samples as render_wavefront: bit-identical images."""
    from . import _lib, tcnn
    from ._lib import lib, ptr, check, stream
    _lib.require_device()
    N_rays, dev = len(rays_o), rays_o.device
    esf = float(kwargs.get("exp_step_factor", 0.))
    T_thr = float(kwargs.get("T_threshold", 1e-4))
    sched = _round_schedule(int(kwargs.get("max_samples", MAX_SAMPLES)))
    geo = (model.cascades, float(model.scale), esf, model.grid_size, MAX_SAMPLES)
    enc, snet, cnet = model.xyz_encoder, model.sigma_net, model.rgb_net
    g = enc.grid
    tile_bytes = int(lib.ngp_feature_tile_bytes(g.n_levels, g.n_features))
    table, sp, cp = enc.params.detach(), snet.params.detach(), cnet.params.detach()
    field = (tcnn._aabb_arg(model.aabb()), ptr(table), 0, *g.args(), ptr(sp), ptr(cp), snet.mlp.width, cnet.mlp.n_hidden)
    bitfield = ptr(model.density_bitfield)
    R = len(sched)
    if dev not in _PIPE:
        _PIPE[dev] = (torch.cuda.Stream(device=dev), torch.zeros(64, 2, dtype=torch.int32).pin_memory())
    copy_stream, host = _PIPE[dev]
    counters = torch.zeros(R + 1, 2, dtype=torch.int32, device=dev)
    ws = torch.empty(int(lib.ngp_render_workspace_bytes(N_rays)), dtype=torch.uint8, device=dev)
    alive_bufs = [torch.empty(N_rays, dtype=torch.int64, device=dev) for _ in range(2)]
    main = torch.cuda.current_stream()
    st = stream()
    cbase = counters.data_ptr()
    alive_in, n_alive, prev = None, N_rays, (None,) * 5
    total, last = 0, -1
    none3 = (None, None, None)
    for k, n_next in enumerate(sched):
        cap = n_alive * n_next
        alive_out = alive_bufs[k & 1]
        rays_a = torch.empty(n_alive, 3, dtype=torch.int64, device=dev)
        xyzs = torch.empty(cap, 3, device=dev); dirs = torch.empty(cap, 3, device=dev)
        deltas = torch.empty(cap, device=dev); ts = torch.empty(cap, device=dev)
        tiles = torch.empty((cap + 127) // 128 * tile_bytes, dtype=torch.uint8, device=dev)
        h = torch.empty(cap, 16, device=dev); sig = torch.empty(cap, device=dev); col = torch.empty(cap, 3, device=dev)
        check(lib.ngp_render_advance_full(ptr(rays_o), ptr(rays_d), ptr(hits_t), ptr(alive_in), n_alive, ptr(prev[0]), ptr(prev[1]),
                                          ptr(prev[2]), ptr(prev[3]), ptr(prev[4]), T_thr, bitfield, *geo, n_next,
                                          ptr(opacity), ptr(depth), ptr(rgb), ptr(alive_out), cbase + 8 * k, ptr(ws),
                                          *none3, 0, *none3, st), "render_advance")
        check(lib.ngp_render_emit(ptr(rays_o), ptr(rays_d), ptr(hits_t), ptr(alive_out), n_alive, bitfield, *geo, ptr(ws), cap,
                                  ptr(rays_a), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(ts), cbase + 8 * k, st), "render_emit")
        ev = torch.cuda.Event(); ev.record(main)
        check(lib.ngp_field_compact_fw(ptr(xyzs), ptr(dirs), cap, cbase + 8 * k + 4, *field, ptr(tiles), ptr(h), ptr(sig), ptr(col), st),
              "field_compact_fw")
        with torch.cuda.stream(copy_stream):               # the counters leave as soon as the emit pass is done, the field keeps the GPU busy
            copy_stream.wait_event(ev)
            host[k % 64].copy_(counters[k], non_blocking=True)
            done = torch.cuda.Event(); done.record(copy_stream)
        done.synchronize()
        n_alive_out, n_pts = int(host[k % 64, 0]), int(host[k % 64, 1])
        total += n_pts
        prev, alive_in, last = (rays_a, sig, col, deltas, ts), alive_out, k
        n_alive = n_alive_out
        if n_alive_out == 0 or n_pts == 0:
            break
    if last >= 0 and n_alive > 0:        # composite the last marched round (n_next = 0: no march)
        check(lib.ngp_render_advance_full(ptr(rays_o), ptr(rays_d), ptr(hits_t), ptr(alive_in), n_alive, ptr(prev[0]), ptr(prev[1]),
                                          ptr(prev[2]), ptr(prev[3]), ptr(prev[4]), T_thr, bitfield, *geo, 0,
                                          ptr(opacity), ptr(depth), ptr(rgb), ptr(alive_bufs[(last + 1) & 1]), cbase + 8 * R, ptr(ws),
                                          *none3, 0, *none3, st), "render_advance")
    main.wait_stream(copy_stream)
    return total


@torch.no_grad()
def _render_rays_test(model, rays_o, rays_d, hits_t, **kwargs):
    """rendering.py:135-190."""
    hits_t = hits_t[:, 0, :].contiguous()
    classes = kwargs.get("num_classes", 7)
    N_rays, device = len(rays_o), rays_o.device
    if kwargs.get("renderer", "wavefront") == "wavefront" and classes <= 32:
        opacity = torch.zeros(N_rays, device=device); depth = torch.zeros(N_rays, device=device)
        rgb = torch.zeros(N_rays, 3, device=device)
        if getattr(model, "has_normals", True):          # fields with normal / semantic heads: all six accumulators
            normal_pred = torch.zeros(N_rays, 3, device=device); normal_raw = torch.zeros(N_rays, 3, device=device)
            sem = torch.zeros(N_rays, classes, device=device)
            total = render_wavefront(model, rays_o, rays_d, hits_t, opacity, depth, rgb, normal_pred, normal_raw, sem, **kwargs)
            semantic = torch.argmax(sem, dim=-1, keepdim=True) if classes > 0 else torch.zeros(N_rays, 1, dtype=torch.long, device=device)
            normal_pred, normal_raw = F.normalize(normal_pred, dim=-1), F.normalize(normal_raw, dim=-1)
        else:
            # device_loop: rounds enqueued from bounds, counters read one round late (render_wavefront_compact).  Opt-in: measured on
            # B200 (tools/render_probe.py, medians over 12-20 frames, profiles/r02d_render_probe.txt) the host-driven loop below is the
            # faster one at every frame size — 1080p 11.5 vs 13.6 ms, 4K 43.0 vs 52.9 ms, a 1/8-frame tile 3.87 vs 3.99 ms: sizing
            # every round by a bound of the alive count costs more (buffers, idle threads) than the two-counter read-back it avoids.
            compact = (kwargs.get("device_loop", False) and getattr(model, "fused_density", False) and hasattr(model, "sigma_net")
                       and model.xyz_encoder.params.dtype == torch.float32 and model.sigma_net.mlp.n_hidden == 1 and model.sigma_net.mlp.n_out == 16
                       and model.rgb_net.mlp.n_out == 3 and model.rgb_net.mlp.width == model.sigma_net.mlp.width)
            fused_field = (getattr(model, "fused_density", False) and hasattr(model, "sigma_net")
                           and model.xyz_encoder.params.dtype == torch.float32 and model.sigma_net.mlp.n_hidden == 1 and model.sigma_net.mlp.n_out == 16
                           and model.rgb_net.mlp.n_out == 3 and model.rgb_net.mlp.width == model.sigma_net.mlp.width)
            # pipelined rounds pay for bound-sized field launches (grids for n_alive * n_next slots, most of them empty): measured
            # (profiles/r02d_render_probe.txt) +8 % on a 259 k-ray tile, -5 % at 2 M rays, -30 % at 8 M rays -> small batches only
            pipelined = fused_field and kwargs.get("pipelined", N_rays <= 600_000) and not compact
            fn = render_wavefront_compact if compact else (render_wavefront_pipelined if pipelined else render_wavefront)
            total = fn(model, rays_o, rays_d, hits_t, opacity, depth, rgb, **kwargs)
            normal_pred = normal_raw = torch.zeros(N_rays, 3, device=device)
            semantic = torch.zeros(N_rays, 1, dtype=torch.long, device=device)
        if kwargs.get("use_skybox", False):              # rendering.py:126-131
            rgb += model.forward_skybox(rays_d) * (1 - opacity)[:, None]
        return {"opacity": opacity, "depth": depth, "rgb": rgb, "normal_pred": normal_pred, "normal_raw": normal_raw,
                "semantic": semantic, "total_samples": torch.tensor(total, device=device),
                "points": rays_o + rays_d * depth.unsqueeze(-1), "mask": torch.zeros(N_rays, device=device)}
    opacity = torch.zeros(N_rays, device=device)
    depth = torch.zeros(N_rays, device=device)
    rgb = torch.zeros(N_rays, 3, device=device)
    normal_pred = torch.zeros(N_rays, 3, device=device)
    normal_raw = torch.zeros(N_rays, 3, device=device)
    sem = torch.zeros(N_rays, classes, device=device)
    total_samples = volume_render(model, rays_o, rays_d, hits_t, opacity, depth, rgb, normal_pred, normal_raw, sem,
                                  **kwargs)
    return {
        "opacity": opacity, "depth": depth, "rgb": rgb,
        "normal_pred": F.normalize(normal_pred, dim=-1), "normal_raw": F.normalize(normal_raw, dim=-1),
        "semantic": torch.argmax(sem, dim=-1, keepdim=True) if classes > 0 else torch.zeros(N_rays, 1, dtype=torch.long, device=device),
        "total_samples": total_samples,
        "points": rays_o + rays_d * depth.unsqueeze(-1),
        "mask": torch.zeros(N_rays, device=device),
    }


def _render_rays_train(model, rays_o, rays_d, hits_t, **kwargs):
    """rendering.py:193-251."""
    esf = kwargs.get("exp_step_factor", 0.)
    T_thr = kwargs.get("T_threshold", 1e-4)
    classes = kwargs.get("num_classes", 7)
    results = {}
    with torch.no_grad():
        rays_a, xyzs, dirs, results["deltas"], results["ts"], total_samples = RayMarcher.apply(
            rays_o, rays_d, hits_t[:, 0].contiguous(), model.density_bitfield, model.cascades, model.scale, esf,
            model.grid_size, MAX_SAMPLES)
    results["rays_a"] = rays_a
    results["total_samples"] = total_samples

    kw = dict(kwargs)
    for k, v in kwargs.items():          # per-ray tensors are repeated per sample (rendering.py:217-219)
        if isinstance(v, torch.Tensor):
            if v.is_cuda and v.dtype == torch.float32 and v.dim() == 2 and v.shape[1] <= 32 and rays_a.is_contiguous():
                kw[k] = ExpandPerRay.apply(v, rays_a, xyzs.shape[0])
            else:
                kw[k] = torch.repeat_interleave(v[rays_a[:, 0]], rays_a[:, 2], 0, output_size=xyzs.shape[0])
    sigmas, rgbs, normals_raw, normals_pred, sems = model(xyzs, dirs, **kw)
    results["sigma"], results["xyzs"] = sigmas, xyzs
    lite = normals_pred is None          # field without normal / semantic heads (NGPCompact)

    if lite:
        results["vr_samples"], results["opacity"], results["depth"], results["rgb"], results["ws"] = \
            VolumeRendererLite.apply(sigmas.contiguous(), rgbs.contiguous(), results["deltas"], results["ts"], rays_a, T_thr)
    else:
        (results["vr_samples"], results["opacity"], results["depth"], results["rgb"], results["normal_pred"],
         results["semantic"], results["ws"]) = VolumeRenderer.apply(
            sigmas.contiguous(), rgbs.contiguous(), normals_pred.contiguous(), sems.contiguous(), results["deltas"],
            results["ts"], rays_a, T_thr, classes)

    if kwargs.get("use_skybox", False):
        rgb_bg = model.forward_skybox(rays_d)
    elif esf != 0 and kwargs.get("random_bg", False):
        rgb_bg = torch.rand(3, device=rays_o.device)
    else:
        rgb_bg = None                    # black background: rgb + 0 * (1 - opacity) is rgb itself (rendering.py:236-241), no kernels for it
    if rgb_bg is not None:
        results["rgb"] = results["rgb"] + rgb_bg * (1 - results["opacity"])[:, None]

    if lite:
        return results
    # Ref-NeRF regularisers (rendering.py:243-249)
    if normals_raw.is_cuda and normals_raw.dtype == torch.float32 and normals_pred.dtype == torch.float32:
        normals_diff, normals_ori = RefLossPrep.apply(normals_raw, normals_pred, dirs)
    else:
        normals_diff = (normals_raw - normals_pred) ** 2
        view = F.normalize(dirs, p=2, dim=-1, eps=1e-6)
        normals_ori = torch.clamp(torch.sum(normals_raw * view, dim=-1), min=0.) ** 2
    results["Ro"], results["Rp"] = RefLoss.apply(sigmas.detach().contiguous(), normals_diff.contiguous(),
                                                 normals_ori.contiguous(), results["deltas"], results["ts"], rays_a,
                                                 T_thr)
    return results
