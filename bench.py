#!/usr/bin/env python
"""Headline benchmark: training rays/s (fw+bw+Adam) of the instant-ngp-pp hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Lego-shaped procedural scene (800x800, 100 views, scale 0.5),
ngp_pl-shaped field (16-level F2 hash grid T=2^19, 64-wide MLPs), 2^18 rays per GPU per step,
fp32 tables / bf16 tensor-core operands with fp32 accumulation.  One step = occupancy update (every
16th) + AABB + march + hash encode + MLPs + composite + losses + backward + (all-reduce) + Adam.
`value` times the step on device-resident batches; `e2e` times the same step fed as the reference's loader feeds it
— (img_idxs, pix_idxs, rgb) from pinned host memory, copied one step ahead on a copy stream into a double buffer, rays
generated on the device (ngp_get_rays), the step's loss read back on the host every step.
--workload street | playground run the BASELINE.json configs[3] / configs[2] shapes (not the headline).
Prints ONE JSON line (see README / DESIGN.md §6 for every key).
"""
import argparse
import json
import os
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")   # per-step sample counts vary: no cudaMalloc/cudaFree churn
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "instant-ngp-pp_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

R_PER_GPU = 1 << 18
WORKLOAD = "lego-shaped 800x800x100 views, scale 0.5, hashgrid L16 F2 T2^19 + 64-wide MLPs, 2^18 rays/GPU/step"
# --workload street: BASELINE.json configs[3] shape (not the headline): unbounded street scene, scale 8 -> 5 cascades,
# exponential stepping 1/256, T = 2^22 table (363 MB fp32: does NOT fit the L2), distortion loss
WORKLOADS = {
    "lego": dict(scene="lego", scale=0.5, log2_T=19, esf=0.0, lr=1e-2, views=100, name=WORKLOAD),
    "street": dict(scene="street", scale=8.0, log2_T=22, esf=1.0 / 256, lr=2e-3, views=128,
                   name="street-shaped (KITTI-360-1538 shape) 1408x376x128 views, scale 8 (5 cascades), exp_step 1/256, "
                        "hashgrid L16 F2 T2^22 + 64-wide MLPs, distortion loss, 2^18 rays/GPU/step"),
    # --workload playground: BASELINE.json configs[2] shape: the reference's LITERAL field (networks.py:13-163: two F=8 grids
    # T=2^19 / 2^21 = 174 + 588 MiB, Softplus density net with autograd normals, rgb / normal / semantic heads), scale 8,
    # appearance embedding 8, 7 classes, normal + semantic compositing, Ref-NeRF + semantic + distortion losses,
    # 2^18 rays over 8 GPUs = 2^15 rays per GPU
    "playground": dict(scene="street", scale=8.0, esf=1.0 / 256, lr=2e-3, views=128, field="ngp", rays=1 << 15, classes=7, embed_a_len=8,
                       name="playground-shaped (TanksAndTemples-BG shape: unbounded, scale 8, 5 cascades, exp_step 1/256), reference-literal "
                            "NGP field (2 hash grids L16 F8 T2^19/T2^21, density net + autograd normals, rgb/normal/semantic heads), "
                            "appearance embedding 8, 7 classes, normal+semantic compositing, RefLoss+CE+distortion, 2^15 rays/GPU/step"),
}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
CPU_WORKLOAD = ("BASELINE.json configs[0]: lego-shaped 800x800x100 views, scale 0.5, 8192 rays/batch through the noCUDA path "
                "(rendering_noCUDA coarse 64 + fine 128 samples/ray, two networks_noCUDA fields), fw+bw+Adam on the host cores")


def cpu_arm(steps, warmup, rays_per_step=512):
    """The reference's CPU-shaped implementation of the path (BASELINE.json configs[0]: models/rendering_noCUDA.py +
    models/networks_noCUDA.py), restated in oracle/nocuda_pipeline.py because the files themselves import tinycudann / vren
    and call .cuda().  Each step is a BOUNDED SAMPLE of one 8192-ray batch (`rays_per_step` rays of it) so
    that the run ends within minutes; rays/s does not depend on the sample size (cost is linear in rays).  Imports nothing
    from the product package: no libngp_b200.so is mapped by this arm."""
    from oracle.nocuda_pipeline import NoCUDAPipeline, aabb_hits
    from synth_scenes import BoxScene
    sc = BoxScene("lego")
    pipe = NoCUDAPipeline(scale=0.5)
    poses = sc.poses(100)
    gen = torch.Generator().manual_seed(20220806)
    batches = []
    for _ in range(2):
        ro, rd = sc.sample_rays(rays_per_step * 2, poses, gen)
        _, hit = aabb_hits(ro, rd, 0.5)                      # the noCUDA sampler needs rays that hit the box (0/0 otherwise)
        ro, rd = ro[hit][:rays_per_step].contiguous(), rd[hit][:rays_per_step].contiguous()
        rgb, *_ = sc.shade(ro, rd)
        batches.append((ro, rd, rgb))
    n = batches[0][0].shape[0]
    for i in range(warmup):
        pipe.train_step(*batches[i % 2])
    t0 = time.perf_counter()
    samples = 0
    for i in range(steps):
        _, s = pipe.train_step(*batches[i % 2]); samples += s
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": "rays/s", "cores": pipe.threads, "kind": "port",
            "sample": f"{steps} steps x {n} rays (a {n}/8192 sample of one configs[0] batch per step), {samples // max(steps, 1) // n} samples/ray "
                      f"(coarse 64 + fine 128), fw+bw+Adam, torch CPU ops on {pipe.threads} threads; restatement of rendering_noCUDA.py:103-214 + "
                      "networks_noCUDA.py:49-369 (oracle/nocuda_pipeline.py)"}, dt / max(steps, 1)


# ------------------------------------------------------------------------------------------ kernel timing
def time_kernel(fn, iters=10):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def kernel_breakdown_ngp(model, xyzs, dirs):
    """Per-kernel times for the reference-literal field (tcnn-API kernels: fp32 (N, L*F) feature / gradient matrices)."""
    from ngp_b200 import tcnn
    S = xyzs.shape[0]
    aabb = model.aabb()
    xw = xyzs.contiguous()
    out = {}
    for tag, enc in (("xyz", model.xyz_encoder), ("rgb", model.rgb_encoder)):
        g, table = enc.grid, enc.params.detach()
        LF = g.n_levels * g.n_features
        dy = torch.randn(S, LF, device=xw.device)
        dtab = torch.zeros_like(table)
        out[f"hashgrid_fw_{tag}"] = (time_kernel(lambda: tcnn.grid_forward(xw, table, g, aabb), 3), S * (12 + 8 * LF * 4 + LF * 4), "B")
        out[f"hashgrid_bw_params_{tag}"] = (time_kernel(lambda: tcnn.grid_backward_params(xw, dy, g, out=dtab, aabb=aabb), 3), S * (12 + LF * 4 + 16 * LF * 4), "B")
        if tag == "xyz":
            out["hashgrid_bw_input_xyz"] = (time_kernel(lambda: tcnn.grid_backward_input(xw, dy, table, g, aabb), 3), S * (12 + 8 * LF * 4 + LF * 4 + 12), "B")
        del dy, dtab
    feat = tcnn.grid_forward(xw, model.rgb_encoder.params.detach(), model.rgb_encoder.grid, aabb)
    m2, p2 = model.rgb_net.mlp, model.rgb_net.params.detach()
    emb = torch.randn(S, model.rgb_net.n_input_dims - 16 - feat.shape[1], device=xw.device)
    segs = [(dirs, 16, 1), (feat, feat.shape[1], 0)] + ([(emb, emb.shape[1], 0)] if emb.shape[1] else [])
    k0 = model.rgb_net.n_input_dims
    flops = lambda m, k0: 2 * (k0 * m.width + (m.n_hidden - 1) * m.width ** 2 + m.width * m.n_out)      # useful flops: real output columns only
    rgb = tcnn.mlp_forward(segs, p2, m2)
    out["mlp_rgb_fw"] = (time_kernel(lambda: tcnn.mlp_forward(segs, p2, m2), 3), S * flops(m2, k0), "F")
    drgb = torch.randn_like(rgb)
    out["mlp_rgb_bw"] = (time_kernel(lambda: tcnn.mlp_backward(segs, p2, m2, drgb, [False, True] + [True] * (len(segs) - 2), saved_out=rgb), 3),
                         S * flops(m2, k0) * 3, "F")
    return out


def kernel_breakdown(model, xyzs, dirs):
    """Times the individual hot kernels of the step — the ones the fused density path launches — on the step's real
    sample set; returns {name: (seconds, algorithmic bytes or flops, unit)}."""
    from ngp_b200 import tcnn
    from ngp_b200._lib import lib, ptr, check, stream
    S = xyzs.shape[0]
    g = model.xyz_encoder.grid
    LF = g.n_levels * g.n_features
    k0p = (LF + 15) // 16 * 16
    aabb = model.aabb()
    xw = xyzs.contiguous()
    table = model.xyz_encoder.params.detach()
    tiles = tcnn.grid_forward_tiles(xw, table, g, aabb)
    dy_tiles = torch.randn((S + 127) // 128 * 128 * k0p, device=xw.device)
    dtab = torch.zeros_like(table)
    out = {}
    # fw 12 + 8*L*F*s_p + L*F*s_o (s_o = 2: bf16 operand tiles) ; bw(params) 12 + L*F*s_o + 2*8*L*F*s_g
    out["hashgrid_fw"] = (time_kernel(lambda: tcnn.grid_forward_tiles(xw, table, g, aabb)), S * (12 + 8 * LF * 4 + LF * 2), "B")
    scatter = lambda: check(lib.ngp_hashgrid_bw_params_tiles(ptr(xw), tcnn._aabb_arg(aabb), ptr(dy_tiles), *g.args(), S, ptr(dtab), stream()), "bw")
    out["hashgrid_bw_params"] = (time_kernel(scatter), S * (12 + LF * 4 + 16 * LF * 4), "B")
    m1, m2 = model.sigma_net.mlp, model.rgb_net.mlp
    p1, p2 = model.sigma_net.params.detach(), model.rgb_net.params.detach()
    h, _ = tcnn.mlp_forward([(tiles, LF, 2)], p1, m1, aux_exp=True, n=S)
    flops = lambda m, k0: 2 * (k0 * m.width + (m.n_hidden - 1) * m.width ** 2 + m.width * m.n_out)      # useful flops: real output columns only
    out["mlp_sigma_fw"] = (time_kernel(lambda: tcnn.mlp_forward([(tiles, LF, 2)], p1, m1, aux_exp=True, n=S)), S * flops(m1, LF), "F")
    dh = torch.randn_like(h); ds = torch.randn(S, device=xw.device)
    out["mlp_sigma_bw"] = (time_kernel(lambda: tcnn.mlp_backward([(tiles, LF, 2)], p1, m1, dh, [True], d_aux=ds, n=S, dseg_numel=dy_tiles.numel(), saved_out=h)),
                           S * flops(m1, LF) * 3, "F")
    segs = [(dirs, 16, 1), (h, 16, 0)]
    rgb = tcnn.mlp_forward(segs, p2, m2)
    out["mlp_rgb_fw"] = (time_kernel(lambda: tcnn.mlp_forward(segs, p2, m2)), S * flops(m2, 32), "F")
    drgb = torch.randn_like(rgb)
    out["mlp_rgb_bw"] = (time_kernel(lambda: tcnn.mlp_backward(segs, p2, m2, drgb, [False, True], saved_out=rgb)), S * flops(m2, 32) * 3, "F")
    return out


def render_bench(model, scene, poses, frames=9, wh=(1920, 1080), chunk=1 << 20, rank=0, world=1, esf=0.0, num_classes=0, extra=None):
    """Test-time rendering (BASELINE.json configs[4]): full frames through raymarching_test +
    composite_test_fw rounds, T_threshold 1e-2 (render.py:125); Mrays/s for both round schedules.
    With world > 1 every frame's rays are split into `world` contiguous tiles, one per rank, no collective on
    the data path; the frame time is the max over ranks."""
    import torch.distributed as dist
    from ngp_b200.rendering import render
    out = {}
    full = getattr(model, "has_normals", True)        # fields with normal / semantic heads: ngp_render_advance_full composites those streams too
    extra = extra or {}
    scheds = (("wavefront", "reference") if full else ("wavefront", "geometric", "reference")) if world == 1 else ("wavefront",)
    with torch.no_grad():
        for sched in scheds:
            def frame(i):
                W, H = wh
                n = W * H
                # this rank's share of the frame: image rows rank, rank + world, ... — whole rows keep neighbouring rays together (they
                # walk the same cells), interleaving balances the load (contiguous stripes give the ranks that see the object 4-5x the
                # samples of the ranks that see background, and the frame time is the max over ranks)
                rows = torch.arange(rank, H, world, device="cuda")
                px = (rows[:, None] * W + torch.arange(W, device="cuda")[None, :]).reshape(-1)
                a0, a1 = 0, px.numel()
                sc = scene.img_wh[0] / W
                u, v = (px % W).float() * sc + (sc - 1) / 2, (px // W).float() * sc + (sc - 1) / 2     # = BoxScene.image_rays on the tile
                ro, rd = scene.rays_from_pixels(poses[i % poses.shape[0]][None], torch.zeros(a1 - a0, dtype=torch.long, device="cuda"), u, v)
                tot = 0
                for a in range(0, ro.shape[0], chunk):
                    r = render(model, ro[a:a + chunk], rd[a:a + chunk], exp_step_factor=esf, num_classes=num_classes, test_time=True,
                               T_threshold=1e-2, sample_schedule=sched if sched != "wavefront" else "geometric",
                               renderer="wavefront" if sched == "wavefront" else "loop", **extra)
                    tot += int(r["total_samples"])
                return tot, n
            # warm-up: two frames for the product path (allocator pools, lazy kernel attributes), one for the comparison loops
            for w in range(2 if sched == scheds[0] else 1):
                frame(w); torch.cuda.synchronize()
            nf = frames if sched == scheds[0] and not full else 1
            per_frame = []
            for i in range(nf):
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                tot, nr = frame(i + 2)
                torch.cuda.synchronize()
                per_frame.append(time.perf_counter() - t0)
            per_frame.sort()
            dt = torch.tensor([per_frame[len(per_frame) // 2]], device="cuda", dtype=torch.float64)      # median frame (wall clock: a host hiccup must not decide)
            tt = torch.tensor([float(tot)], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX); dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            dt, tot = float(dt), float(tt)
            out[sched] = {"Mrays_per_s": nr / dt / 1e6, "ms_per_frame": dt * 1e3, "samples_per_ray": tot / nr, "frames": nf,
                          "ms_min": per_frame[0] * 1e3, "ms_max": per_frame[-1] * 1e3}
    return {"metric": "render Mrays/s", "frame": f"{wh[0]}x{wh[1]}", "value": out[scheds[0]]["Mrays_per_s"], "unit": "Mrays/s",
            "n_gpus": world, "sharding": "interleaved image rows per rank (row r -> rank r % N), no collective" if world > 1 else "single GPU",
            "legend": "wavefront = fused advance kernel per round; geometric / reference = reference-style loop over "
                      "raymarching_test + composite_test_fw with 4,8,16.. / the reference's own round sizes",
            "timing": "wall clock per frame incl. the per-round host read-backs, rays generated on device; median over `frames` frames after warm-up, max over ranks", **{k: v for k, v in out.items()}}


# ------------------------------------------------------------------------------------------ GPU arm
def gpu_arm(args):
    import torch.distributed as dist
    from ngp_b200 import _lib, vren
    from ngp_b200.networks import NGP, NGPCompact
    from synth_scenes import BoxScene, scene_density_grid
    from ngp_b200.trainer import Trainer, psnr
    from ngp_b200.rendering import render

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    for attempt in range(6):            # a freshly leased box can refuse the very first driver init for a few seconds
        try:
            torch.cuda.init()
            break
        except RuntimeError:
            if attempt == 5:
                raise
            time.sleep(3)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(20220806); np.random.seed(20220806)          # train.py:402-404

    wl = WORKLOADS[args.workload]
    scene = BoxScene(wl["scene"], device=dev)
    poses = scene.poses(wl["views"])
    full = wl.get("field") == "ngp"                                  # the reference-literal field with normal / semantic heads
    if full:
        model = NGP(scale=wl["scale"], embed_a=True, embed_a_len=wl["embed_a_len"], classes=wl["classes"]).to(dev)
        emb = torch.nn.Embedding(wl["views"], wl["embed_a_len"]).to(dev)              # train.py:117-119
        rkw = dict(exp_step_factor=wl["esf"], num_classes=wl["classes"], normal_ref=True, semantic=True)
    else:
        model = NGPCompact(scale=wl["scale"], log2_T=wl["log2_T"]).to(dev)
        emb = None
        rkw = dict(exp_step_factor=wl["esf"], num_classes=0)
    model.density_grid.copy_(scene_density_grid(scene))             # converged-occupancy proxy; maintained by update_density_grid afterwards
    vren.packbits(model.density_grid, 0.5, model.density_bitfield)
    tr = Trainer(model, lr=wl["lr"], render_kwargs=rkw, world_size=world, max_grad_norm=50.0 if full else None,
                 extra_params=emb.parameters() if full else ())

    R = wl.get("rays", R_PER_GPU)
    if args.scaling == "strong":      # total rays per step fixed at the 1-GPU batch, sharded over the ranks (configs[2]: 2^18 rays over 8 GPUs)
        R = R_PER_GPU // world
    n_batches = 8
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)       # each rank draws its own shard of the global batch
    W_, H_ = scene.img_wh
    from ngp_b200 import ray_utils
    K_ = [[scene.focal, 0.0, W_ / 2], [0.0, scene.focal, H_ / 2], [0.0, 0.0, 1.0]]
    directions = ray_utils.get_ray_directions(H_, W_, K_, device=dev)       # train.py:103 self.directions (device buffer)
    pool, host = [], []
    for _ in range(n_batches):
        img = torch.randint(poses.shape[0], (R,), device=dev, generator=gen)
        pix = torch.randint(W_ * H_, (R,), device=dev, generator=gen)
        ro, rd = ray_utils.get_rays_indexed(directions, poses, img, pix)
        c, _, _, lab = scene.shade(ro, rd)
        pool.append((ro, rd, c, img, lab))                           # resident batch: rays already generated
        # what the reference's loader hands a step (datasets/*: img_idxs, pix_idxs, rgb [, label]) in pinned host memory
        host.append(tuple(t.cpu().pin_memory() for t in ((img, pix, c, lab) if full else (img, pix, c))))
    pool_o, pool_d = [b[0] for b in pool], [b[1] for b in pool]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    sample_log = []

    def step(o, d, c, img=None, lab=None, host_loss=False):
        if full:
            return tr.train_step(o, d, c, target={"label": lab}, embedding_a=emb(img), host_loss=host_loss)
        return tr.train_step(o, d, c, host_loss=host_loss)

    def step_resident(i):
        out = step(*(pool[i % n_batches] if full else pool[i % n_batches][:3]))
        sample_log.append(tr.last_samples)
        return out

    # end to end: every step's (img_idxs, pix_idxs, rgb[, label]) travel host -> device inside the timed region, on a copy
    # stream one step ahead of the compute stream (double buffering); rays are generated on the device from the indices
    # (train.py:136-156), and the step's loss is read back to the host.
    copy_stream = torch.cuda.Stream(device=dev)
    dev_bufs = [tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host[0]) for _ in range(2)]   # double buffer, allocated once
    free_ev = [None, None]               # recorded on the compute stream when the step that read buffer b is fully enqueued
    pending, seq = [], [0]

    def prefetch():
        b = seq[0] % 2
        with torch.cuda.stream(copy_stream):
            if free_ev[b] is not None:
                copy_stream.wait_event(free_ev[b])
            for dst, src in zip(dev_bufs[b], host[seq[0] % n_batches]):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_stream)
        pending.append((b, ev))
        seq[0] += 1

    def step_e2e(i):
        if not pending:
            prefetch()
        b, ev = pending.pop(0)
        main = torch.cuda.current_stream()
        main.wait_event(ev)
        prefetch()                                                   # next step's inputs fly while this step computes
        bufs = dev_bufs[b]
        img, pix, c = bufs[:3]
        o, d = ray_utils.get_rays_indexed(directions, poses, img, pix)
        hl = os.environ.get("NGP_BENCH_BLOCKING_LOSS", "0") != "1"     # 1: float(loss) after the step (drains the GPU every step)
        loss, _ = step(o, d, c, img, bufs[3], host_loss=hl) if full else step(o, d, c, host_loss=hl)
        free_ev[b] = torch.cuda.Event(); free_ev[b].record(main)
        return float(loss)                                                  # python float: this step's loss, read back on the host (Trainer.train_step)

    def eval_psnr():
        """PSNR on 2^15 held-out rays (generator seed 4321: the same rays baseline/ref_train.py evaluates the reference on)."""
        with torch.no_grad():
            g2 = torch.Generator(device=dev).manual_seed(4321)
            ro, rd = scene.sample_rays(1 << 15, poses, g2)
            gt, *_ = scene.shade(ro, rd)
            kw_e = {"embedding_a": emb(torch.zeros(1 << 15, dtype=torch.long, device=dev))} if full else {}
            return float(psnr(render(model, ro, rd, **rkw, **kw_e)["rgb"], gt))

    # pre-train so that the occupancy grid / sample count are at their steady state
    psnr_at_ref_steps = None
    for i in range(args.pretrain):
        step_resident(i)
        if i + 1 == args.ref_steps:
            psnr_at_ref_steps = eval_psnr()
    for i in range(args.warmup):
        step_resident(i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        # nvidia-smi's start-up (NVML init, device enumeration) stalls kernel launches for ~100 ms on a fresh box:
        # keep stepping (untimed) until its first samples have arrived, THEN time; it keeps sampling through the region
        t_wait, k = time.time(), 0
        while True:
            fn(k); k += 1
            torch.cuda.synchronize()
            ready = torch.tensor([1 if (rank != 0 or len(sampler.lines) >= 2 or time.time() - t_wait > 8.0) else 0], device=dev)
            if world > 1:
                dist.all_reduce(ready, op=dist.ReduceOp.MIN)
            if int(ready):
                break
        barrier()
        calls0 = _lib.lib_calls() if hasattr(_lib, "lib_calls") else 0
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn(i)
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        calls = (_lib.lib_calls() if hasattr(_lib, "lib_calls") else 0) - calls0
        return float(ms) * 1e-3, (sampler.stop() if rank == 0 else None), calls

    del sample_log[:]
    t_res, clocks, launches = timed(step_resident, args.steps)
    spr_timed = [float(x) / R for x in sample_log]
    t_e2e, _, _ = timed(step_e2e, args.steps)

    # steady-state quality + sample statistics (outside the timed regions)
    with torch.no_grad():
        img = torch.randint(poses.shape[0], (1 << 15,), device=dev, generator=gen)
        ro, rd = ray_utils.get_rays_indexed(directions, poses, img, torch.randint(W_ * H_, (1 << 15,), device=dev, generator=gen))
        gt, *_ = scene.shade(ro, rd)
        out = render(model, ro, rd, **rkw, **({"embedding_a": emb(img)} if full else {}))
        q = float(psnr(out["rgb"], gt))
        spr = float(out["total_samples"]) / ro.shape[0]
        rays_a, xyzs, dirs = out["rays_a"], out["xyzs"], None
    rend = None
    if not args.no_render and full:      # reference-literal field: chunks of 131 072 rays as the reference renders (render.py:33-48), test embedding 0 (train.py:150-151)
        with torch.no_grad():
            e0 = emb(torch.zeros(1, dtype=torch.long, device=dev))
        rend = render_bench(model, scene, poses, frames=1, chunk=1 << 17, rank=rank, world=world, esf=wl["esf"], num_classes=wl["classes"],
                            extra={"embedding_a": e0})
    if not args.no_render and not full:
        rend = render_bench(model, scene, poses, rank=rank, world=world, esf=wl["esf"])
        if args.render_4k:
            rend["4k"] = render_bench(model, scene, poses, frames=5, wh=(3840, 2160), rank=rank, world=world, esf=wl["esf"])
    dp_exchange = None
    if world > 1:     # what the gradient exchange costs: the same step with the exchange switched off (replicas diverge: timing only, last thing done)
        tr._exchange = "none"
        t_none, _, _ = timed(step_resident, min(args.steps, 20))
        per, per_none = t_res / args.steps * 1e3, t_none / min(args.steps, 20) * 1e3
        dp_exchange = {"ms_per_step_with_exchange": per, "ms_per_step_without_exchange": per_none, "exposed_ms": per - per_none,
                       "frac_of_step": (per - per_none) / per, "mode": "per-tensor NCCL all-reduce after backward (see profiles/r02b_dp_exchange_probe.txt)"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # per-kernel roofline on this step's real samples (rank 0, after the timed regions)
    from ngp_b200.custom_functions import RayMarcher
    from ngp_b200.rendering import MAX_SAMPLES
    with torch.no_grad():
        _, hits_t, _ = vren.ray_aabb_intersect(pool_o[0], pool_d[0], model.center, model.half_size, 1)
        ra, xyzs, dirs, deltas, ts, tot = RayMarcher.apply(pool_o[0], pool_d[0], hits_t[:, 0].contiguous(), model.density_bitfield,
                                                           model.cascades, model.scale, wl["esf"], model.grid_size, MAX_SAMPLES)
        kb = (kernel_breakdown_ngp if full else kernel_breakdown)(model, xyzs, dirs)
    pk, pk_src = peaks()
    kern = {}
    for k, (sec, work, unit) in kb.items():
        if unit == "B":
            kern[k] = {"ms": sec * 1e3, "achieved_GBs": work / sec / 1e9, "frac_hbm": work / sec / 1e9 / pk["hbm_gbs"]}
        else:
            kern[k] = {"ms": sec * 1e3, "achieved_TFLOPs": work / sec / 1e12, "frac_tensor": work / sec / 1e12 / pk["bf16_tflops_sustained"]}
    dom = max(kb, key=lambda k: kb[k][0])
    sec, work, unit = kb[dom]
    S_l = int(xyzs.shape[0])
    if unit == "F":
        roof = {"kernel": dom, "bound": "tensor", "achieved": work / sec / 1e12, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "peak_source": pk_src, "traffic": None, "samples_per_launch": S_l}
        roof["frac"] = roof["achieved"] / roof["peak"]
    else:
        # HBM-bound class.  `achieved` is the kernel's MEASURED DRAM traffic per launch (dram__bytes_read+write of the committed
        # ncu --set full capture of this kernel, per sample, x this launch's samples) over its live CUDA-event time: a physical
        # fraction of the HBM peak.  The ALGORITHMIC bytes are reported beside it: for a table that lives in the
        # 126 MB L2 they exceed what HBM could deliver, which is the point of keeping the table resident — the kernel is then
        # bound by the L2's reduction request rate, whose peak is re-measured in this run (ngp_probe_l2_reduction).
        roof = {"kernel": dom, "bound": "hbm", "achieved": None, "peak": pk["hbm_gbs"], "unit": "GB/s", "peak_source": pk_src,
                "frac": None, "traffic": None, "samples_per_launch": S_l,
                "algorithmic": {"bytes_per_launch": work, "GBs": work / sec / 1e9, "over_hbm_peak": work / sec / 1e9 / pk["hbm_gbs"],
                                "formula": "SURVEY 8d: scatter 12 + L*F*4 + 2*8*L*F*4, gather 12 + 8*L*F*4 + L*F*s_out bytes per sample"}}
        try:
            import glob
            cands = [json.load(open(f)) for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))]
            cands = [t for t in cands if t["kernel"] == dom and t.get("workload", "lego") == args.workload]
            tr_ = cands[-1]                                                    # newest capture of this kernel on this workload
            roof["traffic"] = tr_["dram_bytes_per_sample"] * S_l
            roof["traffic_source"] = tr_["source"]
            roof["achieved"] = roof["traffic"] / sec / 1e9
            roof["frac"] = roof["achieved"] / roof["peak"]
            if "red_sectors_per_sample" in tr_:
                from ngp_b200._lib import lib as _l, ptr as _p, stream as _s
                table_b = (model.rgb_encoder if (full and dom.endswith("_rgb")) else model.xyz_encoder).params.numel() * 4
                buf = torch.zeros(table_b // 4, device=dev)
                probe = lambda: _l.ngp_probe_l2_reduction(_p(buf), table_b, 64, _s())
                n_req = probe(); torch.cuda.synchronize()
                peak_req = n_req / time_kernel(probe, 3) / 1e9
                rate = tr_["red_sectors_per_sample"] * S_l / sec / 1e9
                roof["limiter"] = {"resource": "L2 reduction sector requests (one per 32-byte sector per instruction)", "achieved_G_per_s": rate,
                                   "peak_G_per_s": peak_req, "frac": rate / peak_req, "sectors_per_sample": tr_["red_sectors_per_sample"],
                                   "peak_source": f"ngp_probe_l2_reduction timed in this run on a zeroed {table_b >> 20} MB buffer (the table's size)"}
                del buf
        except Exception as ex:
            roof["traffic_note"] = "no committed ncu capture for this kernel / workload: " + repr(ex)[:120]
    cpu = None if (args.no_cpu or world > 1) else cpu_arm(steps=3, warmup=1)[0]       # the CPU leg is timed at N=1 only (rank 0)
    value = world * R * args.steps / t_res
    # reference GPU path on the same box, same run: the reference's unmodified glue + its own CUDA kernels + a
    # torch-op tinycudann stand-in, same scene / batch recipe / lr / Adam eps, trained args.ref_steps steps, last 10 timed
    ref_gpu = None
    if args.workload == "lego" and world == 1 and args.ref_steps > 0:
        del tr, pool, pool_o, pool_d
        torch.cuda.empty_cache()
        try:
            from baseline import ref_train
            ref_gpu = ref_train.run_reference_gpu(scene_kind="lego", field="ngp_pl", rays=R, steps_total=args.ref_steps, timed_last=10,
                                                  lr=wl["lr"], eps=1e-15, views=wl["views"], log2_T=wl["log2_T"])
            ref_gpu["ours_psnr_after_same_steps"] = psnr_at_ref_steps
            ref_gpu["ours_over_reference_gpu"] = value / ref_gpu["value"]
            torch.cuda.empty_cache()
            # the same reference glue with BOTH extension imports bound to libngp_b200.so (zero-edit drop-in)
            drop = ref_train.run_reference_gpu(scene_kind="lego", field="ngp_pl", rays=R, steps_total=args.ref_steps, timed_last=10,
                                               lr=wl["lr"], eps=1e-15, views=wl["views"], log2_T=wl["log2_T"], vren="ours", tcnn="ours")
            ref_gpu["reference_glue_on_libngp_b200"] = {k: drop[k] for k in ("value", "unit", "ms_per_step", "ms_per_step_median", "ms_per_step_max", "psnr_after_steps", "samples_per_ray")}
        except Exception as ex:            # baseline/_ref or vren_ref.so not present on this box
            ref_gpu = {"unavailable": repr(ex)[:300]} if ref_gpu is None else ref_gpu
    others = None
    if args.workload == "lego" and world == 1 and args.other_configs:
        # BASELINE.json configs[3] / configs[2] shapes, short runs of this same script (their own processes: fresh allocator, own
        # model), so that the driver-run line carries them too; each is a full line of its own under --workload
        others = {}
        for wname in ("street", "playground"):
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", wname, "--steps", "8", "--warmup", "3", "--pretrain", "40",
                                      "--ref-steps", "0", "--no-cpu", "--no-render"], capture_output=True, text=True, timeout=600)
                sub = json.loads(out.stdout.strip().splitlines()[-1])
                others[wname] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "gpu_launches", "roofline")}
                others[wname].update({"workload": sub["config"]["workload"], "samples_per_ray": sub["config"]["samples_per_ray"],
                                      "rays_per_gpu": sub["config"]["rays_per_gpu"], "psnr_after_pretrain": sub["config"]["psnr_after_pretrain"],
                                      "pretrain_steps": sub["config"]["pretrain_steps"], "e2e": sub["e2e"]})
            except Exception as ex:
                others[wname] = {"error": repr(ex)[:300]}
    line = {
        "metric": "train rays/s (fw+bw)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "fp32 (bf16 tensor-core operands, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": wl["name"], "rays_per_gpu": R, "global_batch_rays": world * R, "samples_per_ray": spr, "samples_per_ray_timed_steps": {"min": min(spr_timed), "max": max(spr_timed), "mean": sum(spr_timed) / len(spr_timed)},
                   "pretrain_steps": args.pretrain, "psnr_after_pretrain": q, "l2": "inputs_exceed_l2 (>250 MB of samples per step)",
                   "parallelism": f"ray-sharded dp{world} ({args.scaling} scaling), NCCL all-reduce of table+MLP gradients after backward" if world > 1 else "single GPU",
                   "occupancy": "analytic voxelisation at step 0, then update_density_grid every 16 steps (inside the timed region)"},
        "e2e": {"value": world * R * args.steps / t_e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": t_e2e / args.steps * 1e3,
                "path": "pinned host (img_idxs i64, pix_idxs i64, rgb f32[, label]) -> H2D on a copy stream one step ahead -> ngp_get_rays -> "
                        "Trainer.train_step(host_loss=True): loss -> pinned host right after the forward pass, waited for after the step is enqueued; every step"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernels": kern, "cpu_baseline": cpu, "reference_gpu": ref_gpu, "render": rend, "configs": others, "dp_exchange": dp_exchange,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--pretrain", type=int, default=400)
    ap.add_argument("--ref-steps", type=int, default=100, help="steps the reference-GPU arm trains (PSNR is compared after the same number of steps); 0 = skip")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lego", choices=sorted(WORKLOADS), help="lego = BASELINE.json configs[1] (the headline); street = configs[3] shape")
    ap.add_argument("--no-render", action="store_true", help="skip the test-time render sweep")
    ap.add_argument("--no-render-4k", dest="render_4k", action="store_false", help="skip the 3840x2160 frames of the render sweep (BASELINE.json configs[4])")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: 2^18 rays per GPU (default); strong: 2^18 rays per step in total, sharded over the ranks")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-other-configs", dest="other_configs", action="store_false",
                    help="skip the short street (configs[3]) / playground (configs[2]) runs appended to the headline line as `configs`")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) != 0:
            return
        cpu, per_step = cpu_arm(steps=args.steps, warmup=args.warmup)
        print(json.dumps({
            "impl": "reference", "metric": "train rays/s (fw+bw)", "value": cpu["value"], "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": {"workload": CPU_WORKLOAD, "sample": cpu["sample"]},
            "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    gpu_arm(args)


if __name__ == "__main__":
    main()
