"""REFERENCE-ARM INFRASTRUCTURE — the reference's optimiser step without Lightning (pytorch_lightning is not installed):
a restatement of train.py:128-132 (density_grid / grid_coords buffers), :244 (Adam), :268-310 (training_step: occupancy
update every 16 steps with a 256-step warm-up, `render`, `NeRFLoss`, `sum(lo.mean())`) and :435 (global-norm clip)
that calls the reference's OWN unmodified glue — models/rendering.py:render, losses.py:NeRFLoss,
models/networks.py:NGP (incl. its update_density_grid), models/custom_functions.py — loaded by baseline/ref_harness.py.

Two field shapes:
  'ngp'    the reference's literal model, models/networks.py:13-163 (2 hash grids L16 F8, density net, 3 heads);
  'ngp_pl' the ngp_pl-shaped model BASELINE.json's headline config names (16-level F2 hash grid T=2^19, 64-wide MLPs):
           upstream kwea123/ngp_pl's field (xyz_encoder = NetworkWithInputEncoding(HashGrid -> 64 -> 16), sigma =
           TruncExp(h0), rgb_net(SH4(d) | h) 32 -> 64 -> 64 -> 3), written as a subclass of the reference's NGP so that the
           occupancy-grid code, `render` and `NeRFLoss` are the reference's own.  The reference's renderer expects five
           field outputs; the normal / semantic ones are zeros with num_classes = 0.
"""
import time

import numpy as np
import torch
from torch import nn

from . import ref_harness

MAX_SAMPLES = 1024


def ngp_pl_class(glue):
    base, tcnn, TruncExp = glue.networks.NGP, glue.tcnn, glue.custom_functions.TruncExp

    class NGPpl(base):
        def __init__(self, scale, rgb_act="Sigmoid", log2_T=19, L=16, F=2, N_min=16, width=64):
            nn.Module.__init__(self)
            self.rgb_act, self.scale, self.use_skybox, self.embed_a = rgb_act, scale, False, False
            self.register_buffer("center", torch.zeros(1, 3))
            self.register_buffer("xyz_min", -torch.ones(1, 3) * scale)
            self.register_buffer("xyz_max", torch.ones(1, 3) * scale)
            self.register_buffer("half_size", (self.xyz_max - self.xyz_min) / 2)
            self.cascades = max(1 + int(np.ceil(np.log2(2 * scale))), 1)
            self.grid_size = 128
            self.register_buffer("density_bitfield", torch.zeros(self.cascades * self.grid_size ** 3 // 8, dtype=torch.uint8))
            b = float(np.exp(np.log(2048 * scale / N_min) / (L - 1)))
            self.xyz_encoder = tcnn.NetworkWithInputEncoding(
                n_input_dims=3, n_output_dims=16,
                encoding_config={"otype": "HashGrid", "n_levels": L, "n_features_per_level": F, "log2_hashmap_size": log2_T,
                                 "base_resolution": N_min, "per_level_scale": b},
                network_config={"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None",
                                "n_neurons": width, "n_hidden_layers": 1})
            self.dir_encoder = tcnn.Encoding(n_input_dims=3, encoding_config={"otype": "SphericalHarmonics", "degree": 4})
            self.rgb_net = tcnn.Network(n_input_dims=32, n_output_dims=3,
                                        network_config={"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": rgb_act,
                                                        "n_neurons": width, "n_hidden_layers": 2})

        def density(self, x, return_feat=False, **_):
            x = (x - self.xyz_min) / (self.xyz_max - self.xyz_min)
            h = self.xyz_encoder(x)
            sigmas = TruncExp.apply(h[:, 0])
            return (sigmas, h) if return_feat else sigmas

        def forward(self, x, d, **kwargs):
            sigmas, h = self.density(x, return_feat=True)
            d = d / torch.norm(d, dim=1, keepdim=True)
            rgbs = self.rgb_net(torch.cat([self.dir_encoder((d + 1) / 2), h], 1))
            z3 = torch.zeros(x.shape[0], 3, device=x.device)
            return sigmas, rgbs, z3, z3, torch.zeros(x.shape[0], kwargs.get("num_classes", 0), device=x.device)

        def forward_test(self, x, d, **kwargs):
            return self.forward(x, d, **kwargs)

    return NGPpl


def make_model(glue, field="ngp_pl", scale=0.5, device="cuda", **kw):
    """Field + the two buffers train.py:128-132 registers on it."""
    model = (ngp_pl_class(glue)(scale, **kw) if field == "ngp_pl" else glue.networks.NGP(scale=scale, **kw)).to(device)
    G = model.grid_size
    model.register_buffer("density_grid", torch.zeros(model.cascades, G ** 3, device=device))
    ax = torch.arange(G, dtype=torch.int32, device=device)
    model.register_buffer("grid_coords", torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3))
    return model


class RefTrainer:
    """training_step of train.py:268-310 + configure_optimizers :244 + gradient_clip_val :435."""

    def __init__(self, glue, model, lr=1e-2, eps=1e-8, render_kwargs=None, loss_kwargs=None, max_grad_norm=None,
                 extra_params=(), update_interval=16, warmup_steps=256, density_threshold=0.01):
        self.glue, self.model = glue, model
        self.loss = glue.losses.NeRFLoss()
        self.extra = [p for p in extra_params]
        self.params = [p for p in model.parameters() if p.numel() > 0] + self.extra
        self.opt = torch.optim.Adam(self.params, lr, eps=eps)
        self.render_kwargs, self.loss_kwargs = dict(render_kwargs or {}), dict(loss_kwargs or {})
        self.max_grad_norm = max_grad_norm
        self.update_interval, self.warmup_steps, self.density_threshold = update_interval, warmup_steps, density_threshold
        self.step = 0
        self.last_samples = None

    def train_step(self, rays_o, rays_d, batch, update_grid=True, **step_kwargs):
        m = self.model
        if update_grid and self.step % self.update_interval == 0:
            m.update_density_grid(self.density_threshold * MAX_SAMPLES / 3 ** 0.5, warmup=self.step < self.warmup_steps, erode=False)
        results = self.glue.rendering.render(m, rays_o, rays_d, **{**self.render_kwargs, **step_kwargs})
        loss_d = self.loss(results, batch, **self.loss_kwargs)
        loss = sum(lo.mean() for lo in loss_d.values())
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.max_grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.max_grad_norm)
        self.opt.step()
        self.step += 1
        self.last_samples = results["total_samples"]
        return loss.detach(), results


def run_reference_gpu(scene_kind="lego", field="ngp_pl", rays=1 << 18, steps_total=100, timed_last=10, lr=1e-2, eps=1e-15,
                      tf32=True, seed=20220806, views=100, log2_T=19, vren="ref", tcnn="standin", eval_rays=1 << 15, device="cuda"):
    """Trains `steps_total` steps of the reference pipeline on the procedural scene (same scene, camera set, ray-batch
    recipe, initial occupancy, learning rate and Adam eps as bench.py's own arm), times the last `timed_last` of them with
    CUDA events and evaluates PSNR on held-out rays.  -> dict for bench.py's `reference_gpu` object."""
    from synth_scenes import BoxScene, scene_density_grid
    glue = ref_harness.load(vren=vren, tcnn=tcnn)
    dev = torch.device(device)
    prev_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    try:
        torch.manual_seed(seed); np.random.seed(seed)
        scene = BoxScene(scene_kind, device=dev)
        poses = scene.poses(views)
        model = make_model(glue, field, scale=scene.scale, device=dev, **({"log2_T": log2_T} if field == "ngp_pl" else {}))
        model.density_grid.copy_(scene_density_grid(scene))
        glue.vren.packbits(model.density_grid, 0.5, model.density_bitfield)
        rkw = dict(exp_step_factor=scene.exp_step_factor, num_classes=0 if field == "ngp_pl" else 7)
        tr = RefTrainer(glue, model, lr=lr, eps=eps, render_kwargs=rkw)
        gen = torch.Generator(device=dev).manual_seed(1234)
        pool = []
        for _ in range(8):
            ro, rd = scene.sample_rays(rays, poses, gen)
            c, *_ = scene.shade(ro, rd)
            pool.append((ro, rd, {"rgb": c}))
        n_un = steps_total - timed_last
        for i in range(n_un):
            tr.train_step(*pool[i % 8])
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(timed_last + 1)]     # per-step times: the median is reported beside the mean
        t0 = time.perf_counter()
        s.record()
        samples = 0
        marks[0].record()
        for i in range(n_un, steps_total):
            _, res = tr.train_step(*pool[i % 8])
            samples += int(res["total_samples"])
            marks[i - n_un + 1].record()
        e.record(); torch.cuda.synchronize()
        sec = s.elapsed_time(e) * 1e-3
        wall = time.perf_counter() - t0
        per_step = sorted(marks[k].elapsed_time(marks[k + 1]) for k in range(timed_last))
        median_ms = per_step[timed_last // 2] if timed_last % 2 else 0.5 * (per_step[timed_last // 2 - 1] + per_step[timed_last // 2])
        with torch.no_grad():          # held-out rays: the same set bench.py evaluates its own arm on (seed 4321)
            ro, rd = scene.sample_rays(eval_rays, poses, torch.Generator(device=dev).manual_seed(4321))
            gt, *_ = scene.shade(ro, rd)
            out = glue.rendering.render(model, ro, rd, **rkw)
            psnr = float(glue.metrics.psnr(out["rgb"], gt))
        return {"value": rays * timed_last / sec, "unit": "rays/s", "ms_per_step": sec / timed_last * 1e3, "wall_ms_per_step": wall / timed_last * 1e3,
                "ms_per_step_median": median_ms, "ms_per_step_max": per_step[-1],
                "rays_per_step": rays, "steps_trained": steps_total, "steps_timed": timed_last, "psnr_after_steps": psnr,
                "samples_per_ray": samples / timed_last / rays, "field": field, "vren": vren, "tinycudann": tcnn, "tf32_mlp": bool(tf32),
                "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30,
                "what": "the reference's unmodified models/rendering.py:render + losses.py:NeRFLoss + models/networks.py occupancy update + "
                        "models/custom_functions.py over the reference's own CUDA kernels (vren_ref, compiled in place from models/csrc) "
                        "and a plain-torch-op tinycudann stand-in (tiny-cuda-nn itself is absent offline) — NOT a tcnn measurement"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev_tf32
