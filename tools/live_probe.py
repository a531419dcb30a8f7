"""Fraction of marched samples that are still live (before the compositor's T < 1e-4 termination) after training."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "instant-ngp-pp_b200")]
import torch
from ngp_b200 import vren
from ngp_b200.networks import NGPCompact
from ngp_b200.synthetic import BoxScene, scene_density_grid
from ngp_b200.trainer import Trainer
dev = torch.device("cuda", 0)
scene = BoxScene("lego", device=dev); poses = scene.poses(100)
model = NGPCompact(scale=0.5).to(dev)
model.density_grid.copy_(scene_density_grid(scene)); vren.packbits(model.density_grid, 0.5, model.density_bitfield)
tr = Trainer(model, render_kwargs=dict(exp_step_factor=0.0, num_classes=0))
gen = torch.Generator(device=dev).manual_seed(1)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 400):
    ro, rd = scene.sample_rays(1 << 18, poses, gen)
    rgb, *_ = scene.shade(ro, rd)
    loss, res = tr.train_step(ro, rd, rgb)
    if i % 100 == 99 or i < 3:
        S = int(res["total_samples"]); live = int(res["vr_samples"].sum())
        ws = res["ws"]
        print(f"step {i}: samples {S} ({S / (1 << 18):.1f}/ray)  live {live} ({100 * live / S:.1f} %)  ws>0: {100 * float((ws > 0).float().mean()):.1f} %  ws>1e-4: {100 * float((ws > 1e-4).float().mean()):.1f} %", flush=True)
