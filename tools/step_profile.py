"""Kernel-level breakdown of one training step (torch.profiler / CUPTI): shares only, never a bench value."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "instant-ngp-pp_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from ngp_b200 import vren
from ngp_b200.networks import NGPCompact
from synth_scenes import BoxScene, scene_density_grid
from ngp_b200.trainer import Trainer

dev = torch.device("cuda", 0)
scene = BoxScene("lego", device=dev)
poses = scene.poses(100)
model = NGPCompact(scale=0.5).to(dev)
model.density_grid.copy_(scene_density_grid(scene))
vren.packbits(model.density_grid, 0.5, model.density_bitfield)
tr = Trainer(model, render_kwargs=dict(exp_step_factor=0.0, num_classes=0))
R = 1 << 18
ro, rd = scene.sample_rays(R, poses)
rgb, *_ = scene.shade(ro, rd)
pre = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for i in range(pre):
    tr.train_step(ro, rd, rgb)
torch.cuda.synchronize()
tr.step = 1                     # keep the occupancy update out of the 3 profiled steps
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        tr.train_step(ro, rd, rgb)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
