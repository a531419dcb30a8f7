"""Kernel-level breakdown of one training step of the reference-literal field (bench --workload playground):
torch.profiler / CUPTI shares only, never a bench value."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "instant-ngp-pp_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from ngp_b200 import vren
from ngp_b200.networks import NGP
from synth_scenes import BoxScene, scene_density_grid
from ngp_b200.trainer import Trainer

dev = torch.device("cuda", 0)
scene = BoxScene("street", device=dev)
poses = scene.poses(128)
model = NGP(scale=8.0, embed_a=True, embed_a_len=8, classes=7).to(dev)
emb = torch.nn.Embedding(128, 8).to(dev)
model.density_grid.copy_(scene_density_grid(scene))
vren.packbits(model.density_grid, 0.5, model.density_bitfield)
kw = dict(exp_step_factor=1.0 / 256, num_classes=7, normal_ref=True, semantic=True)
tr = Trainer(model, lr=2e-3, render_kwargs=kw, max_grad_norm=50.0, extra_params=emb.parameters())
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 15
img = torch.randint(128, (R,), device=dev)
u = torch.randint(scene.img_wh[0], (R,), device=dev).float(); v = torch.randint(scene.img_wh[1], (R,), device=dev).float()
ro, rd = scene.rays_from_pixels(poses, img, u, v)
rgb, _, _, lab = scene.shade(ro, rd)
tr.step = 1                     # keep the occupancy update out
for i in range(3):
    tr.train_step(ro, rd, rgb, target={"label": lab}, embedding_a=emb(img))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(2):
        tr.train_step(ro, rd, rgb, target={"label": lab}, embedding_a=emb(img))
    torch.cuda.synchronize()
print("samples", int(tr.last_samples))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
