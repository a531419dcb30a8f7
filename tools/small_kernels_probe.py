"""Which small kernels does one headline training step launch?  (each dependent launch costs the stream ~8 us whatever its work)
usage: python tools/small_kernels_probe.py"""
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
model.density_grid.copy_(scene_density_grid(scene)); vren.packbits(model.density_grid, 0.5, model.density_bitfield)
tr = Trainer(model, render_kwargs=dict(exp_step_factor=0.0, num_classes=0))
ro, rd = scene.sample_rays(1 << 18, poses)
rgb, *_ = scene.shade(ro, rd)
for i in range(40):
    tr.train_step(ro, rd, rgb)
torch.cuda.synchronize()
tr.step = 1
N = 4
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=False) as prof:
    for i in range(N):
        tr.train_step(ro, rd, rgb, host_loss=True)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_type.name == "CUDA" or e.self_device_time_total > 0]
rows = sorted(((e.key, e.count / N, e.self_device_time_total / N) for e in prof.key_averages() if e.self_device_time_total > 0 and not e.key.startswith(("autograd", "_", "Optimizer")) and "::" in e.key or e.key.startswith("void") or e.key.startswith("ngp::")), key=lambda r: -r[1])
tot_small = 0
print(f"{'kernel':100s} launches/step   us/step")
for k, c, t in rows:
    if k.startswith("aten::") or k.startswith("cuda"):
        continue
    print(f"{k[:100]:100s} {c:8.1f} {t:10.1f}")
