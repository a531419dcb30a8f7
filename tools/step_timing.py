import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "instant-ngp-pp_b200"))
import torch
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
pool = []
for _ in range(4):
    ro, rd = scene.sample_rays(R, poses); rgb, *_ = scene.shade(ro, rd); pool.append((ro, rd, rgb))
for i in range(60):
    tr.train_step(*pool[i % 4])
torch.cuda.synchronize()

def timeit(fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

print("step, same rays, no grid update   : %.2f ms" % timeit(lambda i: tr.train_step(*pool[0], update_grid=False)))
print("step, rotating rays, no grid update: %.2f ms" % timeit(lambda i: tr.train_step(*pool[i % 4], update_grid=False), 8))
print("update_density_grid(warmup=True)   : %.2f ms" % timeit(lambda i: model.update_density_grid(5.9, warmup=True), 3))
print("update_density_grid(warmup=False)  : %.2f ms" % timeit(lambda i: model.update_density_grid(5.9, warmup=False), 3))
tr.step = 64
print("16 steps incl. one update (per step): %.2f ms" % (timeit(lambda i: [tr.train_step(*pool[j % 4]) for j in range(16)], 2) / 16))
