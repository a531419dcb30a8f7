"""Training marcher with / without empty-ray culling (NGP_MARCH_CULL=0|1, read once per process) on a Lego-shaped batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "instant-ngp-pp_b200")]
import torch
from ngp_b200 import vren
from synth_scenes import BoxScene, scene_density_grid
dev = torch.device("cuda", 0)
scene = BoxScene("lego", device=dev); poses = scene.poses(100)
grid = scene_density_grid(scene)
bf = torch.zeros(128 ** 3 // 8, dtype=torch.uint8, device=dev); vren.packbits(grid, 0.5, bf)
ro, rd = scene.sample_rays(1 << 18, poses)
_, hits_t, _ = vren.ray_aabb_intersect(ro, rd, torch.zeros(1, 3, device=dev), torch.full((1, 3), 0.5, device=dev), 1)
h = hits_t[:, 0].contiguous()
noise = torch.rand(1 << 18, device=dev)
def run():
    return vren.raymarching_train_count(ro, rd, h, bf, 1, 0.5, 0.0, noise, 128, 1024)
for _ in range(3): plan = run()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20): plan = run()
e.record(); torch.cuda.synchronize()
print(f"NGP_MARCH_CULL={os.environ.get('NGP_MARCH_CULL', '1')}: count pass {s.elapsed_time(e) / 20:.3f} ms for 2^18 rays")
