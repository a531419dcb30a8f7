"""Training marcher's count pass on a Lego-shaped batch (seeded rays): x-major bitfield lookups (NGP_MARCH_LINEAR=0|1) and resident
CTAs per SM (NGP_MARCH_CTAS_PER_SM).  The switches are read once per process: run once per setting and compare the lines; the
checksum of the per-ray counts must not move."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "instant-ngp-pp_b200")]
import torch
from ngp_b200 import vren
from synth_scenes import BoxScene, scene_density_grid
dev = torch.device("cuda", 0)
R = 1 << 18
scene = BoxScene("lego", device=dev); poses = scene.poses(100)
grid = scene_density_grid(scene)
bf = torch.zeros(128 ** 3 // 8, dtype=torch.uint8, device=dev); vren.packbits(grid, 0.5, bf)
torch.manual_seed(11)
ro, rd = scene.sample_rays(R, poses)
_, hits_t, _ = vren.ray_aabb_intersect(ro, rd, torch.zeros(1, 3, device=dev), torch.full((1, 3), 0.5, device=dev), 1)
h = hits_t[:, 0].contiguous()
torch.manual_seed(7)
noise = torch.rand(R, device=dev)
def run():
    return vren.raymarching_train_count(ro, rd, h, bf, 1, 0.5, 0.0, noise, 128, 1024)
for _ in range(3): plan = run()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20): plan = run()
e.record(); torch.cuda.synchronize()
n = plan.workspace[:R * 4].view(torch.int32).long()
chk = int((n * (torch.arange(R, device=dev) % 65521 + 1)).sum())
cfg = " ".join(f"{k}={os.environ.get(k, 'default')}" for k in ("NGP_MARCH_LINEAR", "NGP_MARCH_CTAS_PER_SM"))
print(f"{cfg}: count pass {s.elapsed_time(e) / 20:.3f} ms for 2^18 rays, {int(n.sum())} samples, {int((n > 0).sum())} rays with samples, "
      f"max {int(n.max())}, checksum {chk}")
