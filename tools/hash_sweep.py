"""Sweep NGP_HASH_SPT for hashgrid_bw_params on the step's real sample set."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "instant-ngp-pp_b200")]
import torch
from ngp_b200 import vren, tcnn
from ngp_b200.networks import NGPCompact
from synth_scenes import BoxScene, scene_density_grid
from ngp_b200.custom_functions import RayMarcher
dev = torch.device("cuda", 0)
scene = BoxScene("lego", device=dev); poses = scene.poses(100)
model = NGPCompact(scale=0.5).to(dev)
model.density_grid.copy_(scene_density_grid(scene)); vren.packbits(model.density_grid, 0.5, model.density_bitfield)
ro, rd = scene.sample_rays(1 << 18, poses)
with torch.no_grad():
    _, hits_t, _ = vren.ray_aabb_intersect(ro, rd, model.center, model.half_size, 1)
    ra, xyzs, dirs, deltas, ts, tot = RayMarcher.apply(ro, rd, hits_t[:, 0].contiguous(), model.density_bitfield, 1, 0.5, 0.0, 128, 1024)
g = model.xyz_encoder.grid
xn = ((xyzs - model.xyz_min) / (model.xyz_max - model.xyz_min)).contiguous()
dy = torch.randn(xn.shape[0], 32, device=dev)
ref = None
def tm(fn, it=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it
import itertools
for lc, spt in itertools.product((4, 2), (16, 24, 32)):
    os.environ["NGP_HASH_SPT"] = str(spt); os.environ["NGP_HASH_LC"] = str(lc)
    dt = torch.zeros(g.n_params, device=dev)
    t = tm(lambda: tcnn.grid_backward_params(xn, dy, g, out=dt))
    dt.zero_(); tcnn.grid_backward_params(xn, dy, g, out=dt)
    if ref is None: ref = dt.clone()
    print(f"LC {lc} SPT {spt}: {t:.3f} ms  samples {xn.shape[0]}  rel diff vs SPT8 {float((dt-ref).norm()/ref.norm()):.2e}", flush=True)
