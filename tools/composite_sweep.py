"""Lanes-per-ray sweep of the compositors + distortion loss on a trained-state sample set."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "instant-ngp-pp_b200")]
import torch
from ngp_b200 import vren
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
S = xyzs.shape[0]; R = ra.shape[0]
n = ra[:, 2]
print(f"rays {R} samples {S}: empty rays {100 * float((n == 0).float().mean()):.1f} %, mean over non-empty {float(n[n > 0].float().mean()):.1f}, p50 {float(n[n > 0].float().median()):.0f}, max {int(n.max())}")
sig = torch.rand(S, device=dev) * 4; rgb = torch.rand(S, 3, device=dev)
z0 = torch.zeros(S, 0, device=dev)
def tm(fn, it=10):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it
from ngp_b200.custom_functions import VolumeRendererLite
from ngp_b200.losses import DistortionLoss
# second axis (round 2): threads per CTA of the group kernels (NGP_COMPOSITE_BLOCK; CTA slots come free when the longest ray is done)
# third axis: NGP_COMPOSITE_TILED (G = 32: a warp walks 8 consecutive rays instead of one launch slot per ray)
for G, B, T in ((16, 256, 0), (32, 256, 0), (32, 64, 0), (32, 64, 1), (32, 128, 1), (32, 256, 1)) if len(sys.argv) < 2 else ((32, 64, 0), (32, 64, 1), (32, 128, 1)):
    os.environ["NGP_COMPOSITE_G"] = str(G)
    os.environ["NGP_COMPOSITE_BLOCK"] = str(B)
    os.environ["NGP_COMPOSITE_TILED"] = str(T)
    s1 = sig.clone().requires_grad_(True); c1 = rgb.clone().requires_grad_(True)
    def fw():
        return VolumeRendererLite.apply(s1, c1, deltas, ts, ra, 1e-4)
    out = fw()
    loss = out[1].sum() + out[3].sum()
    t_fw = tm(fw)
    t_bw = tm(lambda: torch.autograd.grad(loss, (s1, c1), retain_graph=True))
    ws = out[4].detach().clone().requires_grad_(True)
    t_dfw = tm(lambda: DistortionLoss.apply(ws, deltas, ts, ra))
    dl = DistortionLoss.apply(ws, deltas, ts, ra).sum()
    t_dbw = tm(lambda: torch.autograd.grad(dl, (ws,), retain_graph=True))
    print(f"G={G} block={B} tiled={T}: composite fw {t_fw:.3f} ms  bw (incl. autograd glue) {t_bw:.3f} ms  distortion fw {t_dfw:.3f} ms  bw {t_dbw:.3f} ms", flush=True)
