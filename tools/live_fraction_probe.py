"""What fraction of a step's samples lies before their ray's termination (T > T_threshold)?  Samples past it have zero
weight and zero gradient (volumerendering.cu:111-114 breaks out of the loop), yet are encoded and pushed through the MLPs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "instant-ngp-pp_b200")]
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch
from ngp_b200 import vren
from ngp_b200.networks import NGPCompact
from synth_scenes import BoxScene, scene_density_grid
from ngp_b200.trainer import Trainer
dev = torch.device("cuda", 0)
for kind, scale, T, esf, lr in (("lego", 0.5, 19, 0.0, 1e-2), ("street", 8.0, 22, 1 / 256, 2e-3)):
    scene = BoxScene(kind, device=dev); poses = scene.poses(100)
    model = NGPCompact(scale=scale, log2_T=T).to(dev)
    model.density_grid.copy_(scene_density_grid(scene)); vren.packbits(model.density_grid, 0.5, model.density_bitfield)
    tr = Trainer(model, lr=lr, render_kwargs=dict(exp_step_factor=esf, num_classes=0))
    for i in range(401):
        ro, rd = scene.sample_rays(1 << 18 if kind == "lego" else 1 << 16, poses); c, *_ = scene.shade(ro, rd)
        loss, res = tr.train_step(ro, rd, c)
        if i in (0, 50, 100, 200, 400):
            S = int(res["total_samples"]); live = int(res["vr_samples"].sum()); nz = int((res["ws"] > 0).sum())
            ra = res["rays_a"]; hit = int((ra[:, 2] > 0).sum())
            print(f"{kind} step {i}: samples {S} ({S / ro.shape[0]:.1f}/ray), composited before termination {live} = {live / S:.3f}, w>0 {nz / S:.3f}, rays with samples {hit / ro.shape[0]:.3f}", flush=True)
    del model, tr
    torch.cuda.empty_cache()
