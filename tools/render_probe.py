"""Per-frame times of the test-time renderer on the headline field (after a short training): wavefront with the device-resident
round loop (render_wavefront_compact) vs the host-driven wavefront (device_loop=False) vs the reference-style loop.
usage: python tools/render_probe.py [frames=12] [W=1920] [H=1080]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "instant-ngp-pp_b200"))
import torch
from ngp_b200 import vren
from ngp_b200.networks import NGPCompact
from ngp_b200.trainer import Trainer
from ngp_b200.rendering import render
from synth_scenes import BoxScene, scene_density_grid

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 12
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
dev = torch.device("cuda", 0)
scene = BoxScene("lego", device=dev)
poses = scene.poses(100)
model = NGPCompact(scale=0.5).to(dev)
model.density_grid.copy_(scene_density_grid(scene)); vren.packbits(model.density_grid, 0.5, model.density_bitfield)
tr = Trainer(model, render_kwargs=dict(exp_step_factor=0.0, num_classes=0))
g = torch.Generator(device=dev).manual_seed(1)
for i in range(300):
    ro, rd = scene.sample_rays(1 << 18, poses, g)
    rgb, *_ = scene.shade(ro, rd)
    tr.train_step(ro, rd, rgb)
torch.cuda.synchronize()


def frame_rays(i):
    n = W * H
    px = torch.arange(n, device=dev)
    sc = scene.img_wh[0] / W
    u, v = (px % W).float() * sc + (sc - 1) / 2, (px // W).float() * sc + (sc - 1) / 2
    return scene.rays_from_pixels(poses[i % poses.shape[0]][None], torch.zeros(n, dtype=torch.long, device=dev), u, v)


variants = {"wavefront, pipelined rounds (default)": dict(renderer="wavefront", sample_schedule="geometric"),
            "wavefront, device-resident rounds": dict(renderer="wavefront", device_loop=True, sample_schedule="geometric"),
            "wavefront, host-driven rounds": dict(renderer="wavefront", pipelined=False, sample_schedule="geometric"),
            "reference-style loop, reference schedule": dict(renderer="loop", sample_schedule="reference")}
with torch.no_grad():
    for name, kw in variants.items():
        times, spr = [], 0
        for i in range(frames + 2):
            ro, rd = frame_rays(i)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = render(model, ro, rd, exp_step_factor=0.0, num_classes=0, test_time=True, T_threshold=1e-2, **kw)
            tot = int(r["total_samples"]); torch.cuda.synchronize()
            if i >= 2:
                times.append((time.perf_counter() - t0) * 1e3); spr = tot / (W * H)
        times.sort()
        med = times[len(times) // 2]
        print(f"{W}x{H} {name:42s}: median {med:7.2f} ms/frame = {W * H / med / 1e3:6.1f} Mrays/s   min {times[0]:7.2f}  max {times[-1]:7.2f}   {spr:.1f} samples/ray", flush=True)
print("peak memory GB", torch.cuda.max_memory_allocated() / 2 ** 30)
