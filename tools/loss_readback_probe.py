"""How should a training step hand its loss to the host every step?  Times the headline step (resident inputs) with:
A no read-back, B float(loss) after the step, C Trainer(host_loss=True): pinned copy after the forward + event wait after
the step is enqueued, D pinned copy, value of the PREVIOUS step read (no wait at all)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "instant-ngp-pp_b200")]
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch
from ngp_b200 import vren
from ngp_b200.networks import NGPCompact
from synth_scenes import BoxScene, scene_density_grid
from ngp_b200.trainer import Trainer

dev = torch.device("cuda", 0)
scene = BoxScene("lego", device=dev); poses = scene.poses(100)
model = NGPCompact(scale=0.5).to(dev)
model.density_grid.copy_(scene_density_grid(scene)); vren.packbits(model.density_grid, 0.5, model.density_bitfield)
tr = Trainer(model, render_kwargs=dict(exp_step_factor=0.0, num_classes=0))
batches = []
for _ in range(4):
    ro, rd = scene.sample_rays(1 << 18, poses); c, *_ = scene.shade(ro, rd); batches.append((ro, rd, c))
for i in range(100):
    tr.train_step(*batches[i % 4])
torch.cuda.synchronize()
pinned = torch.empty((), dtype=torch.float32).pin_memory()
print("pinned 0-dim is_pinned:", pinned.is_pinned(), flush=True)


def run(name, fn, steps=30):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); s.record()
    for i in range(steps):
        fn(i)
    e.record(); torch.cuda.synchronize()
    print(f"{name}: {s.elapsed_time(e) / steps:.3f} ms/step (device), {(time.perf_counter() - t0) / steps * 1e3:.3f} ms/step (host)", flush=True)


def A(i): tr.train_step(*batches[i % 4])
def B(i): return float(tr.train_step(*batches[i % 4])[0])
def C(i): return tr.train_step(*batches[i % 4], host_loss=True)[0]
prev = [None]
def D(i):
    loss, _ = tr.train_step(*batches[i % 4])
    v = float(pinned) if prev[0] is not None else 0.0
    pinned.copy_(loss, non_blocking=True); prev[0] = True
    return v
for name, fn in (("A none", A), ("B float(loss) after the step", B), ("C host_loss", C), ("D deferred pinned", D), ("A none", A), ("C host_loss", C)):
    run(name, fn)

# ---- with the per-step host -> device input copy (bench.py e2e): indices + rgb from pinned host memory
from ngp_b200 import ray_utils
W_, H_ = scene.img_wh
K_ = [[scene.focal, 0.0, W_ / 2], [0.0, scene.focal, H_ / 2], [0.0, 0.0, 1.0]]
directions = ray_utils.get_ray_directions(H_, W_, K_, device=dev)
host = []
for _ in range(4):
    img = torch.randint(poses.shape[0], (1 << 18,), device=dev); pix = torch.randint(W_ * H_, (1 << 18,), device=dev)
    ro, rd = ray_utils.get_rays_indexed(directions, poses, img, pix); c, *_ = scene.shade(ro, rd)
    host.append(tuple(t.cpu().pin_memory() for t in (img, pix, c)))
copy_stream = torch.cuda.Stream(device=dev)
dev_bufs = [tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host[0]) for _ in range(2)]
free_ev, pending, seq = [None, None], [], [0]


def prefetch():
    b = seq[0] % 2
    with torch.cuda.stream(copy_stream):
        if free_ev[b] is not None:
            copy_stream.wait_event(free_ev[b])
        for dst, src in zip(dev_bufs[b], host[seq[0] % 4]):
            dst.copy_(src, non_blocking=True)
        ev = torch.cuda.Event(); ev.record(copy_stream)
    pending.append((b, ev)); seq[0] += 1


def e2e(host_loss, pre):
    def fn(i):
        main = torch.cuda.current_stream()
        if pre:
            if not pending:
                prefetch()
            b, ev = pending.pop(0); main.wait_event(ev); prefetch()
            img, pix, c = dev_bufs[b]
        else:
            img, pix, c = (t.to(dev, non_blocking=True) for t in host[i % 4])
        o, d = ray_utils.get_rays_indexed(directions, poses, img, pix)
        loss, _ = tr.train_step(o, d, c, host_loss=host_loss)
        if pre:
            free_ev[b] = torch.cuda.Event(); free_ev[b].record(main)
        return float(loss)
    return fn
for name, fn in (("E prefetch + float(loss)", e2e(False, True)), ("F prefetch + host_loss", e2e(True, True)),
                 ("G same-stream copy + host_loss", e2e(True, False)), ("H same-stream copy + float(loss)", e2e(False, False)),
                 ("F prefetch + host_loss", e2e(True, True)), ("A none", A)):
    run(name, fn)
