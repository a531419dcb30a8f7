"""Cost of one occupancy update (torch-op path of round 1 vs the kernel chain) and of the density-net kernels alone."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "instant-ngp-pp_b200"))
import torch
from ngp_b200 import vren
from ngp_b200.networks import NGP, NGPCompact, _dn_fw, _dn_bw
from synth_scenes import BoxScene, scene_density_grid

dev = torch.device("cuda", 0)
THR = 0.01 * 1024 / 3 ** 0.5


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time(); s.record(); fn(); e.record(); torch.cuda.synchronize()
        out.append((s.elapsed_time(e), (time.time() - t0) * 1e3))
    return min(o[0] for o in out), min(o[1] for o in out)


for name, make, kind in (("lego/NGPCompact", lambda: NGPCompact(scale=0.5), "lego"), ("playground/NGP", lambda: NGP(scale=8.0, embed_a=True, embed_a_len=8, classes=7), "street")):
    scene = BoxScene(kind, device=dev)
    m = make().to(dev)
    m.density_grid.copy_(scene_density_grid(scene)); vren.packbits(m.density_grid, 0.5, m.density_bitfield)
    for fused in (False, True):
        m.fused_update = fused
        for warm in (True, False):
            gpu, wall = timeit(lambda: m.update_density_grid(THR, warmup=warm))
            print(f"{name:18s} update_density_grid fused={fused!s:5s} warmup={warm!s:5s}: {gpu:8.2f} ms GPU  {wall:8.2f} ms wall", flush=True)

n = 14_000_000
g = torch.Generator(device=dev).manual_seed(0)
e = torch.randn(n, 128, device=dev, generator=g) * 0.3
W1 = torch.randn(128, 128, device=dev, generator=g) * 0.1; b1 = torch.randn(128, device=dev, generator=g)
W2 = torch.randn(1, 128, device=dev, generator=g) * 0.3; b2 = torch.randn(1, device=dev, generator=g)
gpu, _ = timeit(lambda: _dn_fw(e, W1, b1, W2, b2, True)); print(f"density_net_fw  (g_e)    n={n}: {gpu:.2f} ms")
gpu, _ = timeit(lambda: _dn_fw(e, W1, b1, W2, b2, False)); print(f"density_net_fw  (no g_e) n={n}: {gpu:.2f} ms")
sig, s2, ge = _dn_fw(e, W1, b1, W2, b2, True)
dge = torch.randn(n, 128, device=dev, generator=g); ds = torch.randn(n, device=dev, generator=g)
gpu, _ = timeit(lambda: _dn_bw(e, dge, ge, ds, s2, W1, b1, W2, True)); print(f"density_net_bw  (both)   n={n}: {gpu:.2f} ms")
gpu, _ = timeit(lambda: _dn_bw(e, None, None, ds, s2, W1, b1, W2, True)); print(f"density_net_bw  (dsigma) n={n}: {gpu:.2f} ms")
