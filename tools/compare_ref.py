"""Kernel-by-kernel timing of OUR C-ABI path against the REFERENCE'S OWN CUDA kernels (oracle/_ref/vren_ref.so,
compiled in place from /root/reference/models/csrc) on identical tensors at the headline size (2^18 rays of the
Lego-shaped scene).  Times include the host-side allocation / zero-fill each API performs, because that is what a
caller pays (the reference zero-fills N_rays*1024 rows per marching call).  CUDA events, 3 warm-ups, 10 iterations."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "instant-ngp-pp_b200"))
import torch
from ngp_b200 import vren
from synth_scenes import BoxScene, scene_density_grid
from oracle import build_ref

ref = build_ref.load()
assert ref is not None
dev = torch.device("cuda", 0)
scene = BoxScene("lego", device=dev)
poses = scene.poses(100)
R = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 18)
ro, rd = scene.sample_rays(R, poses)
grid = scene_density_grid(scene)
bf = torch.zeros(grid.numel() // 8, dtype=torch.uint8, device=dev)
vren.packbits(grid, 0.5, bf)
center = torch.zeros(1, 3, device=dev); half = torch.full((1, 3), 0.5, device=dev)


def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

def device_ms(fn, iters=10):
    """GPU time per call from CUPTI kernel records (torch.profiler): every kernel / memset the call launches, without the
    host-side launch overhead that dominates the event timing of the 10-microsecond ops."""
    from torch.profiler import profile, ProfilerActivity
    for _ in range(3): fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(iters): fn()
        torch.cuda.synchronize()
    return sum(e.self_device_time_total for e in prof.key_averages()) / iters / 1e3


rows = []
def cmp(name, f_ours, f_ref, iters=10):
    a, b = timeit(f_ours, iters), timeit(f_ref, iters)
    da, db = device_ms(f_ours, iters), device_ms(f_ref, iters)
    rows.append({"op": name, "ours_ms": a, "reference_ms": b, "speedup": b / a, "ours_device_ms": da, "reference_device_ms": db, "device_speedup": db / da})
    print(f"{name:36s} call: ours {a:8.3f} ms  reference {b:8.3f} ms  x{b / a:5.1f}   |   GPU time: ours {da:8.4f} ms  reference {db:8.4f} ms  x{db / da:5.1f}", flush=True)

coords = torch.randint(0, 128, (128 ** 3, 3), dtype=torch.int32, device=dev)
cmp("morton3D (128^3)", lambda: vren.morton3D(coords), lambda: ref.morton3D(coords))
idx = vren.morton3D(coords)
cmp("morton3D_invert (128^3)", lambda: vren.morton3D_invert(idx), lambda: ref.morton3D_invert(idx))
g5 = torch.randn(5, 128 ** 3, device=dev); b5 = torch.zeros(5 * 128 ** 3 // 8, dtype=torch.uint8, device=dev)
cmp("packbits (5 x 128^3)", lambda: vren.packbits(g5, 0.1, b5), lambda: ref.packbits(g5, 0.1, b5))
cmp("ray_aabb_intersect", lambda: vren.ray_aabb_intersect(ro, rd, center, half, 1), lambda: ref.ray_aabb_intersect(ro, rd, center, half, 1))
_, hits_t, _ = vren.ray_aabb_intersect(ro, rd, center, half, 1)
h = hits_t[:, 0].contiguous()
noise = torch.rand(R, device=dev)

def ours_march():
    return vren.raymarching_train(ro, rd, h, bf, 1, 0.5, 0.0, noise, 128, 1024)
def ref_march():
    out = ref.raymarching_train(ro, rd, h, bf, 1, 0.5, 0.0, noise, 128, 1024)
    tot = int(out[5][0])               # the reference's Python does this read-back too (custom_functions.py:93)
    return out
cmp("raymarching_train (+count readback)", ours_march, ref_march, iters=5)
rays_a, xyzs, dirs, deltas, ts, counter = ours_march()
S = xyzs.shape[0]; C = 7
sig = torch.rand(S, device=dev) * 30; rgbs = torch.rand(S, 3, device=dev); nrm = torch.randn(S, 3, device=dev)
sems = torch.rand(S, C, device=dev)
cmp("composite_train_fw (C=7)", lambda: vren.composite_train_fw(sig, rgbs, nrm, sems, deltas, ts, rays_a, 1e-4, C),
    lambda: ref.composite_train_fw(sig, rgbs, nrm, sems, deltas, ts, rays_a, 1e-4, C))
tot, op, dep, rgb, nr, sm, ws = vren.composite_train_fw(sig, rgbs, nrm, sems, deltas, ts, rays_a, 1e-4, C)
g = [torch.randn_like(t) for t in (op, dep, rgb, nr, sm, ws)]
cmp("composite_train_bw (C=7)", lambda: vren.composite_train_bw(*g, sig, rgbs, nrm, ws, deltas, ts, rays_a, op, dep, rgb, nr, 1e-4, C),
    lambda: ref.composite_train_bw(*g, sig, rgbs, nrm, ws, deltas, ts, rays_a, op, dep, rgb, nr, 1e-4, C))
nd = torch.rand(S, 3, device=dev); no = torch.rand(S, device=dev)
cmp("composite_refloss_fw", lambda: vren.composite_refloss_fw(sig, nd, no, deltas, ts, rays_a, 1e-4),
    lambda: ref.composite_refloss_fw(sig, nd, no, deltas, ts, rays_a, 1e-4))
lo, lp = vren.composite_refloss_fw(sig, nd, no, deltas, ts, rays_a, 1e-4)
glo, glp = torch.randn_like(lo), torch.randn_like(lp)
cmp("composite_refloss_bw", lambda: vren.composite_refloss_bw(glo, glp, sig, nd, no, deltas, ts, rays_a, lo, lp, 1e-4),
    lambda: ref.composite_refloss_bw(glo, glp, sig, nd, no, deltas, ts, rays_a, lo, lp, 1e-4))
cmp("distortion_loss_fw", lambda: vren.distortion_loss_fw(ws, deltas, ts, rays_a), lambda: ref.distortion_loss_fw(ws, deltas, ts, rays_a))
loss, wi, wti = vren.distortion_loss_fw(ws, deltas, ts, rays_a)
gl = torch.randn_like(loss)
cmp("distortion_loss_bw", lambda: vren.distortion_loss_bw(gl, wi, wti, ws, deltas, ts, rays_a), lambda: ref.distortion_loss_bw(gl, wi, wti, ws, deltas, ts, rays_a))
# test-time round: 8 samples for every ray
alive = torch.arange(R, device=dev)
def ours_test():
    ht = h.clone(); return vren.raymarching_test(ro, rd, ht, alive, bf, 1, 0.5, 0.0, 128, 1024, 8)
def ref_test():
    ht = h.clone(); return ref.raymarching_test(ro, rd, ht, alive, bf, 1, 0.5, 0.0, 128, 1024, 8)
cmp("raymarching_test (8 samples)", ours_test, ref_test)
x8, d8, dl8, ts8, neff = ours_test()
s8 = torch.rand(R, 8, device=dev) * 30; c8 = torch.rand(R, 8, 3, device=dev); n8 = torch.randn(R, 8, 3, device=dev); m8 = torch.rand(R, 8, C, device=dev)
def mk():
    return [torch.zeros(R, device=dev), torch.zeros(R, device=dev), torch.zeros(R, 3, device=dev), torch.zeros(R, 3, device=dev), torch.zeros(R, 3, device=dev), torch.zeros(R, C, device=dev)]
st = mk()
cmp("composite_test_fw (8 samples, C=7)", lambda: vren.composite_test_fw(s8, c8, n8, n8, m8, dl8, ts8, h, alive.clone(), 1e-2, C, neff, *st),
    lambda: ref.composite_test_fw(s8, c8, n8, n8, m8, dl8, ts8, h, alive.clone(), 1e-2, C, neff, *st))
print(json.dumps({"rays": R, "samples": S, "rows": rows}))
