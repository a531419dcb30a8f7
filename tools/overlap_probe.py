"""Does the L2-atomic-bound hash scatter overlap with the issue-bound MLP backward when both are in flight
(two streams)?  Serial vs concurrent timing on the step's real sample set."""
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
S = xn.shape[0]; H = S // 2
table = model.xyz_encoder.params.detach()
y = tcnn.grid_forward(xn, table, g)
dy = torch.randn(S, 32, device=dev); dh = torch.randn(S, 16, device=dev); drgb = torch.randn(S, 3, device=dev)
h = torch.randn(S, 16, device=dev)
dt = torch.zeros(g.n_params, device=dev)
m1, m2 = model.sigma_net.mlp, model.rgb_net.mlp
p1, p2 = model.sigma_net.params.detach(), model.rgb_net.params.detach()
hi = torch.cuda.Stream(priority=-1); lo = torch.cuda.Stream(priority=0)

def tm(fn, it=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it

sl = [slice(0, H), slice(H, S)]
hash_bw = lambda k: tcnn.grid_backward_params(xn[sl[k]], dy[sl[k]], g, out=dt)
hash_fw = lambda k: tcnn.grid_forward(xn[sl[k]], table, g)
sig_bw = lambda k: tcnn.mlp_backward([(y[sl[k]], 32, 0)], p1, m1, dh[sl[k]], [True])
rgb_bw = lambda k: tcnn.mlp_backward([(dirs[sl[k]], 16, 1), (h[sl[k]], 16, 0)], p2, m2, drgb[sl[k]], [False, True])
sig_fw = lambda k: tcnn.mlp_forward([(y[sl[k]], 32, 0)], p1, m1)
rgb_fw = lambda k: tcnn.mlp_forward([(dirs[sl[k]], 16, 1), (h[sl[k]], 16, 0)], p2, m2)

def serial(a, b):
    def f():
        a(0); b(1)
    return f
def conc(a, b, a_hi=True):
    def f():
        cur = torch.cuda.current_stream()
        hi.wait_stream(cur); lo.wait_stream(cur)
        with torch.cuda.stream(hi if a_hi else lo): a(0)
        with torch.cuda.stream(lo if a_hi else hi): b(1)
        cur.wait_stream(hi); cur.wait_stream(lo)
    return f
for slots in (None, 2, 3):
    if slots: os.environ["NGP_MLP_SLOTS_BW"] = str(slots); os.environ["NGP_MLP_SLOTS_FW"] = str(slots + 1)
    print("slots override", slots)
    for name, a, b in (("sigma_bw | hash_bw", sig_bw, hash_bw), ("rgb_bw | hash_bw", rgb_bw, hash_bw), ("sigma_fw | hash_fw", sig_fw, hash_fw), ("rgb_fw | hash_fw", rgb_fw, hash_fw)):
        ta, tb = tm(lambda: a(0)), tm(lambda: b(1))
        print(f"  {name:22s} alone {ta:.3f} + {tb:.3f} = {ta+tb:.3f}  serial {tm(serial(a, b)):.3f}  concurrent(mlp hi) {tm(conc(a, b, True)):.3f}  concurrent(mlp lo) {tm(conc(a, b, False)):.3f}", flush=True)
