"""Launches the individual hot kernels on a realistic sample set (for ncu captures).
usage: python tools/run_kernel.py [n_rays_log2=17]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "instant-ngp-pp_b200"))
import torch
from ngp_b200 import vren, tcnn
from ngp_b200.networks import NGPCompact
from synth_scenes import BoxScene, scene_density_grid
from ngp_b200.custom_functions import RayMarcher

dev = torch.device("cuda", 0)
scene = BoxScene("lego", device=dev)
poses = scene.poses(100)
model = NGPCompact(scale=0.5).to(dev)
model.density_grid.copy_(scene_density_grid(scene))
vren.packbits(model.density_grid, 0.5, model.density_bitfield)
R = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 17)
ro, rd = scene.sample_rays(R, poses)
with torch.no_grad():
    _, hits_t, _ = vren.ray_aabb_intersect(ro, rd, model.center, model.half_size, 1)
    for _ in range(2):
        ra, xyzs, dirs, deltas, ts, tot = RayMarcher.apply(ro, rd, hits_t[:, 0].contiguous(), model.density_bitfield, 1, 0.5, 0.0, 128, 1024)
    S = xyzs.shape[0]
    from ngp_b200._lib import lib, ptr, check, stream
    g = model.xyz_encoder.grid
    aabb = model.aabb()
    xw = xyzs.contiguous()
    table = model.xyz_encoder.params.detach()
    for _ in range(2):
        tiles = tcnn.grid_forward_tiles(xw, table, g, aabb)
    dy_tiles = torch.randn((S + 127) // 128 * 128 * 32, device=dev); dtab = torch.zeros_like(table)
    for _ in range(2):
        check(lib.ngp_hashgrid_bw_params_tiles(ptr(xw), tcnn._aabb_arg(aabb), ptr(dy_tiles), *g.args(), S, ptr(dtab), stream()), "bw")
    m1, m2 = model.sigma_net.mlp, model.rgb_net.mlp
    for _ in range(2):
        h, sig = tcnn.mlp_forward([(tiles, 32, 2)], model.sigma_net.params.detach(), m1, aux_exp=True, n=S)
    dh = torch.randn_like(h); ds = torch.randn_like(sig)
    for _ in range(2):
        tcnn.mlp_backward([(tiles, 32, 2)], model.sigma_net.params.detach(), m1, dh, [True], d_aux=ds, n=S, dseg_numel=dy_tiles.numel(), saved_out=h)
    segs = [(dirs, 16, 1), (h, 16, 0)]
    for _ in range(2):
        rgb = tcnn.mlp_forward(segs, model.rgb_net.params.detach(), m2)
    drgb = torch.randn_like(rgb)
    for _ in range(2):
        tcnn.mlp_backward(segs, model.rgb_net.params.detach(), m2, drgb, [False, True], saved_out=rgb)
    sigmas = torch.rand(S, device=dev) * 20
    z0 = torch.zeros(S, 0, device=dev)
    for _ in range(2):
        out = vren.composite_train_fw(sigmas, rgb, None, z0, deltas, ts, ra, 1e-4, 0)
    tot, opacity, depth, rgb_r, _, sem_r, ws = out
    g1 = torch.randn(R, device=dev); g3 = torch.randn(R, 3, device=dev); gws = torch.randn(S, device=dev)
    for _ in range(2):
        vren.composite_train_bw(g1, g1, g3, None, torch.zeros(R, 0, device=dev), gws, sigmas, rgb, None, ws, deltas, ts, ra,
                                opacity, depth, rgb_r, None, 1e-4, 0)
torch.cuda.synchronize()
print("rays", R, "samples", S)
