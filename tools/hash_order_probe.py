"""Block order of the hash-grid gather / scatter (GridMeta::chunk_major, hashgrid.cu) on the two table regimes:
lego shape (T=2^19: 44 MB, L2 resident) and street shape (T=2^22: 363 MB, beyond the 126 MB L2).
NGP_HASH_ORDER=0 level-chunk fastest, 1 level-chunk slowest; the default picks by table size.
    python tools/hash_order_probe.py [lego|street ...]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "instant-ngp-pp_b200")]
import torch
from ngp_b200 import vren, tcnn
from ngp_b200._lib import lib, ptr, check, stream
from ngp_b200.networks import NGPCompact
from synth_scenes import BoxScene, scene_density_grid
from ngp_b200.custom_functions import RayMarcher

dev = torch.device("cuda", 0)
SHAPES = {"lego": dict(scene="lego", scale=0.5, log2_T=19, esf=0.0, views=100),
          "street": dict(scene="street", scale=8.0, log2_T=22, esf=1.0 / 256, views=128)}


def tm(fn, it=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it


for name in (sys.argv[1:] or ["lego", "street"]):
    wl = SHAPES[name]
    scene = BoxScene(wl["scene"], device=dev); poses = scene.poses(wl["views"])
    model = NGPCompact(scale=wl["scale"], log2_T=wl["log2_T"]).to(dev)
    model.density_grid.copy_(scene_density_grid(scene)); vren.packbits(model.density_grid, 0.5, model.density_bitfield)
    torch.nn.init.uniform_(model.xyz_encoder.params, -1.0, 1.0)       # distinguishable entries for the equality check
    ro, rd = scene.sample_rays(1 << 18, poses)
    with torch.no_grad():
        _, hits_t, _ = vren.ray_aabb_intersect(ro, rd, model.center, model.half_size, 1)
        ra, xyzs, dirs, deltas, ts, tot = RayMarcher.apply(ro, rd, hits_t[:, 0].contiguous(), model.density_bitfield,
                                                           model.cascades, model.scale, wl["esf"], 128, 1024)
    del dirs, deltas, ts
    g = model.xyz_encoder.grid
    S = xyzs.shape[0]
    LF = g.n_levels * g.n_features
    k0p = (LF + 15) // 16 * 16
    aabb = model.aabb()
    table = model.xyz_encoder.params.detach()
    dy_tiles = torch.randn((S + 127) // 128 * 128 * k0p, device=dev)
    dtab = torch.zeros_like(table)
    print(f"== {name}: {S} samples ({S / (1 << 18):.1f}/ray), table {table.numel() * 4 / 2**20:.0f} MB", flush=True)
    ref_f = ref_b = None
    for order in ("0", "1", None):
        if order is None:
            os.environ.pop("NGP_HASH_ORDER", None)
        else:
            os.environ["NGP_HASH_ORDER"] = order
        fw = lambda: tcnn.grid_forward_tiles(xyzs, table, g, aabb)
        bw = lambda: check(lib.ngp_hashgrid_bw_params_tiles(ptr(xyzs), tcnn._aabb_arg(aabb), ptr(dy_tiles), *g.args(), S, ptr(dtab), stream()), "bw")
        tf, tb = tm(fw), tm(bw)
        tiles = fw(); dtab.zero_(); bw(); torch.cuda.synchronize()
        tv = tiles.view(-1, k0p // 8, 128 * 16 + 64)[:, :, :128 * 16]       # the 64 pad bytes per chunk are never written
        if ref_f is None:
            ref_f, ref_b = tv.clone(), dtab.clone()
        same_f = bool((tv == ref_f).all())
        rel_b = float((dtab - ref_b).norm() / ref_b.norm())
        fw_gbs = S * (12 + 8 * LF * 4 + LF * 2) / tf / 1e6
        bw_gbs = S * (12 + LF * 4 + 16 * LF * 4) / tb / 1e6
        print(f"order {order if order is not None else 'auto'}: gather {tf:8.3f} ms ({fw_gbs:7.0f} GB/s algorithmic)   scatter {tb:8.3f} ms ({bw_gbs:7.0f} GB/s)"
              f"   gather bits equal {same_f}   scatter rel diff {rel_b:.1e}", flush=True)
    del dy_tiles, dtab, xyzs, model
    torch.cuda.empty_cache()
