"""Where do the hash-grid scatter's L2 reduction requests come from?  Times ngp_hashgrid_bw_params_tiles_range per level
pair (the lane pair's unit of work) on the headline sample set and, run under
  ncu --metrics lts__t_requests_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_red.sum,gpu__time_duration.sum
gives the reduction requests per pair.  VERDICT r1 item 4 proposed privatising the dense coarse levels in shared memory;
this probe shows what share of the requests those levels carry.  usage: python tools/scatter_levels_probe.py [log2_rays=18]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "instant-ngp-pp_b200"))
import torch
from ngp_b200 import vren, tcnn
from ngp_b200.networks import NGPCompact
from ngp_b200._lib import lib, ptr, check, stream
from synth_scenes import BoxScene, scene_density_grid
from ngp_b200.custom_functions import RayMarcher

dev = torch.device("cuda", 0)
scene = BoxScene("lego", device=dev)
model = NGPCompact(scale=0.5).to(dev)
model.density_grid.copy_(scene_density_grid(scene)); vren.packbits(model.density_grid, 0.5, model.density_bitfield)
R = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 18)
ro, rd = scene.sample_rays(R, scene.poses(100))
with torch.no_grad():
    _, hits_t, _ = vren.ray_aabb_intersect(ro, rd, model.center, model.half_size, 1)
    ra, xyzs, dirs, deltas, ts, tot = RayMarcher.apply(ro, rd, hits_t[:, 0].contiguous(), model.density_bitfield, 1, 0.5, 0.0, 128, 1024)
S = xyzs.shape[0]
g, aabb = model.xyz_encoder.grid, model.aabb()
xw = xyzs.contiguous()
dy = torch.randn((S + 127) // 128 * 128 * 32, device=dev)
dtab = torch.zeros_like(model.xyz_encoder.params)
res = list(g.resolutions)


def run(lo, hi, reps):
    for _ in range(2):
        check(lib.ngp_hashgrid_bw_params_tiles_range(ptr(xw), tcnn._aabb_arg(aabb), ptr(dy), *g.args(), S, ptr(dtab), lo, hi, stream()), "bw")
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        check(lib.ngp_hashgrid_bw_params_tiles_range(ptr(xw), tcnn._aabb_arg(aabb), ptr(dy), *g.args(), S, ptr(dtab), lo, hi, stream()), "bw")
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


reps = 1 if os.environ.get("NCU") else 10
print(f"rays {R} samples {S}  ({S / R:.1f} per ray)")
full = run(0, g.n_levels, reps)
print(f"levels  0-{g.n_levels - 1:2d}: {full:7.3f} ms  (one launch)")
tot = 0.0
for lo in range(0, g.n_levels, 2):
    t = run(lo, lo + 2, reps); tot += t
    kind = ["dense" if g.dense[l] else "hashed" for l in (lo, lo + 1)]
    print(f"levels {lo:2d}-{lo + 1:2d}: {t:7.3f} ms  res {res[lo]:5d},{res[lo + 1]:5d}  entries {g.sizes[lo]:7d},{g.sizes[lo + 1]:7d}  {kind[0]},{kind[1]}")
print(f"sum of the pairs: {tot:.3f} ms")
for lo in (2, 4, 6, 8):
    print(f"levels {lo:2d}-{g.n_levels - 1:2d}: {run(lo, g.n_levels, reps):7.3f} ms  (one launch, coarse levels below {lo} left out)")


def run_full(reps):
    for _ in range(2):
        check(lib.ngp_hashgrid_bw_params_tiles(ptr(xw), tcnn._aabb_arg(aabb), ptr(dy), *g.args(), S, ptr(dtab), stream()), "bw")
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        check(lib.ngp_hashgrid_bw_params_tiles(ptr(xw), tcnn._aabb_arg(aabb), ptr(dy), *g.args(), S, ptr(dtab), stream()), "bw")
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


