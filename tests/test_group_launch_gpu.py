"""Launch shapes of the per-ray group kernels (compositors, distortion / Ref-NeRF losses: csrc/scan.cuh for_each_ray).
At G = 32 a warp walks `tile` consecutive rays (1..8, chosen from the mean samples per ray) instead of owning one launch slot
per ray.  The walk is the same kernel whatever the tile and the CTA size, so every output must be BIT-identical under every
tiled shape — including ragged ray counts (not a multiple of the tile), batches that are mostly empty rays and rays longer than
the 4 x 32 samples a trip keeps in flight.  The one-slot-per-ray shape is a second instantiation of the same source (nothing
obliges the compiler to contract its multiply-adds the same way), so it is held to 1e-5 of each output's scale instead.
The switches are read at every call (tuning / A-B only)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vren():
    from ngp_b200 import vren as v
    return v


def packed_rays(n_rays, mean_len, empty_frac, seed, max_len=700):
    """rays_a (R,3) i64 [ray_idx, start, N] in ray order with a given share of empty rays; -> rays_a, S"""
    g = torch.Generator().manual_seed(seed)
    n = torch.poisson(torch.full((n_rays,), float(mean_len)), generator=g).long().clamp_(1, max_len)
    if n_rays > 3:
        n[3] = max_len                                      # one long ray: several trips, early termination inside
    n[torch.rand(n_rays, generator=g) < empty_frac] = 0
    start = torch.cumsum(n, 0) - n
    ra = torch.stack([torch.arange(n_rays), start, n], 1).contiguous()
    return ra.cuda(), int(n.sum())


def make_inputs(ra, S, classes, seed):
    """One fixed set of inputs for every launch shape (built once: torch.cumsum on the GPU is a decoupled look-back scan whose
    float roundings differ from call to call)."""
    R = ra.shape[0]
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s: torch.rand(*s, device="cuda", generator=g)
    dl = r(S) * 0.01 + 1e-3
    return dict(ra=ra, classes=classes, sig=r(S) * 30.0, rgb=r(S, 3), nrm=r(S, 3) - 0.5, sem=r(S, classes), dl=dl, ts=torch.cumsum(dl, 0),
                g_opa=r(R), g_dep=r(R), g_rgb=r(R, 3), g_nrm=r(R, 3), g_sem=r(R, classes), g_ws=r(S) - 0.5, g_dist=r(R),
                nd=r(S, 3) - 0.5, no=r(S), g_lo=r(R), g_lp=r(R, 3))


def run_all(vren, x):
    ra, C, sig, dl, ts = x["ra"], x["classes"], x["sig"], x["dl"], x["ts"]
    out = vren.composite_train_fw(sig, x["rgb"], x["nrm"], x["sem"], dl, ts, ra, 1e-4, C)
    tot, opa, dep, col, nor, se, ws = out
    bw = vren.composite_train_bw(x["g_opa"], x["g_dep"], x["g_rgb"], x["g_nrm"], x["g_sem"], x["g_ws"], sig, x["rgb"], x["nrm"], ws, dl, ts, ra,
                                 opa, dep, col, nor, 1e-4, C)
    al = vren.composite_alpha_fw(sig, dl, ra, 1e-4)
    dfw = vren.distortion_loss_fw(ws, dl, ts, ra)
    dbw = vren.distortion_loss_bw(x["g_dist"], dfw[1], dfw[2], ws, dl, ts, ra)
    rfw = vren.composite_refloss_fw(sig, x["nd"], x["no"], dl, ts, ra, 1e-4)
    rbw = vren.composite_refloss_bw(x["g_lo"], x["g_lp"], sig, x["nd"], x["no"], dl, ts, ra, rfw[0], rfw[1], 1e-4)
    res = list(out) + list(bw) + list(al) + list(dfw) + [dbw] + list(rfw) + list(rbw)
    torch.cuda.synchronize()
    return [t.clone() for t in res]


@pytest.mark.parametrize("n_rays,mean_len,empty_frac,classes", [
    (1000, 60, 0.75, 0),      # Lego-shaped: mostly empty rays, default tile 8 (ragged: 1000 = 125 x 8, 1003 below)
    (1003, 60, 0.75, 7),
    (517, 140, 0.3, 3),       # default tile 2-3
    (129, 400, 0.0, 0),       # street-shaped: every ray long, default tile 1
    (7, 50, 0.5, 2),          # fewer rays than one tile
])
def test_tiled_launch_shapes_are_bit_identical(vren, n_rays, mean_len, empty_frac, classes):
    ra, S = packed_rays(n_rays, mean_len, empty_frac, seed=n_rays)
    x = make_inputs(ra, S, classes, seed=5)
    keep = {k: os.environ.get(k) for k in ("NGP_COMPOSITE_G", "NGP_COMPOSITE_TILED", "NGP_COMPOSITE_BLOCK")}
    try:
        os.environ["NGP_COMPOSITE_G"] = "32"
        os.environ["NGP_COMPOSITE_TILED"] = "1"; os.environ["NGP_COMPOSITE_BLOCK"] = "128"
        ref = run_all(vren, x)
        shapes = [(str(t), b) for t in (2, 3, 5, 8) for b in ("64", "128")] + [("0", "256"), ("0", "64"), (None, None)]   # last: the library's own choice
        for tile, block in shapes:
            for k, v in (("NGP_COMPOSITE_TILED", tile), ("NGP_COMPOSITE_BLOCK", block)):
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
            if tile is None:
                os.environ.pop("NGP_COMPOSITE_G", None)
                if S / n_rays <= 12.0:                        # the library would pick G < 32 here: a different summation order
                    continue
            got = run_all(vren, x)
            for i, (a, b) in enumerate(zip(ref, got)):
                if tile == "0":                               # one launch slot per ray: the other instantiation
                    a64, b64 = a.double(), b.double()
                    tol = 1e-5 * float(a64.abs().max()) if a.numel() else 0.0
                    assert a.shape == b.shape and (a.numel() == 0 or float((a64 - b64).abs().max()) <= tol), \
                        f"output {i} differs beyond rounding under tile={tile} block={block}"
                else:
                    assert torch.equal(a, b), f"output {i} differs under tile={tile} block={block}"
    finally:
        for k, v in keep.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_empty_batch_and_all_empty_rays(vren):
    z = lambda *s: torch.zeros(*s, device="cuda")
    ra = torch.stack([torch.arange(40), torch.zeros(40, dtype=torch.long), torch.zeros(40, dtype=torch.long)], 1).contiguous().cuda()
    out = vren.composite_train_fw(z(0), z(0, 3), z(0, 3), z(0, 0), z(0), z(0), ra, 1e-4, 0)
    assert all(float(t.float().abs().sum()) == 0.0 for t in out[:6]) and out[6].numel() == 0
    loss, wi, wti = vren.distortion_loss_fw(z(0), z(0), z(0), ra)
    assert loss.shape[0] == 40 and float(loss.abs().sum()) == 0.0
