"""Ray generation (reference datasets/ray_utils.py:10-72; called from train.py:136-156 every step).

get_ray_directions — camera-space pixel directions, un-normalised, pixel centres at +0.5 (ray_utils.py:10-46);
get_rays           — same signature as the reference's (directions (N,3), c2w (3,4) | (N,3,4)) -> rays_o, rays_d;
get_rays_indexed   — the training-step form: the two gathers `self.poses[img_idxs]`, `self.directions[pix_idxs]`
                     (train.py:136-137) and get_rays as ONE kernel, so a step's host input is
                     the batch's (img_idxs, pix_idxs, rgb) exactly as in the reference's data loader.
"""
import torch

from . import _lib
from ._lib import lib, ptr, check, stream


def get_ray_directions(H, W, K, device="cpu", flatten=True):
    """(H*W,3) directions ((u-cx+.5)/fx, (v-cy+.5)/fy, 1) of a pinhole camera K (3,3) (ray_utils.py:10-46)."""
    fx, fy, cx, cy = float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])
    v, u = torch.meshgrid(torch.arange(H, device=device, dtype=torch.float32),
                          torch.arange(W, device=device, dtype=torch.float32), indexing="ij")
    d = torch.stack([(u - cx + 0.5) / fx, (v - cy + 0.5) / fy, torch.ones_like(u)], -1)
    return d.reshape(-1, 3) if flatten else d


def _launch(directions, poses, img_idxs, pix_idxs, n):
    _lib.require_device()
    rays_o = torch.empty(n, 3, dtype=torch.float32, device=directions.device)
    rays_d = torch.empty_like(rays_o)
    check(lib.ngp_get_rays(ptr(directions), ptr(poses), ptr(img_idxs), ptr(pix_idxs), n, ptr(rays_o), ptr(rays_d), stream()),
          "get_rays")
    return rays_o, rays_d


def get_rays(directions, c2w):
    """rays_o, rays_d (N,3) in world coordinates (ray_utils.py:49-72)."""
    directions = directions.float().contiguous()
    n = directions.shape[0]
    c2w = c2w.float().contiguous()
    if c2w.ndim == 2:
        return _launch(directions, c2w, None, None, n)
    idx = torch.arange(n, device=directions.device)          # one pose per ray
    return _launch(directions, c2w, idx, None, n)


def get_rays_indexed(directions, poses, img_idxs, pix_idxs):
    """get_rays(directions[pix_idxs], poses[img_idxs]) without materialising either gather (train.py:136-156)."""
    return _launch(directions.float().contiguous(), poses.float().contiguous(), img_idxs.long().contiguous(),
                   pix_idxs.long().contiguous(), img_idxs.shape[0])
