"""Procedural stand-ins for the datasets the reference trains on (none exist offline): a
"Lego-shaped" bounded object (NeRF-synthetic geometry: 800x800, 100 inward views on a sphere of
radius 1.5, scale 0.5 — datasets/nerf.py:26-27,51,60, opt.py:23) and a "street-shaped" unbounded
scene (scale 8, 5 cascades, forward-moving rig — configs/kitti360_1538.txt).  Rays follow the
reference's pinhole convention (datasets/ray_utils.py:8-74): camera-space direction
((u-cx+.5)/fx, (v-cy+.5)/fy, 1), NOT normalised, rotated by the pose; origin = camera centre.

Ground truth is analytic: the colour of the first box surface a ray hits (Lambert + a small
view-dependent term), black background with zero opacity — so PSNR-after-N-iterations is well
defined without any image files.  Everything is torch ops on the caller's device; this module is
data plumbing, not the hot path.
"""
import math

import numpy as np
import torch


def _lego_boxes():
    """~40 axis-aligned boxes inside [-0.35, 0.35]^3: base plate, body, cab, studs, boom, tracks."""
    B = []
    add = lambda c, h: B.append((c, h))
    add((0.0, 0.0, -0.20), (0.30, 0.22, 0.03))            # base plate
    add((-0.05, 0.0, -0.10), (0.20, 0.14, 0.07))          # body
    add((0.14, 0.0, -0.02), (0.07, 0.10, 0.10))           # cab
    add((-0.18, 0.0, 0.02), (0.05, 0.05, 0.12))           # exhaust tower
    for i in range(6):                                    # boom segments
        t = i / 5.0
        add((-0.10 + 0.30 * t, 0.0, 0.08 + 0.16 * t - 0.20 * t * t), (0.035, 0.025, 0.025))
    add((0.24, 0.0, 0.02), (0.05, 0.08, 0.03))            # bucket
    for sx in (-1, 1):                                    # tracks
        add((0.0, 0.19 * sx, -0.15), (0.26, 0.035, 0.05))
        for i in range(5):
            add((-0.20 + 0.10 * i, 0.19 * sx, -0.09), (0.03, 0.035, 0.012))
    for i in range(4):                                    # studs on the body
        for j in range(3):
            add((-0.17 + 0.08 * i, -0.08 + 0.08 * j, -0.015), (0.02, 0.02, 0.015))
    return B


def _street_boxes(rng):
    B = [((0.0, 0.0, -0.55), (7.5, 7.5, 0.05))]          # ground slab
    for _ in range(48):                                   # buildings / objects, denser near the centre
        r = abs(rng.normal()) * 2.5 + 0.3
        a = rng.uniform(0, 2 * math.pi)
        h = (rng.uniform(0.1, 0.6), rng.uniform(0.1, 0.6), rng.uniform(0.2, 1.2))
        B.append(((r * math.cos(a), r * math.sin(a), -0.5 + h[2]), h))
    return B


class BoxScene:
    def __init__(self, kind="lego", device="cpu", seed=20220806):
        rng = np.random.RandomState(seed)
        if kind == "lego":
            boxes, self.scale, self.exp_step_factor = _lego_boxes(), 0.5, 0.0
            self.img_wh, self.cam_radius = (800, 800), 1.5
            self.focal = 0.5 * 800 / math.tan(0.5 * 0.6911112070083618)   # datasets/nerf.py:26-27
        elif kind == "street":
            boxes, self.scale, self.exp_step_factor = _street_boxes(rng), 8.0, 1.0 / 256
            self.img_wh, self.cam_radius = (1408, 376), 3.0
            self.focal = 552.55
        else:
            raise ValueError(kind)
        self.kind = kind
        self.device = torch.device(device)
        self.centers = torch.tensor([b[0] for b in boxes], dtype=torch.float32, device=self.device)
        self.halves = torch.tensor([b[1] for b in boxes], dtype=torch.float32, device=self.device)
        self.albedo = torch.tensor(rng.uniform(0.25, 0.95, size=(len(boxes), 3)), dtype=torch.float32,
                                   device=self.device)
        self.semantic = torch.tensor(rng.randint(0, 7, size=len(boxes)), dtype=torch.int64, device=self.device)
        self.light = torch.nn.functional.normalize(torch.tensor([0.4, 0.3, 0.85], device=self.device), dim=0)
        self.cascades = max(1 + int(np.ceil(np.log2(2 * self.scale))), 1)

    # ------------------------------------------------------------------ cameras / rays
    def poses(self, n_views=100, seed=0):
        """(n,3,4) camera-to-world, looking at the origin (lego) or along +x on a lane (street)."""
        rng = np.random.RandomState(seed)
        P = []
        for i in range(n_views):
            if self.kind == "lego":
                th = rng.uniform(0, 2 * math.pi); ph = rng.uniform(math.radians(10), math.radians(80))
                c = self.cam_radius * np.array([math.cos(th) * math.sin(ph), math.sin(th) * math.sin(ph), math.cos(ph)])
                fwd = -c / np.linalg.norm(c)
            else:
                c = np.array([-3.0 + 6.0 * i / max(n_views - 1, 1), 0.15 * (1 if i % 2 else -1), -0.25])
                fwd = np.array([1.0, 0.0, 0.0])
            up = np.array([0.0, 0.0, 1.0])
            right = np.cross(fwd, up); right /= np.linalg.norm(right)
            down = np.cross(fwd, right)
            P.append(np.stack([right, down, fwd, c], 1))      # camera x=right, y=down, z=forward
        return torch.tensor(np.stack(P), dtype=torch.float32, device=self.device)

    def rays_from_pixels(self, poses, img_idx, u, v):
        W, H = self.img_wh
        dirs = torch.stack([(u - W / 2 + 0.5) / self.focal, (v - H / 2 + 0.5) / self.focal, torch.ones_like(u)], -1)
        R = poses[img_idx, :, :3]
        rays_d = torch.einsum("nij,nj->ni", R, dirs)
        rays_o = poses[img_idx, :, 3]
        return rays_o.contiguous(), rays_d.contiguous()

    def sample_rays(self, n, poses, generator=None):
        W, H = self.img_wh
        dev = self.device
        img_idx = torch.randint(poses.shape[0], (n,), device=dev, generator=generator)
        u = torch.randint(W, (n,), device=dev, generator=generator).float()
        v = torch.randint(H, (n,), device=dev, generator=generator).float()
        return self.rays_from_pixels(poses, img_idx, u, v)

    def image_rays(self, pose, wh=None):
        W, H = wh or self.img_wh
        s = self.img_wh[0] / W
        v, u = torch.meshgrid(torch.arange(H, device=self.device).float(), torch.arange(W, device=self.device).float(),
                              indexing="ij")
        idx = torch.zeros(H * W, dtype=torch.long, device=self.device)
        return self.rays_from_pixels(pose[None], idx, u.reshape(-1) * s + (s - 1) / 2, v.reshape(-1) * s + (s - 1) / 2)

    # ------------------------------------------------------------------ analytic ground truth
    @torch.no_grad()
    def shade(self, rays_o, rays_d, chunk=1 << 18):
        """-> rgb (N,3), opacity (N), depth t (N), label (N) of the first surface hit."""
        outs = []
        for i in range(0, rays_o.shape[0], chunk):
            o, d = rays_o[i:i + chunk, None], rays_d[i:i + chunk, None]
            inv = 1.0 / d
            t0 = (self.centers - self.halves - o) * inv
            t1 = (self.centers + self.halves - o) * inv
            tn = torch.minimum(t0, t1); tf = torch.maximum(t0, t1)
            tnear, axis = tn.max(-1)
            tfar = tf.min(-1).values
            hit = (tnear <= tfar) & (tfar > 0) & (tnear > 0)
            tnear = torch.where(hit, tnear, torch.full_like(tnear, float("inf")))
            t, b = tnear.min(-1)
            any_hit = torch.isfinite(t)
            ax = axis.gather(1, b[:, None])[:, 0]
            dd = rays_d[i:i + chunk]
            n = torch.zeros_like(dd)
            n.scatter_(1, ax[:, None], -torch.sign(dd.gather(1, ax[:, None])))
            lambert = (n * self.light).sum(-1).clamp(min=0) * 0.6 + 0.4
            vdir = torch.nn.functional.normalize(dd, dim=-1)
            spec = ((n * -vdir).sum(-1).clamp(min=0) ** 8) * 0.15
            rgb = (self.albedo[b] * lambert[:, None] + spec[:, None]).clamp(0, 1)
            rgb = torch.where(any_hit[:, None], rgb, torch.zeros_like(rgb))
            outs.append((rgb, any_hit.float(), torch.where(any_hit, t, torch.zeros_like(t)),
                         torch.where(any_hit, self.semantic[b], torch.full_like(b, 4))))
        return tuple(torch.cat(x) for x in zip(*outs))


def morton_encode(x, y, z):
    """30-bit Morton interleave of int64 tensors (torch ops; matches raymarching.cu:35-50)."""
    def expand(v):
        v = (v * 0x00010001) & 0xFF0000FF
        v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3
        v = (v * 0x00000005) & 0x49249249
        return v
    return expand(x) | (expand(y) << 1) | (expand(z) << 2)


@torch.no_grad()
def scene_density_grid(scene: BoxScene, grid_size=128, dilate_cells=1.0, value=100.0):
    """(cascades, G^3) float32 grid in Morton order: `value` in cells overlapping a dilated box, 0
    elsewhere.  Cascade c covers [-s, s]^3 with s = min(2^(c-1), scale) (networks.py:388-392)."""
    G, dev = grid_size, scene.device
    ax = torch.arange(G, device=dev)
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    coords = torch.stack([X, Y, Z], -1).reshape(-1, 3)
    idx = morton_encode(coords[:, 0].long(), coords[:, 1].long(), coords[:, 2].long())
    grids = []
    for c in range(scene.cascades):
        s = min(2.0 ** (c - 1), scene.scale)
        cell = 2 * s / G
        ctr = (coords.float() + 0.5) / G * 2 * s - s
        occ = torch.zeros(G ** 3, dtype=torch.bool, device=dev)
        pad = cell * (0.5 + dilate_cells)
        for b in range(scene.centers.shape[0]):
            occ |= ((ctr - scene.centers[b]).abs() <= scene.halves[b] + pad).all(-1)
        g = torch.zeros(G ** 3, device=dev)
        g[idx] = occ.float() * value
        grids.append(g)
    return torch.stack(grids)


def pack_bitfield_torch(density_grid, thr):
    """torch restatement of packbits for host-side scene setup on CPU (bit i of byte n <-> cell 8n+i)."""
    bits = (density_grid.reshape(-1, 8) > thr).to(torch.uint8)
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=density_grid.device)
    return (bits * w).sum(-1).to(torch.uint8)
