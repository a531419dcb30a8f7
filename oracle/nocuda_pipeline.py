"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's "noCUDA" path, BASELINE.json configs[0]:
models/rendering_noCUDA.py:103-214 (coarse -> fine sampling, `raw2outputs` compositing) over two
models/networks_noCUDA.py:49-369 fields, on host tensors.

Why a restatement: the two files cannot run on a CPU — they `import tinycudann`, `import vren`, call `.cuda()` on every
temporary and `from .rendering_old import NEAR_DISTANCE`, a module that does not exist (networks_noCUDA.py:5-10;
rendering_noCUDA.py:7,134,139) — and nothing in the tree calls them.  "noCUDA" means "no custom
ray-marching kernel": vanilla-NeRF stratified sampling + `sample_pdf` + `raw2outputs` in torch ops.  Everything here is
device-agnostic torch ops (tcnn operators from oracle/tcnn_oracle.py), so the same code is timed on the host cores
(bench.py `cpu_baseline` / `--impl reference`) and checked on the GPU against the reference's own `sample_pdf`,
`raw2outputs` and `rendering_noCUDA.render` (tests/test_nocuda_gpu.py) — that is what pins this file.

Stated assumptions where the reference leaves a hole:
  * `samples=[64,128]`: kwargs['samples'] has no caller in the tree (rendering_noCUDA.py:116); NeRF's standard counts.
  * coarse model: rendering_noCUDA.py:177 unpacks THREE outputs `sigmas, rgbs, sems = model(...)` while
    networks_noCUDA.NGP.forward returns six — the coarse class is not in the tree; here it is the same field evaluated
    without normals (density, colour, semantics), i.e. forward_test's outputs (networks_noCUDA.py:327-363).
  * loss: train.py never drives this path; rgb MSE of both levels + the opacity entropy and Rp terms of losses.py:94-109.
  * rays that miss the box have near = far = -1 and produce 0/0 in the exp-warp (rendering_noCUDA.py:143-144); the
    workload draws rays that hit the box.
"""
import math
import os

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import tcnn_oracle

NEAR_DISTANCE = 0.01      # rendering_noCUDA.py:11


def sample_pdf(bins, weights, n_samples):
    """models/custom_functions.py:248-278 with det=True (the only mode rendering_noCUDA.py:126 uses)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.linspace(0., 1., steps=n_samples, device=bins.device).expand(list(cdf.shape[:-1]) + [n_samples]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    inds_g = torch.stack([below, above], -1)
    shape = [inds_g.shape[0], inds_g.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(shape), 2, inds_g)
    bins_g = torch.gather(bins.unsqueeze(1).expand(shape), 2, inds_g)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    return bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])


def raw2outputs(raw, z_vals, rays_d):
    """models/custom_functions.py:280-321: raw (R,S,10+C) = [sigma, rgb(3), normal_raw(3), normal_pred(3), sem(C)]."""
    sigmas, rgbs, n_raw, n_pred, sems = raw[..., 0], raw[..., 1:4], raw[..., 4:7], raw[..., 7:10], raw[..., 10:]
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    alpha = 1. - torch.exp(-sigmas * dists)
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    w = alpha * T
    return (torch.sum(w, -1), torch.sum(w[..., None] * rgbs, -2), torch.sum(w[..., None] * n_raw, -2),
            torch.sum(w[..., None] * n_pred, -2), torch.sum(w[..., None] * sems, -2), w, torch.sum(w * z_vals, -1))


_LAYOUTS = {}


def grid_encode_fast(x, table, n_levels, n_features, log2_T, base_res, per_level_scale):
    """tcnn_oracle.grid_encode (same layout, hash and weights) with all levels and corners in ONE gather, so that the
    table gradient is one index_add into one table-sized buffer instead of 8*L of them (the per-corner indexing of the
    oracle allocates a table-sized zero tensor per corner and level in backward: 90 s per step on the host at this
    shape; checked equal to grid_encode in tests/test_oracle_cpu.py)."""
    key = (n_levels, n_features, log2_T, base_res, per_level_scale, x.device)
    if key not in _LAYOUTS:
        levels, total = tcnn_oracle.grid_layout(n_levels, n_features, log2_T, base_res, per_level_scale)
        t = lambda k, dt: torch.tensor([lv[k] for lv in levels], dtype=dt, device=x.device)
        offs = torch.tensor([[(k >> d) & 1 for d in range(3)] for k in range(8)], dtype=torch.int64, device=x.device)
        _LAYOUTS[key] = (t("scale", torch.float32), t("res", torch.int64), t("size", torch.int64), t("offset", torch.int64),
                         t("dense", torch.bool), offs, total)
    scale, res, size, offset, dense, offs, total = _LAYOUTS[key]
    pos = x[:, None, :] * scale.to(x.dtype)[None, :, None] + 0.5                     # (N,L,3)
    cell = torch.floor(pos)
    w = pos - cell
    c = cell.detach().to(torch.int64)[:, :, None, :] + offs[None, None]              # (N,L,8,3)
    r = res[None, :, None]
    m = 0xFFFFFFFF
    i_dense = c[..., 0] + c[..., 1] * r + c[..., 2] * r * r
    i_hash = ((c[..., 0] * tcnn_oracle.PRIMES[0]) & m) ^ ((c[..., 1] * tcnn_oracle.PRIMES[1]) & m) ^ ((c[..., 2] * tcnn_oracle.PRIMES[2]) & m)
    idx = torch.where(dense[None, :, None], i_dense, i_hash) % size[None, :, None] + offset[None, :, None]
    wk = torch.where(offs.bool()[None, None], w[:, :, None, :], 1 - w[:, :, None, :]).prod(-1)           # (N,L,8)
    vals = table.view(total, n_features).index_select(0, idx.reshape(-1)).view(x.shape[0], n_levels, 8, n_features)
    return (wk[..., None] * vals).sum(2).reshape(x.shape[0], n_levels * n_features)


class NGPNoCUDA(nn.Module):
    """networks_noCUDA.py:49-369: xyz grid L16 F2 T2^19 -> Linear(32,128) Softplus Linear(128,1) -> Softplus(beta=100);
    rgb grid L32 F2 T2^21; rgb_net([x, SH4(d), feat_rgb(, embed_a)]) 128-wide; 32-wide normal / semantic heads."""

    def __init__(self, scale=0.5, embed_a_len=0, classes=7, seed=1337, log2_T_xyz=19, log2_T_rgb=21):
        super().__init__()
        self.scale, self.classes, self.embed_a_len = scale, classes, embed_a_len
        self.g_xyz = (16, 2, log2_T_xyz, 16, float(np.exp(np.log(2048 * scale / 16) / 15)))          # :71-72
        self.g_rgb = (32, 2, log2_T_rgb, 16, float(np.exp(np.log(2048 * scale / 16) / 31)))          # :96-97
        g = torch.Generator().manual_seed(seed)
        table = lambda cfg: nn.Parameter((torch.rand(tcnn_oracle.grid_layout(*cfg)[1] * cfg[1], generator=g) * 2 - 1) * 1e-4)
        self.xyz_table, self.rgb_table = table(self.g_xyz), table(self.g_rgb)
        self.xyz_net = nn.Sequential(nn.Linear(32, 128), nn.Softplus(), nn.Linear(128, 1))                # :89-93
        xav = lambda shapes: nn.Parameter(torch.cat([((torch.rand(o, i, generator=g) * 2 - 1) * math.sqrt(6 / (i + o))).reshape(-1)
                                                     for o, i in shapes]))
        self.rgb_in = 64 + 16 + 3 + embed_a_len                                                            # :138
        self.rgb_p = xav(tcnn_oracle.mlp_layer_shapes(self.rgb_in, 128, 1, 3))
        self.norm_p = xav(tcnn_oracle.mlp_layer_shapes(64, 32, 1, 3))
        self.sem_p = xav(tcnn_oracle.mlp_layer_shapes(64, 32, 1, classes))

    def density(self, x, return_feat=False):                                                               # :217-237
        xn = (x + self.scale) / (2 * self.scale)
        h = self.xyz_net(grid_encode_fast(xn, self.xyz_table, *self.g_xyz))
        sigmas = F.softplus(h[:, 0], beta=100)
        if return_feat:
            return sigmas, grid_encode_fast(xn, self.rgb_table, *self.g_rgb)
        return sigmas

    def _heads(self, x, d, feat, embed_a):
        n_pred = -F.normalize(tcnn_oracle.mlp_forward(feat, self.norm_p, 64, 32, 1, 3), p=2, dim=-1, eps=1e-6)
        sem = torch.softmax(tcnn_oracle.mlp_forward(feat, self.sem_p, 64, 32, 1, self.classes), -1)
        dn = F.normalize(d, p=2, dim=-1, eps=1e-6)
        inp = [x, tcnn_oracle.sh_encode((dn + 1) / 2, 4), feat] + ([embed_a] if self.embed_a_len else [])
        rgb = tcnn_oracle.mlp_forward(torch.cat(inp, 1), self.rgb_p, self.rgb_in, 128, 1, 3, "ReLU", "Sigmoid")
        return rgb, n_pred, sem

    def forward(self, x, d, embed_a=None):                                                                 # :264-325
        with torch.enable_grad():
            x = x.requires_grad_(True)
            sigmas, feat = self.density(x, return_feat=True)
            (grads,) = torch.autograd.grad(sigmas, x, torch.ones_like(sigmas), retain_graph=True)         # :255-262
        n_raw = -F.normalize(grads.detach(), p=2, dim=-1, eps=1e-6)
        rgb, n_pred, sem = self._heads(x, d, feat, embed_a)
        return sigmas, rgb, n_raw, n_pred, sem

    def forward_coarse(self, x, d, embed_a=None):
        """density, colour, semantics without normals (see the module docstring)."""
        sigmas, feat = self.density(x, return_feat=True)
        rgb, _, sem = self._heads(x, d, feat, embed_a)
        return sigmas, rgb, sem


def render_rays_train(models, rays_o, rays_d, hits_t, samples=(64, 128), classes=7, embedding_a=(None, None), t_rand=None):
    """rendering_noCUDA.py:103-214 (`__render_rays_train`), bg = zeros.  hits_t (R,1,2)."""
    res, R = {"total_samples": 0}, rays_o.shape[0]
    for i, S in enumerate(samples):
        model = models[0] if i < len(samples) - 1 else models[1]
        if i > 0:                                                                                          # :124-131
            zp = res[f"z_vals{i - 1}"]
            mids = .5 * (zp[:, 1:] + zp[:, :-1])
            z_vals = sample_pdf(mids.detach(), res[f"ws{i - 1}"][..., 1:-1].detach(), S)
            z_vals, _ = torch.sort(z_vals, -1)
            z_vals = z_vals.detach()
        else:                                                                                              # :132-149
            t_vals = torch.linspace(0., 1. - 1e-3, steps=S, device=rays_o.device)
            near = hits_t[:, 0, 0].unsqueeze(-1).repeat(1, S)
            far = hits_t[:, 0, 1].unsqueeze(-1).repeat(1, S)
            t_vals = t_vals.unsqueeze(0).repeat(R, 1)
            tr = 5e-5 * (torch.rand(R, device=rays_o.device) if t_rand is None else t_rand)
            t_vals = t_vals + tr[:, None]
            z_vals = near * (1. - t_vals) + far * t_vals
            ef = 1. + 1 / 16
            tp = (ef ** (z_vals - near) - 1.) / (ef ** (far - near) - 1.)
            z_vals = near * (1. - tp) + far * tp
            z_vals = z_vals + tr[:, None]
        xyzs = (rays_o[:, None, :] + rays_d[:, None, :] * z_vals[:, :, None]).reshape(-1, 3).detach()
        dirs = rays_d.unsqueeze(1).repeat(1, S, 1).reshape(-1, 3).detach()
        res["total_samples"] += S
        emb = torch.repeat_interleave(embedding_a[i], S, 0) if embedding_a[i] is not None else None         # :163
        if i == len(samples) - 1:
            sigmas, rgbs, n_raw, n_pred, sems = model(xyzs, dirs, emb)
        else:
            sigmas, rgbs, sems = model.forward_coarse(xyzs, dirs, emb)
            n_raw = torch.zeros_like(rgbs); n_pred = torch.zeros_like(rgbs)
        n_raw = n_raw.detach()
        raw = torch.cat([sigmas.reshape(R, S, 1), rgbs.reshape(R, S, 3), n_raw.reshape(R, S, 3), n_pred.reshape(R, S, 3),
                         sems.reshape(R, S, classes)], -1)
        (res[f"opacity{i}"], res[f"rgb{i}"], res[f"normal_raw{i}"], res[f"normal_pred{i}"], res[f"semantic{i}"], res[f"ws{i}"],
         res[f"depth{i}"]) = raw2outputs(raw, z_vals, rays_d)
        res[f"z_vals{i}"] = z_vals
        nd = torch.sum((n_raw - n_pred) ** 2, dim=-1)
        res[f"Rp{i}"] = (nd.reshape(R, S) * res[f"ws{i}"]).reshape(-1)                                     # :192-194
    return res


def aabb_hits(rays_o, rays_d, scale):
    """intersection.cu:5-22 + rendering.py:28-30 through the C oracle -> hits_t (R,1,2) torch, hit mask."""
    from . import oracle
    o, d = np.ascontiguousarray(rays_o.cpu().numpy(), np.float32), np.ascontiguousarray(rays_d.cpu().numpy(), np.float32)
    _, ht, _ = oracle.ray_aabb_intersect(o, d, np.zeros((1, 3), np.float32), np.full((1, 3), scale, np.float32), 1)
    ht = ht.copy()
    m = (ht[:, 0, 0] >= 0) & (ht[:, 0, 0] < NEAR_DISTANCE)
    ht[m, 0, 0] = NEAR_DISTANCE
    return torch.from_numpy(ht), torch.from_numpy(ht[:, 0, 1] > 0)


class NoCUDAPipeline:
    """train_step on CPU tensors: configs[0] (Lego-shaped, scale 0.5, 8192-ray batches, samples [64,128])."""

    def __init__(self, scale=0.5, lr=1e-2, threads=None, samples=(64, 128), classes=7, **model_kw):
        torch.set_num_threads(threads or os.cpu_count())
        self.threads = torch.get_num_threads()
        self.scale, self.samples, self.classes = scale, tuple(samples), classes
        self.models = [NGPNoCUDA(scale, classes=classes, seed=1337, **model_kw), NGPNoCUDA(scale, classes=classes, seed=1338, **model_kw)]
        self.opt = torch.optim.Adam([p for m in self.models for p in m.parameters()], lr=lr, eps=1e-8)

    def train_step(self, rays_o, rays_d, rgb_gt):
        hits_t, hit = aabb_hits(rays_o, rays_d, self.scale)
        assert bool(hit.all()), "the noCUDA restatement takes rays that hit the box (module docstring)"
        res = render_rays_train(self.models, rays_o, rays_d, hits_t, self.samples, self.classes)
        last = len(self.samples) - 1
        o = res[f"opacity{last}"] + 1e-10
        loss = sum(((res[f"rgb{i}"] - rgb_gt) ** 2).mean() for i in range(len(self.samples))) \
            + 2e-4 * (-o * torch.log(o)).mean() + 1e-3 * res[f"Rp{last}"].mean()
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return float(loss.detach()), res["total_samples"] * rays_o.shape[0]
