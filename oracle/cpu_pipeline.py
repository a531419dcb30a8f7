"""TEST INFRASTRUCTURE ONLY — the hot path end to end on the HOST CPU, built from the oracle pieces:
AABB / marching / compositing / distortion from ngp_oracle.c (single thread, one serial loop per ray
like the reference's one CUDA thread per ray), hash grid + SH + MLPs from tcnn_oracle.py on CPU torch
tensors (intra-op threads = all host cores), torch autograd + Adam.

Used by bench.py for the `cpu_baseline` object and the `--impl reference` arm, and by
__graft_entry__.smoke() / tests as an end-to-end checker.  Two field shapes:
  'compact' — BASELINE.json configs[1] model (L16 F2 T2^19 grid, 64-wide MLPs, sigma=exp) so the CPU arm
              runs the SAME workload as the CUDA arm;
  'noCUDA'  — the reference's CPU-shaped path (BASELINE.json configs[0]: models/networks_noCUDA.py +
              models/rendering_noCUDA.py — coarse/fine stratified sampling, sample_pdf, raw2outputs),
              restated because the files themselves import tinycudann/vren and call .cuda()
              (SURVEY.md §0 finding 5).
"""
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import oracle, tcnn_oracle


class _CompositeFn(torch.autograd.Function):
    """oracle composite_train_fw/bw as an autograd op on CPU tensors (custom_functions.py:117-163)."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, ts, rays_a, T_thr):
        S = sigmas.shape[0]
        z3 = np.zeros((S, 3), np.float32); z0 = np.zeros((S, 0), np.float32)
        total, op, dep, rgb, nrm, sem, ws = oracle.composite_train_fw(sigmas.numpy(), rgbs.numpy(), z3, z0, deltas.numpy(),
                                                                      ts.numpy(), rays_a.numpy(), T_thr, 0)
        ctx.T_thr = T_thr
        t = torch.from_numpy
        ctx.save_for_backward(sigmas, rgbs, deltas, ts, rays_a, t(op), t(dep), t(rgb), t(ws))
        return t(op), t(dep), t(rgb), t(ws)

    @staticmethod
    def backward(ctx, g_op, g_dep, g_rgb, g_ws):
        sigmas, rgbs, deltas, ts, rays_a, op, dep, rgb, ws = ctx.saved_tensors
        S, R = sigmas.shape[0], rays_a.shape[0]
        z3 = np.zeros((S, 3), np.float32)
        dsig, drgb, _, _ = oracle.composite_train_bw(
            g_op.contiguous().numpy(), g_dep.contiguous().numpy(), g_rgb.contiguous().numpy(), np.zeros((R, 3), np.float32),
            np.zeros((R, 0), np.float32), g_ws.contiguous().numpy(), sigmas.numpy(), rgbs.numpy(), z3, ws.numpy(),
            deltas.numpy(), ts.numpy(), rays_a.numpy(), op.numpy(), dep.numpy(), rgb.numpy(), np.zeros((R, 3), np.float32),
            ctx.T_thr, 0)
        return torch.from_numpy(dsig), torch.from_numpy(drgb), None, None, None, None


class _DistortionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ws, deltas, ts, rays_a):
        loss, wi, wti = oracle.distortion_loss_fw(ws.numpy(), deltas.numpy(), ts.numpy(), rays_a.numpy())
        ctx.save_for_backward(torch.from_numpy(wi), torch.from_numpy(wti), ws, deltas, ts, rays_a)
        return torch.from_numpy(loss)

    @staticmethod
    def backward(ctx, g):
        wi, wti, ws, deltas, ts, rays_a = ctx.saved_tensors
        return torch.from_numpy(oracle.distortion_loss_bw(g.contiguous().numpy(), wi.numpy(), wti.numpy(), ws.numpy(),
                                                          deltas.numpy(), ts.numpy(), rays_a.numpy())), None, None, None


class CompactFieldCPU(torch.nn.Module):
    """NGPCompact on CPU: same parameter layout (flat tables / flat MLP params) as the CUDA modules."""

    def __init__(self, scale=0.5, L=16, Fe=2, log2_T=19, base=16, width=64, seed=1337):
        super().__init__()
        self.scale = scale
        self.grid = (L, Fe, log2_T, base, float(np.exp(np.log(2048 * scale / base) / (L - 1))))
        _, total = tcnn_oracle.grid_layout(*self.grid)
        g = torch.Generator().manual_seed(seed)
        self.table = torch.nn.Parameter((torch.rand(total * Fe, generator=g) * 2 - 1) * 1e-4)
        def xavier(shapes):
            return torch.cat([((torch.rand(o, i, generator=g) * 2 - 1) * math.sqrt(6 / (i + o))).reshape(-1) for o, i in shapes])
        self.width = width
        self.sigma_p = torch.nn.Parameter(xavier(tcnn_oracle.mlp_layer_shapes(L * Fe, width, 1, 16)))
        self.rgb_p = torch.nn.Parameter(xavier(tcnn_oracle.mlp_layer_shapes(32, width, 2, 3)))

    def forward(self, x, d):
        xn = (x + self.scale) / (2 * self.scale)
        enc = tcnn_oracle.grid_encode(xn, self.table, *self.grid)
        h = tcnn_oracle.mlp_forward(enc, self.sigma_p, enc.shape[1], self.width, 1, 16)
        sigmas = torch.exp(h[:, 0])
        dn = F.normalize(d, dim=-1, eps=1e-6)
        rgbs = tcnn_oracle.mlp_forward(torch.cat([tcnn_oracle.sh_encode((dn + 1) / 2, 4), h], 1), self.rgb_p, 32, self.width,
                                       2, 3, "ReLU", "Sigmoid")
        return sigmas, rgbs


class CPUPipeline:
    """train_step(rays_o, rays_d, rgb_gt) on numpy/torch CPU data, configs[1] workload."""

    def __init__(self, bitfield, scale=0.5, cascades=1, esf=0.0, lr=1e-2, threads=None, **field_kw):
        torch.set_num_threads(threads or os.cpu_count())
        self.threads = torch.get_num_threads()
        self.bitfield = np.ascontiguousarray(bitfield, np.uint8)
        self.scale, self.cascades, self.esf = scale, cascades, esf
        self.field = CompactFieldCPU(scale=scale, **field_kw)
        self.opt = torch.optim.Adam(self.field.parameters(), lr=lr, eps=1e-15)
        self.rng = np.random.RandomState(0)

    def render(self, rays_o, rays_d):
        cnt, ht, _ = oracle.ray_aabb_intersect(rays_o, rays_d, np.zeros((1, 3), np.float32), np.full((1, 3), self.scale, np.float32), 1)
        h = ht[:, 0, :].copy()
        m = (h[:, 0] >= 0) & (h[:, 0] < 0.01); h[m, 0] = 0.01
        noise = self.rng.rand(rays_o.shape[0]).astype(np.float32)
        ra, xyzs, dirs, deltas, ts, counter = oracle.raymarching_train(rays_o, rays_d, h, self.bitfield, self.cascades, self.scale,
                                                                       self.esf, noise, 128, 1024)
        t = torch.from_numpy
        sigmas, rgbs = self.field(t(xyzs), t(dirs))
        op, dep, rgb, ws = _CompositeFn.apply(sigmas, rgbs.contiguous(), t(deltas), t(ts), t(ra), 1e-4)
        return dict(opacity=op, depth=dep, rgb=rgb, ws=ws, deltas=t(deltas), ts=t(ts), rays_a=t(ra), total_samples=int(counter[0]))

    def train_step(self, rays_o, rays_d, rgb_gt):
        res = self.render(np.ascontiguousarray(rays_o, np.float32), np.ascontiguousarray(rays_d, np.float32))
        o = res["opacity"] + 1e-10
        loss = ((res["rgb"] - torch.from_numpy(np.ascontiguousarray(rgb_gt, np.float32))) ** 2).mean() \
            + 2e-4 * (-o * torch.log(o)).mean() \
            + 3e-4 * _DistortionFn.apply(res["ws"], res["deltas"], res["ts"], res["rays_a"]).mean()
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return float(loss.detach()), res["total_samples"]
