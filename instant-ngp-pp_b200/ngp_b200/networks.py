"""Field models of the hot path.

NGP         — the reference's literal model (models/networks.py:13-425): two F=8 hash grids
              (xyz T=2^19, rgb T=2^21), torch density MLP 128->128->1 with Softplus (double
              backward for autograd normals), bias-free heads rgb_net / norm_pred_header /
              semantic_header, optional skybox; same attribute and state-dict key names
              (xyz_encoder.params, xyz_net.0.weight, rgb_net.params, density_bitfield, ...).
NGPCompact  — the ngp_pl-shaped model BASELINE.json's headline config names ("16-level hash
              grid, 2^19 table, 64-wide MLP"): one F=2 grid -> 64-wide MLP -> 16 features,
              sigma = exp(h0); rgb = MLP(SH4(d) | h).

Both expose density / forward / forward_test / update_density_grid with the reference's
argument meaning.  What changed relative to the reference's glue:
  * the input gradient of the density (normals) is obtained with ONE extra input-gradient kernel
    of the xyz grid instead of a generic autograd.grad through both encoders — the rgb grid
    never computes dy/dx (the reference does and discards it);
  * SH + concatenation are assembled inside the MLP kernel (no dir_encoder / torch.cat pass);
  * update_density_grid has no host sync: the mean density is reduced on the device and the
    packbits threshold is read by the kernel's caller from a 0-dim tensor only once per update.
"""
import math
import os

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import tcnn, vren
from .custom_functions import TruncExp
from .tcnn import _GridBwFn

NEAR_DISTANCE = 0.01   # models/rendering.py:10


def _cascades(scale):
    return max(1 + int(np.ceil(np.log2(2 * scale))), 1)     # networks.py:29


class _OccupancyMixin:
    """Occupancy-grid maintenance shared by both models (networks.py:284-408)."""

    grid_size = 128

    def _init_occupancy(self, scale):
        self.scale = scale
        self.cascades = _cascades(scale)
        G = self.grid_size
        self.register_buffer("center", torch.zeros(1, 3))
        self.register_buffer("xyz_min", -torch.ones(1, 3) * scale)
        self.register_buffer("xyz_max", torch.ones(1, 3) * scale)
        self.register_buffer("half_size", (self.xyz_max - self.xyz_min) / 2)
        self.register_buffer("density_bitfield", torch.zeros(self.cascades * G ** 3 // 8, dtype=torch.uint8))
        # the reference registers these two from train.py:128-132
        self.register_buffer("density_grid", torch.zeros(self.cascades, G ** 3))
        ax = torch.arange(G, dtype=torch.int32)
        self.register_buffer("grid_coords", torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3))

    def aabb(self):
        """(lo xyz, range xyz) as python floats for the fused normalisation of the grid kernels; read from the
        xyz_min / xyz_max buffers once (they are fixed by `scale`, networks.py:23-27)."""
        cached = getattr(self, "_aabb_host", None)
        if cached is None:
            lo = self.xyz_min.detach().float().cpu().reshape(3)
            rng = (self.xyz_max - self.xyz_min).detach().float().cpu().reshape(3)
            cached = self._aabb_host = tuple(float(v) for v in lo) + tuple(float(v) for v in rng)
        return cached

    @torch.no_grad()
    def get_all_cells(self):
        """[(indices, coords)] * cascades — every cell of the 128^3 lattice (networks.py:294-305)."""
        indices = vren.morton3D(self.grid_coords).long()
        return [(indices, self.grid_coords)] * self.cascades

    @torch.no_grad()
    def sample_uniform_and_occupied_cells(self, M, density_threshold):
        """M uniform + M occupied cells per cascade (networks.py:308-333)."""
        G, dev = self.grid_size, self.density_grid.device
        cells = []
        for c in range(self.cascades):
            coords1 = torch.randint(G, (M, 3), dtype=torch.int32, device=dev)
            indices1 = vren.morton3D(coords1).long()
            # M draws (with replacement) among the cells above the threshold — the reference's
            # nonzero()[randint(len)] without the host round trip: rank r of the draw -> r-th occupied cell through
            # a running count.  No occupied cell -> the reference keeps an empty list; here the M draws all land on
            # the last cell, which just gets one more (legitimate) density sample.
            occ = self.density_grid[c] > density_threshold
            count = torch.cumsum(occ, 0, dtype=torch.int32)
            total = count[-1]
            r = (torch.rand(M, device=dev) * total).to(torch.int32).clamp_(max=(total - 1).clamp(min=0))
            indices2 = torch.searchsorted(count, r + 1).clamp_(max=G ** 3 - 1)
            coords2 = vren.morton3D_invert(indices2.int())
            cells.append((torch.cat([indices1, indices2]), torch.cat([coords1, coords2])))
        return cells

    @torch.no_grad()
    def mark_invisible_cells(self, K, poses, img_wh, chunk=64 ** 3):
        """density_grid = -1 for cells no camera sees (networks.py:336-376)."""
        N_cams = poses.shape[0]
        self.count_grid = torch.zeros_like(self.density_grid)
        w2c_R = poses[:, :3, :3].transpose(1, 2)
        w2c_T = -w2c_R @ poses[:, :3, 3:]
        cells = self.get_all_cells()
        G = self.grid_size
        for c in range(self.cascades):
            indices, coords = cells[c]
            s = min(2 ** (c - 1), self.scale)
            hgs = s / G
            for i in range(0, len(indices), chunk):
                xyzs = coords[i:i + chunk] / (G - 1) * 2 - 1
                xyzs_w = (xyzs * (s - hgs)).T
                uvd = K @ (w2c_R @ xyzs_w + w2c_T)
                uv = uvd[:, :2] / uvd[:, 2:]
                in_img = (uvd[:, 2] >= 0) & (uv[:, 0] >= 0) & (uv[:, 0] < img_wh[0]) & (uv[:, 1] >= 0) & (uv[:, 1] < img_wh[1])
                covered = (uvd[:, 2] >= NEAR_DISTANCE) & in_img
                count = covered.sum(0) / N_cams
                self.count_grid[c, indices[i:i + chunk]] = count
                too_near = ((uvd[:, 2] < NEAR_DISTANCE) & in_img).any(0)
                valid = (count > 0) & (~too_near)
                self.density_grid[c, indices[i:i + chunk]] = torch.where(valid, 0., -1.)

    fused_update = True       # occupancy update as one kernel chain (csrc/occupancy.cu: ngp_occupancy_sample / _update)

    def _occ_workspace(self):
        from ._lib import lib
        ws = getattr(self, "_occ_ws", None)
        if ws is None or ws.device != self.density_grid.device:
            ws = self._occ_ws = torch.zeros(int(lib.ngp_occupancy_workspace_bytes(self.cascades)), dtype=torch.uint8,
                                            device=self.density_grid.device)
        return ws

    @torch.no_grad()
    def sample_cells_fused(self, density_threshold, warmup=False, seed=None):
        """-> (indices (Cc*n) i32 with -1 = void draw, xyzs (Cc*n, 3), n): the cells of sample_uniform_and_occupied_cells /
        get_all_cells for ALL cascades with their jittered positions, from one 3-kernel chain (no nonzero, no host sync).
        The random stream is counter based: `seed` (default: one draw from torch's CPU generator, so torch.manual_seed
        fixes it) and the draw's index determine every sample."""
        from ._lib import lib, ptr, check, stream
        G, Cc, dev = self.grid_size, self.cascades, self.density_grid.device
        M = G ** 3 // 4
        n = G ** 3 if warmup else 2 * M
        if seed is None:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
        indices = torch.empty(Cc * n, dtype=torch.int32, device=dev)
        xyzs = torch.empty(Cc * n, 3, device=dev)
        check(lib.ngp_occupancy_sample(ptr(self.density_grid), Cc, float(self.scale), float(density_threshold), M, int(bool(warmup)),
                                       seed & 0xFFFFFFFFFFFFFFFF, ptr(self._occ_workspace()), ptr(indices), ptr(xyzs), stream()),
              "occupancy_sample/warmup" if warmup else "occupancy_sample")
        return indices, xyzs, n

    @torch.no_grad()
    def update_density_grid(self, density_threshold, warmup=False, decay=0.95, erode=False, seed=None):
        """EMA-max update of the occupancy grid + repack of the bitfield (networks.py:379-408)."""
        if self.fused_update and self.density_grid.is_cuda and self.density_grid.dtype == torch.float32:
            from ._lib import lib, ptr, check, stream
            indices, xyzs, n = self.sample_cells_fused(density_threshold, warmup, seed)
            sig = self.density(xyzs).float().contiguous()
            cg = self.count_grid.contiguous() if erode else None
            check(lib.ngp_occupancy_update(ptr(self.density_grid), self.cascades, ptr(indices), ptr(sig), n, float(decay), ptr(cg),
                                           float(density_threshold), ptr(self._occ_workspace()), ptr(self.density_bitfield), stream()),
                  "occupancy_update")
            return
        G = self.grid_size
        tmp = torch.zeros_like(self.density_grid)
        cells = self.get_all_cells() if warmup else self.sample_uniform_and_occupied_cells(G ** 3 // 4, density_threshold)
        for c in range(self.cascades):
            indices, coords = cells[c]
            s = min(2 ** (c - 1), self.scale)
            hgs = s / G
            xyzs_w = (coords / (G - 1) * 2 - 1) * (s - hgs)
            xyzs_w += (torch.rand_like(xyzs_w) * 2 - 1) * hgs
            tmp[c, indices] = self.density(xyzs_w)
        if erode:
            decay = torch.clamp(decay ** (1 / self.count_grid), 0.1, 0.95)
        self.density_grid = torch.where(self.density_grid < 0, self.density_grid,
                                        torch.maximum(self.density_grid * decay, tmp))
        pos = self.density_grid > 0
        mean_density = (self.density_grid * pos).sum() / pos.sum().clamp(min=1)
        thr = torch.clamp(mean_density, max=density_threshold).float()     # min(mean, thr): stays on the device
        vren.packbits_dthr(self.density_grid, thr, self.density_bitfield)


class _tf32_matmul:
    """`with _tf32_matmul(on):` — fp32 GEMMs of the enclosed torch calls run on the tensor cores as TF32; the process-wide
    torch switch is restored on exit, so nothing outside the density net changes precision."""

    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        if self.on:
            torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


def _dn_fw(e, W1, b1, W2, b2, want_ge=True):
    """The density net on tcgen05 (csrc/density_net.cu): e (N,128) -> sigma (N), s2 (N), g_e (N,128) | None."""
    from ._lib import lib, ptr, check, stream
    n = e.shape[0]
    sigma = torch.empty(n, device=e.device); s2 = torch.empty(n, device=e.device)
    g_e = torch.empty(n, W1.shape[1], device=e.device) if want_ge else None
    check(lib.ngp_density_net_fw(ptr(e), ptr(W1.detach().contiguous()), ptr(b1.detach().contiguous()), ptr(W2.detach().reshape(-1).contiguous()),
                                 ptr(b2.detach().contiguous()), n, e.shape[1], W1.shape[0], ptr(sigma), ptr(s2), ptr(g_e), stream()),
          "density_net_fw")
    return sigma, s2, g_e


def _dn_bw(e, d_ge, g_e, dsigma, s2, W1, b1, W2, need_de=True):
    """-> de (N,128) | None, dW1 (128,128), db1 (128), dW2 (1,128), db2 (1) for upstream (dsigma | None, d_ge | None)."""
    from ._lib import lib, ptr, check, stream
    n, W = e.shape[0], W1.shape[0]
    de = torch.empty_like(e) if need_de else None
    acc = torch.zeros(W * W + 2 * W + 4, device=e.device)         # one memset for all four parameter gradients
    dW1, db1, dw2, db2 = acc[:W * W].view(W, W), acc[W * W:W * W + W], acc[W * W + W:W * W + 2 * W], acc[W * W + 2 * W:W * W + 2 * W + 1]
    check(lib.ngp_density_net_bw(ptr(e), ptr(d_ge), ptr(g_e) if d_ge is not None else None, ptr(dsigma), ptr(s2), ptr(W1.detach().contiguous()),
                                 ptr(b1.detach().contiguous()), ptr(W2.detach().reshape(-1).contiguous()), n, e.shape[1], W,
                                 ptr(de), ptr(dW1), ptr(db1), ptr(dw2), ptr(db2), stream()), "density_net_bw")
    return de, dW1, db1, dw2[None], db2


class _DensityNormalsFn(torch.autograd.Function):
    """(sigma (N), g_e (N,D) = d sigma / d e) of the reference's torch density net
        sigma = Softplus(Linear(W,1)(Softplus(Linear(D,W)(e))))                       (networks.py:54-59,172-181)
    as ONE autograd node whose backward covers both outputs — what the reference obtains from
    torch.autograd.grad(sigmas, x, create_graph=True) + loss.backward() (networks.py:186-196) through ~30 elementwise
    torch kernels on (N,W) tensors.  Here: one fused elementwise kernel per direction (csrc/density_head.cu, formulas
    there) between the GEMMs with W1, which stay on cuBLAS."""

    @staticmethod
    def forward(ctx, e, W1, b1, W2, b2, tf32=True):
        from . import _lib
        from ._lib import lib, ptr, check, stream
        _lib.require_device()
        e = e.contiguous()
        n, W = e.shape[0], W1.shape[0]
        ctx.tf32 = tf32
        ctx.set_materialize_grads(False)
        if tf32 == "tc05":                          # one tcgen05 kernel per direction (csrc/density_net.cu)
            sigma, s2, g_e = _dn_fw(e, W1, b1, W2, b2)
            ctx.save_for_backward(e, W1, b1, W2, s2, g_e)
            return sigma, g_e
        with _tf32_matmul(tf32):
            z1 = torch.addmm(b1, e, W1.t())
        sigma = torch.empty(n, device=e.device); s2 = torch.empty(n, device=e.device)
        t = torch.empty(n, W, device=e.device)
        w2 = W2.reshape(-1).contiguous()
        check(lib.ngp_density_head_fw(ptr(z1), ptr(w2), ptr(b2.contiguous()), n, W, ptr(sigma), ptr(s2), ptr(t), stream()),
              "density_head_fw")
        with _tf32_matmul(tf32):
            g_e = t @ W1
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(e, W1, w2, z1, s2, t)
        return sigma, g_e

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dsigma, dg):
        from ._lib import lib, ptr, check, stream
        if ctx.tf32 == "tc05":
            e, W1, b1, W2, s2, g_e = ctx.saved_tensors
            de, dW1, db1, dW2, db2 = _dn_bw(e, dg.contiguous() if dg is not None else None, g_e,
                                            dsigma.contiguous() if dsigma is not None else None, s2, W1, b1, W2, ctx.needs_input_grad[0])
            return de, dW1, db1, dW2, db2, None
        e, W1, w2, z1, s2, t = ctx.saved_tensors
        n, W = z1.shape
        with _tf32_matmul(ctx.tf32):
            v = (dg.contiguous() @ W1.t()) if dg is not None else None
        dz1 = torch.empty_like(z1); dz2 = torch.empty(n, device=z1.device)
        dw2 = torch.zeros(W, device=z1.device); db1 = torch.zeros(W, device=z1.device)
        check(lib.ngp_density_head_bw(ptr(z1), ptr(v), ptr(s2), ptr(dsigma.contiguous() if dsigma is not None else None), ptr(w2), n, W,
                                      ptr(dz1), ptr(dz2), ptr(dw2), ptr(db1), stream()), "density_head_bw")
        with _tf32_matmul(ctx.tf32):
            de = dz1 @ W1 if ctx.needs_input_grad[0] else None
            dW1 = dz1.t() @ e
            if dg is not None:
                dW1.addmm_(t.t(), dg)
        return de, dW1, db1, dw2[None], dz2.sum()[None], None


class _DensityFieldNormalsFn(torch.autograd.Function):
    """x (N,3) world -> (sigma (N), d sigma / d x_unit (N,3)): hash-grid encoding, density net and the analytic input
    gradient (models/networks.py:186-196 `grad`: encode, xyz_net, sigma_act, autograd.grad(create_graph=True)) as ONE
    autograd node.  Forward = gather, GEMM, density_head_fw, GEMM, input-gradient kernel.  Backward, given the upstream
    of both outputs: double-backward gather (d/d g_e), density_head_bw between the W1 GEMMs, and ONE dual scatter for
    both terms of the table gradient (ngp_hashgrid_bw_params_dual: the first- and second-order scatters hit the same
    corners) instead of two.  The result is d sigma / d x in UNIT-CUBE coordinates, like _GridBwFn."""

    @staticmethod
    def forward(ctx, x, table, W1, b1, W2, b2, grid, aabb, tf32):
        from . import _lib
        from ._lib import lib, ptr, check, stream
        _lib.require_device()
        x = x.contiguous()
        n, W = x.shape[0], W1.shape[0]
        tb = table.detach()
        enc = tcnn.grid_forward(x, tb, grid, aabb)
        if tf32 == "tc05":
            sigma, s2, g_e = _dn_fw(enc, W1, b1, W2, b2)
            g_x = tcnn.grid_backward_input(x, g_e, tb, grid, aabb)
            ctx.grid, ctx.aabb, ctx.tf32 = grid, aabb, tf32
            ctx.set_materialize_grads(False)
            ctx.save_for_backward(x, table, W1, b1, W2, enc, s2, g_e)      # 1 KB / sample (the torch-GEMM path keeps 2.5 KB)
            return sigma, g_x
        with _tf32_matmul(tf32):
            z1 = torch.addmm(b1, enc, W1.t())
        sigma = torch.empty(n, device=x.device); s2 = torch.empty(n, device=x.device)
        t = torch.empty(n, W, device=x.device)
        w2 = W2.reshape(-1).contiguous()
        check(lib.ngp_density_head_fw(ptr(z1), ptr(w2), ptr(b2.contiguous()), n, W, ptr(sigma), ptr(s2), ptr(t), stream()),
              "density_head_fw")
        with _tf32_matmul(tf32):
            g_e = t @ W1
        g_x = tcnn.grid_backward_input(x, g_e, tb, grid, aabb)
        ctx.grid, ctx.aabb, ctx.tf32 = grid, aabb, tf32
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, table, W1, w2, enc, z1, s2, t, g_e)
        return sigma, g_x

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dsigma, dgx):
        from ._lib import lib, ptr, check, stream
        tc = ctx.tf32 == "tc05"
        if tc:
            x, table, W1, b1, W2, enc, s2, g_e = ctx.saved_tensors
        else:
            x, table, W1, w2, enc, z1, s2, t, g_e = ctx.saved_tensors
        g, aabb = ctx.grid, ctx.aabb
        n, W = enc.shape[0], W1.shape[0]
        tdt = 0 if table.dtype == torch.float32 else 1
        d_ge = v = None
        if dgx is not None:
            dgx = dgx.contiguous()
            d_ge = torch.empty_like(enc)            # d(dgx . g_x)/d(g_e): the gather with the input-gradient coefficients
            check(lib.ngp_hashgrid_bwbw_input(ptr(x), tcnn._aabb_arg(aabb), ptr(dgx), None, ptr(table.detach()), tdt, *g.args(), n,
                                              None, ptr(d_ge), stream()), "hashgrid_bwbw_input")
        if tc:
            de, dW1, db1, dW2, db2 = _dn_bw(enc, d_ge, g_e, dsigma.contiguous() if dsigma is not None else None, s2, W1, b1, W2, True)
            dtable = None
            if ctx.needs_input_grad[1]:
                dtable = torch.zeros(g.n_params, dtype=torch.float32, device=x.device)
                check(lib.ngp_hashgrid_bw_params_dual(ptr(x), tcnn._aabb_arg(aabb), ptr(de), ptr(dgx), ptr(g_e) if dgx is not None else None,
                                                      *g.args(), n, ptr(dtable), stream()), "hashgrid_bw_params_dual")
            return None, dtable, dW1, db1, dW2, db2, None, None, None
        if d_ge is not None:
            with _tf32_matmul(ctx.tf32):
                v = d_ge @ W1.t()
        dz1 = torch.empty_like(z1); dz2 = torch.empty(n, device=z1.device)
        dw2 = torch.zeros(W, device=z1.device); db1 = torch.zeros(W, device=z1.device)
        check(lib.ngp_density_head_bw(ptr(z1), ptr(v), ptr(s2), ptr(dsigma.contiguous() if dsigma is not None else None), ptr(w2), n, W,
                                      ptr(dz1), ptr(dz2), ptr(dw2), ptr(db1), stream()), "density_head_bw")
        with _tf32_matmul(ctx.tf32):
            de = dz1 @ W1
            dW1 = dz1.t() @ enc
            if d_ge is not None:
                dW1.addmm_(t.t(), d_ge)
        dtable = None
        if ctx.needs_input_grad[1]:
            dtable = torch.zeros(g.n_params, dtype=torch.float32, device=x.device)
            check(lib.ngp_hashgrid_bw_params_dual(ptr(x), tcnn._aabb_arg(aabb), ptr(de), ptr(dgx), ptr(g_e) if dgx is not None else None,
                                                  *g.args(), n, ptr(dtable), stream()), "hashgrid_bw_params_dual")
        return None, dtable, dW1, db1, dw2[None], dz2.sum()[None], None, None, None


class _TwoHeadsFn(torch.autograd.Function):
    """norm_pred_header(feat) and semantic_header(feat) (networks.py:101-123: two bias-free MLPs  D -> 32 -> 3  and
    D -> 32 -> C  on the SAME feature matrix) evaluated as ONE block-structured MLP  D -> 64 -> (3 + C):
        W1 = [W1_norm ; W1_sem]          Wout = [[Wout_norm, 0], [0, Wout_sem]]
    — bit for bit the same products and sums per output (the zero blocks add exact zeros), one pass over the (S, D)
    features instead of two in each direction, and ONE dL/dfeat instead of two matrices that autograd then adds.
    The gradient of the merged parameter vector is sliced back into the two heads' own `.params` (the cross blocks'
    gradients are discarded: those weights do not exist)."""

    @staticmethod
    def _merge(pn, ps, D, H, C):
        nop = 16
        W1 = torch.cat([pn[:H * D].view(H, D), ps[:H * D].view(H, D)], 0)
        Wo = torch.zeros(nop, 2 * H, device=pn.device, dtype=pn.dtype)
        Wo[:3, :H] = pn[H * D:].view(nop, H)[:3]
        Wo[3:3 + C, H:] = ps[H * D:].view(nop, H)[:C]
        return torch.cat([W1.reshape(-1), Wo.reshape(-1)])

    @staticmethod
    def forward(ctx, feat, pn, ps, m_norm, m_sem):
        from . import _lib
        _lib.require_device()
        feat = feat.contiguous()
        D, H, C = feat.shape[1], m_norm.width, m_sem.n_out
        merged = _TwoHeadsFn._merge(pn.detach(), ps.detach(), D, H, C)
        m = tcnn.MlpConfig(D, 3 + C, {"otype": "CutlassMLP", "activation": "ReLU", "output_activation": "None",
                                      "n_neurons": 2 * H, "n_hidden_layers": 1})
        out = tcnn.mlp_forward([(feat, D, 0)], merged, m)
        ctx.m, ctx.dims = m, (D, H, C)
        ctx.save_for_backward(feat, merged, out)
        return out[:, :3], out[:, 3:]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dn, ds):
        feat, merged, out = ctx.saved_tensors
        D, H, C = ctx.dims
        dout = torch.cat([dn if dn is not None else torch.zeros_like(out[:, :3]),
                          ds if ds is not None else torch.zeros_like(out[:, 3:])], 1)
        dmerged, dsegs = tcnn.mlp_backward([(feat, D, 0)], merged, ctx.m, dout, [ctx.needs_input_grad[0]], saved_out=out)
        dW1 = dmerged[:2 * H * D].view(2 * H, D)
        dWo = dmerged[2 * H * D:].view(16, 2 * H)
        dpn = torch.zeros(H * D + 16 * H, device=feat.device); dps = torch.zeros_like(dpn)
        dpn[:H * D] = dW1[:H].reshape(-1); dps[:H * D] = dW1[H:].reshape(-1)
        dpn[H * D:].view(16, H)[:3] = dWo[:3, :H]
        dps[H * D:].view(16, H)[:C] = dWo[3:3 + C, H:]
        return dsegs[0], dpn, dps, None, None


class NGP(nn.Module, _OccupancyMixin):
    def __init__(self, scale, rgb_act="Sigmoid", use_skybox=False, embed_a=False, embed_a_len=12, classes=7,
                 grid_levels=16, grid_features=8, log2_T_xyz=19, log2_T_rgb=21, base_res=16, density_net_tf32=True,
                 density_net_tc=True):
        super().__init__()
        # The density net is the one part of the field the reference keeps in torch (nn.Linear 128 -> 128 -> 1 with
        # Softplus, double-differentiated for the normals, networks.py:54-59): its fp32 matmuls run on the SIMT pipe by
        # default (76 ms of a 340 ms step at 14 M samples, tools/step_profile_ngp.py).  TF32 puts them on the tensor
        # cores — the precision class SURVEY.md a12 states for this net; every other head already rounds to bf16.
        # torch's switch is process-wide, so it is flipped only around this net's own GEMMs (_tf32_matmul) and restored.
        self.density_net_tf32 = density_net_tf32
        # density_net_tc: the whole net (both GEMM pairs, Softplus epilogues, closed-form double backward) as one tcgen05 kernel per
        # direction (csrc/density_net.cu; bf16 operands like every other head).  Needs the reference's 128 -> 128 -> 1 shape and
        # fp32 (TF32 off) NOT requested: density_net_tf32=False keeps meaning "exact fp32 GEMMs" for the parity tests.
        import os
        self.density_net_tc = density_net_tc and density_net_tf32 and os.environ.get("NGP_DENSITY_NET_TC", "1") != "0"   # env: A/B runs only
        self.rgb_act = rgb_act
        self.use_skybox = use_skybox
        self.embed_a = embed_a
        self.classes = classes
        self._init_occupancy(scale)

        L, Fe = grid_levels, grid_features
        b = float(np.exp(np.log(2048 * scale / base_res) / (L - 1)))     # networks.py:37,64
        grid_cfg = lambda log2_T, otype: {"otype": otype, "type": "Hash", "n_levels": L, "n_features_per_level": Fe,
                                          "log2_hashmap_size": log2_T, "base_resolution": base_res,
                                          "per_level_scale": b, "interpolation": "Linear"}
        self.xyz_encoder = tcnn.Encoding(3, grid_cfg(log2_T_xyz, "Grid"))
        self.xyz_net = nn.Sequential(nn.Linear(self.xyz_encoder.n_output_dims, 128), nn.Softplus(), nn.Linear(128, 1))
        self.sigma_act = nn.Softplus()
        self.rgb_encoder = tcnn.Encoding(3, grid_cfg(log2_T_rgb, "HashGrid"))
        self.dir_encoder = tcnn.Encoding(3, {"otype": "SphericalHarmonics", "degree": 4})

        feat = self.rgb_encoder.n_output_dims
        rgb_in = feat + self.dir_encoder.n_output_dims + (embed_a_len if embed_a else 0)
        head = lambda n_in, n_out, width, out_act: tcnn.Network(n_in, n_out, {
            "otype": "CutlassMLP", "activation": "ReLU", "output_activation": out_act, "n_neurons": width,
            "n_hidden_layers": 1})
        self.rgb_net = head(rgb_in, 3, 128, rgb_act)
        self.norm_pred_header = head(feat, 3, 32, "None")
        self.semantic_header = head(feat, classes, 32, "None")
        self.semantic_act = nn.Softmax(dim=-1)
        if use_skybox:
            self.skybox_dir_encoder = tcnn.Encoding(3, {"otype": "SphericalHarmonics", "degree": 3})
            self.skybox_rgb_net = head(9, 3, 32, rgb_act)
        if rgb_act == "None":
            for i in range(3):
                setattr(self, f"tonemapper_net_{i}", head(1, 1, 64, "Sigmoid"))

    # ------------------------------------------------------------------ density / normals
    fused_density_head = True     # sigma and d sigma / d(encoding) through _DensityNormalsFn (csrc/density_head.cu)
    fused_density_field = True    # ... and the encoding around it in the same node (_DensityFieldNormalsFn): one dual scatter

    def _normalise(self, x):
        return (x - self.xyz_min) / (self.xyz_max - self.xyz_min)

    def _density_tc(self):
        l0, l2 = self.xyz_net[0], self.xyz_net[2]
        return (self.density_net_tc and l0.weight.is_cuda and l0.in_features == 128 and l0.out_features == 128 and l2.out_features == 1
                and l0.weight.dtype == torch.float32 and self.xyz_encoder.params.dtype == torch.float32)

    def density(self, x, return_feat=False, grad=True, grad_feat=True):
        """sigmas (N) [, feat_rgb (N, L*F)] for x (N,3) in [-scale, scale]  (networks.py:165-184)."""
        x, ab = x.contiguous(), self.aabb()              # (x - xyz_min) / (xyz_max - xyz_min) happens inside the grid kernels
        if self._density_tc() and not (grad and torch.is_grad_enabled()):
            # no graph wanted (occupancy update, test-time density): encoder + ONE forward kernel, no g_e
            with torch.no_grad():
                l0, l2 = self.xyz_net[0], self.xyz_net[2]
                sigmas = _dn_fw(self.xyz_encoder(x, ab), l0.weight, l0.bias, l2.weight, l2.bias, want_ge=False)[0]
        else:
            with torch.set_grad_enabled(grad and torch.is_grad_enabled()), _tf32_matmul(self.density_net_tf32):
                sigmas = self.sigma_act(self.xyz_net(self.xyz_encoder(x, ab))[:, 0])
        if not return_feat:
            return sigmas
        with torch.set_grad_enabled(grad_feat and torch.is_grad_enabled()):
            feat_rgb = self.rgb_encoder(x, ab)
        return sigmas, feat_rgb

    @torch.enable_grad()
    def grad(self, x, unit=False):
        """sigmas, feat_rgb, d sigma / d x (N,3), differentiable w.r.t. the parameters
        (networks.py:186-196).  unit=True: the gradient is left in unit-cube coordinates (the caller folds 1 / (xyz_max - xyz_min)
        into its next kernel, see _normals)."""
        x, ab = x.detach().contiguous(), self.aabb()
        l0, l2 = self.xyz_net[0], self.xyz_net[2]
        mode = "tc05" if self._density_tc() else self.density_net_tf32
        if (self.fused_density_field and self.fused_density_head and x.is_cuda and self.xyz_encoder.params.dtype == torch.float32
                and l0.out_features % 128 == 0 and l0.out_features <= 512):
            sigmas, g_xn = _DensityFieldNormalsFn.apply(x, self.xyz_encoder.params, l0.weight, l0.bias, l2.weight, l2.bias,
                                                        self.xyz_encoder.grid, ab, mode)
            return sigmas, self.rgb_encoder(x, ab), (g_xn if unit else g_xn / (self.xyz_max - self.xyz_min))
        enc = self.xyz_encoder(x, ab)
        if self.fused_density_head and enc.is_cuda and enc.dtype == torch.float32 and l0.out_features % 128 == 0 and l0.out_features <= 512:
            sigmas, g_enc = _DensityNormalsFn.apply(enc, l0.weight, l0.bias, l2.weight, l2.bias, mode)
        else:       # any other density net: generic autograd double backward, as the reference does it
            sigmas = self.sigma_act(self.xyz_net(enc)[:, 0])
            (g_enc,) = torch.autograd.grad(sigmas, enc, torch.ones_like(sigmas), create_graph=True)
        # input gradient of the grid as a differentiable op of (g_enc, table): its backward is the
        # double-backward kernel
        g_xn, _ = _GridBwFn.apply(g_enc.contiguous(), x, self.xyz_encoder.params, self.xyz_encoder.grid,
                                  True, False, ab)
        grads = g_xn if unit else g_xn / (self.xyz_max - self.xyz_min)
        feat_rgb = self.rgb_encoder(x, ab)
        return sigmas, feat_rgb, grads

    # ------------------------------------------------------------------ heads
    def _rgb(self, d, feat_rgb, kwargs):
        tensors, kinds = [d, feat_rgb], [1, 0]
        if self.embed_a:
            embed_a = kwargs["embedding_a"]
            if embed_a.size(0) < feat_rgb.size(0):
                embed_a = torch.repeat_interleave(embed_a, int(feat_rgb.size(0) / embed_a.size(0)), 0)
            tensors.append(embed_a); kinds.append(0)
        rgbs = self.rgb_net.forward_segments(tensors, kinds)
        if self.rgb_act == "None":
            rgbs = TruncExp.apply(rgbs) if kwargs.get("output_radiance", False) else self.log_radiance_to_rgb(rgbs, **kwargs)
        return rgbs

    fused_aux_heads = True        # normal + semantic heads as one block-structured MLP (_TwoHeadsFn)

    def _aux_heads(self, feat_rgb):
        """-> (norm_pred_header(feat), semantic_header(feat)) raw outputs."""
        a, b = self.norm_pred_header, self.semantic_header
        ma, mb = a.mlp, b.mlp
        if (self.fused_aux_heads and feat_rgb.is_cuda and ma.width == mb.width and ma.n_hidden == mb.n_hidden == 1 and 2 * ma.width <= 128
                and ma.act_h == mb.act_h == tcnn.ACT["ReLU"] and ma.act_o == mb.act_o == tcnn.ACT["None"] and 3 + mb.n_out <= 16):
            return _TwoHeadsFn.apply(feat_rgb, a.params, b.params, ma, mb)
        return a(feat_rgb), b(feat_rgb)

    fused_normals = True          # -normalize(.) of both normal outputs as one kernel per direction (csrc/normals.cu)

    def _normals(self, grads, n_out, fused):
        """-> normals_raw, normals_pred (networks.py:209,222-223).  fused: `grads` is in unit-cube coordinates and the world scale
        1 / (xyz_max - xyz_min) is applied inside the kernel."""
        if fused:
            from .custom_functions import NegNormalize
            inv_range = tuple(1.0 / r for r in self.aabb()[3:])
            return NegNormalize.apply(grads, inv_range, 1e-6), NegNormalize.apply(n_out, None, 1e-6)
        return -F.normalize(grads, p=2, dim=-1, eps=1e-6), -F.normalize(n_out, p=2, dim=-1, eps=1e-6)

    def log_radiance_to_rgb(self, log_radiances, **kwargs):
        out = [getattr(self, f"tonemapper_net_{i}")(log_radiances[:, i:i + 1]) for i in range(3)]
        return torch.cat(out, 1)

    def forward(self, x, d, **kwargs):
        """-> sigmas (N), rgbs (N,3), normals_raw (N,3), normals_pred (N,3), semantic (N,C)  (networks.py:198-240)"""
        fused = self.fused_normals and x.is_cuda and os.environ.get("NGP_FUSED_NORMALS", "1") != "0"       # env: A/B runs only
        sigmas, feat_rgb, grads = self.grad(x, unit=fused)
        n_out, s_out = self._aux_heads(feat_rgb)
        normals_raw, normals_pred = self._normals(grads, n_out, fused)
        semantic = self.semantic_act(s_out)
        rgbs = self._rgb(d, feat_rgb, kwargs)
        return sigmas, rgbs, normals_raw, normals_pred, semantic

    def forward_test(self, x, d, **kwargs):
        """Same quantities in the reference's test-time order: sigmas, rgbs, normals_pred,
        normals_raw, semantic (networks.py:242-282).  No graph is kept."""
        fused = self.fused_normals and x.is_cuda
        with torch.enable_grad():
            sigmas, feat_rgb, grads = self.grad(x, unit=fused)
        sigmas, feat_rgb, grads = sigmas.detach(), feat_rgb.detach(), grads.detach()
        with torch.no_grad():
            n_out, s_out = self._aux_heads(feat_rgb)
            normals_raw, normals_pred = self._normals(grads, n_out, fused)
            semantic = self.semantic_act(s_out)
            rgbs = self._rgb(d, feat_rgb, kwargs)
        return sigmas, rgbs, normals_pred, normals_raw, semantic

    def forward_skybox(self, d):
        if not self.use_skybox:
            return None
        d = d / torch.norm(d, dim=1, keepdim=True)
        return self.skybox_rgb_net(self.skybox_dir_encoder((d + 1) / 2))


class NGPCompact(nn.Module, _OccupancyMixin):
    """ngp_pl-shaped field: hash grid (L16,F2,T2^19) -> 64-wide MLP -> 16 ; rgb = MLP(SH4(d)|h)."""

    def __init__(self, scale, rgb_act="Sigmoid", n_levels=16, n_features=2, log2_T=19, base_res=16, width=64,
                 classes=0):
        super().__init__()
        self.rgb_act = rgb_act
        self.classes = classes
        self.use_skybox = False
        self.embed_a = False
        self._init_occupancy(scale)
        b = float(np.exp(np.log(2048 * scale / base_res) / (n_levels - 1)))
        self.xyz_encoder = tcnn.Encoding(3, {"otype": "HashGrid", "n_levels": n_levels,
                                             "n_features_per_level": n_features, "log2_hashmap_size": log2_T,
                                             "base_resolution": base_res, "per_level_scale": b})
        mlp = lambda n_hidden, out_act: {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": out_act,
                                         "n_neurons": width, "n_hidden_layers": n_hidden}
        self.sigma_net = tcnn.Network(self.xyz_encoder.n_output_dims, 16, mlp(1, "None"))
        self.rgb_net = tcnn.Network(32, 3, mlp(2, rgb_act))

    def _normalise(self, x):
        return (x - self.xyz_min) / (self.xyz_max - self.xyz_min)

    has_normals = False     # no normal / semantic heads: the renderer takes the lite compositor path
    fused_density = True    # encoder -> bf16 operand tiles -> MLP -> gradient tiles -> scatter (tcnn._DensityFieldFn)

    def density(self, x, return_feat=False):
        if self.fused_density and self.xyz_encoder.params.dtype == torch.float32:
            h, sigmas = self.sigma_net.forward_density_field(x, self.xyz_encoder, self.aabb())
        else:
            h, sigmas = self.sigma_net.forward_density_head(self.xyz_encoder(x.contiguous(), self.aabb()))
        return (sigmas, h) if return_feat else sigmas

    def forward(self, x, d, **kwargs):
        """-> sigmas (N), rgbs (N,3), None, None, None (no normal / semantic heads)."""
        sigmas, h = self.density(x, return_feat=True)
        rgbs = self.rgb_net.forward_segments([d, h], [1, 0])
        return sigmas, rgbs, None, None, None

    def forward_test(self, x, d, **kwargs):
        with torch.no_grad():
            sigmas, rgbs, _, _, _ = self.forward(x, d, **kwargs)
        z3 = torch.zeros(x.shape[0], 3, device=x.device)
        return sigmas, rgbs, z3, z3, torch.zeros(x.shape[0], kwargs.get("num_classes", self.classes), device=x.device)

    def forward_skybox(self, d):
        return None
