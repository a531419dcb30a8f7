"""tinycudann-shaped modules over the libngp_b200 C ABI.

Mirrors the slice of tiny-cuda-nn's PyTorch binding the reference uses
(models/networks.py:40-162, models/implicit_mask.py:16-27):

    Encoding(n_input_dims, encoding_config, seed=1337, dtype=None)
    Network(n_input_dims, n_output_dims, network_config, seed=1337)
    NetworkWithInputEncoding(n_input_dims, n_output_dims, encoding_config, network_config)

each an nn.Module with .n_input_dims, .n_output_dims, a single flat fp32 nn.Parameter `.params`
(empty for SphericalHarmonics) and forward(x:(N,n_in)) -> (N,n_out) fp32 — the
TCNN_HALF_PRECISION=0 build the reference's README.md:24 asks for.  The Grid encoding supports
first- and second-order autograd w.r.t. its input (normals need it, models/networks.py:186-196).

tiny-cuda-nn itself is not in /root/reference; semantics follow SURVEY.md Appendix B and parity
for this half is pinned only against oracle/tcnn_oracle.py ("parity unpinned" by the reference).
"""
import ctypes
import math

import numpy as np
import torch
from torch import nn

from . import _lib
from ._lib import lib, ptr, check, stream

ACT = {"None": 0, "ReLU": 1, "Sigmoid": 2, "Exponential": 3}


# --------------------------------------------------------------------------------------- grid
class GridConfig:
    __slots__ = ("n_levels", "n_features", "log2_T", "base_res", "per_level_scale", "n_params", "offsets", "sizes",
                 "resolutions", "scales", "dense")

    def __init__(self, cfg):
        self.n_levels = int(cfg.get("n_levels", 16))
        self.n_features = int(cfg.get("n_features_per_level", 2))
        self.log2_T = int(cfg.get("log2_hashmap_size", 19))
        self.base_res = int(cfg.get("base_resolution", 16))
        self.per_level_scale = float(cfg.get("per_level_scale", 2.0))
        if cfg.get("interpolation", "Linear") != "Linear":
            raise NotImplementedError("ngp_b200 Grid encoding: only Linear interpolation (the reference's choice)")
        if cfg.get("type", "Hash") != "Hash":
            raise NotImplementedError("ngp_b200 Grid encoding: only type=Hash (the reference's choice)")
        L = self.n_levels
        off = (ctypes.c_uint32 * (L + 1))()
        siz = (ctypes.c_uint32 * L)()
        res = (ctypes.c_uint32 * L)()
        sc = (ctypes.c_float * L)()
        dn = (ctypes.c_uint8 * L)()
        n = lib.ngp_hashgrid_layout(L, self.n_features, self.log2_T, self.base_res, self.per_level_scale,
                                    ctypes.cast(off, ctypes.c_void_p), ctypes.cast(siz, ctypes.c_void_p),
                                    ctypes.cast(res, ctypes.c_void_p), ctypes.cast(sc, ctypes.c_void_p),
                                    ctypes.cast(dn, ctypes.c_void_p))
        if n < 0:
            raise ValueError(_lib.last_error())
        self.n_params = int(n)
        self.offsets, self.sizes, self.resolutions = list(off), list(siz), list(res)
        self.scales, self.dense = list(sc), [bool(d) for d in dn]

    def args(self):
        return (self.n_levels, self.n_features, self.log2_T, self.base_res, self.per_level_scale)


def _aabb_arg(aabb):
    """aabb = None | 6 python floats (lo xyz, range xyz) -> HOST float[6] for the C ABI (kernels then take world x)."""
    return None if aabb is None else (ctypes.c_float * 6)(*[float(v) for v in aabb])


def grid_forward(x, table, g: GridConfig, aabb=None):
    n = x.shape[0]
    y = torch.empty(n, g.n_levels * g.n_features, dtype=torch.float32, device=x.device)
    check(lib.ngp_hashgrid_fw(ptr(x), _aabb_arg(aabb), ptr(table), 0 if table.dtype == torch.float32 else 1, *g.args(), n, ptr(y),
                              stream()), "hashgrid_fw")
    return y


def grid_backward_params(x, dy, g: GridConfig, out=None, aabb=None):
    dtable = torch.zeros(g.n_params, dtype=torch.float32, device=x.device) if out is None else out
    check(lib.ngp_hashgrid_bw_params(ptr(x), _aabb_arg(aabb), ptr(dy), *g.args(), x.shape[0], ptr(dtable), stream()),
          "hashgrid_bw_params")
    return dtable


def grid_backward_input(x, dy, table, g: GridConfig, aabb=None):
    dx = torch.empty_like(x)
    check(lib.ngp_hashgrid_bw_input(ptr(x), _aabb_arg(aabb), ptr(dy), ptr(table), 0 if table.dtype == torch.float32 else 1, *g.args(),
                                    x.shape[0], ptr(dx), stream()), "hashgrid_bw_input")
    return dx


class _GridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, table, g, aabb=None):
        _lib.require_device()
        x = x.contiguous()
        ctx.g, ctx.aabb = g, aabb
        ctx.save_for_backward(x, table)
        return grid_forward(x.detach(), table.detach(), g, aabb)

    @staticmethod
    def backward(ctx, dy):
        x, table = ctx.saved_tensors
        need_dx, need_dt = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx, dt = _GridBwFn.apply(dy.contiguous(), x, table, ctx.g, need_dx, need_dt, ctx.aabb)
        if need_dx and ctx.aabb is not None:        # chain rule of the fused (x - lo) / range
            dx = dx / dx.new_tensor(ctx.aabb[3:])
        return (dx if need_dx else None), (dt if need_dt else None), None, None


class _GridBwFn(torch.autograd.Function):
    """First-order backward as a differentiable op, so that autograd.grad(..., create_graph=True)
    through the encoding (models/networks.py:189-195) has a double backward."""

    @staticmethod
    def forward(ctx, dy, x, table, g, need_dx, need_dt, aabb=None):
        """dx is w.r.t. the UNIT-CUBE coordinate (also when aabb maps world x inside the kernels)."""
        ctx.g, ctx.aabb = g, aabb
        ctx.save_for_backward(dy, x, table)
        dx = grid_backward_input(x.detach(), dy.detach(), table.detach(), g, aabb) if need_dx else torch.zeros_like(x)
        dt = grid_backward_params(x.detach(), dy.detach(), g, aabb=aabb) if need_dt else torch.zeros(0, device=x.device)
        return dx, dt

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_dx, g_dt):
        # only the dL/dx branch is differentiated again (g_dt: nobody differentiates the table
        # gradient); outputs: d/d(dy), d/dx (not propagated), d/dtable
        dy, x, table = ctx.saved_tensors
        g = ctx.g
        n = x.shape[0]
        need_ddy, need_dt = ctx.needs_input_grad[0], ctx.needs_input_grad[2]
        d_dy = torch.empty_like(dy) if need_ddy else None
        dt2 = torch.zeros(g.n_params, dtype=torch.float32, device=x.device) if need_dt else None
        if need_ddy or need_dt:
            check(lib.ngp_hashgrid_bwbw_input(ptr(x), _aabb_arg(ctx.aabb), ptr(g_dx.contiguous()), ptr(dy), ptr(table),
                                              0 if table.dtype == torch.float32 else 1, *g.args(), n, ptr(dt2),
                                              ptr(d_dy), stream()), "hashgrid_bwbw_input")
        return d_dy, None, dt2, None, None, None, None


# --------------------------------------------------------------------------------------- SH
def sh_forward(v, degree):
    _lib.require_device()
    v = v.contiguous()
    out = torch.empty(v.shape[0], degree * degree, dtype=torch.float32, device=v.device)
    check(lib.ngp_sh_fw(ptr(v), int(degree), v.shape[0], ptr(out), stream()), "sh_fw")
    return out


class Encoding(nn.Module):
    """tcnn.Encoding for otype Grid / HashGrid / SphericalHarmonics (models/networks.py:40-85,128-135)."""

    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__()
        self.n_input_dims = int(n_input_dims)
        self.encoding_config = dict(encoding_config)
        self.seed = seed
        otype = encoding_config["otype"]
        if otype in ("Grid", "HashGrid"):
            if self.n_input_dims != 3:
                raise NotImplementedError("ngp_b200 Grid encoding: 3-D inputs only (the hot path's case)")
            self.kind = "grid"
            self.grid = GridConfig(encoding_config)
            self.n_output_dims = self.grid.n_levels * self.grid.n_features
            gen = torch.Generator().manual_seed(seed)
            init = (torch.rand(self.grid.n_params, generator=gen) * 2 - 1) * 1e-4   # tcnn: U(-1e-4, 1e-4)
            self.params = nn.Parameter(init)
        elif otype == "SphericalHarmonics":
            self.kind = "sh"
            self.degree = int(encoding_config.get("degree", 4))
            self.n_output_dims = self.degree ** 2
            self.params = nn.Parameter(torch.zeros(0))
        else:
            raise NotImplementedError(f"ngp_b200.tcnn.Encoding: otype {otype!r} is not on the hot path")

    def forward(self, x, aabb=None):
        """aabb (extension, grid only): 6 floats (lo xyz, range xyz); x is then world-space and the kernels apply
        (x - lo) / range themselves — same bits as normalising with tensor ops first, two passes over x fewer."""
        if self.kind == "grid":
            return _GridFn.apply(x.float(), self.params, self.grid, aabb)
        return sh_forward(x.float().detach(), self.degree)


# --------------------------------------------------------------------------------------- MLP
class MlpConfig:
    __slots__ = ("n_in", "width", "n_hidden", "n_out", "act_h", "act_o", "n_params")

    def __init__(self, n_in, n_out, cfg):
        self.n_in, self.n_out = int(n_in), int(n_out)
        self.width = int(cfg.get("n_neurons", 128))
        self.n_hidden = int(cfg.get("n_hidden_layers", 1))
        self.act_h = ACT[cfg.get("activation", "ReLU")]
        self.act_o = ACT[cfg.get("output_activation", "None")]
        self.n_params = int(lib.ngp_mlp_param_count(self.n_in, self.width, self.n_hidden, self.n_out))

    def layer_shapes(self):
        nop = (self.n_out + 15) // 16 * 16
        return [(self.width, self.n_in)] + [(self.width, self.width)] * (self.n_hidden - 1) + [(nop, self.width)]


def _seg_arrays(segs):
    """segs: list of (tensor, width, kind) -> ctypes arrays (ptrs, widths, kinds, strides)."""
    k = len(segs)
    P = (ctypes.c_void_p * k)(*[ptr(t) for t, _, _ in segs])
    W = (ctypes.c_int * k)(*[int(w) for _, w, _ in segs])
    K = (ctypes.c_int * k)(*[int(kd) for _, _, kd in segs])
    S = (ctypes.c_int64 * k)(*[int(t.stride(0)) if kd != 2 else 0 for t, _, kd in segs])
    return P, W, K, S


def mlp_forward(segs, params, m: MlpConfig, aux_exp=False, n=None):
    """segs: [(tensor (N,w) fp32 | dirs (N,3), width, kind)]; kind 0 plain, 1 SH4-of-normalised-dirs, 2 bf16 feature
    tiles (pass n).  aux_exp=True additionally returns exp(out[:,0]) (N) from the same epilogue."""
    n = segs[0][0].shape[0] if n is None else n
    out = torch.empty(n, m.n_out, dtype=torch.float32, device=params.device)
    aux = torch.empty(n, dtype=torch.float32, device=params.device) if aux_exp else None
    P, W, K, S = _seg_arrays(segs)
    check(lib.ngp_mlp_fw(len(segs), P, W, K, S, ptr(params), m.width, m.n_hidden, m.n_out, m.act_h, m.act_o, n,
                         ptr(out), out.stride(0), ptr(aux), stream()), "mlp_fw")
    return (out, aux) if aux_exp else out


def mlp_backward(segs, params, m: MlpConfig, dout, need_dseg, d_aux=None, n=None, dseg_numel=None, saved_out=None):
    n = segs[0][0].shape[0] if n is None else n
    dparams = torch.zeros_like(params)
    P, W, K, S = _seg_arrays(segs)
    dsegs = [(torch.empty(dseg_numel, dtype=torch.float32, device=params.device) if kd == 2 else
              torch.empty(n, w, dtype=torch.float32, device=params.device)) if (nd and kd != 1) else None
             for (_, w, kd), nd in zip(segs, need_dseg)]
    k = len(segs)
    DP = (ctypes.c_void_p * k)(*[ptr(d) for d in dsegs])
    DS = (ctypes.c_int64 * k)(*[int(d.stride(0)) if (d is not None and d.dim() == 2) else 0 for d in dsegs])
    dout = dout.contiguous()
    check(lib.ngp_mlp_bw(k, P, W, K, S, ptr(params), m.width, m.n_hidden, m.n_out, m.act_h, m.act_o, n, ptr(dout),
                         dout.stride(0), ptr(dparams), DP, DS, ptr(d_aux.contiguous()) if d_aux is not None else None,
                         ptr(saved_out) if saved_out is not None else None, saved_out.stride(0) if saved_out is not None else 0,
                         stream()), "mlp_bw")
    return dparams, dsegs


class _MlpFn(torch.autograd.Function):
    """out = MLP(cat(segments)); segments are passed flat as (t0, t1, ...) after the static args."""

    @staticmethod
    def forward(ctx, params, m, kinds, *tensors):
        _lib.require_device()
        tensors = tuple(t.contiguous() for t in tensors)
        segs = [(t.detach(), (16 if kd == 1 else t.shape[1]), kd) for t, kd in zip(tensors, kinds)]
        ctx.m, ctx.kinds = m, kinds
        out = mlp_forward(segs, params.detach(), m)
        ctx.save_for_backward(params, out, *tensors)        # the output lets the backward skip the output-layer recompute
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        params, out, *tensors = ctx.saved_tensors
        m, kinds = ctx.m, ctx.kinds
        segs = [(t, (16 if kd == 1 else t.shape[1]), kd) for t, kd in zip(tensors, kinds)]
        need = [ctx.needs_input_grad[3 + i] for i in range(len(tensors))]
        dparams, dsegs = mlp_backward(segs, params, m, dout, need, saved_out=out)
        return (dparams if ctx.needs_input_grad[0] else None, None, None, *dsegs)


class _MlpDensityHeadFn(torch.autograd.Function):
    """(h, sigma) = (MLP(x), exp(MLP(x)[:,0])) with TruncExp's backward (clamp to +-7) — the ngp_pl-shaped
    density head in one kernel each way (no strided select / exp / select_backward passes)."""

    @staticmethod
    def forward(ctx, params, m, x):
        _lib.require_device()
        x = x.contiguous()
        ctx.m = m
        h, sigma = mlp_forward([(x.detach(), x.shape[1], 0)], params.detach(), m, aux_exp=True)
        ctx.save_for_backward(params, x, h)
        return h, sigma

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dh, dsigma):
        params, x, h = ctx.saved_tensors
        dparams, dsegs = mlp_backward([(x, x.shape[1], 0)], params, ctx.m, dh, [ctx.needs_input_grad[2]], d_aux=dsigma, saved_out=h)
        return (dparams if ctx.needs_input_grad[0] else None), None, dsegs[0]


# --------------------------------------------------------------------------------------- fused density path
def grid_forward_tiles(x, table, g: GridConfig, aabb=None):
    """bf16 feature tiles (uint8 blob, ceil(N/128) tiles) in the MLP's tcgen05 operand layout."""
    n = x.shape[0]
    tb = int(lib.ngp_feature_tile_bytes(g.n_levels, g.n_features))
    tiles = torch.empty((n + 127) // 128 * tb, dtype=torch.uint8, device=x.device)
    check(lib.ngp_hashgrid_fw_tiles(ptr(x), _aabb_arg(aabb), ptr(table), 0 if table.dtype == torch.float32 else 1, *g.args(), n,
                                    ptr(tiles), stream()), "hashgrid_fw_tiles")
    return tiles


class GradSink:
    """In-place destination of a hash table's gradient for data-parallel training (ngp_b200.trainer.Trainer): the scatter
    writes straight into `buf` (which the trainer has installed as table.grad) one level range at a time, finest levels
    first, and `on_ready(a, b)` is called as soon as the range's launch is enqueued with the flat slice [a, b) of `buf`
    it completes — the trainer starts that slice's all-reduce there, so it runs under the next range's scatter.
    Registered per table storage in GRAD_SINKS; without an entry the backward returns a fresh dense gradient as before."""

    def __init__(self, buf, grid, on_ready, n_pieces=4):
        self.buf, self.on_ready = buf, on_ready
        F = grid.n_features
        lc = 1 if F >= 4 else 4 // F                           # levels per lane pair of the scatter kernel
        sizes, L = grid.sizes, grid.n_levels
        target = sum(sizes) / max(n_pieces - 1, 1)
        self.ranges, hi, acc = [], L, 0
        for l in range(L - 1, -1, -1):                         # finest first: they take the most scatter time, the coarse rest is small
            acc += sizes[l]
            if acc >= target and l % lc == 0 and l > 0:
                self.ranges.append((l, hi)); hi, acc = l, 0
        if hi > 0:            # the remainder: split off the dense coarse levels (a few MB) so that the LAST, exposed all-reduce is the smallest
            d = next((l for l in range(hi) if sizes[l] == max(sizes)), hi)
            d -= d % lc
            if 0 < d < hi:
                self.ranges.append((d, hi)); hi = d
            self.ranges.append((0, hi))
        self.flat = [(grid.offsets[a] * F, grid.offsets[b] * F) for a, b in self.ranges]


GRAD_SINKS = {}     # table.data_ptr() -> GradSink


class _DensityFieldFn(torch.autograd.Function):
    """(h, sigma) = density_head(MLP(encode(x))) of the ngp_pl-shaped field as ONE autograd node over three kernels
    per direction: the encoder writes bf16 operand tiles, the MLP bulk-copies them (no fp32 feature matrix, no
    conversion pass), the MLP backward writes dL/dy as gradient tiles that the scatter reads sector by sector.
    Same arithmetic as Encoding -> Network.forward_density_head (the MLP rounds its operands to bf16 either way)."""

    @staticmethod
    def forward(ctx, x, table, params, g, m, aabb):
        _lib.require_device()
        x = x.contiguous()
        n = x.shape[0]
        tiles = grid_forward_tiles(x.detach(), table.detach(), g, aabb)
        ctx.g, ctx.m, ctx.aabb = g, m, aabb
        h, sigma = mlp_forward([(tiles, g.n_levels * g.n_features, 2)], params.detach(), m, aux_exp=True, n=n)
        ctx.save_for_backward(x, table, params, tiles, h)
        return h, sigma

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dh, dsigma):
        x, table, params, tiles, h = ctx.saved_tensors
        g, m = ctx.g, ctx.m
        n = x.shape[0]
        k0p = (g.n_levels * g.n_features + 15) // 16 * 16
        need_table = ctx.needs_input_grad[1]
        dparams, dsegs = mlp_backward([(tiles, g.n_levels * g.n_features, 2)], params, m, dh, [need_table], d_aux=dsigma, n=n,
                                      dseg_numel=(n + 127) // 128 * 128 * k0p, saved_out=h)
        dtable = None
        sink = GRAD_SINKS.get(table.data_ptr()) if need_table else None
        if sink is not None:            # data-parallel: scatter level ranges into table.grad itself, hand each finished slice to the all-reduce
            for (la, lb), (fa, fb) in zip(sink.ranges, sink.flat):
                check(lib.ngp_hashgrid_bw_params_tiles_range(ptr(x), _aabb_arg(ctx.aabb), ptr(dsegs[0]), *g.args(), n, ptr(sink.buf), la, lb,
                                                             stream()), "hashgrid_bw_params_tiles")
                sink.on_ready(fa, fb)
        elif need_table:
            dtable = torch.zeros(g.n_params, dtype=torch.float32, device=x.device)
            check(lib.ngp_hashgrid_bw_params_tiles(ptr(x), _aabb_arg(ctx.aabb), ptr(dsegs[0]), *g.args(), n, ptr(dtable), stream()),
                  "hashgrid_bw_params_tiles")
        return None, dtable, (dparams if ctx.needs_input_grad[2] else None), None, None, None


def xavier_uniform_flat(shapes, seed):
    gen = torch.Generator().manual_seed(seed)
    parts = []
    for (o, i) in shapes:
        s = math.sqrt(6.0 / (i + o))
        parts.append(((torch.rand(o, i, generator=gen) * 2 - 1) * s).reshape(-1))
    return torch.cat(parts)


class Network(nn.Module):
    """tcnn.Network with otype CutlassMLP / FullyFusedMLP (models/networks.py:89-162): bias-free."""

    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        self.n_input_dims, self.n_output_dims = int(n_input_dims), int(n_output_dims)
        self.network_config = dict(network_config)
        self.seed = seed
        self.mlp = MlpConfig(n_input_dims, n_output_dims, network_config)
        self.params = nn.Parameter(xavier_uniform_flat(self.mlp.layer_shapes(), seed))

    def forward(self, x):
        return _MlpFn.apply(self.params, self.mlp, (0,), x.float())

    def forward_density_head(self, x):
        """-> (out (N,n_out), exp(out[:,0]) (N))"""
        return _MlpDensityHeadFn.apply(self.params, self.mlp, x.float())

    def forward_density_field(self, x, encoding, aabb=None):
        """-> (out (N,n_out), exp(out[:,0]) (N)) = forward_density_head(encoding(x, aabb)) through the fused
        feature-tile path (one autograd node; the fp32 feature matrix is never materialised)."""
        return _DensityFieldFn.apply(x.float(), encoding.params, self.params, encoding.grid, self.mlp, aabb)

    def forward_segments(self, tensors, kinds):
        """Fused input assembly: MLP(cat(segments)) without materialising the concatenation
        (replaces torch.cat + dir_encoder at models/networks.py:221-231)."""
        return _MlpFn.apply(self.params, self.mlp, tuple(kinds), *tensors)


class NetworkWithInputEncoding(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=1337):
        super().__init__()
        self.encoding = Encoding(n_input_dims, encoding_config, seed=seed)
        self.network = Network(self.encoding.n_output_dims, n_output_dims, network_config, seed=seed)
        self.n_input_dims, self.n_output_dims = int(n_input_dims), int(n_output_dims)

    def forward(self, x):
        return self.network(self.encoding(x))
