"""TEST INFRASTRUCTURE ONLY — plain-PyTorch restatement of the tiny-cuda-nn operators the
reference's hot path calls (models/networks.py:40-162): Grid/HashGrid encoding (Linear
interpolation), SphericalHarmonics (degree <= 4) and the bias-free CutlassMLP.

PARITY UNPINNED: the algorithm lives in NVlabs/tiny-cuda-nn, an un-vendored, un-pinned dependency
of the reference (README.md:14-26; `git clone --recursive` of master, TCNN_HALF_PRECISION=0) that
is absent from /root/reference and from this image, and the reference holds no golden vectors or
tests for it.  This file restates tcnn's published algorithm: it pins OUR
kernels to that restatement, not to a tcnn binary.

Everything is differentiable torch ops on whatever device/dtype the inputs live on (float32 or
float64), so autograd supplies forward, parameter/input backward and double backward.
"""
import math

import numpy as np
import torch

PRIMES = (1, 2654435761, 805459861)


def grid_layout(n_levels, n_features, log2_T, base_res, per_level_scale):
    """-> list of dict(offset, size, res, scale, dense) per level, total entries.  float32 maths as in
    tcnn's grid_scale()/grid_resolution()."""
    levels, off = [], 0
    log2_pls = np.float32(np.log2(np.float32(per_level_scale)))
    for l in range(n_levels):
        scale = np.float32(np.exp2(np.float32(l) * log2_pls) * np.float32(base_res) - np.float32(1.0))
        res = int(np.ceil(scale)) + 1
        dense_n = res ** 3
        size = min((dense_n + 7) // 8 * 8, 1 << log2_T)
        levels.append(dict(offset=off, size=size, res=res, scale=float(scale), dense=dense_n <= size))
        off += size
    return levels, off


def grid_encode(x, table, n_levels, n_features, log2_T, base_res, per_level_scale):
    """x (N,3) in [0,1]; table flat (n_entries*F) -> (N, L*F)."""
    levels, total = grid_layout(n_levels, n_features, log2_T, base_res, per_level_scale)
    tab = table.reshape(total, n_features)
    outs = []
    for lv in levels:
        pos = x * lv["scale"] + 0.5
        cell = torch.floor(pos)
        w = pos - cell
        ci = cell.detach().to(torch.int64)
        acc = 0
        for k in range(8):
            b = [(k >> d) & 1 for d in range(3)]
            c = [ci[:, d] + b[d] for d in range(3)]
            if lv["dense"]:
                idx = (c[0] + c[1] * lv["res"] + c[2] * lv["res"] ** 2) % lv["size"]
            else:
                m = 0xFFFFFFFF
                h = ((c[0] * PRIMES[0]) & m) ^ ((c[1] * PRIMES[1]) & m) ^ ((c[2] * PRIMES[2]) & m)
                idx = h % lv["size"]
            wk = 1
            for d in range(3):
                wk = wk * (w[:, d] if b[d] else 1 - w[:, d])
            acc = acc + wk[:, None] * tab[lv["offset"] + idx]
        outs.append(acc)
    return torch.cat(outs, 1)


def sh_encode(v, degree=4):
    """v (N,3) in [0,1] -> (N, degree^2) of the direction 2v-1."""
    x, y, z = (2 * v - 1).unbind(-1)
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    o = [torch.full_like(x, 0.28209479177387814)]
    if degree > 1:
        o += [-0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x]
    if degree > 2:
        o += [1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999,
              -1.0925484305920792 * xz, 0.54627421529603959 * x2 - 0.54627421529603959 * y2]
    if degree > 3:
        o += [0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
              0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0),
              0.45704579946446572 * x * (1.0 - 5.0 * z2), 1.4453057213202769 * z * (x2 - y2),
              0.59004358992664352 * x * (-x2 + 3.0 * y2)]
    return torch.stack(o, -1)


_ACT = {"None": lambda z: z, "ReLU": torch.relu, "Sigmoid": torch.sigmoid, "Exponential": torch.exp}


def mlp_layer_shapes(n_in, width, n_hidden, n_out):
    nop = (n_out + 15) // 16 * 16
    return [(width, n_in)] + [(width, width)] * (n_hidden - 1) + [(nop, width)]


def mlp_forward(x, params, n_in, width, n_hidden, n_out, activation="ReLU", output_activation="None",
                operand_dtype=None):
    """Bias-free MLP; params flat, row-major (out,in) per layer.  operand_dtype=torch.bfloat16 rounds
    every GEMM operand (activations and weights) to bf16 and accumulates in fp32 — the arithmetic
    of the tcgen05 kind::f16 path."""
    def rnd(t):
        return t.to(operand_dtype).to(t.dtype) if operand_dtype is not None else t
    off, h = 0, x
    shapes = mlp_layer_shapes(n_in, width, n_hidden, n_out)
    for li, (o, i) in enumerate(shapes):
        W = params[off:off + o * i].reshape(o, i); off += o * i
        z = rnd(h) @ rnd(W).t()
        h = _ACT[activation](z) if li < len(shapes) - 1 else _ACT[output_activation](z[:, :n_out])
    return h


# ------------------------------------------------------------------------------------------ density net + normals
def density_head_reference(e, W1, b1, W2, b2):
    """The reference's formulation (models/networks.py:54-59,172-181,186-196): xyz_net = Linear -> Softplus -> Linear(.,1),
    sigma_act = Softplus, and d sigma / d e by autograd with create_graph=True.  -> (sigma (N), g_e (N,D))."""
    import torch.nn.functional as F
    sig = F.softplus(F.linear(F.softplus(F.linear(e, W1, b1)), W2, b2))[:, 0]
    (g,) = torch.autograd.grad(sig, e, torch.ones_like(sig), create_graph=True)
    return sig, g


def density_head_closed_form(e, W1, b1, W2, b2, dsigma, dg):
    """Closed forms that csrc/density_head.cu evaluates (its header comment), in plain torch ops:
    -> (sigma, g_e, (de, dW1, db1, dW2, db2)) for upstream gradients (dsigma (N), dg (N,D)) of the two outputs."""
    import torch.nn.functional as F
    z1 = torch.addmm(b1, e, W1.t())
    s1, a1, w2 = torch.sigmoid(z1), F.softplus(z1), W2[0]
    z2 = a1 @ w2 + b2
    sigma, s2 = F.softplus(z2), torch.sigmoid(z2)
    t = s2[:, None] * s1 * w2
    g_e = t @ W1
    v = dg @ W1.t()
    uv = s1 * v
    dz2 = (uv @ w2) * s2 * (1 - s2) + dsigma * s2
    dz1 = w2 * (uv * s2[:, None] * (1 - s1) + dz2[:, None] * s1)
    return sigma, g_e, (dz1 @ W1, t.t() @ dg + dz1.t() @ e, dz1.sum(0), (s2[:, None] * uv + dz2[:, None] * a1).sum(0)[None],
                        dz2.sum()[None])
