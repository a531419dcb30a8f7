"""GPU numerics of the tcnn-shaped half against oracle/tcnn_oracle.py
(PyTorch restatement of tiny-cuda-nn's published algorithm — parity UNPINNED by the reference, see
that file's header).  Tolerances: hash grid / SH fp32: rtol 1e-5 + atol 1e-6 (same fp32 maths, different
summation order); table gradients (fp32 atomics): rtol 1e-4; MLP: operands are rounded to bf16 for
the tcgen05 kind::f16 path, so it is compared (a) with an oracle that rounds the same operands
(rtol 2e-3) and (b) with the exact fp32 oracle (rtol 3e-2 of the output scale)."""
import numpy as np
import pytest
import torch

from oracle import tcnn_oracle

pytestmark = pytest.mark.gpu

GRIDS = [  # L, F, log2_T, base, scale-of-scene
    (16, 2, 19, 16, 0.5),     # BASELINE shape
    (16, 8, 19, 16, 0.5),     # reference xyz_encoder
    (16, 8, 21, 16, 8.0),     # reference rgb_encoder at scale 8
    (8, 4, 14, 8, 1.0),
    (5, 1, 12, 4, 1.0),
]


def _grid(L, F, T, base, scale):
    from ngp_b200 import tcnn
    b = float(np.exp(np.log(2048 * scale / base) / (L - 1)))
    enc = tcnn.Encoding(3, {"otype": "HashGrid", "n_levels": L, "n_features_per_level": F, "log2_hashmap_size": T,
                            "base_resolution": base, "per_level_scale": b}).cuda()
    with torch.no_grad():
        enc.params.copy_(torch.randn_like(enc.params) * 0.5)
    return enc, (L, F, T, base, b)


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("cfg", GRIDS, ids=lambda c: f"L{c[0]}F{c[1]}T{c[2]}")
def test_hashgrid_forward_and_first_order(cfg):
    """Against the oracle evaluated in float64 on the same float32 inputs.  The encoding's condition
    number grows with the level resolution (pos = x*scale + 0.5 is rounded at magnitude `scale`), so the
    forward bound is per level: |err| <= 16 * eps_fp32 * scale_l * max|table| + 1e-6 (3 weights, each
    off by up to eps*scale_l, times up to 8 corners); gradients are compared in relative L2 (1e-3: the same
    weight error eps*scale_l ~ 1e-4 at the finest level enters every scattered term)."""
    enc, args = _grid(*cfg)
    L, F = cfg[0], cfg[1]
    g = torch.Generator(device="cuda").manual_seed(0)
    n = 3001
    x = torch.rand(n, 3, device="cuda", generator=g)
    x[:8] = torch.tensor([[0, 0, 0], [1, 1, 1], [1, 0, 0], [0, 1, 0], [0.5, 0.5, 0.5], [0, 0, 1], [1, 1, 0], [0.999999, 0, 1]], device="cuda")
    x.requires_grad_(True)
    y = enc(x)
    xo = x.detach().double().requires_grad_(True)
    po = enc.params.detach().double().requires_grad_(True)
    yo = tcnn_oracle.grid_encode(xo, po, *args)
    assert y.shape == yo.shape == (n, L * F)
    vmax = float(enc.params.abs().max())
    scales = torch.tensor(enc.grid.scales, device="cuda").repeat_interleave(F)
    bound = 16 * 6e-8 * scales * vmax + 1e-6
    err = (y.double() - yo).abs()
    assert bool((err <= bound[None, :]).all()), float((err / bound[None, :]).max())
    dy = torch.randn(y.shape, device="cuda", generator=g)
    dy[::5] = 0                                              # exactly-zero upstream rows take the skip branch
    gx, gp = torch.autograd.grad(y, (x, enc.params), dy)
    gxo, gpo = torch.autograd.grad(yo, (xo, po), dy.double())
    assert rel(gp, gpo) < 1e-3, rel(gp, gpo)
    # dy/dx is piecewise constant and DISCONTINUOUS at cell faces: a sample whose fp32 position rounds
    # into the neighbouring cell of a fine level (P ~ eps*scale per axis per level) gets that cell's
    # slope.  Robust statistics: median per-sample error, and the fraction of such outliers.
    e = (gx.double() - gxo).norm(dim=1) / (gxo.norm(dim=1) + 1e-12)
    assert float(e[8:].median()) < 2e-3, float(e[8:].median())
    assert float((e[8:] > 2e-2).float().mean()) < 1e-2, float((e[8:] > 2e-2).float().mean())


@pytest.mark.parametrize("F", [4, 8, 2])
def test_hashgrid_double_backward(F):
    """Normals-style second order: loss = |d softplus(y.W)/dx|^2 + sum; gradient w.r.t. the table goes
    through the double-backward kernels (the H2 variants of the gather / scatter; F = 8 takes the feature-split
    scatter).  Low resolutions keep the fp32 problem well conditioned."""
    from ngp_b200 import tcnn
    from ngp_b200.tcnn import _GridFn
    L, T, base, b = 6, 11, 4, 1.45
    enc = tcnn.Encoding(3, {"otype": "HashGrid", "n_levels": L, "n_features_per_level": F, "log2_hashmap_size": T,
                            "base_resolution": base, "per_level_scale": b}).cuda()
    assert enc.grid.dense[0] and not enc.grid.dense[-1]
    with torch.no_grad():
        enc.params.copy_(torch.randn_like(enc.params) * 0.5)
    args = (L, F, T, base, b)
    g = torch.Generator(device="cuda").manual_seed(1)
    n = 4000
    x = torch.rand(n, 3, device="cuda", generator=g)
    W = torch.randn(L * F, 1, device="cuda", generator=g) * 0.3

    def normals_loss(encode, params, xin, Wm):
        xin = xin.clone().requires_grad_(True)
        s = torch.nn.functional.softplus(encode(xin, params) @ Wm)[:, 0]
        (dx,) = torch.autograd.grad(s, xin, torch.ones_like(s), create_graph=True)
        return (dx ** 2).sum() + s.sum()

    l1 = normals_loss(lambda xi, p: _GridFn.apply(xi, p, enc.grid), enc.params, x, W)
    (g1,) = torch.autograd.grad(l1, enc.params)
    po = enc.params.detach().double().requires_grad_(True)
    l2 = normals_loss(lambda xi, p: tcnn_oracle.grid_encode(xi, p, *args), po, x.double(), W.double())
    (g2,) = torch.autograd.grad(l2, po)
    assert abs(float(l1) - float(l2)) <= 1e-4 * abs(float(l2))
    assert rel(g1, g2) < 1e-3, rel(g1, g2)


def test_sh4_and_sh3():
    from ngp_b200 import tcnn
    g = torch.Generator(device="cuda").manual_seed(2)
    d = torch.nn.functional.normalize(torch.randn(5000, 3, device="cuda", generator=g), dim=-1)
    for deg in (4, 3, 2, 1):
        enc = tcnn.Encoding(3, {"otype": "SphericalHarmonics", "degree": deg})
        y = enc((d + 1) / 2)
        assert torch.allclose(y, tcnn_oracle.sh_encode((d + 1) / 2, deg), rtol=1e-5, atol=1e-6)


MLPS = [  # n_in, width, n_hidden, n_out, act, out_act
    (32, 64, 1, 16, "ReLU", "None"),        # BASELINE sigma net
    (32, 64, 2, 3, "ReLU", "Sigmoid"),      # BASELINE rgb net
    (144, 128, 1, 3, "ReLU", "Sigmoid"),    # reference rgb_net
    (152, 128, 1, 3, "ReLU", "None"),       # reference rgb_net + embedding, HDR
    (128, 32, 1, 3, "ReLU", "None"),        # norm_pred_header
    (128, 32, 1, 7, "ReLU", "None"),        # semantic_header
    (9, 32, 1, 3, "ReLU", "Sigmoid"),       # skybox
    (1, 64, 1, 1, "ReLU", "Sigmoid"),       # tonemapper
    (16, 16, 3, 5, "ReLU", "None"),
]


@pytest.mark.parametrize("cfg", MLPS, ids=lambda c: f"{c[0]}x{c[1]}x{c[2]}x{c[3]}")
@pytest.mark.parametrize("n", [1, 127, 128, 5000])
def test_mlp_forward(cfg, n):
    from ngp_b200 import tcnn
    n_in, width, nh, n_out, act, oact = cfg
    net = tcnn.Network(n_in, n_out, {"otype": "CutlassMLP", "activation": act, "output_activation": oact,
                                     "n_neurons": width, "n_hidden_layers": nh}).cuda()
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.randn(n, n_in, device="cuda", generator=g)
    with torch.no_grad():
        y = net(x)
        yb = tcnn_oracle.mlp_forward(x, net.params, n_in, width, nh, n_out, act, oact, operand_dtype=torch.bfloat16)
        yf = tcnn_oracle.mlp_forward(x, net.params, n_in, width, nh, n_out, act, oact)
    assert y.shape == (n, n_out) and torch.isfinite(y).all()
    scale = float(yf.abs().max()) + 1e-3
    assert torch.allclose(y, yb, rtol=2e-3, atol=2e-3 * scale), float((y - yb).abs().max())
    assert float((y - yf).abs().max()) <= 3e-2 * scale


@pytest.mark.parametrize("cfg", MLPS, ids=lambda c: f"{c[0]}x{c[1]}x{c[2]}x{c[3]}")
def test_mlp_backward(cfg):
    from ngp_b200 import tcnn
    n_in, width, nh, n_out, act, oact = cfg
    net = tcnn.Network(n_in, n_out, {"otype": "CutlassMLP", "activation": act, "output_activation": oact,
                                     "n_neurons": width, "n_hidden_layers": nh}).cuda()
    g = torch.Generator(device="cuda").manual_seed(7)
    n = 128 * 37 + 5
    x = torch.randn(n, n_in, device="cuda", generator=g, requires_grad=True)
    dy = torch.randn(n, n_out, device="cuda", generator=g)
    y = net(x)
    gx, gp = torch.autograd.grad(y, (x, net.params), dy)
    # the forward operands are rounded to bf16 exactly as the kernel does (straight-through gradient), so
    # both sides see the same ReLU masks; what remains is the bf16 rounding of the backward operands.
    xo = x.detach().clone().requires_grad_(True); po = net.params.detach().clone().requires_grad_(True)
    yo = tcnn_oracle.mlp_forward(xo, po, n_in, width, nh, n_out, act, oact, operand_dtype=torch.bfloat16)
    gxo, gpo = torch.autograd.grad(yo, (xo, po), dy)
    assert rel(gx, gxo) < 1.5e-2, rel(gx, gxo)
    assert rel(gp, gpo) < 1.5e-2, rel(gp, gpo)
    # and against exact fp32 maths (mask flips near z=0 dominate: sqrt(fraction flipped))
    yf = tcnn_oracle.mlp_forward(xo, po, n_in, width, nh, n_out, act, oact)
    gxf, gpf = torch.autograd.grad(yf, (xo, po), dy)
    assert rel(gx, gxf) < 8e-2 and rel(gp, gpf) < 8e-2, (rel(gx, gxf), rel(gp, gpf))
    pad = gp.numel() - (width * n_in + (nh - 1) * width * width + n_out * width)
    if pad:
        assert float(gp[-pad:].abs().max()) == 0.0          # padded output rows never receive gradient


def test_mlp_segments_fuse_sh_and_concat():
    from ngp_b200 import tcnn
    net = tcnn.Network(32, 3, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "Sigmoid",
                               "n_neurons": 64, "n_hidden_layers": 2}).cuda()
    g = torch.Generator(device="cuda").manual_seed(3)
    n = 4099
    dirs = torch.randn(n, 3, device="cuda", generator=g) * 3
    h = torch.randn(n, 16, device="cuda", generator=g, requires_grad=True)
    y = net.forward_segments([dirs, h], [1, 0])
    d = torch.nn.functional.normalize(dirs, dim=-1)
    x = torch.cat([tcnn_oracle.sh_encode((d + 1) / 2, 4), h], 1)
    yo = tcnn_oracle.mlp_forward(x, net.params, 32, 64, 2, 3, "ReLU", "Sigmoid", operand_dtype=torch.bfloat16)
    assert torch.allclose(y, yo, rtol=2e-3, atol=2e-3)
    dy = torch.randn(n, 3, device="cuda", generator=g)
    (gh,) = torch.autograd.grad(y, h, dy)
    ho = h.detach().clone().requires_grad_(True)
    yo = tcnn_oracle.mlp_forward(torch.cat([tcnn_oracle.sh_encode((d + 1) / 2, 4), ho], 1), net.params.detach(), 32, 64, 2, 3, "ReLU", "Sigmoid")
    (gho,) = torch.autograd.grad(yo, ho, dy)
    assert float((gh - gho).norm() / gho.norm()) < 8e-2


def test_mlp_density_head_fused_exp():
    """(h, sigma=exp(h[:,0])) from one kernel; backward applies TruncExp's clamp(+-7) rule."""
    from ngp_b200 import tcnn
    net = tcnn.Network(32, 16, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None",
                                "n_neurons": 64, "n_hidden_layers": 1}).cuda()
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 128 * 9 + 17
    x = (torch.randn(n, 32, device="cuda", generator=g) * 0.5).requires_grad_(True)
    h, sigma = net.forward_density_head(x)
    ho = tcnn_oracle.mlp_forward(x, net.params, 32, 64, 1, 16, "ReLU", "None", operand_dtype=torch.bfloat16)
    assert torch.allclose(h, ho, rtol=2e-3, atol=2e-3)
    assert torch.allclose(sigma, torch.exp(ho[:, 0]), rtol=5e-3, atol=1e-4)
    dh = torch.randn(n, 16, device="cuda", generator=g); ds = torch.randn(n, device="cuda", generator=g)
    gx, gp = torch.autograd.grad((h, sigma), (x, net.params), (dh, ds))
    xo = x.detach().clone().requires_grad_(True); po = net.params.detach().clone().requires_grad_(True)
    ho = tcnn_oracle.mlp_forward(xo, po, 32, 64, 1, 16, "ReLU", "None", operand_dtype=torch.bfloat16)
    dh_eff = dh.clone(); dh_eff[:, 0] += ds * torch.exp(ho[:, 0].detach().clamp(-7, 7))
    gxo, gpo = torch.autograd.grad(ho, (xo, po), dh_eff)
    assert rel(gx, gxo) < 1.5e-2 and rel(gp, gpo) < 1.5e-2, (rel(gx, gxo), rel(gp, gpo))


@pytest.mark.parametrize("half", [8.0, 1.5])          # range 16 (power of two: exact reciprocal path) and 3 (IEEE division path)
def test_hashgrid_fused_aabb_is_bit_identical_to_normalising_first(half):
    """The (x - xyz_min) / (xyz_max - xyz_min) pass of models/networks.py:174 fused into the grid kernels: IEEE sub + div
    in the kernel give the same bits as the two tensor ops, for the forward, both first-order gradients and the
    double backward."""
    from ngp_b200 import tcnn
    enc, _ = _grid(16, 2, 19, 16, 8.0)
    g = torch.Generator(device="cuda").manual_seed(3)
    n = 20011
    lo = torch.full((3,), -half, device="cuda"); rng = torch.full((3,), 2 * half, device="cuda")
    xw = (torch.rand(n, 3, device="cuda", generator=g) * 2 - 1) * half
    xn = ((xw - lo) / rng).contiguous()
    aabb = tuple(lo.tolist()) + tuple(rng.tolist())
    table = enc.params.detach()
    y0 = tcnn.grid_forward(xn, table, enc.grid); y1 = tcnn.grid_forward(xw, table, enc.grid, aabb)
    assert torch.equal(y0, y1)
    dy = torch.randn(n, 32, device="cuda", generator=g)
    assert torch.equal(tcnn.grid_backward_input(xn, dy, table, enc.grid), tcnn.grid_backward_input(xw, dy, table, enc.grid, aabb))
    t0 = tcnn.grid_backward_params(xn, dy, enc.grid); t1 = tcnn.grid_backward_params(xw, dy, enc.grid, aabb=aabb)
    assert rel(t1, t0) < 1e-6          # same terms, atomic order differs run to run
    # through autograd: d/dx of the world-space input carries the 1 / range chain-rule factor
    xa = xw.clone().requires_grad_(True); xb = xn.clone().requires_grad_(True)
    tcnn._GridFn.apply(xa, enc.params, enc.grid, aabb).square().sum().backward()
    ga, pa = xa.grad.clone(), enc.params.grad.clone(); enc.params.grad = None
    tcnn._GridFn.apply(xb, enc.params, enc.grid).square().sum().backward()
    assert torch.allclose(ga, xb.grad / rng, rtol=1e-6, atol=0) and rel(pa, enc.params.grad) < 1e-6


@pytest.mark.parametrize("n", [1, 127, 128, 129, 5000, 40001])
def test_fused_density_field_matches_unfused_path(n):
    """Encoder -> bf16 operand tiles -> MLP (one bulk copy per tile) -> gradient tiles -> scatter must equal
    Encoding -> Network.forward_density_head: the forward bit for bit (the MLP rounds the same fp32 features to the
    same bf16 operands either way), the gradients up to fp32 atomic ordering."""
    from ngp_b200 import tcnn
    enc, _ = _grid(16, 2, 19, 16, 0.5)
    net = tcnn.Network(32, 16, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None",
                                "n_neurons": 64, "n_hidden_layers": 1}).cuda()
    g = torch.Generator(device="cuda").manual_seed(n)
    x = (torch.rand(n, 3, device="cuda", generator=g) - 0.5)
    aabb = (-0.5, -0.5, -0.5, 1.0, 1.0, 1.0)
    gh = torch.randn(n, 16, device="cuda", generator=g); gs = torch.randn(n, device="cuda", generator=g)
    outs = []
    for fused in (False, True):
        enc.params.grad = None; net.params.grad = None
        if fused:
            h, sig = net.forward_density_field(x, enc, aabb)
        else:
            h, sig = net.forward_density_head(enc(x, aabb))
        ((h * gh).sum() + (sig * gs).sum()).backward()
        outs.append((h.detach().clone(), sig.detach().clone(), enc.params.grad.clone(), net.params.grad.clone()))
    (h0, s0, dt0, dp0), (h1, s1, dt1, dp1) = outs
    assert torch.equal(h0, h1) and torch.equal(s0, s1)
    assert rel(dp1, dp0) < 1e-5 and rel(dt1, dt0) < 1e-5


@pytest.mark.parametrize("shape", [(32, 64, 2, 3, "Sigmoid"), (32, 64, 1, 16, "None"), (144, 128, 1, 3, "Sigmoid"), (128, 32, 1, 7, "None")])
def test_mlp_backward_from_saved_output_equals_recompute(shape):
    """ngp_mlp_bw(saved_out=forward output) skips the output-layer recompute; the gradients must not change."""
    from ngp_b200 import tcnn
    n_in, width, nh, n_out, oact = shape
    net = tcnn.Network(n_in, n_out, {"otype": "CutlassMLP", "activation": "ReLU", "output_activation": oact,
                                     "n_neurons": width, "n_hidden_layers": nh}).cuda()
    g = torch.Generator(device="cuda").manual_seed(11)
    n = 128 * 9 + 77
    x = torch.randn(n, n_in, device="cuda", generator=g)
    dy = torch.randn(n, n_out, device="cuda", generator=g)
    p = net.params.detach()
    aux = oact == "None"
    out = tcnn.mlp_forward([(x, n_in, 0)], p, net.mlp, aux_exp=aux)
    y = out[0] if aux else out
    da = torch.randn(n, device="cuda", generator=g) if aux else None
    dp0, dx0 = tcnn.mlp_backward([(x, n_in, 0)], p, net.mlp, dy, [True], d_aux=da)
    dp1, dx1 = tcnn.mlp_backward([(x, n_in, 0)], p, net.mlp, dy, [True], d_aux=da, saved_out=y)
    assert torch.equal(dx0[0], dx1[0])
    assert rel(dp1, dp0) < 1e-5


@pytest.mark.parametrize("F", [2, 8])
def test_hashgrid_block_order_does_not_change_results(F, monkeypatch):
    """GridMeta::chunk_major (hashgrid.cu): tables beyond the L2 walk level chunks slowest.  The order only permutes
    CTAs: the gather is bit-identical, the scatter (fp32 atomics, already order-dependent) agrees to 1e-5 relative."""
    from ngp_b200 import tcnn
    enc, _ = _grid(16, F, 15, 16, 0.5)
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 20011
    x = torch.rand(n, 3, device="cuda", generator=g)
    dy = torch.randn(n, 16 * F, device="cuda", generator=g)
    table = enc.params.detach()
    out = {}
    for order in ("0", "1"):
        monkeypatch.setenv("NGP_HASH_ORDER", order)
        y = tcnn.grid_forward(x, table, enc.grid)
        dt = tcnn.grid_backward_params(x, dy, enc.grid)
        tiles = tcnn.grid_forward_tiles(x, table, enc.grid) if F == 2 else None
        torch.cuda.synchronize()
        out[order] = (y, dt, tiles)
    assert torch.equal(out["0"][0], out["1"][0])
    assert rel(out["1"][1], out["0"][1]) < 1e-5
    if F == 2:
        k0p = 32
        v = lambda t: t.view(-1, k0p // 8, 128 * 16 + 64)[:, :, :128 * 16]
        assert torch.equal(v(out["0"][2]), v(out["1"][2]))


@pytest.mark.parametrize("n", [129, 40001])
def test_grad_sink_level_ranges_equal_full_scatter(n):
    """tcnn.GradSink (data-parallel gradient exchange): the table gradient scattered level range by level range, finest
    first, straight into table.grad must equal the one-launch scatter (the ranges partition the levels; within a level
    the same reductions are issued), and every range must be announced exactly once with its flat slice."""
    from ngp_b200 import tcnn
    enc, _ = _grid(16, 2, 15, 16, 0.5)
    net = tcnn.Network(32, 16, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None",
                                "n_neurons": 64, "n_hidden_layers": 1}).cuda()
    g = torch.Generator(device="cuda").manual_seed(n)
    x = (torch.rand(n, 3, device="cuda", generator=g) - 0.5)
    aabb = (-0.5, -0.5, -0.5, 1.0, 1.0, 1.0)
    gh = torch.randn(n, 16, device="cuda", generator=g); gs = torch.randn(n, device="cuda", generator=g)

    def run():
        enc.params.grad = None; net.params.grad = None
        h, sig = net.forward_density_field(x, enc, aabb)
        return h, sig
    h, sig = run()
    ((h * gh).sum() + (sig * gs).sum()).backward()
    full = enc.params.grad.clone()
    buf = torch.zeros_like(enc.params)
    seen = []
    sink = tcnn.GradSink(buf, enc.grid, lambda a, b: seen.append((a, b)))
    assert sink.ranges[0][1] == 16 and sink.ranges[-1][0] == 0 and all(a[0] == b[1] for a, b in zip(sink.ranges, sink.ranges[1:]))
    tcnn.GRAD_SINKS[enc.params.data_ptr()] = sink
    try:
        h, sig = run()
        enc.params.grad = buf
        ((h * gh).sum() + (sig * gs).sum()).backward()
    finally:
        tcnn.GRAD_SINKS.clear()
    assert enc.params.grad is buf and seen == sink.flat
    assert sorted(seen)[0][0] == 0 and sorted(seen)[-1][1] == buf.numel()
    assert rel(buf, full) < 1e-5
