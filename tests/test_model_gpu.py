"""End-to-end GPU checks of the composed path: the reference-literal NGP
field (two hash grids, torch density MLP, normals via double backward, semantic / normal heads,
appearance embedding) and the ngp_pl-shaped NGPCompact, through render() in train and test mode."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import tcnn_oracle

pytestmark = pytest.mark.gpu


def _small_ngp(**kw):
    from ngp_b200.networks import NGP
    torch.manual_seed(0)
    m = NGP(scale=0.5, grid_levels=8, grid_features=8, log2_T_xyz=15, log2_T_rgb=16, **kw).cuda()
    with torch.no_grad():
        m.xyz_encoder.params.mul_(3000.0)          # tcnn's U(-1e-4,1e-4) init gives ~zero features; make them O(0.3)
        m.rgb_encoder.params.mul_(3000.0)
    return m


def _oracle_field(m, x, d, embed=None):
    """The same field evaluated with plain torch ops (tcnn_oracle) from the module's own parameters."""
    g1, g2 = m.xyz_encoder.grid, m.rgb_encoder.grid
    x = x.clone().requires_grad_(True)
    xn = (x - m.xyz_min) / (m.xyz_max - m.xyz_min)
    enc = tcnn_oracle.grid_encode(xn, m.xyz_encoder.params, g1.n_levels, g1.n_features, g1.log2_T, g1.base_res, g1.per_level_scale)
    sig = F.softplus(m.xyz_net(enc)[:, 0])
    (grads,) = torch.autograd.grad(sig, x, torch.ones_like(sig), create_graph=True)
    feat = tcnn_oracle.grid_encode(xn, m.rgb_encoder.params, g2.n_levels, g2.n_features, g2.log2_T, g2.base_res, g2.per_level_scale)
    n_raw = -F.normalize(grads, dim=-1, eps=1e-6)
    n_pred = -F.normalize(tcnn_oracle.mlp_forward(feat, m.norm_pred_header.params, feat.shape[1], 32, 1, 3), dim=-1, eps=1e-6)
    sem = torch.softmax(tcnn_oracle.mlp_forward(feat, m.semantic_header.params, feat.shape[1], 32, 1, m.classes), -1)
    dn = F.normalize(d, dim=-1, eps=1e-6)
    inp = [tcnn_oracle.sh_encode((dn + 1) / 2, 4), feat] + ([embed] if embed is not None else [])
    inp = torch.cat(inp, 1)
    rgb = tcnn_oracle.mlp_forward(inp, m.rgb_net.params, inp.shape[1], 128, 1, 3, "ReLU", "Sigmoid")
    return sig, rgb, n_raw, n_pred, sem


@pytest.mark.parametrize("embed_a", [False, True])
def test_ngp_field_matches_torch_restatement(embed_a):
    m = _small_ngp(embed_a=embed_a, embed_a_len=8, classes=7, density_net_tf32=False)     # fp32 density GEMMs on both sides
    g = torch.Generator(device="cuda").manual_seed(1)
    n = 3000
    x = (torch.rand(n, 3, device="cuda", generator=g) - 0.5) * 0.98
    d = torch.randn(n, 3, device="cuda", generator=g)
    kw = {}
    emb = None
    if embed_a:
        emb = torch.randn(n, 8, device="cuda", generator=g) * 0.3
        kw["embedding_a"] = emb
    sig, rgb, n_raw, n_pred, sem = m(x, d, **kw)
    o_sig, o_rgb, o_nraw, o_npred, o_sem = _oracle_field(m, x, d, emb)
    assert torch.allclose(sig, o_sig, rtol=1e-3, atol=1e-4)
    cos = (n_raw * o_nraw).sum(-1)
    assert float(cos.median()) > 0.9999 and float((cos < 0.99).float().mean()) < 0.02     # cell-face flips only
    assert float((rgb - o_rgb).abs().max()) < 3e-2 and float((rgb - o_rgb).abs().mean()) < 4e-3   # bf16 heads
    assert float((sem - o_sem).abs().max()) < 3e-2
    assert float(((n_pred * o_npred).sum(-1)).median()) > 0.999
    # gradients of a scalar of every output w.r.t. the parameters that feed it, incl. the double backward
    w = torch.randn(n, device="cuda", generator=g)
    loss = (sig * w).sum() + (rgb.sum(-1) * w).sum() + (n_raw[:, 0] * w).sum() + (sem[:, 1] * w).sum()
    o_loss = (o_sig * w).sum() + (o_rgb.sum(-1) * w).sum() + (o_nraw[:, 0] * w).sum() + (o_sem[:, 1] * w).sum()
    ps = [m.xyz_encoder.params, m.rgb_encoder.params, m.xyz_net[0].weight, m.rgb_net.params]
    g1 = torch.autograd.grad(loss, ps)
    g2 = torch.autograd.grad(o_loss, ps)
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-20))
    assert rel(g1[2], g2[2]) < 2e-2, rel(g1[2], g2[2])      # torch density MLP weights (through normals too)
    assert rel(g1[0], g2[0]) < 5e-2, rel(g1[0], g2[0])      # xyz table: first + second order scatter
    assert rel(g1[1], g2[1]) < 5e-2, rel(g1[1], g2[1])      # rgb table through three bf16 heads
    assert rel(g1[3], g2[3]) < 5e-2, rel(g1[3], g2[3])


def _scene_and_model(model_kind):
    from ngp_b200 import vren
    from ngp_b200.networks import NGPCompact
    from synth_scenes import BoxScene, scene_density_grid
    scene = BoxScene("lego", device="cuda")
    if model_kind == "compact":
        torch.manual_seed(0)
        m = NGPCompact(scale=0.5).cuda()
    else:
        m = _small_ngp(classes=7)
    m.density_grid.copy_(scene_density_grid(scene))
    vren.packbits(m.density_grid, 0.5, m.density_bitfield)
    return scene, m


@pytest.mark.parametrize("kind", ["compact", "ngp"])
def test_training_converges_and_test_render_agrees(kind):
    from ngp_b200.rendering import render
    from ngp_b200.trainer import Trainer, psnr
    scene, m = _scene_and_model(kind)
    classes = 0 if kind == "compact" else 7
    kw = dict(exp_step_factor=0.0, num_classes=classes)
    if kind == "ngp":
        kw["normal_ref"] = True
    tr = Trainer(m, lr=1e-2, render_kwargs=kw, max_grad_norm=50.0)
    poses = scene.poses(20)
    gen = torch.Generator(device="cuda").manual_seed(3)
    losses = []
    for it in range(60):
        ro, rd = scene.sample_rays(8192, poses, gen)
        rgb, *_ = scene.shade(ro, rd)
        loss, res = tr.train_step(ro, rd, rgb, update_grid=False)
        losses.append(float(loss))
    assert np.isfinite(losses).all()
    assert np.mean(losses[-5:]) < 0.5 * np.mean(losses[:3]), (losses[:3], losses[-5:])
    # the adaptive test-time renderer and the training renderer see the same field
    ro, rd = scene.sample_rays(4096, poses, gen)
    gt, *_ = scene.shade(ro, rd)
    with torch.no_grad():
        a = render(m, ro, rd, exp_step_factor=0.0, num_classes=classes)
        b = render(m, ro, rd, exp_step_factor=0.0, num_classes=classes, test_time=True, T_threshold=1e-4)
    assert float(psnr(a["rgb"], gt)) > 15
    assert float(psnr(b["rgb"], a["rgb"])) > 30, float(psnr(b["rgb"], a["rgb"]))
    assert torch.allclose(a["opacity"], b["opacity"], atol=5e-2)
    assert b["depth"].shape == (4096,) and b["semantic"].shape == (4096, 1)
    if kind == "ngp":
        assert torch.isfinite(b["normal_raw"]).all() and b["normal_pred"].shape == (4096, 3)


def test_occupancy_update_runs_and_keeps_bitfield_consistent():
    from ngp_b200 import vren
    scene, m = _scene_and_model("compact")
    before = int(torch.count_nonzero(m.density_bitfield))
    for warm in (True, False):
        m.update_density_grid(0.01 * 1024 / 3 ** 0.5, warmup=warm)
    thr_grid = m.density_grid.clone()
    pos = thr_grid > 0
    mean = float((thr_grid * pos).sum() / pos.sum().clamp(min=1))
    ref = torch.zeros_like(m.density_bitfield)
    vren.packbits(thr_grid, min(mean, 0.01 * 1024 / 3 ** 0.5), ref)
    assert torch.equal(ref, m.density_bitfield) and before > 0


def test_geometric_schedule_renders_the_same_image_as_the_reference_schedule():
    from ngp_b200.rendering import render
    from ngp_b200.trainer import Trainer
    scene, m = _scene_and_model("compact")
    tr = Trainer(m, lr=1e-2, render_kwargs=dict(exp_step_factor=0.0, num_classes=0))
    poses = scene.poses(20)
    gen = torch.Generator(device="cuda").manual_seed(5)
    for it in range(40):
        ro, rd = scene.sample_rays(8192, poses, gen)
        rgb, *_ = scene.shade(ro, rd)
        tr.train_step(ro, rd, rgb, update_grid=False)
    ro, rd = scene.image_rays(poses[0], wh=(100, 100))
    with torch.no_grad():
        a = render(m, ro, rd, exp_step_factor=0.0, num_classes=0, test_time=True, T_threshold=1e-2, sample_schedule="reference", renderer="loop")
        b = render(m, ro, rd, exp_step_factor=0.0, num_classes=0, test_time=True, T_threshold=1e-2, sample_schedule="geometric", renderer="loop")
        c = render(m, ro, rd, exp_step_factor=0.0, num_classes=0, test_time=True, T_threshold=1e-2, device_loop=True)    # wavefront kernels, device-resident round loop
        d = render(m, ro, rd, exp_step_factor=0.0, num_classes=0, test_time=True, T_threshold=1e-2)                      # ... host-driven rounds, exact counts (default)
    # the device-resident loop launches from bounds instead of exact counts: same kernels on the same samples, bit for bit
    assert torch.equal(c["rgb"], d["rgb"]) and torch.equal(c["depth"], d["depth"]) and torch.equal(c["opacity"], d["opacity"])
    assert int(c["total_samples"]) == int(d["total_samples"])
    with torch.no_grad():
        big_o, big_d = scene.image_rays(poses[1], wh=(400, 300))                                           # > 2048 rays: round-0 culling engaged
        e = render(m, big_o, big_d, exp_step_factor=0.0, num_classes=0, test_time=True, T_threshold=1e-2)
        e2 = render(m, big_o, big_d, exp_step_factor=0.0, num_classes=0, test_time=True, T_threshold=1e-2, device_loop=True)
        assert torch.equal(e["rgb"], e2["rgb"]) and int(e["total_samples"]) == int(e2["total_samples"])
        # default = pipelined rounds (field enqueued behind emit, counters fetched on a side stream); pipelined=False = plain host-driven rounds
        e3 = render(m, big_o, big_d, exp_step_factor=0.0, num_classes=0, test_time=True, T_threshold=1e-2, pipelined=False)
        assert torch.equal(e["rgb"], e3["rgb"]) and torch.equal(e["depth"], e3["depth"]) and torch.equal(e["opacity"], e3["opacity"])
        assert int(e["total_samples"]) == int(e3["total_samples"])
        f = render(m, big_o, big_d, exp_step_factor=0.0, num_classes=0, test_time=True, T_threshold=1e-2, sample_schedule="geometric", renderer="loop")
    assert torch.allclose(e["rgb"], f["rgb"], atol=2e-5) and torch.allclose(e["opacity"], f["opacity"], atol=2e-5)
    assert int(e["total_samples"]) > 0
    assert torch.allclose(a["rgb"], b["rgb"], atol=2e-5) and torch.allclose(a["depth"], b["depth"], atol=2e-4)
    assert torch.allclose(a["opacity"], b["opacity"], atol=2e-5)
    assert torch.allclose(a["rgb"], c["rgb"], atol=2e-5) and torch.allclose(a["depth"], c["depth"], atol=2e-4)
    assert torch.allclose(a["opacity"], c["opacity"], atol=2e-5) and int(c["total_samples"]) > 0
    assert int(b["total_samples"]) > 0 and int(a["total_samples"]) > 0


def test_street_shaped_scene_cascades_semantics_normals_embedding():
    """BASELINE.json configs[2]/[3] in miniature: unbounded-style scene (scale 8 -> 5 cascades, exponential
    stepping 1/256), appearance embedding, 10 semantic classes, normal_ref + semantic + distortion losses,
    global-norm clip 50 — the KITTI-360 / Playground recipe (configs/kitti360_1538.txt, configs/Playground.txt)."""
    from ngp_b200 import vren
    from ngp_b200.networks import NGP
    from ngp_b200.rendering import render
    from synth_scenes import BoxScene, scene_density_grid
    from ngp_b200.trainer import Trainer
    scene = BoxScene("street", device="cuda")
    torch.manual_seed(0)
    m = NGP(scale=8.0, grid_levels=8, grid_features=8, log2_T_xyz=15, log2_T_rgb=16, embed_a=True, embed_a_len=8, classes=10).cuda()
    assert m.cascades == 5 and m.density_bitfield.numel() == 5 * 128 ** 3 // 8
    with torch.no_grad():
        m.xyz_encoder.params.mul_(3000.0); m.rgb_encoder.params.mul_(3000.0)
    m.density_grid.copy_(scene_density_grid(scene))
    vren.packbits(m.density_grid, 0.5, m.density_bitfield)
    emb = torch.nn.Embedding(16, 8).cuda()
    kw = dict(exp_step_factor=1.0 / 256, num_classes=10, normal_ref=True, semantic=True, random_bg=True)
    tr = Trainer(torch.nn.ModuleList([m, emb]), lr=2e-3, render_kwargs=kw, max_grad_norm=50.0)
    tr.model = m                                              # the trainer renders with the field; the optimiser also owns emb
    poses = scene.poses(16)
    gen = torch.Generator(device="cuda").manual_seed(3)
    losses = []
    for it in range(30):
        img = torch.randint(16, (4096,), device="cuda", generator=gen)
        u = torch.randint(scene.img_wh[0], (4096,), device="cuda", generator=gen).float()
        v = torch.randint(scene.img_wh[1], (4096,), device="cuda", generator=gen).float()
        ro, rd = scene.rays_from_pixels(poses, img, u, v)
        rgb, _, _, label = scene.shade(ro, rd)
        res = render(m, ro, rd, embedding_a=emb(img), **kw)
        d = tr.loss_fn(res, {"rgb": rgb, "label": label}, **kw)
        loss = sum(x.mean() for x in d.values())
        tr.opt.zero_grad(set_to_none=True); loss.backward(); tr.opt.step()
        losses.append(float(loss))
        assert res["semantic"].shape == (4096, 10) and res["normal_pred"].shape == (4096, 3)
        assert res["Ro"].shape == (4096,) and res["Rp"].shape == (4096, 3)
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:3])
    assert emb.weight.grad is not None and float(emb.weight.grad.abs().sum()) > 0
    with torch.no_grad():
        out = render(m, ro[:2048], rd[:2048], embedding_a=emb(img[:1]), exp_step_factor=1.0 / 256, num_classes=10,
                     test_time=True, T_threshold=1e-2)
    assert torch.isfinite(out["rgb"]).all() and out["semantic"].max() < 10


@pytest.mark.parametrize("width", [128, 256])
def test_density_head_matches_autograd_double_backward(width):
    """_DensityNormalsFn (csrc/density_head.cu) against the reference's formulation — torch.autograd.grad with
    create_graph=True through Linear -> Softplus -> Linear -> Softplus (networks.py:54-59,186-196) — evaluated in
    fp64: both outputs and all five gradients of a scalar that uses both.  fp32 kernels + fp32 (non-TF32) GEMMs:
    relative error 1e-4 of each tensor's norm."""
    from ngp_b200.networks import _DensityNormalsFn
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = torch.Generator(device="cuda").manual_seed(7)
        n, D = 4099, 96
        rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
        e, W1, b1, W2, b2 = rnd(n, D), rnd(width, D) * 0.2, rnd(width), rnd(1, width) * 0.3, rnd(1)
        e[0] = 40.0                                   # drives pre-activations past softplus' linear threshold (20)
        ds, dg = rnd(n), rnd(n, D)
        ps = [t.clone().requires_grad_(True) for t in (e, W1, b1, W2, b2)]
        sig, ge = _DensityNormalsFn.apply(*ps, False)          # fp32 GEMMs
        grads = torch.autograd.grad((sig * ds).sum() + (ge * dg).sum(), ps)
        po = [t.double().clone().requires_grad_(True) for t in (e, W1, b1, W2, b2)]
        o_sig = F.softplus(F.linear(F.softplus(F.linear(po[0], po[1], po[2])), po[3], po[4]))[:, 0]
        (o_ge,) = torch.autograd.grad(o_sig, po[0], torch.ones_like(o_sig), create_graph=True)
        o_grads = torch.autograd.grad((o_sig * ds.double()).sum() + (o_ge * dg.double()).sum(), po)
        rel = lambda a, b: float((a.double() - b).norm() / (b.norm() + 1e-30))
        assert rel(sig, o_sig) < 1e-5 and rel(ge, o_ge) < 1e-4, (rel(sig, o_sig), rel(ge, o_ge))
        for name, a, b in zip(("e", "W1", "b1", "W2", "b2"), grads, o_grads):
            assert a.shape == b.shape and rel(a, b) < 1e-4, (name, rel(a, b))
        # sigma-only upstream (density evaluation without normals) and normals-only upstream
        sig, ge = _DensityNormalsFn.apply(*ps, False)
        (g_w1,) = torch.autograd.grad((sig * ds).sum(), ps[1])
        (o_w1,) = torch.autograd.grad((F.softplus(F.linear(F.softplus(F.linear(po[0], po[1], po[2])), po[3], po[4]))[:, 0] * ds.double()).sum(), po[1])
        assert rel(g_w1, o_w1) < 1e-4
        # the default: the net's own GEMMs as TF32 (10-bit mantissa operands) — 3e-3 of each tensor's norm — and the
        # process-wide switch is back to what it was afterwards
        sig, ge = _DensityNormalsFn.apply(*ps)
        grads = torch.autograd.grad((sig * ds).sum() + (ge * dg).sum(), ps)
        assert torch.backends.cuda.matmul.allow_tf32 is False
        assert rel(sig, o_sig) < 3e-3 and rel(ge, o_ge) < 3e-3
        for name, a, b in zip(("e", "W1", "b1", "W2", "b2"), grads, o_grads):
            assert rel(a, b) < 3e-3, (name, rel(a, b))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32


def test_fused_aux_heads_equal_the_two_separate_heads():
    """NGP._aux_heads: norm_pred_header + semantic_header as one block-structured MLP (_TwoHeadsFn) against the two
    separate tcnn.Network calls the reference makes (networks.py:221-224): same outputs (the zero blocks add exact
    zeros: 1e-6 absolute) and the same gradients for the features and for BOTH heads' own parameter vectors."""
    m = _small_ngp(classes=7)
    g = torch.Generator(device="cuda").manual_seed(4)
    n = 5000
    feat = (torch.randn(n, m.rgb_encoder.n_output_dims, device="cuda", generator=g) * 0.5).requires_grad_(True)
    wn, ws = torch.randn(n, 3, device="cuda", generator=g), torch.randn(n, 7, device="cuda", generator=g)
    ps = [feat, m.norm_pred_header.params, m.semantic_header.params]
    m.fused_aux_heads = True
    a1, b1 = m._aux_heads(feat)
    g1 = torch.autograd.grad((a1 * wn).sum() + (b1 * ws).sum(), ps)
    m.fused_aux_heads = False
    a0, b0 = m._aux_heads(feat)
    g0 = torch.autograd.grad((a0 * wn).sum() + (b0 * ws).sum(), ps)
    assert a1.shape == (n, 3) and b1.shape == (n, 7)
    assert float((a1 - a0).abs().max()) < 1e-6 and float((b1 - b0).abs().max()) < 1e-6
    rel = lambda x, y: float((x - y).norm() / (y.norm() + 1e-20))
    assert rel(g1[0], g0[0]) < 2e-3          # dL/dfeat: one bf16-operand backward instead of the sum of two
    assert g1[1].shape == g0[1].shape and rel(g1[1], g0[1]) < 1e-4
    assert g1[2].shape == g0[2].shape and rel(g1[2], g0[2]) < 1e-4
    # only one head used downstream (the other upstream gradient is None)
    a1, b1 = (setattr(m, "fused_aux_heads", True), m._aux_heads(feat))[1]
    (gs,) = torch.autograd.grad((b1 * ws).sum(), m.semantic_header.params)
    assert rel(gs, g0[2]) < 1e-4


def test_density_field_single_node_equals_the_three_node_path():
    """NGP.grad through _DensityFieldNormalsFn (gather + density net + input gradient as one autograd node, one dual
    scatter in the backward) against the same quantities through _GridFn -> _DensityNormalsFn -> _GridBwFn (two scatters):
    identical forward kernels (bit-equal outputs), and the table / density-net gradients of a scalar that uses sigma AND
    the normals agree to 1e-4 of their norms (atomics order; the dual kernel adds both terms in registers)."""
    m = _small_ngp(classes=7, density_net_tf32=False)
    g = torch.Generator(device="cuda").manual_seed(9)
    n = 6000
    x = (torch.rand(n, 3, device="cuda", generator=g) - 0.5) * 0.98
    ws, wg = torch.randn(n, device="cuda", generator=g), torch.randn(n, 3, device="cuda", generator=g)
    ps = [m.xyz_encoder.params, m.xyz_net[0].weight, m.xyz_net[0].bias, m.xyz_net[2].weight, m.xyz_net[2].bias]
    outs, grads = {}, {}
    for fused in (True, False):
        m.fused_density_field = fused
        sig, feat, gr = m.grad(x)
        outs[fused] = (sig, gr)
        grads[fused] = torch.autograd.grad((sig * ws).sum() + (gr * wg).sum(), ps)
        (grads[fused + 2],) = torch.autograd.grad((m.grad(x)[0] * ws).sum(), ps[0])      # sigma only: no normals upstream
    assert torch.equal(outs[True][0], outs[False][0]) and torch.equal(outs[True][1], outs[False][1])
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-20))
    for a, b in zip(grads[True], grads[False]):
        assert a.shape == b.shape and rel(a, b) < 1e-4, rel(a, b)
    assert rel(grads[3], grads[2]) < 1e-4


@pytest.mark.parametrize("kind", ["ngp", "street"])
def test_wavefront_renderer_with_normals_and_semantics_equals_the_loop(kind):
    """Fields with normal / semantic heads through the fused wavefront rounds (ngp_render_advance_full: rgb, depth, opacity,
    normal_pred, normal_raw and the C semantic channels composited in the advance kernel) against the reference-style loop
    over raymarching_test + composite_test_fw on the same field: same samples, same per-ray recurrence."""
    from ngp_b200.rendering import render
    from ngp_b200.networks import NGP
    from ngp_b200 import vren
    from synth_scenes import BoxScene, scene_density_grid
    if kind == "ngp":
        scene, m = _scene_and_model("ngp")
        kw = dict(exp_step_factor=0.0, num_classes=7)
    else:
        scene = BoxScene("street", device="cuda")
        torch.manual_seed(0)
        m = NGP(scale=8.0, grid_levels=8, grid_features=8, log2_T_xyz=15, log2_T_rgb=16, embed_a=True, embed_a_len=8, classes=10).cuda()
        with torch.no_grad():
            m.xyz_encoder.params.mul_(3000.0); m.rgb_encoder.params.mul_(3000.0)
        m.density_grid.copy_(scene_density_grid(scene))
        vren.packbits(m.density_grid, 0.5, m.density_bitfield)
        kw = dict(exp_step_factor=1.0 / 256, num_classes=10, embedding_a=torch.randn(1, 8, device="cuda") * 0.3)
    ro, rd = scene.image_rays(scene.poses(4)[1], wh=(96, 64) if kind == "street" else (80, 80))
    with torch.no_grad():
        a = render(m, ro, rd, test_time=True, T_threshold=1e-2, sample_schedule="geometric", renderer="loop", **kw)
        b = render(m, ro, rd, test_time=True, T_threshold=1e-2, **kw)            # wavefront
    assert int(a["total_samples"]) > 0 and int(b["total_samples"]) > 0
    for k in ("rgb", "opacity"):
        assert torch.allclose(a[k], b[k], atol=1e-4), (k, float((a[k] - b[k]).abs().max()))
    assert torch.allclose(a["depth"], b["depth"], rtol=1e-4, atol=1e-3)
    hit = a["opacity"] > 0.05
    assert hit.any()
    for k in ("normal_pred", "normal_raw"):
        assert float((a[k][hit] - b[k][hit]).abs().max()) < 1e-3, k
    assert float((a["semantic"][hit] == b["semantic"][hit]).float().mean()) > 0.999
