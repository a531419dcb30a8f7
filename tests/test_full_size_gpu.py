"""Parity at BASELINE.json's FULL size (configs[1]: 2^18 rays of the Lego-shaped scene, ~9 M samples, the L16 F2 T2^19 grid
and the 64-wide MLPs) — the small-case tests (tests/test_vren_gpu.py, tests/test_tcnn_gpu.py) run 512-20 000 rays.

  * marcher vs the reference's own kernel (vren_ref, live) on the same 2^18 rays: per-ray counts, ts, deltas, xyzs
    bit-exact after canonical ordering; plus size-independent properties (counter = sum of counts, ts strictly
    increasing inside a ray, samples inside the box);
  * compositor fw / bw, distortion fw / bw vs vren_ref on the marcher's own ~9 M samples: rtol 2e-4 + atol 2e-5
    (ex2.approx vs __expf, scan re-association), on >= 99.99 % of the elements;
  * hash grid fw / table gradient vs the plain-torch restatement on the same 9 M points: forward within the per-level conditioning bound
    16*eps*scale_l*max|table|, table gradient relative L2 < 1e-3, linearity of the scatter in dL/dy;
  * MLP fw vs the bf16-operand restatement: |err| < 2e-3 of the output scale.
"""
import numpy as np
import pytest
import torch

from oracle import build_ref, tcnn_oracle

pytestmark = pytest.mark.gpu
R = 1 << 18


@pytest.fixture(scope="module")
def batch():
    from ngp_b200 import vren
    from synth_scenes import BoxScene, scene_density_grid
    scene = BoxScene("lego", device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(20220806)
    ro, rd = scene.sample_rays(R, scene.poses(100), gen)
    grid = scene_density_grid(scene)
    bf = torch.zeros(grid.numel() // 8, dtype=torch.uint8, device="cuda")
    vren.packbits(grid, 0.5, bf)
    c, h = torch.zeros(1, 3, device="cuda"), torch.full((1, 3), 0.5, device="cuda")
    _, hits_t, _ = vren.ray_aabb_intersect(ro, rd, c, h, 1)
    t1 = hits_t[:, 0, 0]
    hits_t[:, 0, 0] = torch.where((t1 >= 0) & (t1 < 0.01), torch.full_like(t1, 0.01), t1)
    noise = torch.rand(R, device="cuda", generator=gen)
    ht = hits_t[:, 0].contiguous()
    out = vren.raymarching_train(ro, rd, ht, bf, 1, 0.5, 0.0, noise, 128, 1024)
    return dict(ro=ro, rd=rd, ht=ht, bf=bf, noise=noise, march=out, c=c, h=h)


def _close(a, b, rtol=2e-4, atol=2e-5, frac=0.9999):
    ok = (a - b).abs() <= atol + rtol * b.abs()
    assert float(ok.float().mean()) >= frac, (float(1 - ok.float().mean()), float((a - b).abs().max()))


def test_marcher_full_size_vs_reference_kernel(batch):
    vref = build_ref.load()
    rays_a, xyzs, dirs, deltas, ts, counter = batch["march"]
    S = int(counter[0])
    assert S == xyzs.shape[0] == int(rays_a[:, 2].sum()) > 5_000_000 and int(counter[1]) == R
    assert torch.equal(rays_a[:, 0], torch.arange(R, device="cuda")) and torch.equal(rays_a[:, 1], torch.cumsum(rays_a[:, 2], 0) - rays_a[:, 2])
    # properties: ts strictly increasing inside every ray, samples inside the (slightly padded) box
    first = torch.zeros(S, dtype=torch.bool, device="cuda")
    first[rays_a[rays_a[:, 2] > 0, 1]] = True
    assert bool(((ts[1:] > ts[:-1]) | first[1:]).all())
    assert float(xyzs.abs().max()) <= 0.5 + 1e-3
    o_s, d_s = batch["ro"].repeat_interleave(rays_a[:, 2], 0), batch["rd"].repeat_interleave(rays_a[:, 2], 0)
    assert float((xyzs - (o_s + d_s * ts[:, None])).abs().max()) < 1e-6 and torch.equal(dirs, d_s)
    del o_s, d_s
    if vref is None:
        pytest.skip("oracle/_ref/vren_ref.so not built")
    va, vx, vd, vdl, vts, vc = vref.raymarching_train(batch["ro"], batch["rd"], batch["ht"], batch["bf"], 1, 0.5, 0.0, batch["noise"], 128, 1024)
    assert int(vc[0]) == S
    start = torch.zeros(R, dtype=torch.int64, device="cuda"); cnt = torch.zeros_like(start)
    start[va[:, 0]] = va[:, 1]; cnt[va[:, 0]] = va[:, 2]
    assert torch.equal(cnt, rays_a[:, 2])                                       # per-ray counts: bit-exact
    ray = torch.repeat_interleave(torch.arange(R, device="cuda"), cnt)
    src = start[ray] + (torch.arange(S, device="cuda") - rays_a[ray, 1])       # the reference's row of every one of our samples
    for ours, theirs in ((ts, vts), (deltas, vdl), (xyzs, vx), (dirs, vd)):
        assert torch.equal(theirs[src].view(torch.int32), ours.view(torch.int32))
    del va, vx, vd, vdl, vts


def test_compositor_and_distortion_full_size_vs_reference_kernels(batch):
    from ngp_b200 import vren
    vref = build_ref.load()
    if vref is None:
        pytest.skip("oracle/_ref/vren_ref.so not built")
    rays_a, xyzs, dirs, deltas, ts, counter = batch["march"]
    S, C = xyzs.shape[0], 7
    g = torch.Generator(device="cuda").manual_seed(1)
    sig = torch.rand(S, device="cuda", generator=g) ** 3 * 60
    rgbs = torch.rand(S, 3, device="cuda", generator=g); nrm = torch.randn(S, 3, device="cuda", generator=g)
    sems = torch.softmax(torch.randn(S, C, device="cuda", generator=g), -1)
    ours = vren.composite_train_fw(sig, rgbs, nrm, sems, deltas, ts, rays_a, 1e-4, C)
    ref = vref.composite_train_fw(sig, rgbs, nrm, sems, deltas, ts, rays_a, 1e-4, C)
    assert torch.equal(ours[0], ref[0])                                         # total_samples per ray (terminating sample excluded)
    for a, b in zip(ours[1:], ref[1:]):
        _close(a, b)
    tot, op, dep, rgb, nr, sm, ws = ref
    up = [torch.randn(op.shape, device="cuda", generator=g) for op in (op, dep, rgb, nr, sm, ws)]
    go = vren.composite_train_bw(*up, sig, rgbs, nrm, ws, deltas, ts, rays_a, op, dep, rgb, nr, 1e-4, C)
    gr = vref.composite_train_bw(*[u.clone() for u in up], sig, rgbs, nrm, ws, deltas, ts, rays_a, op, dep, rgb, nr, 1e-4, C)
    for a, b in zip(go, gr):
        _close(a, b, rtol=1e-3, atol=1e-4, frac=0.999)
    lo, lr = vren.distortion_loss_fw(ws, deltas, ts, rays_a), vref.distortion_loss_fw(ws, deltas, ts, rays_a)
    for a, b in zip(lo, lr):
        _close(a, b, rtol=1e-3, atol=1e-5, frac=0.999)
    gl = torch.randn(R, device="cuda", generator=g)
    _close(vren.distortion_loss_bw(gl, lr[1], lr[2], ws, deltas, ts, rays_a), vref.distortion_loss_bw(gl, lr[1], lr[2], ws, deltas, ts, rays_a),
           rtol=1e-3, atol=1e-5, frac=0.999)


def test_hashgrid_and_mlp_full_size_vs_restatement(batch):
    from ngp_b200 import tcnn
    from baseline import tcnn_standin
    xyzs, dirs = batch["march"][1], batch["march"][2]
    S = xyzs.shape[0]
    cfg = {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16,
           "per_level_scale": float(np.exp(np.log(2048 * 0.5 / 16) / 15))}
    enc = tcnn.Encoding(3, cfg).cuda()
    with torch.no_grad():
        enc.params.mul_(3000.0)
    ref = tcnn_standin.Encoding(3, cfg).cuda()
    ref.load_state_dict(enc.state_dict())
    aabb = (-0.5, -0.5, -0.5, 1.0, 1.0, 1.0)
    table = enc.params.detach()
    y = tcnn.grid_forward(xyzs, table, enc.grid, aabb)
    xn = (xyzs + 0.5)
    with torch.no_grad():
        y_ref = ref(xn)
    # per-level bound of tests/test_tcnn_gpu.py: pos = x*scale + 0.5 is rounded at magnitude `scale`, so a weight is off by up to
    # eps*scale_l; |err| <= 16 * eps_fp32 * scale_l * max|table| + 1e-6 (both sides are fp32 here, hence a factor 2)
    bound = 2 * 16 * 6e-8 * torch.tensor(enc.grid.scales, device="cuda").repeat_interleave(2) * float(table.abs().max()) + 2e-6
    err = (y - y_ref).abs()
    assert bool((err <= bound[None, :]).all()), float((err / bound[None, :]).max())
    g = torch.Generator(device="cuda").manual_seed(2)
    dy = torch.randn(S, 32, device="cuda", generator=g)
    dt = tcnn.grid_backward_params(xyzs, dy, enc.grid, aabb=aabb)
    (dt_ref,) = torch.autograd.grad(ref(xn), ref.params, dy)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    assert rel(dt, dt_ref) < 1e-3          # the same weight error eps*scale_l enters every scattered term (tests/test_tcnn_gpu.py)
    # linearity of the scatter (size-independent property): scatter(2 dy) = 2 scatter(dy)
    assert rel(tcnn.grid_backward_params(xyzs, 2 * dy, enc.grid, aabb=aabb), 2 * dt) < 1e-5
    del dy, dt, dt_ref, y_ref
    # fused density path + colour net at full size against the bf16-operand restatement
    sig_net = tcnn.Network(32, 16, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 1}).cuda()
    rgb_net = tcnn.Network(32, 3, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "Sigmoid", "n_neurons": 64, "n_hidden_layers": 2}).cuda()
    with torch.no_grad():
        h, sigma = sig_net.forward_density_field(xyzs, enc, aabb)
        rgb = rgb_net.forward_segments([dirs, h], [1, 0])
        h_ref = tcnn_oracle.mlp_forward(y, sig_net.params, 32, 64, 1, 16, operand_dtype=torch.bfloat16)
        dn = torch.nn.functional.normalize(dirs, dim=-1, eps=1e-6)
        rgb_ref = tcnn_oracle.mlp_forward(torch.cat([tcnn_oracle.sh_encode((dn + 1) / 2, 4), h], 1), rgb_net.params, 32, 64, 2, 3, "ReLU", "Sigmoid",
                                          operand_dtype=torch.bfloat16)
    scale = float(h_ref.abs().max())
    assert float((h - h_ref).abs().max()) < 2e-3 * scale
    assert torch.allclose(sigma, torch.exp(h[:, 0]), rtol=1e-5)
    assert float((rgb - rgb_ref).abs().max()) < 2e-3
