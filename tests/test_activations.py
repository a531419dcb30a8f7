"""TruncExp / TruncTanh / ReLU (reference models/custom_functions.py:200-244) — forward and backward
against the reference's own classes (baseline/_ref, unmodified) and against the closed forms, including the clamp
boundaries (+-7, +-15) and the reference ReLU's constant 1e-6 zero-side gradient."""
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "MANIFEST.json")) or os.path.isdir("/root/reference")


def _x(dev):
    g = torch.Generator().manual_seed(0)
    x = torch.cat([torch.randn(500, 3, generator=g) * 6, torch.tensor([[-7.0, 7.0, 0.0], [-15.0, 15.0, -0.0], [20.0, -20.0, 7.0001], [-7.0001, 15.5, -16.0]])])
    return x.to(dev).requires_grad_(True)


def _check(dev, ref_cf=None):
    from ngp_b200.custom_functions import TruncExp, TruncTanh, ReLU
    up = torch.randn(504, 3, generator=torch.Generator().manual_seed(1)).to(dev)
    closed = {
        "TruncExp": (lambda x: torch.exp(x), lambda x, g: g * torch.exp(x.clamp(-7, 7))),
        "TruncTanh": (lambda x: torch.tanh(x), lambda x, g: g * (1 - torch.tanh(x.clamp(-15, 15)) ** 2)),
        "ReLU": (lambda x: torch.where(x > 0, x, torch.zeros_like(x)), lambda x, g: torch.where(x > 0, g, torch.full_like(g, 1e-6))),
    }
    for name, fn in (("TruncExp", TruncExp), ("TruncTanh", TruncTanh), ("ReLU", ReLU)):
        x = _x(dev)
        y = fn.apply(x)
        (gx,) = torch.autograd.grad(y, x, up)
        fw, bw = closed[name]
        assert torch.equal(y.detach(), fw(x.detach())), name
        assert torch.equal(gx, bw(x.detach(), up)), name
        if ref_cf is not None:                      # the reference's own class on the same tensors (its ReLU needs CUDA tensors)
            xr = _x(dev)
            yr = getattr(ref_cf, name).apply(xr)
            (gr,) = torch.autograd.grad(yr, xr, up)
            assert torch.equal(y.detach(), yr.detach()) and torch.equal(gx, gr), name


def test_activations_closed_forms_cpu():
    _check("cpu")


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed")
def test_truncexp_trunctanh_match_reference_classes_cpu():
    from baseline import ref_harness
    cf = ref_harness.load(vren="ours", tcnn="standin").custom_functions
    from ngp_b200.custom_functions import TruncExp, TruncTanh
    up = torch.randn(504, 3, generator=torch.Generator().manual_seed(1))
    for name, fn in (("TruncExp", TruncExp), ("TruncTanh", TruncTanh)):
        x, xr = _x("cpu"), _x("cpu")
        (g1,) = torch.autograd.grad(fn.apply(x), x, up)
        (g2,) = torch.autograd.grad(getattr(cf, name).apply(xr), xr, up)
        assert torch.equal(g1, g2), name


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed")
def test_activations_match_reference_classes_gpu():
    from baseline import ref_harness
    _check("cuda", ref_harness.load(vren="ours", tcnn="standin").custom_functions)


@pytest.mark.gpu
def test_neg_normalize_matches_torch_normalize_forward_and_backward():
    """NegNormalize (csrc/normals.cu) = -F.normalize(x * scale, eps=1e-6) (models/networks.py:209,222-223), incl. rows below eps (the
    clamp branch: Jacobian -I/eps), zero rows and the folded unit-cube -> world scale."""
    import torch.nn.functional as F
    from ngp_b200.custom_functions import NegNormalize
    g = torch.Generator(device="cuda").manual_seed(0)
    n = 100003
    x = torch.randn(n, 3, device="cuda", generator=g) * 3
    x[:50] *= 1e-8                                  # below eps
    x[50:60] = 0.0
    x[60:70, 1:] = 0.0                              # axis-aligned
    gy = torch.randn(n, 3, device="cuda", generator=g)
    for scale in (None, (1.0 / 16, 1.0 / 16, 1.0 / 16), (0.5, 2.0, 0.125)):
        a = x.clone().requires_grad_(True)
        b = x.clone().requires_grad_(True)
        ya = NegNormalize.apply(a, scale, 1e-6)
        v = b if scale is None else b * torch.tensor(scale, device="cuda")
        yb = -F.normalize(v, p=2, dim=-1, eps=1e-6)
        assert torch.allclose(ya, yb, rtol=1e-6, atol=1e-7)
        (ga,) = torch.autograd.grad(ya, a, gy)
        (gb,) = torch.autograd.grad(yb, b, gy)
        # torch differentiates clamp_min(norm, eps) with a zero gradient on clamped rows, i.e. d(v/eps) = I/eps there as well
        assert torch.allclose(ga, gb, rtol=2e-5, atol=1e-6 * float(gb.abs().max())), float((ga - gb).abs().max())


@pytest.mark.gpu
def test_refloss_prep_matches_the_torch_ops_forward_and_backward():
    """RefLossPrep (csrc/normals.cu) = the reference's per-sample inputs of RefLoss (models/rendering.py:243-246):
    (normals_raw - normals_pred)**2 and clamp(sum(normals_raw * normalize(dirs)), min=0)**2, with their gradients."""
    import torch.nn.functional as F
    from ngp_b200.custom_functions import RefLossPrep
    g = torch.Generator(device="cuda").manual_seed(1)
    n = 70001
    raw = F.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=-1)
    pred = F.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=-1)
    dirs = torch.randn(n, 3, device="cuda", generator=g) * 2
    dirs[:5] = 0.0                                                   # normalize's eps branch
    gd, go = torch.randn(n, 3, device="cuda", generator=g), torch.randn(n, device="cuda", generator=g)
    a = [raw.clone().requires_grad_(True), pred.clone().requires_grad_(True)]
    b = [raw.clone().requires_grad_(True), pred.clone().requires_grad_(True)]
    da, oa = RefLossPrep.apply(a[0], a[1], dirs)
    db = (b[0] - b[1]) ** 2
    ob = torch.clamp(torch.sum(b[0] * F.normalize(dirs, p=2, dim=-1, eps=1e-6), dim=-1), min=0.) ** 2
    assert torch.allclose(da, db, rtol=1e-6, atol=1e-7) and torch.allclose(oa, ob, rtol=1e-5, atol=1e-7)
    ga = torch.autograd.grad((da * gd).sum() + (oa * go).sum(), a)
    gb = torch.autograd.grad((db * gd).sum() + (ob * go).sum(), b)
    for x, y in zip(ga, gb):
        assert torch.allclose(x, y, rtol=1e-5, atol=1e-6)
    (g1,) = torch.autograd.grad((RefLossPrep.apply(a[0], a[1], dirs)[1] * go).sum(), a[0])      # one output unused
    (g2,) = torch.autograd.grad((torch.clamp(torch.sum(b[0] * F.normalize(dirs, p=2, dim=-1, eps=1e-6), dim=-1), min=0.) ** 2 * go).sum(), b[0])
    assert torch.allclose(g1, g2, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_fused_basic_loss_equals_the_sum_of_means():
    """NeRFLoss.total (ngp_basic_loss: photometric + opacity terms and their gradients in one kernel) = sum(v.mean()) over the loss
    dictionary (losses.py:89-96, train.py:310), value and gradients, incl. opacity 0 (log(1e-10)) and 1."""
    from ngp_b200.losses import NeRFLoss
    g = torch.Generator(device="cuda").manual_seed(2)
    n = 262144 + 7
    rgb = torch.rand(n, 3, device="cuda", generator=g)
    tgt = torch.rand(n, 3, device="cuda", generator=g)
    opa = torch.rand(n, device="cuda", generator=g)
    opa[:100] = 0.0; opa[100:200] = 1.0
    fn = NeRFLoss(lambda_opa=2e-4, lambda_distortion=0)
    a = [rgb.clone().requires_grad_(True), opa.clone().requires_grad_(True)]
    b = [rgb.clone().requires_grad_(True), opa.clone().requires_grad_(True)]
    la = fn.total({"rgb": a[0], "opacity": a[1]}, {"rgb": tgt})
    lb = sum(v.mean() for v in fn({"rgb": b[0], "opacity": b[1]}, {"rgb": tgt}).values())
    assert abs(float(la) - float(lb)) < 1e-6 * abs(float(lb)) + 1e-9
    ga = torch.autograd.grad(la * 3.0, a)
    gb = torch.autograd.grad(lb * 3.0, b)
    assert torch.allclose(ga[0], gb[0], rtol=1e-5, atol=1e-12) and torch.allclose(ga[1], gb[1], rtol=1e-5, atol=1e-12)
