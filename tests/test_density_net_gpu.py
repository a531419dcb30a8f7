"""The reference's density net as one tcgen05 kernel per direction (csrc/density_net.cu) against the
reference's own formulation — Linear(128,128) -> Softplus -> Linear(128,1) -> Softplus with the normals taken by
torch.autograd.grad(create_graph=True) and everything back-propagated by autograd (models/networks.py:54-59,172-196) —
evaluated in fp64.  Tolerance: bf16 tensor-core operands (8-bit mantissa: 2e-3 per rounded operand), fp32 accumulation:
1e-2 of each tensor's norm; sigma itself (one rounded product chain) 5e-3."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _oracle(e, W1, b1, W2, b2, ds, dg):
    po = [t.double().clone().requires_grad_(True) for t in (e, W1, b1, W2, b2)]
    sig = F.softplus(F.linear(F.softplus(F.linear(po[0], po[1], po[2])), po[3], po[4]))[:, 0]
    (ge,) = torch.autograd.grad(sig, po[0], torch.ones_like(sig), create_graph=True)
    loss = 0
    if ds is not None:
        loss = loss + (sig * ds.double()).sum()
    if dg is not None:
        loss = loss + (ge * dg.double()).sum()
    return sig.detach(), ge.detach(), torch.autograd.grad(loss, po)


def _inputs(n, seed=7, scale_e=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    e, W1, b1, W2, b2 = rnd(n, 128) * scale_e, rnd(128, 128) * 0.15, rnd(128) * 0.5, rnd(1, 128) * 0.3, rnd(1)
    # dsigma with a positive mean: db2 = sum_rows dz2 must not be a cancelling sum of +-1 terms, or its RELATIVE error measures the
    # cancellation instead of the kernel
    return e, W1, b1, W2, b2, rnd(n) + 1.0, rnd(n, 128)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 4099, 70001])
def test_density_net_both_outputs_and_all_gradients(n):
    from ngp_b200.networks import _DensityNormalsFn
    e, W1, b1, W2, b2, ds, dg = _inputs(n)
    if n > 200:
        e[0] = 6.0                                     # pre-activations past softplus' linear threshold
        e[1] = -6.0
    ps = [t.clone().requires_grad_(True) for t in (e, W1, b1, W2, b2)]
    sig, ge = _DensityNormalsFn.apply(*ps, "tc05")
    grads = torch.autograd.grad((sig * ds).sum() + (ge * dg).sum(), ps)
    o_sig, o_ge, o_grads = _oracle(e, W1, b1, W2, b2, ds, dg)
    assert sig.shape == (n,) and ge.shape == (n, 128)
    assert rel(sig, o_sig) < 5e-3 and rel(ge, o_ge) < 1e-2, (rel(sig, o_sig), rel(ge, o_ge))
    for name, a, b in zip(("e", "W1", "b1", "W2", "b2"), grads, o_grads):
        assert a.shape == b.shape and torch.isfinite(a).all(), name
        assert rel(a, b) < 1e-2, (name, rel(a, b))


@pytest.mark.parametrize("which", ["sigma_only", "normals_only"])
def test_density_net_single_upstream(which):
    """Density evaluation without normals (dsigma only) and the normal losses alone (d g_e only)."""
    from ngp_b200.networks import _DensityNormalsFn
    n = 3001
    e, W1, b1, W2, b2, ds, dg = _inputs(n, seed=3)
    ps = [t.clone().requires_grad_(True) for t in (e, W1, b1, W2, b2)]
    sig, ge = _DensityNormalsFn.apply(*ps, "tc05")
    loss = (sig * ds).sum() if which == "sigma_only" else (ge * dg).sum()
    grads = torch.autograd.grad(loss, ps)
    _, _, o_grads = _oracle(e, W1, b1, W2, b2, ds if which == "sigma_only" else None, dg if which == "normals_only" else None)
    for name, a, b in zip(("e", "W1", "b1", "W2", "b2"), grads, o_grads):
        assert rel(a, b) < 1e-2, (which, name, rel(a, b))


def test_density_net_forward_without_normals_equals_the_full_forward():
    from ngp_b200.networks import _dn_fw
    e, W1, b1, W2, b2, _, _ = _inputs(5000, seed=11)
    s_a, s2_a, ge = _dn_fw(e, W1, b1, W2, b2, want_ge=True)
    s_b, s2_b, none = _dn_fw(e, W1, b1, W2, b2, want_ge=False)
    assert none is None and torch.equal(s_a, s_b) and torch.equal(s2_a, s2_b)
    with pytest.raises(RuntimeError, match="128 -> 128"):
        _dn_fw(e[:, :96].contiguous(), W1[:, :96].contiguous(), b1, W2, b2)


def _field(tc, seed=0):
    from ngp_b200.networks import NGP
    torch.manual_seed(seed)
    m = NGP(scale=0.5, grid_levels=16, grid_features=8, log2_T_xyz=14, log2_T_rgb=14, density_net_tc=tc).cuda()
    with torch.no_grad():
        m.xyz_encoder.params.mul_(3000.0)
        m.rgb_encoder.params.mul_(3000.0)
    return m


def test_reference_literal_field_tc_density_net_matches_the_torch_gemm_path():
    """NGP.forward (sigma, normals from the double backward) and the gradients of every parameter the density branch feeds, tensor-core
    density net vs the TF32 torch-GEMM path of round 1 (itself checked against autograd in test_model_gpu)."""
    a, b = _field(True), _field(False)
    b.load_state_dict(a.state_dict())
    assert a._density_tc() and not b._density_tc()
    g = torch.Generator(device="cuda").manual_seed(1)
    n = 6000
    x = (torch.rand(n, 3, device="cuda", generator=g) - 0.5) * 0.98
    d = torch.randn(n, 3, device="cuda", generator=g)
    wn = torch.randn(n, 3, device="cuda", generator=g)
    outs = []
    for m in (a, b):
        sig, rgb, n_raw, n_pred, sem = m(x, d)
        loss = (sig * 0.1).sum() + (n_raw * wn).sum() + rgb.sum() * 0.01
        m.zero_grad()
        loss.backward()
        outs.append((sig.detach(), n_raw.detach(), {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}))
    (s1, n1, g1), (s2, n2, g2) = outs
    assert rel(s1, s2) < 5e-3
    cos = (n1 * n2).sum(-1)
    assert float(cos.median()) > 0.9995 and float((cos < 0.98).float().mean()) < 0.02
    assert set(g1) == set(g2)
    for k in g2:                        # through normalize(): the normal term amplifies operand rounding where |d sigma / dx| is small
        if k.startswith("xyz_"):
            assert rel(g1[k], g2[k]) < 6e-2, (k, rel(g1[k], g2[k]))
    # the graph-free density() of the occupancy update takes the same kernel
    with torch.no_grad():
        assert rel(a.density(x), s1) < 1e-6
