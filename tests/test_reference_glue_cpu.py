"""CPU checks of the reference-arm harness (baseline/): the installed reference glue is byte-identical to
/root/reference, the import shims work, and the plain-torch `tinycudann` stand-in agrees with the oracle and starts
from the same weights as ngp_b200.tcnn (so matched-PSNR comparisons start from identical models)."""
import os

import numpy as np
import pytest
import torch

from oracle import tcnn_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "MANIFEST.json")) or os.path.isdir("/root/reference")
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed and /root/reference absent")


@needs_ref
def test_installed_reference_glue_is_unmodified():
    from baseline import install_ref
    assert install_ref.install() is not None
    assert install_ref.verify()
    for rel in install_ref.FILES:
        assert os.path.exists(os.path.join(install_ref.OUT, rel))


@needs_ref
def test_reference_modules_import_with_switchable_back_ends():
    from baseline import ref_harness
    g = ref_harness.load(vren="ours", tcnn="standin")
    for name in ("RayAABBIntersector", "RayMarcher", "VolumeRenderer", "RefLoss", "TruncExp"):
        assert hasattr(g.custom_functions, name)
    assert callable(g.rendering.render) and g.rendering.MAX_SAMPLES == 1024
    assert g.losses.NeRFLoss().lambda_distortion == 3e-4
    # the proxies forward to whichever back end is selected
    import ngp_b200.vren as ours
    assert g.vren.morton3D is ours.morton3D
    from baseline import tcnn_standin
    assert g.tcnn.Encoding is tcnn_standin.Encoding
    g.use(vren="ours", tcnn="ours")
    import ngp_b200.tcnn as ours_tcnn
    assert g.tcnn.Encoding is ours_tcnn.Encoding


def test_torch_scatter_shim_segment_csr():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "baseline", "shims"))
    try:
        sys.modules.pop("torch_scatter", None)
        from torch_scatter import segment_csr
    finally:
        sys.path.pop(0)
    src = torch.arange(12.0).reshape(6, 2)
    out = segment_csr(src, torch.tensor([0, 2, 2, 6]))
    assert torch.equal(out, torch.stack([src[0:2].sum(0), torch.zeros(2), src[2:6].sum(0)]))


def test_standin_grid_matches_oracle_and_double_backward():
    from baseline import tcnn_standin
    cfg = {"otype": "HashGrid", "n_levels": 6, "n_features_per_level": 4, "log2_hashmap_size": 12, "base_resolution": 4,
           "per_level_scale": 1.7}
    enc = tcnn_standin.Encoding(3, cfg).double()
    with torch.no_grad():
        enc.params.mul_(3000.0)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(257, 3, generator=g, dtype=torch.float64).requires_grad_(True)
    y = enc(x.float()) if False else enc.grid.encode(x, enc.params)
    y_o = tcnn_oracle.grid_encode(x, enc.params, 6, 4, 12, 4, 1.7)
    assert torch.allclose(y, y_o, rtol=1e-12, atol=1e-14)
    (gx,) = torch.autograd.grad(y.sum(), x, create_graph=True)
    (gx_o,) = torch.autograd.grad(y_o.sum(), x, create_graph=True)
    assert torch.allclose(gx, gx_o, rtol=1e-10, atol=1e-12)
    (gp,) = torch.autograd.grad((gx ** 2).sum(), enc.params)          # the double backward the normals need
    (gp_o,) = torch.autograd.grad((gx_o ** 2).sum(), enc.params)
    assert torch.allclose(gp, gp_o, rtol=1e-9, atol=1e-12)


def test_standin_starts_from_the_same_weights_as_ours():
    from baseline import tcnn_standin
    import ngp_b200.tcnn as ours
    gcfg = {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 14, "base_resolution": 16,
            "per_level_scale": 1.38}
    a, b = tcnn_standin.Encoding(3, gcfg), ours.Encoding(3, gcfg)
    assert a.params.shape == b.params.shape and torch.equal(a.params, b.params)
    for n_in, n_out, hid, act in ((32, 16, 1, "None"), (32, 3, 2, "Sigmoid"), (152, 3, 1, "Sigmoid")):
        ncfg = {"otype": "CutlassMLP", "activation": "ReLU", "output_activation": act, "n_neurons": 64, "n_hidden_layers": hid}
        a, b = tcnn_standin.Network(n_in, n_out, ncfg), ours.Network(n_in, n_out, ncfg)
        assert a.params.shape == b.params.shape and torch.equal(a.params, b.params)


def test_standin_sh_and_mlp_follow_oracle():
    from baseline import tcnn_standin
    g = torch.Generator().manual_seed(1)
    v = torch.rand(100, 3, generator=g)
    assert torch.equal(tcnn_standin.Encoding(3, {"otype": "SphericalHarmonics", "degree": 4})(v), tcnn_oracle.sh_encode(v, 4))
    net = tcnn_standin.Network(24, 5, {"otype": "CutlassMLP", "activation": "ReLU", "output_activation": "Sigmoid",
                                       "n_neurons": 32, "n_hidden_layers": 2})
    x = torch.randn(50, 24, generator=g)
    assert torch.equal(net(x), tcnn_oracle.mlp_forward(x, net.params, 24, 32, 2, 5, "ReLU", "Sigmoid"))
    assert net(x).shape == (50, 5)
