"""Checkpoint plumbing (reference utils.py:7-42): Lightning-style prefixes, ignored sub-prefixes, slimming, and the
padded-first-layer adaptation.  CPU only: modules construct without a GPU, nothing is launched."""
import io

import pytest
import torch

from ngp_b200 import checkpoint
from ngp_b200.networks import NGP


def _model():
    return NGP(scale=0.5, grid_levels=4, grid_features=2, log2_T_xyz=10, log2_T_rgb=11, classes=5)


def test_lightning_checkpoint_roundtrip_with_ignored_prefixes():
    src, dst = _model(), _model()
    with torch.no_grad():
        for p in src.parameters():
            p.add_(torch.randn_like(p) * 0.1)
        src.density_bitfield.random_(0, 255)
    sd = {"model." + k: v.clone() for k, v in src.state_dict().items()}
    sd.update({"directions": torch.zeros(4, 3), "poses": torch.zeros(2, 3, 4), "val_lpips.net.w": torch.zeros(1),
               "embedding_a.weight": torch.randn(7, 8)})
    buf = io.BytesIO(); torch.save({"state_dict": sd, "epoch": 3}, buf); buf.seek(0)
    ckpt = torch.load(buf, map_location="cpu")
    slim = checkpoint.slim_ckpt({"state_dict": dict(ckpt["state_dict"])})
    assert "directions" not in slim and "poses" not in slim and "model.density_grid" not in slim and "val_lpips.net.w" not in slim
    before_grid = dst.density_grid.clone()
    checkpoint.load_ckpt(dst, {"state_dict": slim}, prefixes_to_ignore=["embedding_a", "density_grid", "grid_coords"])   # render.py:67
    for k, v in src.state_dict().items():
        if k in ("density_grid", "grid_coords"):
            continue
        assert torch.equal(dst.state_dict()[k], v), k
    assert torch.equal(dst.density_grid, before_grid)                  # ignored prefix keeps the module's own buffer
    emb = torch.nn.Embedding(7, 8)
    checkpoint.load_ckpt(emb, {"state_dict": slim}, model_name="embedding_a")           # render.py:60-64
    assert torch.equal(emb.weight, sd["embedding_a.weight"])


def test_padded_first_layer_is_adapted_or_refused():
    m = _model()
    mlp = m.norm_pred_header.mlp                      # 8 -> 32 -> 3 here
    ours = m.norm_pred_header.params.detach()
    W0 = ours[: mlp.width * mlp.n_in].reshape(mlp.width, mlp.n_in)
    padded = torch.cat([torch.cat([W0, torch.zeros(mlp.width, 8)], 1).reshape(-1), ours[mlp.width * mlp.n_in:]])
    assert torch.equal(checkpoint.adapt_mlp_params(padded, mlp.n_in, mlp.width, mlp.n_hidden, mlp.n_out), ours)
    sd = {"model." + k: v.clone() for k, v in m.state_dict().items()}
    sd["model.norm_pred_header.params"] = padded
    dst = _model()
    checkpoint.load_ckpt(dst, sd)
    assert torch.equal(dst.norm_pred_header.params.detach(), ours)
    bad = padded.clone(); bad[mlp.n_in] = 1.0          # a non-zero bias column cannot be represented
    with pytest.raises(RuntimeError, match="bias"):
        checkpoint.adapt_mlp_params(bad, mlp.n_in, mlp.width, mlp.n_hidden, mlp.n_out)
    sd["model.xyz_encoder.params"] = torch.zeros(17)
    with pytest.raises(RuntimeError, match="grid configuration"):
        checkpoint.load_ckpt(_model(), sd)


def test_cutlass_mlp_vector_with_8_row_output_padding_loads():
    """a tcnn CutlassMLP (the otype of every head in models/networks.py:89-162) pads its output rows to 8, not 16: the
    reference rgb_net 144 -> 128 -> 3 then has 128*144 + 8*128 values against 128*144 + 16*128 here."""
    import ngp_b200.tcnn as tcnn
    cfg = {"otype": "CutlassMLP", "activation": "ReLU", "output_activation": "Sigmoid", "n_neurons": 128, "n_hidden_layers": 1}
    net = tcnn.Network(144, 3, cfg)
    ours = net.params.detach()
    W0, Wl = ours[:128 * 144], ours[128 * 144:].reshape(16, 128)
    src = torch.cat([W0, Wl[:8].reshape(-1)])                       # tcnn layout: 8 output rows
    got = checkpoint.adapt_mlp_params(src, 144, 128, 1, 3, "CutlassMLP")
    assert got.numel() == ours.numel()
    assert torch.equal(got[:128 * 144], W0) and torch.equal(got[128 * 144:].reshape(16, 128)[:3], Wl[:3])
    assert float(got[128 * 144:].reshape(16, 128)[8:].abs().max()) == 0.0

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.rgb_net = tcnn.Network(144, 3, cfg)
    dst = Holder()
    checkpoint.load_ckpt(dst, {"model.rgb_net.params": src})
    x = torch.randn(5, 144)
    from oracle import tcnn_oracle
    assert torch.equal(tcnn_oracle.mlp_forward(x, dst.rgb_net.params.detach(), 144, 128, 1, 3, "ReLU", "Sigmoid"),
                       tcnn_oracle.mlp_forward(x, ours, 144, 128, 1, 3, "ReLU", "Sigmoid"))
    # a trained head with a padded INPUT (skybox 9 -> 16) carries a bias column: refused with a clear message
    sky = torch.randn(32 * 16 + 8 * 32)
    with pytest.raises(RuntimeError, match="unsupported"):
        checkpoint.adapt_mlp_params(sky, 9, 32, 1, 3, "CutlassMLP")


def test_depth_mono_scale_shift_uses_valid_pixel_count():
    """with a partial mask the solve must run over the valid pixels only (reference losses.py:7-30 on
    results['depth'][mask])."""
    from ngp_b200.losses import compute_scale_and_shift
    g = torch.Generator().manual_seed(0)
    pred = torch.rand(300, generator=g) * 3 + 0.5
    target = 2.0 * pred + 0.7
    w = (torch.arange(300) % 3 != 0).float()
    target = target * w                                  # invalid pixels carry depth 0
    s, t = compute_scale_and_shift(pred, target, weight=w)
    assert abs(float(s) - 2.0) < 1e-4 and abs(float(t) - 0.7) < 1e-4
    m = w.bool()                                         # the reference's formulation on the masked vectors
    p, q = pred[m], target[m]
    a00, a01, a11, b0, b1 = (p * p).sum(), p.sum(), float(p.numel()), (p * q).sum(), q.sum()
    det = a00 * a11 - a01 * a01
    assert torch.allclose(s, (a11 * b0 - a01 * b1) / det, rtol=1e-5) and torch.allclose(t, (-a01 * b0 + a00 * b1) / det, rtol=1e-4)
