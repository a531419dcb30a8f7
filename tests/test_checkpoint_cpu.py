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
