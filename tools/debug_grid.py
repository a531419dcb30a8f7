import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "instant-ngp-pp_b200"))
import numpy as np, torch
from ngp_b200 import tcnn
from oracle import tcnn_oracle
torch.manual_seed(0)
for (L, F, T, base, scale) in [(16,2,19,16,0.5),(16,8,19,16,0.5),(16,4,19,16,0.5),(16,1,19,16,0.5),(1,8,19,16,0.5),(4,8,19,16,0.5),(5,1,12,4,1.0),(8,8,14,8,1.0)]:
    b = float(np.exp(np.log(2048 * scale / base) / max(L - 1, 1)))
    enc = tcnn.Encoding(3, {"otype": "HashGrid", "n_levels": L, "n_features_per_level": F, "log2_hashmap_size": T,
                            "base_resolution": base, "per_level_scale": b}).cuda()
    with torch.no_grad():
        enc.params.copy_(torch.randn_like(enc.params) * 0.5)
    n = 4000
    x = torch.rand(n, 3, device="cuda")
    dy = torch.randn(n, L * F, device="cuda")
    gx = tcnn.grid_backward_input(x, dy, enc.params.detach(), enc.grid)
    xo = x.double().requires_grad_(True)
    yo = tcnn_oracle.grid_encode(xo, enc.params.detach().double(), L, F, T, base, b)
    (gxo,) = torch.autograd.grad(yo, xo, dy.double())
    err = (gx.double() - gxo).norm(dim=1) / (gxo.norm(dim=1) + 1e-9)
    rel = float((gx.double() - gxo).norm() / gxo.norm())
    bad = int((err > 1e-2).sum())
    print(f"L{L} F{F} T{T} base{base}: rel {rel:.2e}  samples with >1% err: {bad}/{n}  median err {float(err.median()):.1e} max {float(err.max()):.2e}")
