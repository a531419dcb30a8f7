"""Launches the density-net kernels once each on 2 M samples (for ncu; never a bench value)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "instant-ngp-pp_b200"))
import torch
from ngp_b200.networks import _dn_fw, _dn_bw
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
g = torch.Generator(device=dev).manual_seed(0)
e = torch.randn(n, 128, device=dev, generator=g) * 0.3
W1 = torch.randn(128, 128, device=dev, generator=g) * 0.1; b1 = torch.randn(128, device=dev, generator=g)
W2 = torch.randn(1, 128, device=dev, generator=g) * 0.3; b2 = torch.randn(1, device=dev, generator=g)
dge = torch.randn(n, 128, device=dev, generator=g); ds = torch.randn(n, device=dev, generator=g)
for _ in range(2):
    sig, s2, ge = _dn_fw(e, W1, b1, W2, b2, True)
    _dn_bw(e, dge, ge, ds, s2, W1, b1, W2, True)
torch.cuda.synchronize()
print("ok")
