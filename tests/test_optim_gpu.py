"""FusedAdam (ngp_adam_step / ngp_grad_sumsq / ngp_clip_coef) against torch.optim.Adam +
torch.nn.utils.clip_grad_norm_ — the reference's optimiser path (train.py:244-251, 435)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("clip,world", [(None, 1), (50.0, 1), (0.5, 1), (0.5, 4)])
def test_fused_adam_matches_torch(clip, world):
    from ngp_b200.optim import FusedAdam
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(1000003,), (64, 32), (7,)]
    p1 = [torch.randn(s, device="cuda", generator=g).requires_grad_(True) for s in shapes]
    p2 = [p.detach().clone().requires_grad_(True) for p in p1]
    o1 = FusedAdam(p1, lr=1e-2, eps=1e-15, max_grad_norm=clip, grad_scale=1.0 / world)
    o2 = torch.optim.Adam(p2, lr=1e-2, eps=1e-15)
    for it in range(5):
        for a, b in zip(p1, p2):
            gr = torch.randn(a.shape, device="cuda", generator=g) * (0.1 if it % 2 else 3.0)
            a.grad = gr.clone() * world                  # summed over ranks
            b.grad = gr.clone()                          # averaged
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(p2, clip)
        o1.step(); o2.step()
        for a, b in zip(p1, p2):
            assert torch.allclose(a, b, rtol=2e-5, atol=2e-6), float((a - b).abs().max())
