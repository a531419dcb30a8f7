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


def test_lazy_adam_skips_untouched_entries_and_matches_dense_on_the_touched_ones():
    """tiny-cuda-nn's rule for encoding parameters (ngp_adam_step_lazy): entries with an exactly-zero gradient keep p, m, v."""
    from ngp_b200.optim import FusedAdam
    g = torch.Generator(device="cuda").manual_seed(0)
    n = (1 << 20) + 12
    p1 = torch.randn(n, device="cuda", generator=g).requires_grad_(True)
    p2 = p1.detach().clone().requires_grad_(True)
    o1 = FusedAdam([p1], lr=1e-2, eps=1e-15, lazy=True)
    o2 = torch.optim.Adam([p2], lr=1e-2, eps=1e-15)
    touched_ever = torch.zeros(n, dtype=torch.bool, device="cuda")
    always = torch.rand(n, device="cuda", generator=g) < 0.3            # entries touched in EVERY step: dense == lazy there
    for it in range(4):
        sometimes = torch.rand(n, device="cuda", generator=g) < 0.2
        gr = torch.randn(n, device="cuda", generator=g) * (always | sometimes)
        p1.grad = gr.clone(); p2.grad = gr.clone()
        before = p1.detach().clone()
        o1.step(); o2.step()
        untouched = gr == 0
        assert torch.equal(p1.detach()[untouched], before[untouched])     # no coasting on old momentum
        touched_ever |= ~untouched
    st = o1.state[p1]
    assert bool((st["exp_avg"][~touched_ever] == 0).all()) and bool((st["exp_avg_sq"][~touched_ever] == 0).all())
    assert torch.allclose(p1.detach()[always], p2.detach()[always], rtol=2e-5, atol=2e-6)


def test_cosine_schedule_drives_the_fused_optimiser():
    """CosineAnnealingLR(net_opt, T_max, lr/30) as train.py:249-251 sets it up: the kernel reads the group's lr every step."""
    from ngp_b200.optim import FusedAdam
    g = torch.Generator(device="cuda").manual_seed(0)
    p1 = torch.randn(4099, device="cuda", generator=g).requires_grad_(True)
    p2 = p1.detach().clone().requires_grad_(True)
    o1 = FusedAdam([p1], lr=1e-2, eps=1e-8); o2 = torch.optim.Adam([p2], lr=1e-2, eps=1e-8)
    s1 = torch.optim.lr_scheduler.CosineAnnealingLR(o1, 6, 1e-2 / 30); s2 = torch.optim.lr_scheduler.CosineAnnealingLR(o2, 6, 1e-2 / 30)
    for it in range(6):
        gr = torch.randn(4099, device="cuda", generator=g)
        p1.grad = gr.clone(); p2.grad = gr.clone()
        o1.step(); o2.step(); s1.step(); s2.step()
        assert torch.allclose(p1, p2, rtol=2e-5, atol=2e-6)
    assert abs(o1.param_groups[0]["lr"] - 1e-2 / 30) < 1e-9
