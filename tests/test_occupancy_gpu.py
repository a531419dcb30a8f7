"""Occupancy update as one kernel chain (csrc/occupancy.cu ngp_occupancy_sample / ngp_occupancy_update)
against a plain-torch restatement of models/networks.py:308-333,379-408 fed with the SAME sampled cells, plus the
properties of the sampler the reference's randint / nonzero()[randint] draws have."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = 128
THR = 0.01 * 1024 / 3 ** 0.5          # train.py:248 density_threshold = 0.01 * MAX_SAMPLES / 3**0.5


def _model(scale, seed=0):
    """Occupancy mixin around an analytic density (a soft ball) with a random grid incl. invisible (-1) and zero cells."""
    from ngp_b200.networks import _OccupancyMixin

    class Field(torch.nn.Module, _OccupancyMixin):
        def __init__(self):
            super().__init__()
            self._init_occupancy(scale)

        def density(self, x):
            return 40.0 * torch.exp(-((x ** 2).sum(-1)) / 0.08)

    m = Field().cuda()
    g = torch.Generator(device="cuda").manual_seed(seed)
    m.density_grid.copy_(torch.rand(m.cascades, G ** 3, device="cuda", generator=g) * 12.0)
    m.density_grid[:, ::7] = -1.0                      # invisible cells (mark_invisible_cells)
    m.density_grid[:, 3::11] = 0.0
    return m


@pytest.mark.parametrize("scale", [0.5, 8.0])
def test_sampler_draws_uniform_and_occupied_cells(scale):
    from ngp_b200 import vren
    m = _model(scale)
    Cc, M = m.cascades, G ** 3 // 4
    idx, xyz, n = m.sample_cells_fused(THR, warmup=False, seed=1234)
    assert n == 2 * M and idx.shape == (Cc * n,) and xyz.shape == (Cc * n, 3)
    idx = idx.view(Cc, n).long(); xyz = xyz.view(Cc, n, 3)
    assert int(idx.min()) >= 0 and int(idx.max()) < G ** 3
    for c in range(Cc):
        occ = m.density_grid[c] > THR
        # second half: every draw is an occupied cell, and the draws spread over ALL occupied cells about evenly
        i2 = idx[c, M:]
        assert bool(occ[i2].all())
        hits = torch.bincount(i2, minlength=G ** 3)[occ].float()
        expect = M / float(occ.sum())
        assert abs(float(hits.mean()) - expect) < 1e-3 * expect + 1e-6
        assert float(hits.std()) < 1.3 * expect ** 0.5 + 0.05        # Poisson-like spread, no preferred cells
        # first half: uniform over the lattice (each coordinate uniform on 0..127)
        c1 = vren.morton3D_invert(idx[c, :M].int()).float()
        assert torch.allclose(c1.mean(0), torch.full((3,), 63.5, device="cuda"), atol=0.35)
        assert int(c1.min()) == 0 and int(c1.max()) == G - 1
        # positions: cell centre of the draw's cell +- half a cell, in this cascade's extent (networks.py:389-395)
        s = min(2 ** (c - 1), scale); hgs = s / G
        coords = vren.morton3D_invert(idx[c].int()).float()
        centre = (coords / (G - 1) * 2 - 1) * (s - hgs)
        off = xyz[c] - centre
        assert float(off.abs().max()) <= hgs * (1 + 1e-4)
        assert abs(float(off.mean())) < 0.01 * hgs and float(off.std()) > 0.55 * hgs     # uniform jitter: std = hgs / sqrt(3)
    # a different seed gives different draws, the same seed the same ones
    a, _, _ = m.sample_cells_fused(THR, seed=1234)
    b, _, _ = m.sample_cells_fused(THR, seed=99)
    assert torch.equal(a.view(Cc, n).long(), idx) and not torch.equal(a, b)


def test_sampler_warmup_visits_every_cell_once_and_empty_cascade_draws_are_void():
    m = _model(8.0)
    Cc = m.cascades
    idx, xyz, n = m.sample_cells_fused(THR, warmup=True, seed=7)
    assert n == G ** 3
    idx = idx.view(Cc, n)
    assert torch.equal(idx, torch.arange(G ** 3, device="cuda", dtype=torch.int32).expand(Cc, -1))
    m.density_grid[2].fill_(0.0)                      # cascade 2 has no occupied cell: the reference's list is empty
    idx, _, n = m.sample_cells_fused(THR, warmup=False, seed=7)
    idx = idx.view(Cc, n)
    assert bool((idx[2, n // 2:] == -1).all()) and bool((idx[2, :n // 2] >= 0).all())
    assert bool((idx[1] >= 0).all()) and bool((idx[3] >= 0).all())


@pytest.mark.parametrize("scale,warmup,erode", [(0.5, False, False), (8.0, False, False), (8.0, True, False), (0.5, False, True)])
def test_update_equals_the_torch_restatement_on_the_same_cells(scale, warmup, erode):
    from ngp_b200 import vren
    m = _model(scale, seed=3)
    Cc = m.cascades
    if erode:
        g = torch.Generator(device="cuda").manual_seed(5)
        m.count_grid = torch.rand(Cc, G ** 3, device="cuda", generator=g).clamp_(min=0.02)
    grid0 = m.density_grid.clone()
    idx, xyz, n = m.sample_cells_fused(THR, warmup=warmup, seed=42)
    sig = m.density(xyz)
    # --- restatement (networks.py:397-408); duplicates of a cell keep the largest density (documented deviation: torch's
    # index_put winner is unspecified)
    tmp = torch.zeros_like(grid0)
    flat = (idx.view(Cc, n).long() + torch.arange(Cc, device="cuda")[:, None] * G ** 3).reshape(-1)
    tmp.view(-1).scatter_reduce_(0, flat, sig, reduce="amax", include_self=True)
    decay = 0.95
    if erode:
        decay = torch.clamp(decay ** (1 / m.count_grid), 0.1, 0.95)
    want = torch.where(grid0 < 0, grid0, torch.maximum(grid0 * decay, tmp))
    mean = float(want[want > 0].mean())
    want_bits = torch.zeros_like(m.density_bitfield)
    # --- the chain
    m.update_density_grid(THR, warmup=warmup, erode=erode, seed=42)
    assert torch.allclose(m.density_grid, want, rtol=1e-5 if erode else 0, atol=0)
    thr = min(mean, THR)
    vren.packbits(m.density_grid, thr, want_bits)
    diff = int((want_bits != m.density_bitfield).sum())
    if diff:                        # only cells within float rounding of the (re-associated) mean may flip
        near = ((m.density_grid - thr).abs() < 1e-5 * abs(thr)).sum()
        assert diff <= int(near)
    assert int(torch.count_nonzero(m.density_bitfield)) > 0


def test_update_is_reusable_and_matches_the_unfused_path_statistically():
    """Two models from the same grid, one updated by the chain, one by the torch-op path (different random streams):
    after a few updates the occupied sets agree up to the sampling noise of the cells near the threshold."""
    a, b = _model(0.5, seed=9), _model(0.5, seed=9)
    b.fused_update = False
    torch.manual_seed(0)
    for it in range(4):
        a.update_density_grid(THR, warmup=it == 0)
        b.update_density_grid(THR, warmup=it == 0)
    pa = np.unpackbits(a.density_bitfield.cpu().numpy())
    pb = np.unpackbits(b.density_bitfield.cpu().numpy())
    assert pa.sum() > 0 and abs(int(pa.sum()) - int(pb.sum())) < 0.02 * pb.sum() + 50
    assert (pa != pb).mean() < 0.02
