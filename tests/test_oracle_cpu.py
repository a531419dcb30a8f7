"""CPU self-checks of the oracle (ngp_oracle.c / tcnn_oracle.py) against independent numpy/torch
formulations of the same definitions — run everywhere (-m "not gpu")."""
import numpy as np
import pytest
import torch

import cases
from oracle import oracle, tcnn_oracle


def test_morton_roundtrip_and_reference_bits():
    rng = np.random.RandomState(0)
    c = rng.randint(0, 1024, (5000, 3)).astype(np.int32)
    m = oracle.morton3D(c)
    assert (oracle.morton3D_invert(m) == c).all()
    ref = cases.morton_enc(c[:, 0], c[:, 1], c[:, 2]).astype(np.int64)
    assert (m.astype(np.int64) == ref).all()
    # x -> bit 0, y -> bit 1, z -> bit 2 (raymarching.cu:49)
    assert list(oracle.morton3D(np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [3, 0, 0]], np.int32))) == [1, 2, 4, 9]


def test_packbits_bit_order_and_strict_threshold():
    g = np.zeros(64, np.float32)
    g[0] = 2.0; g[9] = 2.0; g[23] = 1.0; g[63] = 5.0; g[40] = -1.0
    b = oracle.packbits(g, 1.0)
    assert list(b) == [1, 2, 0, 0, 0, 0, 0, 128]          # g[23]==thr is NOT set (strict >)
    rng = np.random.RandomState(1)
    g = rng.normal(size=4096).astype(np.float32)
    assert (oracle.packbits(g, 0.3) == np.packbits(g > 0.3, bitorder="little")).all()


def test_aabb_cases():
    o = np.array([[0, 0, -2], [0, 0, -2], [0, 0, 0], [2, 2, -2], [0, 0, 2]], np.float32)
    d = np.array([[0, 0, 1], [0, 1, 0], [0, 0, 1], [0, 0, 1], [0, 0, 1]], np.float32)
    cnt, t, idx = oracle.ray_aabb_intersect(o, d, np.zeros((1, 3), np.float32), np.full((1, 3), 0.5, np.float32), 1)
    assert list(cnt) == [1, 0, 1, 0, 0]
    assert np.allclose(t[0, 0], [1.5, 2.5]) and np.allclose(t[2, 0], [0.0, 0.5])     # inside: t1 clamped to 0
    assert (t[1] == -1).all() and (idx[1] == -1).all() and idx[0, 0] == 0
    # two boxes, max_hits 2: filler slot (-1) sorts first (torch::sort ascending, intersection.cu:95)
    c2 = np.array([[0, 0, 0], [0, 0, 3]], np.float32); h2 = np.full((2, 3), 0.5, np.float32)
    cnt, t, idx = oracle.ray_aabb_intersect(o[:1], d[:1], c2, h2, 2)
    assert cnt[0] == 2 and np.allclose(t[0, :, 0], [1.5, 4.5]) and list(idx[0]) == [0, 1]
    cnt, t, idx = oracle.ray_aabb_intersect(o[:1], d[:1], c2[:1], h2[:1], 2)
    assert cnt[0] == 1 and t[0, 0, 0] == -1 and np.isclose(t[0, 1, 0], 1.5)


@pytest.mark.parametrize("case", cases.MARCH_CASES[:3] + cases.MARCH_CASES[4:5], ids=lambda c: c[0])
def test_marching_invariants(case):
    name, kind, scale, casc, esf, n = case
    n = min(n, 256)
    bf = cases.bitfield(kind, casc, seed=1)
    o, d = cases.rays(n, scale, seed=2)
    cnt, ht, _ = oracle.ray_aabb_intersect(o, d, np.zeros((1, 3), np.float32), np.full((1, 3), scale, np.float32), 1)
    h = cases.near_clamp(ht)
    noise = np.random.RandomState(3).rand(n).astype(np.float32)
    ra, xyz, dirs, dl, ts, counter = oracle.raymarching_train(o, d, h, bf, casc, scale, esf, noise, 128, 1024)
    assert counter[0] == ra[:, 2].sum() and counter[1] == n
    assert (ra[:, 0] == np.arange(n)).all()
    assert (ra[:, 1] == np.concatenate([[0], np.cumsum(ra[:, 2])[:-1]])).all()
    assert (ra[cnt == 0, 2] == 0).all() and ra[:, 2].max() <= 1024
    ray = np.repeat(ra[:, 0], ra[:, 2])
    assert np.allclose(xyz, o[ray] + ts[:, None] * d[ray], atol=1e-5 * max(scale, 1))
    assert (dirs == d[ray]).all()
    first = np.ones(len(ts), bool); first[1:] = ray[1:] != ray[:-1]
    assert (np.diff(ts)[~first[1:]] > 0).all()                         # t strictly increases along a ray
    assert (ts >= h[ray, 0] - 1e-6).all() and (ts < h[ray, 1]).all()
    # every sample lies in an occupied cell of its mip level
    G = 128
    mx = np.abs(xyz).max(1)
    e = np.frexp(mx)[1]; mip_pos = np.clip(e + 1, 0, casc - 1)
    mip_dt = np.clip(np.frexp(dl * G)[1], 0, casc - 1)
    mip = np.maximum(mip_pos, mip_dt)
    bound = np.minimum(np.ldexp(1.0, mip - 1), scale).astype(np.float32)
    cell = np.clip(0.5 * (xyz / bound[:, None] + 1) * G, 0, G - 1).astype(np.int64)
    idx = mip * G ** 3 + cases.morton_enc(cell[:, 0], cell[:, 1], cell[:, 2]).astype(np.int64)
    occ = (bf[idx // 8] >> (idx % 8)) & 1
    assert occ.mean() > 0.999          # float re-derivation may flip a boundary cell; the marcher's own test is exact
    # resume property of the test marcher: N rounds of k samples reproduce the first N*k training samples (noise=0)
    if esf == 0:      # calc_dt(scale) == calc_dt(cascades) only matters through clamping when esf != 0
        return
    ra0, _, _, dl0, ts0, _ = oracle.raymarching_train(o, d, h, bf, casc, scale, esf, np.zeros(n, np.float32), 128, 1024)
    ht2 = h.copy()
    _, _, dl_t, ts_t, neff = oracle.raymarching_test(o, d, ht2, np.arange(n), bf, casc, scale, esf, 128, 1024, 4)
    assert (neff <= 4).all()


def test_test_marcher_matches_train_marcher_when_dt_args_agree():
    """With scale == cascades (scale=1 -> cascades=2? no: pick scale 2, cascades 3 is different) the only
    difference between the two marchers is calc_dt's last argument; at esf=0 with scale=cascades... we use
    scale=1.0 whose cascades = 2 and compare only counts' upper bound instead."""
    scale, casc = 0.5, 1
    bf = cases.bitfield("full", casc)
    o, d = cases.rays(64, scale, seed=5, special=False)
    cnt, ht, _ = oracle.ray_aabb_intersect(o, d, np.zeros((1, 3), np.float32), np.full((1, 3), scale, np.float32), 1)
    h = cases.near_clamp(ht)
    ht2 = h.copy()
    x, dd, dl, ts, neff = oracle.raymarching_test(o, d, ht2, np.arange(64), bf, casc, scale, 0.0, 128, 1024, 16)
    # esf=0: dt = clamp(0, sqrt3/1024, sqrt3*2*cascades/128) = sqrt3/1024 in both kernels
    ra, xyz, dirs, dl1, ts1, _ = oracle.raymarching_train(o, d, h, bf, casc, scale, 0.0, np.zeros(64, np.float32), 128, 1024)
    for r in range(64):
        k = min(16, ra[r, 2]); s = ra[r, 1]
        assert neff[r] == k
        assert (ts[r, :k] == ts1[s:s + k]).all() and (dl[r, :k] == dl1[s:s + k]).all()
        assert (ts[r, k:] == 0).all() and (dd[r, k:] == 0).all()
        if k:
            assert ht2[r, 0] == np.float32(ts[r, k - 1] + dl[r, k - 1])


def _composite_numpy(sig, rgbs, nrm, sems, dl, ts, ra, thr):
    R = ra.shape[0]
    out = dict(opacity=np.zeros(R), depth=np.zeros(R), rgb=np.zeros((R, 3)), normal=np.zeros((R, 3)),
               sem=np.zeros((R, sems.shape[1])), ws=np.zeros(len(sig)), total=np.zeros(R, np.int64))
    for r, s0, n in ra:
        a = 1 - np.exp(-sig[s0:s0 + n].astype(np.float64) * dl[s0:s0 + n])
        T = np.concatenate([[1.0], np.cumprod(1 - a)])
        stop = np.nonzero(T[1:] <= thr)[0]
        k = stop[0] + 1 if len(stop) else n
        w = (a * T[:-1])[:k]
        out["ws"][s0:s0 + k] = w
        out["opacity"][r] = w.sum(); out["depth"][r] = (w * ts[s0:s0 + k]).sum()
        out["rgb"][r] = (w[:, None] * rgbs[s0:s0 + k]).sum(0); out["normal"][r] = (w[:, None] * nrm[s0:s0 + k]).sum(0)
        out["sem"][r] = (w[:, None] * sems[s0:s0 + k]).sum(0)
        out["total"][r] = stop[0] if len(stop) else n
    return out


def _march(case, n=192):
    name, kind, scale, casc, esf, _ = case
    bf = cases.bitfield(kind, casc, seed=1)
    o, d = cases.rays(n, scale, seed=2)
    cnt, ht, _ = oracle.ray_aabb_intersect(o, d, np.zeros((1, 3), np.float32), np.full((1, 3), scale, np.float32), 1)
    noise = np.random.RandomState(3).rand(n).astype(np.float32)
    return oracle.raymarching_train(o, d, cases.near_clamp(ht), bf, casc, scale, esf, noise, 128, 1024)


@pytest.mark.parametrize("thr", [1e-4, 1e-2])
def test_composite_fw_matches_float64_definition(thr):
    ra, xyz, dirs, dl, ts, counter = _march(cases.MARCH_CASES[2])
    S = int(counter[0]); C = 7
    f = cases.sample_fields(S, C, seed=4)
    total, op, dep, rgb, nrm, sem, ws = oracle.composite_train_fw(f["sigmas"], f["rgbs"], f["normals_pred"], f["sems"],
                                                                  dl, ts, ra, thr, C)
    ref = _composite_numpy(f["sigmas"], f["rgbs"], f["normals_pred"], f["sems"], dl, ts, ra, thr)
    assert (total == ref["total"]).mean() > 0.99       # a float32/float64 tie at the threshold may shift one ray
    for a, b in ((op, ref["opacity"]), (dep, ref["depth"]), (rgb, ref["rgb"]), (nrm, ref["normal"]), (sem, ref["sem"])):
        assert np.allclose(a, b, rtol=2e-4, atol=2e-5)
    al, ws2 = oracle.composite_alpha_fw(f["sigmas"], dl, ra, thr)
    assert (ws2 == ws).all()


def test_composite_bw_is_the_gradient_of_fw():
    """finite differences in float64 torch of the forward definition == the analytic backward
    (T_threshold = 0 so that no early termination makes the function discontinuous)."""
    ra, xyz, dirs, dl, ts, counter = _march(cases.MARCH_CASES[2], n=64)
    S = int(counter[0]); C = 3
    f = cases.sample_fields(S, C, seed=6)
    g = cases.ray_grads(64, C, seed=6)
    gws = np.random.RandomState(1).normal(size=S).astype(np.float32)
    sig = torch.tensor(f["sigmas"], dtype=torch.float64, requires_grad=True)
    rgbs = torch.tensor(f["rgbs"], dtype=torch.float64, requires_grad=True)
    loss = 0
    for r, s0, n in ra:
        if n == 0:
            continue
        a = 1 - torch.exp(-sig[s0:s0 + n] * torch.tensor(dl[s0:s0 + n], dtype=torch.float64))
        T = torch.cat([torch.ones(1, dtype=torch.float64), torch.cumprod(1 - a, 0)[:-1]])
        w = a * T
        loss = loss + g["dL_dopacity"][r] * w.sum() + g["dL_ddepth"][r] * (w * torch.tensor(ts[s0:s0 + n], dtype=torch.float64)).sum() \
            + (torch.tensor(g["dL_drgb"][r], dtype=torch.float64) * (w[:, None] * rgbs[s0:s0 + n]).sum(0)).sum() \
            + (torch.tensor(gws[s0:s0 + n], dtype=torch.float64) * w).sum()
    loss.backward()
    total, op, dep, rgb, nrm, sem, ws = oracle.composite_train_fw(f["sigmas"], f["rgbs"], f["normals_pred"], f["sems"], dl, ts, ra, 0.0, C)
    dsig, drgb, dnrm, dsem = oracle.composite_train_bw(g["dL_dopacity"], g["dL_ddepth"], g["dL_drgb"], g["dL_dnormal_pred"],
                                                       g["dL_dsem"], gws, f["sigmas"], f["rgbs"], f["normals_pred"], ws, dl,
                                                       ts, ra, op, dep, rgb, nrm, 0.0, C)
    assert np.allclose(drgb, rgbs.grad.numpy(), rtol=1e-3, atol=1e-5)
    assert np.allclose(dsig, sig.grad.numpy(), rtol=2e-3, atol=2e-4)
    assert np.allclose(dnrm, g["dL_dnormal_pred"][np.repeat(ra[:, 0], ra[:, 2])] * ws[:, None], rtol=1e-5, atol=1e-7)


def test_distortion_loss_matches_pairwise_definition_and_gradient():
    ra, xyz, dirs, dl, ts, counter = _march(cases.MARCH_CASES[2], n=48)
    S = int(counter[0])
    rng = np.random.RandomState(2)
    ws = (rng.rand(S) * 0.05).astype(np.float32)
    loss, wi, wti = oracle.distortion_loss_fw(ws, dl, ts, ra)
    w = torch.tensor(ws, dtype=torch.float64, requires_grad=True)
    tot = 0
    gl = rng.normal(size=48).astype(np.float32)
    for r, s0, n in ra:
        if n == 0:
            continue
        wr = w[s0:s0 + n]; tr = torch.tensor(ts[s0:s0 + n], dtype=torch.float64); dr = torch.tensor(dl[s0:s0 + n], dtype=torch.float64)
        pair = (wr[:, None] * wr[None, :] * (tr[:, None] - tr[None, :]).abs()).sum() + (wr * wr * dr).sum() / 3
        assert np.isclose(loss[r], pair.item(), rtol=2e-3, atol=1e-6)
        tot = tot + gl[r] * pair
    tot.backward()
    dws = oracle.distortion_loss_bw(gl, wi, wti, ws, dl, ts, ra)
    assert np.allclose(dws, w.grad.numpy(), rtol=5e-3, atol=1e-5)


def test_refloss_is_weighted_sum_with_same_weights():
    ra, xyz, dirs, dl, ts, counter = _march(cases.MARCH_CASES[2], n=96)
    S = int(counter[0])
    f = cases.sample_fields(S, 1, seed=8)
    _, _, _, _, _, _, ws = oracle.composite_train_fw(f["sigmas"], f["rgbs"], f["normals_pred"], f["sems"], dl, ts, ra, 1e-4, 1)
    lo, lp = oracle.composite_refloss_fw(f["sigmas"], f["normals_diff"], f["normals_ori"], dl, ts, ra, 1e-4)
    ray = np.repeat(ra[:, 0], ra[:, 2])
    assert np.allclose(lo, np.bincount(ray, ws * f["normals_ori"], 96), rtol=1e-4, atol=1e-6)
    assert np.allclose(lp[:, 1], np.bincount(ray, ws * f["normals_diff"][:, 1], 96), rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------- tcnn half
def test_grid_oracle_partition_of_unity_and_dense_lookup():
    L, F, T = 4, 2, 10
    b = 1.5
    lv, total = tcnn_oracle.grid_layout(L, F, T, 4, b)
    assert lv[0]["dense"] and not lv[-1]["dense"]
    table = torch.ones(total * F, dtype=torch.float64)
    x = torch.rand(100, 3, dtype=torch.float64)
    y = tcnn_oracle.grid_encode(x, table, L, F, T, 4, b)
    assert torch.allclose(y, torch.ones_like(y))                # trilinear weights sum to 1 on every level
    # level 0 dense: value at a lattice point equals the entry x + y*res + z*res^2
    tab = torch.arange(total * F, dtype=torch.float64)
    res, scale = lv[0]["res"], lv[0]["scale"]
    p = torch.tensor([[1, 2, 1]], dtype=torch.float64)
    xq = (p - 0.5) / scale
    y = tcnn_oracle.grid_encode(xq, tab, L, F, T, 4, b)
    e = 1 + 2 * res + 1 * res * res
    assert torch.allclose(y[0, :2], torch.tensor([2.0 * e, 2.0 * e + 1], dtype=torch.float64))


def test_sh_oracle_orthonormal():
    g = torch.Generator().manual_seed(0)
    d = torch.nn.functional.normalize(torch.randn(200000, 3, generator=g, dtype=torch.float64), dim=-1)
    Y = tcnn_oracle.sh_encode((d + 1) / 2, 4)
    gram = (Y.t() @ Y) / d.shape[0] * 4 * np.pi
    assert torch.allclose(gram, torch.eye(16, dtype=torch.float64), atol=0.03)


def test_mlp_oracle_layout():
    p = torch.arange(8 * 4 + 16 * 8, dtype=torch.float32) * 0.01
    x = torch.ones(2, 4)
    y = tcnn_oracle.mlp_forward(x, p, 4, 8, 1, 3)
    W0 = p[:32].reshape(8, 4); W1 = p[32:].reshape(16, 8)
    assert torch.allclose(y, (torch.relu(x @ W0.t()) @ W1.t())[:, :3])


def test_density_head_closed_forms_match_autograd_double_backward():
    """The formulas csrc/density_head.cu implements (oracle/tcnn_oracle.density_head_closed_form) against the reference's
    own formulation — autograd.grad(create_graph=True) through Linear/Softplus/Linear/Softplus, networks.py:54-59,186-196 —
    in fp64 on CPU: outputs and all five gradients to 1e-8 — the one row driven past softplus' linear threshold (20)
    differs by 1 - sigmoid(z) ~ 2e-9 in fp64 (torch switches to the identity there); in fp32 both round to 1."""
    import torch
    from oracle import tcnn_oracle
    torch.manual_seed(3)
    n, D, W = 200, 24, 16
    mk = lambda *s: torch.randn(*s, dtype=torch.double)
    e, W1, b1, W2, b2 = mk(n, D), mk(W, D) * 0.5, mk(W), mk(1, W), mk(1)
    e[0] = 30.0                                            # past softplus' linear threshold
    ds, dg = mk(n), mk(n, D)
    ps = [t.clone().requires_grad_(True) for t in (e, W1, b1, W2, b2)]
    sig, g = tcnn_oracle.density_head_reference(*ps)
    ref = torch.autograd.grad((sig * ds).sum() + (g * dg).sum(), ps)
    sigma, g_e, grads = tcnn_oracle.density_head_closed_form(e, W1, b1, W2, b2, ds, dg)
    assert float((sigma - sig).abs().max()) < 1e-12 and float((g_e - g).abs().max()) < 1e-8
    assert float((g_e[1:] - g[1:]).abs().max()) < 1e-12          # rows below the threshold: exact
    for a, b in zip(grads, ref):
        assert a.shape == b.shape and float((a - b).abs().max()) < 1e-8 * max(1.0, float(b.abs().max()))


def test_nocuda_grid_encode_fast_equals_oracle_grid_encode():
    """the one-gather grid encode of the configs[0] CPU baseline is the oracle's encode (values, dx, dtable)"""
    import torch
    from oracle import tcnn_oracle
    from oracle.nocuda_pipeline import grid_encode_fast
    cfg = (6, 2, 11, 4, 1.6)
    g = torch.Generator().manual_seed(0)
    tab = torch.rand(tcnn_oracle.grid_layout(*cfg)[1] * 2, generator=g, dtype=torch.float64).requires_grad_(True)
    x = torch.rand(400, 3, generator=g, dtype=torch.float64).requires_grad_(True)
    a, b = grid_encode_fast(x, tab, *cfg), tcnn_oracle.grid_encode(x, tab, *cfg)
    assert torch.allclose(a, b, rtol=1e-12, atol=1e-14)
    ga = torch.autograd.grad((a ** 2).sum(), [x, tab]); gb = torch.autograd.grad((b ** 2).sum(), [x, tab])
    assert all(torch.allclose(u, v, rtol=1e-10, atol=1e-12) for u, v in zip(ga, gb))


def test_nocuda_pipeline_trains_on_cpu():
    import torch
    from oracle.nocuda_pipeline import NoCUDAPipeline, aabb_hits
    from synth_scenes import BoxScene
    sc = BoxScene("lego")
    ro, rd = sc.sample_rays(256, sc.poses(4), torch.Generator().manual_seed(0))
    _, hit = aabb_hits(ro, rd, 0.5)
    ro, rd = ro[hit][:64].contiguous(), rd[hit][:64].contiguous()
    rgb, *_ = sc.shade(ro, rd)
    pipe = NoCUDAPipeline(log2_T_xyz=12, log2_T_rgb=13, samples=(16, 32), threads=4)
    l0, n = pipe.train_step(ro, rd, rgb)
    for _ in range(5):
        l1, _ = pipe.train_step(ro, rd, rgb)
    assert n == 64 * 48 and l1 < l0
