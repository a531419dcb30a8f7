"""Pins the CPU oracle (oracle/ngp_oracle.c) to outputs of the REFERENCE'S OWN CUDA kernels:
tests/golden/*.npz were minted on a B200 by tests/golden/make_golden.py from oracle/_ref/vren_ref.so
(the reference's models/csrc compiled in place).  Bit-exact for integer / geometry / marching
outputs; rtol 2e-4 + atol 2e-5 for the __expf compositors (expf vs ex2.approx)."""
import hashlib
import os

import numpy as np
import pytest

import cases
from oracle import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
have = os.path.exists(os.path.join(GOLD, "occupancy.npz"))
pytestmark = pytest.mark.skipif(not have, reason="golden vectors not minted yet: run tests/golden/make_golden.py on a GPU box")
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


def close(a, b, rtol=2e-4, atol=2e-5, frac=1.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    ok = np.abs(a - b) <= atol + rtol * np.abs(b)
    assert ok.mean() >= frac, f"mismatch frac {1 - ok.mean():.2e}, max abs {np.abs(a - b).max():.3e}"


def test_occupancy_golden():
    g = np.load(os.path.join(GOLD, "occupancy.npz"))
    ax = np.arange(128, dtype=np.int32)
    lattice = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    m = oracle.morton3D(lattice)
    assert sha(m) == str(g["morton_full_sha"]) and (m[:4096] == g["morton_full_head"]).all()
    assert (oracle.morton3D(g["coords"]) == g["morton"]).all()
    assert (oracle.morton3D_invert(g["inv_in"]) == g["inv_out"]).all()
    assert (oracle.packbits(g["grid"], float(g["thr"])) == g["bits_f32"]).all()
    assert (oracle.packbits(g["grid"].astype(np.float16).astype(np.float32), float(g["thr"])) == g["bits_f16"]).all()


@pytest.mark.parametrize("ci", range(len(cases.MARCH_CASES)), ids=[c[0] for c in cases.MARCH_CASES])
def test_march_golden(ci):
    name, kind, scale, casc, esf, n = cases.MARCH_CASES[ci]
    g = np.load(os.path.join(GOLD, f"march_{name}.npz"))
    bf = cases.bitfield(kind, casc, seed=1)
    o, d = cases.rays(n, scale, seed=2 + ci)
    cnt, ht, idx = oracle.ray_aabb_intersect(o, d, np.zeros((1, 3), np.float32), np.full((1, 3), scale, np.float32), 1)
    assert (cnt == g["hit_cnt"]).all() and (bits(ht) == bits(g["hits_t"])).all() and (idx == g["hits_idx"]).all()
    h = cases.near_clamp(ht)
    noise = np.random.RandomState(3 + ci).rand(n).astype(np.float32)
    ra, xyzs, dirs, deltas, ts, counter = oracle.raymarching_train(o, d, h, bf, casc, scale, esf, noise, 128, 1024)
    assert (counter == g["counter"]).all()
    assert (ra[:, 2] == g["n_samples"]).all()
    assert sha(ts) == str(g["ts_sha"]) and sha(deltas) == str(g["deltas_sha"])
    assert sha(xyzs) == str(g["xyzs_sha"]) and sha(dirs) == str(g["dirs_sha"])
    ht2 = h.copy()
    for rnd in range(3):
        x, dd, dl, tt, neff = oracle.raymarching_test(o, d, ht2, np.arange(n), bf, casc, scale, esf, 128, 1024, 8)
        assert (neff == g[f"test{rnd}_neff"]).all()
        assert (bits(tt) == bits(g[f"test{rnd}_ts"])).all() and (bits(dl) == bits(g[f"test{rnd}_deltas"])).all()
        assert sha(x) == str(g[f"test{rnd}_xyzs_sha"]) and (bits(ht2) == bits(g[f"test{rnd}_hits_t"])).all()
    if "fw_t4_opacity" not in g:
        return
    C, S = 7, int(counter[0])
    f = cases.sample_fields(S, C, seed=5 + ci)
    gr = cases.ray_grads(n, C, seed=5 + ci)
    ray = np.repeat(ra[:, 0], ra[:, 2])
    for tag, thr in (("t4", 1e-4), ("t2", 1e-2)):
        total, op, dep, rgb, nrm, sem, ws = oracle.composite_train_fw(f["sigmas"], f["rgbs"], f["normals_pred"], f["sems"],
                                                                      deltas, ts, ra, thr, C)
        same = total == g[f"fw_{tag}_total"]
        assert same.mean() > 0.995
        for a, k in ((op, "opacity"), (dep, "depth"), (rgb, "rgb"), (nrm, "normal"), (sem, "sem")):
            close(a[same], g[f"fw_{tag}_{k}"][same])
        close(ws[same[ray]], g[f"fw_{tag}_ws"][same[ray]])
        gws = np.random.RandomState(9).normal(size=S).astype(np.float32)
        dsig, drgb, dnrm, dsem = oracle.composite_train_bw(
            gr["dL_dopacity"], gr["dL_ddepth"], gr["dL_drgb"], gr["dL_dnormal_pred"], gr["dL_dsem"], gws, f["sigmas"],
            f["rgbs"], f["normals_pred"], g[f"fw_{tag}_ws"], deltas, ts, ra, g[f"fw_{tag}_opacity"], g[f"fw_{tag}_depth"],
            g[f"fw_{tag}_rgb"], g[f"fw_{tag}_normal"], thr, C)
        m = same[ray]
        close(drgb[m], g[f"bw_{tag}_drgbs"][m]); close(dsem[m], g[f"bw_{tag}_dsems"][m]); close(dnrm[m], g[f"bw_{tag}_dnormals"][m])
        sc = np.abs(g[f"bw_{tag}_dsigmas"]).max() + 1e-6
        close(dsig[m], g[f"bw_{tag}_dsigmas"][m], rtol=2e-3, atol=2e-5 * sc, frac=0.999)
        lo, lp = oracle.composite_refloss_fw(f["sigmas"], f["normals_diff"], f["normals_ori"], deltas, ts, ra, thr)
        close(lo[same], g[f"ref_{tag}_loss_o"][same]); close(lp[same], g[f"ref_{tag}_loss_p"][same])
        rs, rd, ro = oracle.composite_refloss_bw(gr["dL_dloss_o"], gr["dL_dloss_p"], f["sigmas"], f["normals_diff"],
                                                 f["normals_ori"], deltas, ts, ra, g[f"ref_{tag}_loss_o"], g[f"ref_{tag}_loss_p"], thr)
        close(rd[m], g[f"ref_{tag}_ddiff"][m]); close(ro[m], g[f"ref_{tag}_dori"][m])
        close(rs[m], g[f"ref_{tag}_dsigmas"][m], rtol=2e-3, atol=2e-5 * (np.abs(g[f"ref_{tag}_dsigmas"]).max() + 1e-6), frac=0.999)
    al, ws2 = oracle.composite_alpha_fw(f["sigmas"], deltas, ra, 1e-4)
    close(al, g["alpha_t4_alphas"], frac=0.999); close(ws2, g["alpha_t4_ws"], frac=0.999)
    loss, wi, wti = oracle.distortion_loss_fw(g["fw_t4_ws"], deltas, ts, ra)
    close(loss, g["dist_loss"], rtol=1e-3, atol=1e-6); close(wi, g["dist_ws_incl"], rtol=1e-5, atol=1e-7)
    close(wti, g["dist_wts_incl"], rtol=1e-5, atol=1e-7)
    dws = oracle.distortion_loss_bw(gr["dL_dloss"], g["dist_ws_incl"], g["dist_wts_incl"], g["fw_t4_ws"], deltas, ts, ra)
    close(dws, g["dist_dws"], rtol=1e-5, atol=1e-6)
    # test-time compositor
    ht3 = h.copy(); alive = np.arange(n, dtype=np.int64)
    x, dd, dl, tt, neff = oracle.raymarching_test(o, d, ht3, alive, bf, casc, scale, esf, 128, 1024, 8)
    ft = cases.sample_fields(n * 8, C, seed=50 + ci)
    rngs = np.random.RandomState(60 + ci)
    st = dict(opacity=(rngs.rand(n) * 0.5).astype(np.float32), depth=rngs.rand(n).astype(np.float32),
              rgb=rngs.rand(n, 3).astype(np.float32), normal=rngs.rand(n, 3).astype(np.float32),
              normal_raw=rngs.rand(n, 3).astype(np.float32), sem=rngs.rand(n, C).astype(np.float32))
    oracle.composite_test_fw(ft["sigmas"].reshape(n, 8), ft["rgbs"].reshape(n, 8, 3), ft["normals_pred"].reshape(n, 8, 3),
                             ft["normals_raw"].reshape(n, 8, 3), ft["sems"].reshape(n, 8, C), dl, tt, ht3, alive, 1e-2, C,
                             neff, st["opacity"], st["depth"], st["rgb"], st["normal"], st["normal_raw"], st["sem"])
    same = alive == g["ctest_alive"]
    assert same.mean() > 0.995
    for k in st:
        close(st[k][same], g[f"ctest_{k}"][same])
