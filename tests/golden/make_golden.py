"""Mints tests/golden/*.npz from the REFERENCE'S OWN CUDA kernels (oracle/_ref/vren_ref.so, built in
place from /root/reference/models/csrc by oracle/build_ref.py) on seeded inputs (tests/cases.py).

Run on a B200 box (no /root/reference needed there, only the prebuilt .so):

    python tests/golden/make_golden.py [out_dir=gpurun_out/golden]

then copy the .npz files into tests/golden/ and commit them.  The reference's sample order is
nondeterministic (two independent atomics, raymarching.cu:237-241); every packed array is
canonicalised to ray-index order (tests/cases.py:canonical_order) before it is stored.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle import build_ref  # noqa: E402

dev = "cuda"
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
N = lambda t: t.detach().cpu().numpy()
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ref = build_ref.load()
    assert ref is not None, "oracle/_ref/vren_ref.so missing: run python oracle/build_ref.py where /root/reference exists"

    # ---------------------------------------------------------------- morton / packbits
    ax = np.arange(128, dtype=np.int32)
    lattice = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    m_full = N(ref.morton3D(T(lattice)))
    rng = np.random.RandomState(11)
    coords = rng.randint(0, 1024, (4096, 3)).astype(np.int32)       # beyond 128 too: 10-bit inputs
    m_rand = N(ref.morton3D(T(coords)))
    inv_in = rng.randint(0, 2 ** 30, 4096).astype(np.int32)
    inv_in[:4] = [0, -1, 2 ** 30 - 1, -(2 ** 31)]                    # sign-bit cases (arithmetic >>)
    m_inv = N(ref.morton3D_invert(T(inv_in)))
    grid = rng.normal(1.0, 2.0, 128 * 128 * 8).astype(np.float32)
    grid[::7] = -1.0; grid[3::11] = 1.25                             # invisible cells, exact-threshold cells
    bits32 = torch.zeros(grid.size // 8, dtype=torch.uint8, device=dev)
    ref.packbits(T(grid), 1.25, bits32)
    bits16 = torch.zeros(grid.size // 8, dtype=torch.uint8, device=dev)
    ref.packbits(T(grid).half(), 1.25, bits16)
    np.savez_compressed(os.path.join(out_dir, "occupancy.npz"), morton_full_sha=sha(m_full), morton_full_head=m_full[:4096],
                        coords=coords, morton=m_rand, inv_in=inv_in, inv_out=m_inv, grid=grid, thr=np.float32(1.25),
                        bits_f32=N(bits32), bits_f16=N(bits16))

    # ---------------------------------------------------------------- AABB + marching + compositing
    for ci, (name, kind, scale, casc, esf, n) in enumerate(cases.MARCH_CASES):
        bf = cases.bitfield(kind, casc, seed=1)
        o, d = cases.rays(n, scale, seed=2 + ci)
        center = np.zeros((1, 3), np.float32); half = np.full((1, 3), scale, np.float32)
        hit_cnt, hits_t, hits_idx = ref.ray_aabb_intersect(T(o), T(d), T(center), T(half), 1)
        out = dict(hit_cnt=N(hit_cnt), hits_t=N(hits_t), hits_idx=N(hits_idx))
        h = cases.near_clamp(N(hits_t))
        noise = np.random.RandomState(3 + ci).rand(n).astype(np.float32)
        rays_a, xyzs, dirs, deltas, ts, counter = ref.raymarching_train(T(o), T(d), T(h), T(bf), casc, scale, esf, T(noise), 128, 1024)
        tot = int(N(counter)[0])
        ra, xyzs, dirs, deltas, ts = cases.canonical_order(N(rays_a), N(xyzs[:tot]), N(dirs[:tot]), N(deltas[:tot]), N(ts[:tot]))
        out.update(counter=N(counter), n_samples=ra[:, 2].astype(np.int32), ts_sha=sha(ts), deltas_sha=sha(deltas),
                   xyzs_sha=sha(xyzs), dirs_sha=sha(dirs), xyzs_head=xyzs[:2048])
        if ci in (0, 2, 3, 4):
            out.update(ts=ts, deltas=deltas)
        # test-time marcher: three rounds of 8 samples from the same start
        ht = T(h.copy())
        alive = torch.arange(n, device=dev)
        for rnd in range(3):
            x_t, d_t, dl_t, ts_t, neff = ref.raymarching_test(T(o), T(d), ht, alive, T(bf), casc, scale, esf, 128, 1024, 8)
            out[f"test{rnd}_neff"] = N(neff); out[f"test{rnd}_ts"] = N(ts_t); out[f"test{rnd}_deltas"] = N(dl_t)
            out[f"test{rnd}_xyzs_sha"] = sha(N(x_t)); out[f"test{rnd}_hits_t"] = N(ht)
        if name in ("sparse_s05", "sparse_s8"):
            C = 7
            f = cases.sample_fields(tot, C, seed=5 + ci)
            g = cases.ray_grads(n, C, seed=5 + ci)
            for tag, thr in (("t4", 1e-4), ("t2", 1e-2)):
                total, opacity, depth, rgb, normal, sem, ws = ref.composite_train_fw(
                    T(f["sigmas"]), T(f["rgbs"]), T(f["normals_pred"]), T(f["sems"]), T(deltas), T(ts), T(ra), thr, C)
                out.update({f"fw_{tag}_total": N(total), f"fw_{tag}_opacity": N(opacity), f"fw_{tag}_depth": N(depth),
                            f"fw_{tag}_rgb": N(rgb), f"fw_{tag}_normal": N(normal), f"fw_{tag}_sem": N(sem), f"fw_{tag}_ws": N(ws)})
                gws = np.random.RandomState(9).normal(size=tot).astype(np.float32)
                dsig, drgb, dnrm, dsem = ref.composite_train_bw(
                    T(g["dL_dopacity"]), T(g["dL_ddepth"]), T(g["dL_drgb"]), T(g["dL_dnormal_pred"]), T(g["dL_dsem"]), T(gws),
                    T(f["sigmas"]), T(f["rgbs"]), T(f["normals_pred"]), ws, T(deltas), T(ts), T(ra), opacity, depth, rgb, normal, thr, C)
                out.update({f"bw_{tag}_dsigmas": N(dsig), f"bw_{tag}_drgbs": N(drgb), f"bw_{tag}_dnormals": N(dnrm), f"bw_{tag}_dsems": N(dsem)})
                lo, lp = ref.composite_refloss_fw(T(f["sigmas"]), T(f["normals_diff"]), T(f["normals_ori"]), T(deltas), T(ts), T(ra), thr)
                rs, rd, ro = ref.composite_refloss_bw(T(g["dL_dloss_o"]), T(g["dL_dloss_p"]), T(f["sigmas"]), T(f["normals_diff"]),
                                                      T(f["normals_ori"]), T(deltas), T(ts), T(ra), lo, lp, thr)
                out.update({f"ref_{tag}_loss_o": N(lo), f"ref_{tag}_loss_p": N(lp), f"ref_{tag}_dsigmas": N(rs),
                            f"ref_{tag}_ddiff": N(rd), f"ref_{tag}_dori": N(ro)})
                if tag == "t4":
                    al, ws2 = ref.composite_alpha_fw(T(f["sigmas"]), T(deltas), T(ra), thr)
                    out.update(alpha_t4_alphas=N(al), alpha_t4_ws=N(ws2))
                    loss, wi, wti = ref.distortion_loss_fw(ws, T(deltas), T(ts), T(ra))
                    dws = ref.distortion_loss_bw(T(g["dL_dloss"]), wi, wti, ws, T(deltas), T(ts), T(ra))
                    out.update(dist_loss=N(loss), dist_ws_incl=N(wi), dist_wts_incl=N(wti), dist_dws=N(dws))
            # test-time compositor: one round on the dense (n,8) slots of round 0, resumed from a partial state
            ht = T(h.copy()); alive = torch.arange(n, device=dev)
            x_t, d_t, dl_t, ts_t, neff = ref.raymarching_test(T(o), T(d), ht, alive, T(bf), casc, scale, esf, 128, 1024, 8)
            ft = cases.sample_fields(n * 8, C, seed=50 + ci)
            rngs = np.random.RandomState(60 + ci)
            st = dict(opacity=(rngs.rand(n) * 0.5).astype(np.float32), depth=rngs.rand(n).astype(np.float32),
                      rgb=rngs.rand(n, 3).astype(np.float32), normal=rngs.rand(n, 3).astype(np.float32),
                      normal_raw=rngs.rand(n, 3).astype(np.float32), sem=rngs.rand(n, C).astype(np.float32))
            tt = {k: T(v) for k, v in st.items()}
            ref.composite_test_fw(T(ft["sigmas"]).view(n, 8), T(ft["rgbs"]).view(n, 8, 3), T(ft["normals_pred"]).view(n, 8, 3),
                                  T(ft["normals_raw"]).view(n, 8, 3), T(ft["sems"]).view(n, 8, C), dl_t, ts_t, ht, alive, 1e-2, C,
                                  neff, tt["opacity"], tt["depth"], tt["rgb"], tt["normal"], tt["normal_raw"], tt["sem"])
            out.update({f"ctest_{k}": N(v) for k, v in tt.items()})
            out["ctest_alive"] = N(alive)
        np.savez_compressed(os.path.join(out_dir, f"march_{name}.npz"), **out)
        print(name, "samples", tot, flush=True)
    print("golden vectors written to", out_dir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
