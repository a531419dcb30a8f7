"""Seeded input generators shared by the golden-vector script (tests/golden/make_golden.py, run
against the reference's own CUDA kernels on a B200), the CPU oracle tests and the GPU parity tests.
Everything derives from numpy RandomState seeds so the three places see identical bytes."""
import numpy as np

G = 128


def morton_enc(x, y, z):
    def ex(v):
        v = (v * 0x00010001) & 0xFF0000FF
        v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3
        v = (v * 0x00000005) & 0x49249249
        return v
    return ex(x.astype(np.uint64)) | (ex(y.astype(np.uint64)) << 1) | (ex(z.astype(np.uint64)) << 2)


def bitfield(kind, cascades, seed=0):
    """Procedural occupancy bitfields, (cascades*G^3/8) uint8 in Morton order."""
    rng = np.random.RandomState(seed)
    ax = np.arange(G)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    idx = morton_enc(X.ravel(), Y.ravel(), Z.ravel()).astype(np.int64)
    c = (np.stack([X, Y, Z], -1).reshape(-1, 3) + 0.5) / G * 2 - 1        # cell centres in [-1,1]
    out = []
    for ci in range(cascades):
        if kind == "shell":
            r = np.linalg.norm(c, axis=1)
            occ = (r > 0.45) & (r < 0.6)
        elif kind == "boxes":
            occ = np.zeros(G ** 3, bool)
            for _ in range(12):
                ctr = rng.uniform(-0.6, 0.6, 3); h = rng.uniform(0.05, 0.25, 3)
                occ |= (np.abs(c - ctr) < h).all(1)
        elif kind == "sparse":
            occ = rng.rand(G ** 3) < 0.04
        elif kind == "full":
            occ = np.ones(G ** 3, bool)
        else:
            raise ValueError(kind)
        g = np.zeros(G ** 3, bool)
        g[idx] = occ
        out.append(np.packbits(g.reshape(-1, 8)[:, ::-1], axis=1).ravel())
    return np.concatenate(out).astype(np.uint8)


def rays(n, scale, seed=0, special=True):
    """Rays aimed at the [-scale,scale]^3 box from outside and inside, un-normalised directions,
    plus axis-parallel / zero-component / grazing / missing rays when special=True."""
    rng = np.random.RandomState(seed)
    o = rng.normal(size=(n, 3)); o = o / np.linalg.norm(o, axis=1, keepdims=True) * scale * rng.uniform(0.2, 3.0, (n, 1))
    tgt = rng.uniform(-scale, scale, (n, 3)) * 0.8
    d = tgt - o
    d = d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(0.8, 1.3, (n, 1))
    if special and n >= 64:
        d[0:8, 0] = 0.0                      # zero x component (inv_d = inf)
        d[8:16, 1] = -0.0                    # negative zero
        d[16:24] = np.eye(3)[rng.randint(0, 3, 8)] * rng.choice([-1, 1], (8, 1))   # axis parallel
        o[24:32] = rng.uniform(-scale, scale, (8, 3)) * 0.5                        # origin inside
        d[32:40] = -d[32:40]                 # pointing away
        o[40:48] = np.array([scale, scale, 3 * scale])                             # grazing an edge
        d[40:48] = np.array([0, 0, -1.0]) + rng.normal(size=(8, 3)) * 1e-3
    return o.astype(np.float32), d.astype(np.float32)


def near_clamp(hits_t):
    """models/rendering.py:30"""
    h = hits_t[:, 0, :].copy()
    m = (h[:, 0] >= 0) & (h[:, 0] < 0.01)
    h[m, 0] = 0.01
    return h


MARCH_CASES = [
    # name, bitfield kind, scale, cascades, exp_step_factor, n_rays
    ("shell_s05", "shell", 0.5, 1, 0.0, 1024),
    ("boxes_s05", "boxes", 0.5, 1, 0.0, 1024),
    ("sparse_s05", "sparse", 0.5, 1, 0.0, 512),
    ("boxes_s8", "boxes", 8.0, 5, 1.0 / 256, 1024),
    ("sparse_s8", "sparse", 8.0, 5, 1.0 / 256, 512),
    ("shell_s2_esf0", "shell", 2.0, 3, 0.0, 512),
]


def sample_fields(n_samples, classes, seed=0):
    """Random per-sample field outputs for the compositors."""
    rng = np.random.RandomState(seed)
    f = lambda *s: rng.rand(*s).astype(np.float32)
    sig = (rng.gamma(0.6, 30.0, n_samples) * (rng.rand(n_samples) > 0.3)).astype(np.float32)
    nrm = rng.normal(size=(n_samples, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    sems = rng.dirichlet(np.ones(max(classes, 1)), n_samples).astype(np.float32)[:, :classes]
    return dict(sigmas=sig, rgbs=f(n_samples, 3), normals_pred=nrm, sems=np.ascontiguousarray(sems),
                normals_raw=np.roll(nrm, 1, 0).copy(), normals_diff=f(n_samples, 3), normals_ori=f(n_samples))


def ray_grads(n_rays, classes, seed=0):
    rng = np.random.RandomState(seed + 77)
    g = lambda *s: rng.normal(size=s).astype(np.float32)
    return dict(dL_dopacity=g(n_rays), dL_ddepth=g(n_rays), dL_drgb=g(n_rays, 3), dL_dnormal_pred=g(n_rays, 3),
                dL_dsem=g(n_rays, classes), dL_dloss=g(n_rays), dL_dloss_o=g(n_rays), dL_dloss_p=g(n_rays, 3))


def canonical_order(rays_a, *per_sample):
    """Reorder a (possibly atomically ordered) rays_a + packed per-sample arrays into ray-index
    order with start = exclusive scan (SURVEY.md §7 'canonical ordering')."""
    order = np.argsort(rays_a[:, 0], kind="stable")
    ra = rays_a[order]
    n = ra[:, 2]
    new_start = np.concatenate([[0], np.cumsum(n)[:-1]]).astype(np.int64)
    gather = np.concatenate([np.arange(s, s + k) for s, k in zip(ra[:, 1], n)]) if n.sum() else np.zeros(0, np.int64)
    out_ra = np.stack([ra[:, 0], new_start, n], 1).astype(np.int64)
    return (out_ra,) + tuple(a[gather] for a in per_sample)
