"""Pins oracle/nocuda_pipeline.py (the CPU baseline of BASELINE.json configs[0]) to the reference's own code: its
`sample_pdf`, `raw2outputs` (models/custom_functions.py:248-321) and `rendering_noCUDA.render` (models/rendering_noCUDA.py)
only run on a GPU (.cuda() on every temporary), so the restatement — device-agnostic torch ops — is compared with them
here, on the same tensors and the same field.  Tolerance: fp32 re-association only (rtol 1e-5 / atol 1e-6)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def glue():
    from baseline import ref_harness
    return ref_harness.load(vren="ours", tcnn="standin")


def test_sample_pdf_and_raw2outputs_match_reference(glue):
    from oracle import nocuda_pipeline as nc
    g = torch.Generator(device="cuda").manual_seed(0)
    bins = torch.sort(torch.rand(257, 63, device="cuda", generator=g) * 4 + 0.1, -1).values
    w = torch.rand(257, 62, device="cuda", generator=g) ** 4
    assert torch.equal(nc.sample_pdf(bins, w, 128), glue.custom_functions.sample_pdf(bins, w, 128, det=True))
    raw = torch.rand(257, 64, 17, device="cuda", generator=g)
    z = torch.sort(torch.rand(257, 64, device="cuda", generator=g) * 3, -1).values
    d = torch.randn(257, 3, device="cuda", generator=g)
    for a, b in zip(nc.raw2outputs(raw, z, d), glue.custom_functions.raw2outputs(raw, z, d, classes=7)):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)


def test_renderer_restatement_matches_reference_rendering_noCUDA(glue):
    from oracle import nocuda_pipeline as nc
    from synth_scenes import BoxScene
    scene = BoxScene("lego", device="cuda")
    poses = scene.poses(8)
    gen = torch.Generator(device="cuda").manual_seed(5)
    ro, rd = scene.sample_rays(3000, poses, gen)
    fields = [nc.NGPNoCUDA(0.5, seed=1337, log2_T_xyz=14, log2_T_rgb=15).cuda(), nc.NGPNoCUDA(0.5, seed=1338, log2_T_xyz=14, log2_T_rgb=15).cuda()]
    with torch.no_grad():
        for f in fields:
            f.xyz_table.mul_(3000.0); f.rgb_table.mul_(3000.0)

    class Coarse:                       # the reference's coarse call site unpacks three outputs (rendering_noCUDA.py:177)
        center, half_size = torch.zeros(1, 3, device="cuda"), torch.full((1, 3), 0.5, device="cuda")
        def __call__(self, x, d, emb, **kw):
            return fields[0].forward_coarse(x, d, None)

    class Fine:                         # ... and its fine call site six (:170-175)
        def __call__(self, x, d, emb, **kw):
            return (*fields[1](x, d, None), 0)
        def forward_skybox(self, d):
            return None

    # rays that hit the box (the exp-warp is 0/0 otherwise, see oracle/nocuda_pipeline.py)
    hits_t, hit = nc.aabb_hits(ro, rd, 0.5)
    ro, rd, hits_t = ro[hit.cuda()].contiguous(), rd[hit.cuda()].contiguous(), hits_t[hit].cuda()
    R = ro.shape[0]
    emb = torch.zeros(R, 4, device="cuda")
    torch.manual_seed(3)
    ref = glue.rendering_noCUDA.render([Coarse(), Fine()], ro, rd, samples=[64, 128], num_classes=7, embedding_a0=emb, embedding_a1=emb)
    torch.manual_seed(3)
    t_rand = torch.rand(R).cuda()       # the reference draws it on the host generator (rendering_noCUDA.py:139)
    got = nc.render_rays_train(fields, ro, rd, hits_t, (64, 128), 7, t_rand=t_rand)
    assert ref["total_samples"] == got["total_samples"] == 192
    for k in ("z_vals0", "z_vals1"):
        assert torch.allclose(got[k], ref[k], rtol=1e-5, atol=1e-6), k
    for k in ("opacity0", "rgb0", "ws0", "opacity1", "rgb1", "depth1", "ws1", "normal_pred1", "normal_raw1", "semantic1", "Rp1"):
        assert torch.allclose(got[k], ref[k], rtol=2e-4, atol=2e-5), (k, float((got[k] - ref[k]).abs().max()))
