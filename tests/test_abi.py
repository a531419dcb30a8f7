"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/ngp_b200.h declares, and the Python surface mirrors the reference's names."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ngp_b200 import _lib
    assert len(_lib.PROTOTYPES) >= 29
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.PROTOTYPES:
        assert hasattr(raw, name), name
    assert raw.ngp_abi_version() == 1


def test_header_has_no_torch_types_and_cites_reference():
    src = open(os.path.join(ROOT, "include", "ngp_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)          # prototypes only, comments stripped
    assert "torch" not in code.lower() and "Tensor" not in code
    assert 'extern "C"' in src
    for cite in ("binding.cpp:4-16", "binding.cpp:60-81", "binding.cpp:121-145", "binding.cpp:287-298",
                 "networks.py:40-52", "networks.py:89-162", "ray_utils.py:49-72", "rendering.py:217-219", "networks.py:54-59",
                 "networks.py:186-196"):
        assert cite in src, cite


def test_vren_surface_matches_reference_registration():
    """names registered at the reference's binding.cpp:323-342"""
    import vren
    names = ["ray_aabb_intersect", "ray_sphere_intersect", "morton3D", "morton3D_invert", "packbits",
             "raymarching_train", "raymarching_test", "composite_alpha_fw", "composite_train_fw",
             "composite_train_bw", "composite_refloss_fw", "composite_refloss_bw", "composite_test_fw",
             "distortion_loss_fw", "distortion_loss_bw"]
    for n in names:
        assert callable(getattr(vren, n)), n


def test_product_path_never_touches_oracle():
    """oracle/ is test infrastructure: nothing under instant-ngp-pp_b200/ may import or load it."""
    pkg = os.path.join(ROOT, "instant-ngp-pp_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "libngp_oracle" not in txt and "vren_ref" not in txt, f


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import vren
    with pytest.raises(RuntimeError):
        vren.morton3D(torch.zeros(8, 3, dtype=torch.int32))


def test_tcnn_shaped_modules_construct():
    import tinycudann as tcnn
    import numpy as np
    b = float(np.exp(np.log(2048 * 0.5 / 16) / 15))
    enc = tcnn.Encoding(3, {"otype": "Grid", "type": "Hash", "n_levels": 16, "n_features_per_level": 8,
                            "log2_hashmap_size": 19, "base_resolution": 16, "per_level_scale": b,
                            "interpolation": "Linear"})
    assert enc.n_output_dims == 128 and enc.params.numel() == 45780160     # 174.6 MiB fp32
    assert float(enc.params.abs().max()) <= 1e-4
    sh = tcnn.Encoding(3, {"otype": "SphericalHarmonics", "degree": 4})
    assert sh.n_output_dims == 16 and sh.params.numel() == 0
    net = tcnn.Network(144, 3, {"otype": "CutlassMLP", "activation": "ReLU", "output_activation": "Sigmoid",
                                "n_neurons": 128, "n_hidden_layers": 1})
    assert net.params.numel() == 128 * 144 + 16 * 128


def test_grid_layout_matches_oracle_layout():
    from ngp_b200.tcnn import GridConfig
    from oracle import tcnn_oracle
    import numpy as np
    for scale, L, F, T in ((0.5, 16, 8, 19), (0.5, 16, 8, 21), (8.0, 16, 2, 22), (0.5, 16, 2, 19), (1.0, 8, 2, 16)):
        b = float(np.exp(np.log(2048 * scale / 16) / (L - 1)))
        g = GridConfig({"otype": "HashGrid", "n_levels": L, "n_features_per_level": F, "log2_hashmap_size": T,
                        "base_resolution": 16, "per_level_scale": b})
        lv, total = tcnn_oracle.grid_layout(L, F, T, 16, b)
        assert g.n_params == total * F
        assert g.offsets[:-1] == [l["offset"] for l in lv]
        assert g.resolutions == [l["res"] for l in lv]
        assert g.dense == [l["dense"] for l in lv]


def test_ray_directions_follow_the_reference_formula():
    """datasets/ray_utils.py:10-46: ((u - cx + .5)/fx, (v - cy + .5)/fy, 1), row-major over (v, u), un-normalised.
    (get_rays itself is a kernel: tests/test_vren_gpu.py; without a GPU it must refuse, not fall back.)"""
    import torch
    from ngp_b200 import ray_utils
    H, W = 3, 5
    K = [[10.0, 0, 2.5], [0, 20.0, 1.5], [0, 0, 1]]
    d = ray_utils.get_ray_directions(H, W, K)
    assert d.shape == (H * W, 3) and d.dtype == torch.float32
    for v in range(H):
        for u in range(W):
            row = d[v * W + u]
            assert abs(float(row[0]) - (u - 2.5 + 0.5) / 10.0) < 1e-7 and abs(float(row[1]) - (v - 1.5 + 0.5) / 20.0) < 1e-7
            assert float(row[2]) == 1.0
    assert ray_utils.get_ray_directions(H, W, K, flatten=False).shape == (H, W, 3)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            ray_utils.get_rays(d, torch.eye(4)[:3])
