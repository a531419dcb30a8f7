"""Drop-in alias: with instant-ngp-pp_b200/ on sys.path, `import vren` (reference
models/custom_functions.py:2, models/rendering.py:7, models/networks.py:6, losses.py:4) resolves
to the B200 implementation."""
from ngp_b200.vren import *  # noqa: F401,F403
from ngp_b200.vren import (ray_aabb_intersect, ray_sphere_intersect, morton3D, morton3D_invert, packbits,  # noqa: F401
                           raymarching_train, raymarching_test, composite_alpha_fw, composite_train_fw,
                           composite_train_bw, composite_refloss_fw, composite_refloss_bw, composite_test_fw,
                           distortion_loss_fw, distortion_loss_bw)
