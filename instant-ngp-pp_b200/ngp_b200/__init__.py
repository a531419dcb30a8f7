"""ngp_b200 — B200-native (sm_100a) hot path of instant-ngp-pp behind the reference's own
PyTorch surface.  See DESIGN.md / INTEGRATION.md at the repository root.

Importing this package loads instant-ngp-pp_b200/libngp_b200.so through ctypes and fails loudly
if it has not been built; nothing here falls back to PyTorch or CPU code.
"""
from . import _lib  # noqa: F401  (loads the C ABI or raises)
from . import vren, tcnn, custom_functions, losses, networks, rendering  # noqa: F401

__all__ = ["vren", "tcnn", "custom_functions", "losses", "networks", "rendering"]
