"""Drop-in alias: `import tinycudann as tcnn` (reference models/networks.py:5) resolves to the
B200 implementation of the Encoding / Network modules the hot path uses."""
from ngp_b200.tcnn import Encoding, Network, NetworkWithInputEncoding  # noqa: F401
