"""REFERENCE-ARM INFRASTRUCTURE — a plain-PyTorch-op `tinycudann` so that the reference's unmodified
models/networks.py (`import tinycudann as tcnn`, networks.py:5) runs end to end on this box.

tiny-cuda-nn (NVlabs, master, TCNN_HALF_PRECISION=0 — README.md:14-26) is an un-vendored, un-pinned dependency that is
absent from /root/reference and from this image, so the reference's tcnn kernels cannot be timed or compared here.
Everything below is torch ops (autograd supplies backward, input backward and the double backward the normals need),
following SURVEY.md Appendix B through oracle/tcnn_oracle.py's layout and SH definitions.  Numbers measured with it
are labelled "vren_ref + torch stand-in", never "tcnn".

Parameter layout and initialisation are identical to instant-ngp-pp_b200/ngp_b200/tcnn.py (flat fp32 `.params`; grid:
level-major, entry, F features, U(-1e-4,1e-4) from torch.Generator(seed); MLP: per layer row-major (out,in), output rows
padded to 16, Xavier-uniform from torch.Generator(seed)), so that a model built on either back end starts from the same
weights and state dicts are interchangeable (tests/test_reference_glue_cpu.py).
"""
import math

import torch
from torch import nn

from oracle import tcnn_oracle


class _Grid:
    def __init__(self, cfg):
        self.L = int(cfg.get("n_levels", 16)); self.F = int(cfg.get("n_features_per_level", 2))
        self.log2_T = int(cfg.get("log2_hashmap_size", 19)); self.base = int(cfg.get("base_resolution", 16))
        self.pls = float(cfg.get("per_level_scale", 2.0))
        self.levels, self.total = tcnn_oracle.grid_layout(self.L, self.F, self.log2_T, self.base, self.pls)
        self._offs = {}

    def corner_offsets(self, dev):
        if dev not in self._offs:
            self._offs[dev] = torch.tensor([[(k >> d) & 1 for d in range(3)] for k in range(8)], dtype=torch.int64, device=dev)
        return self._offs[dev]

    def encode(self, x, table):
        """x (N,3) in [0,1] -> (N, L*F); one gather of 8 corners per level (index_select: its backward is an
        atomic index_add, not the sort-based index_put)."""
        tab = table.view(self.total, self.F)
        offs = self.corner_offsets(x.device)                  # (8,3) in {0,1}
        offs_b = offs.bool()
        outs = []
        for lv in self.levels:
            pos = x * lv["scale"] + 0.5
            cell = torch.floor(pos)
            w = pos - cell                                    # (N,3)
            c = cell.detach().to(torch.int64)[:, None, :] + offs[None]          # (N,8,3)
            if lv["dense"]:
                idx = c[..., 0] + c[..., 1] * lv["res"] + c[..., 2] * (lv["res"] ** 2)
            else:
                m = 0xFFFFFFFF
                idx = ((c[..., 0] * tcnn_oracle.PRIMES[0]) & m) ^ ((c[..., 1] * tcnn_oracle.PRIMES[1]) & m) ^ ((c[..., 2] * tcnn_oracle.PRIMES[2]) & m)
            idx = idx % lv["size"] + lv["offset"]
            wk = torch.where(offs_b[None], w[:, None, :], 1 - w[:, None, :]).prod(-1)           # (N,8)
            vals = tab.index_select(0, idx.reshape(-1)).view(x.shape[0], 8, self.F)
            outs.append((wk[..., None] * vals).sum(1))
        return torch.cat(outs, 1)


class Encoding(nn.Module):
    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__()
        self.n_input_dims = int(n_input_dims)
        self.encoding_config = dict(encoding_config)
        otype = encoding_config["otype"]
        if otype in ("Grid", "HashGrid"):
            self.kind = "grid"
            self.grid = _Grid(encoding_config)
            self.n_output_dims = self.grid.L * self.grid.F
            gen = torch.Generator().manual_seed(seed)
            self.params = nn.Parameter((torch.rand(self.grid.total * self.grid.F, generator=gen) * 2 - 1) * 1e-4)
        elif otype == "SphericalHarmonics":
            self.kind = "sh"
            self.degree = int(encoding_config.get("degree", 4))
            self.n_output_dims = self.degree ** 2
            self.params = nn.Parameter(torch.zeros(0))
        elif otype == "Frequency":
            self.kind = "freq"
            self.n_freq = int(encoding_config.get("n_frequencies", 6))
            self.n_output_dims = self.n_input_dims * 2 * self.n_freq
            self.params = nn.Parameter(torch.zeros(0))
        else:
            raise NotImplementedError(otype)

    def forward(self, x):
        x = x.float()
        if self.kind == "grid":
            return self.grid.encode(x, self.params)
        if self.kind == "sh":
            return tcnn_oracle.sh_encode(x, self.degree)
        outs = []
        for k in range(self.n_freq):
            a = x * (2.0 ** k * math.pi)
            outs += [torch.sin(a), torch.cos(a)]
        return torch.stack(outs, -1).reshape(x.shape[0], -1)


class Network(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        self.n_input_dims, self.n_output_dims = int(n_input_dims), int(n_output_dims)
        self.network_config = dict(network_config)
        self.width = int(network_config.get("n_neurons", 128))
        self.n_hidden = int(network_config.get("n_hidden_layers", 1))
        self.act = network_config.get("activation", "ReLU")
        self.out_act = network_config.get("output_activation", "None")
        self.shapes = tcnn_oracle.mlp_layer_shapes(self.n_input_dims, self.width, self.n_hidden, self.n_output_dims)
        gen = torch.Generator().manual_seed(seed)
        parts = [((torch.rand(o, i, generator=gen) * 2 - 1) * math.sqrt(6.0 / (i + o))).reshape(-1) for o, i in self.shapes]
        self.params = nn.Parameter(torch.cat(parts))

    def forward(self, x):
        # fp32 GEMMs; whether they run as TF32 on the tensor cores is the process-wide torch switch, which the harness
        # sets for the duration of a reference run (baseline/ref_train.py: tf32=True is the class of tcnn's fp32 CutlassMLP)
        return tcnn_oracle.mlp_forward(x.float(), self.params, self.n_input_dims, self.width, self.n_hidden,
                                       self.n_output_dims, self.act, self.out_act)


class NetworkWithInputEncoding(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=1337):
        super().__init__()
        self.encoding = Encoding(n_input_dims, encoding_config, seed=seed)
        self.network = Network(self.encoding.n_output_dims, n_output_dims, network_config, seed=seed)
        self.n_input_dims, self.n_output_dims = int(n_input_dims), int(n_output_dims)

    def forward(self, x):
        return self.network(self.encoding(x))
