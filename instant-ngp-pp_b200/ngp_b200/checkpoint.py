"""Checkpoint plumbing of the hot path's models (reference utils.py:7-42).

Same three entry points, same semantics: a checkpoint is either a bare state dict or a PyTorch-Lightning file with a
'state_dict' member whose keys carry the attribute name of the LightningModule as prefix ('model.', 'embedding_a.',
train.py:117-134); `load_ckpt` strips the prefix, drops the listed sub-prefixes and loads the rest over the model's own
state dict (missing keys keep their current values, exactly like the reference's dict.update + load_state_dict).

Parameter layouts are the reference's: every tcnn-shaped module owns ONE flat fp32 `.params` — the hash grid
level-major / entry / feature, the MLPs layer after layer, row-major (out, in), output rows
padded to 16 HERE.  tiny-cuda-nn pads output rows to 16 (FullyFusedMLP) or 8 (CutlassMLP, the reference's heads) and may pad
the input width too: `adapt_mlp_params` converts such a vector (see its docstring for what is and is not supported).
tiny-cuda-nn itself is not available offline, so that converter is unverified against a real tcnn file.
"""
import torch


def extract_model_state_dict(ckpt_path, model_name="model", prefixes_to_ignore=()):
    """utils.py:7-21.  ckpt_path may also be an already loaded dict."""
    checkpoint = torch.load(ckpt_path, map_location="cpu") if isinstance(ckpt_path, (str, bytes)) or hasattr(ckpt_path, "read") \
        else ckpt_path
    if "state_dict" in checkpoint:            # a pytorch-lightning checkpoint
        checkpoint = checkpoint["state_dict"]
    out = {}
    for k, v in checkpoint.items():
        if not k.startswith(model_name):
            continue
        k = k[len(model_name) + 1:]
        if any(k.startswith(p) for p in prefixes_to_ignore):
            continue
        out[k] = v
    return out


def load_ckpt(model, ckpt_path, model_name="model", prefixes_to_ignore=()):
    """utils.py:24-29, plus the layout adaptation a tcnn-written MLP vector may need (see module docstring)."""
    if not ckpt_path:
        return
    model_dict = model.state_dict()
    incoming = extract_model_state_dict(ckpt_path, model_name, prefixes_to_ignore)
    for k, v in list(incoming.items()):
        if k in model_dict and k.endswith(".params") and v.numel() != model_dict[k].numel():
            mod = model
            for part in k.split(".")[:-1]:
                mod = getattr(mod, part)
            mlp = getattr(mod, "mlp", None)
            if mlp is None:
                raise RuntimeError(f"{k}: checkpoint has {v.numel()} parameters, the module {model_dict[k].numel()} "
                                   "(different grid configuration: levels / features / log2_hashmap_size / scale)")
            incoming[k] = adapt_mlp_params(v, mlp.n_in, mlp.width, mlp.n_hidden, mlp.n_out, getattr(mod, 'network_config', {}).get('otype'))
    model_dict.update(incoming)
    model.load_state_dict(model_dict)


def slim_ckpt(ckpt_path, save_poses=False):
    """utils.py:32-42: drop what inference does not need."""
    ckpt = torch.load(ckpt_path, map_location="cpu") if isinstance(ckpt_path, (str, bytes)) else ckpt_path
    sd = ckpt["state_dict"]
    keys_to_pop = ["directions", "model.density_grid", "model.grid_coords"]
    if not save_poses:
        keys_to_pop += ["poses"]
    keys_to_pop += [k for k in sd if k.startswith("val_lpips")]
    for k in keys_to_pop:
        sd.pop(k, None)
    return sd


def _pad(n, a):
    return (n + a - 1) // a * a


def adapt_mlp_params(flat, n_in, width, n_hidden, n_out, otype=None):
    """A bias-free tcnn MLP vector -> this library's layout (width x n_in | (n_hidden-1) x width x width | 16-row-padded
    n_out x width).

    tiny-cuda-nn pads an MLP's OUTPUT rows to 16 (FullyFusedMLP) or to its tensor-core width 8 (CutlassMLP — the otype of
    every head in models/networks.py:89-162) and may pad the INPUT width to the same alignment (the padded columns see the
    constant 1, i.e. they act as a bias).  The source padding is taken from `otype` when it explains the vector's length,
    otherwise every (input pad, output pad) in {none, 8, 16} x {8, 16} is tried; the output rows are zero-padded or
    truncated to this library's 16 (rows >= n_out are never read).  Padded input columns are dropped only when they are all
    zero: a trained tcnn head with padded inputs (skybox 9 -> 16, tonemapper 1 -> 8) carries a learnt bias there that this
    library's bias-free networks cannot represent — such checkpoints are reported as unsupported instead of silently
    changing the function.  tiny-cuda-nn is not available offline, so the padding rules are from memory of its source
    (SURVEY.md Appendix B items marked with a dagger) and unverified against a real tcnn file."""
    hidden = (n_hidden - 1) * width * width
    pref = {"CutlassMLP": 8, "FullyFusedMLP": 16}.get(otype)
    out_pads = [a for a in ((pref,) if pref else ()) + (16, 8) if a]
    cands = []
    for oa in dict.fromkeys(out_pads):
        for ia in dict.fromkeys([1, oa, 8, 16]):
            ni, no = _pad(n_in, ia), _pad(n_out, oa)
            if width * ni + hidden + no * width == flat.numel():
                cands.append((ni, no))
    if not cands:
        raise RuntimeError(f"MLP parameter vector of {flat.numel()} values does not fit {n_in}->{width}x{n_hidden}->{n_out} "
                           f"under any tcnn padding (otype {otype})")
    ni, no = cands[0]
    W0 = flat[:width * ni].reshape(width, ni)
    if ni > n_in and float(W0[:, n_in:].abs().max()) != 0.0:
        raise RuntimeError(f"MLP first layer carries {ni - n_in} non-zero padded input columns (a tcnn bias column): unsupported — "
                           "this library's networks are bias-free at the padded inputs")
    Wl = flat[width * ni + hidden:].reshape(no, width)
    nop = _pad(n_out, 16)
    Wl16 = torch.zeros(nop, width, dtype=flat.dtype)
    Wl16[:min(no, nop)] = Wl[:min(no, nop)]
    return torch.cat([W0[:, :n_in].reshape(-1), flat[width * ni:width * ni + hidden], Wl16.reshape(-1)])
