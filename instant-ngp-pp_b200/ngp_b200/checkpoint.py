"""Checkpoint plumbing of the hot path's models (reference utils.py:7-42).

Same three entry points, same semantics: a checkpoint is either a bare state dict or a PyTorch-Lightning file with a
'state_dict' member whose keys carry the attribute name of the LightningModule as prefix ('model.', 'embedding_a.',
train.py:117-134); `load_ckpt` strips the prefix, drops the listed sub-prefixes and loads the rest over the model's own
state dict (missing keys keep their current values, exactly like the reference's dict.update + load_state_dict).

Parameter layouts are the reference's: every tcnn-shaped module owns ONE flat fp32 `.params` — the hash grid
level-major / entry / feature, the MLPs layer after layer, row-major (out, in), output rows
padded to 16.  tiny-cuda-nn pads an MLP's INPUT width to its alignment as well (16 for FullyFusedMLP, 8 for CutlassMLP;
the padded columns see the constant 1): `adapt_mlp_params` converts such a vector to this library's unpadded first
layer.  tiny-cuda-nn itself is not available offline, so that converter is unverified against a real tcnn file.
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
            incoming[k] = adapt_mlp_params(v, mlp.n_in, mlp.width, mlp.n_hidden, mlp.n_out)
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


def adapt_mlp_params(flat, n_in, width, n_hidden, n_out):
    """A bias-free MLP vector whose FIRST layer was stored with a padded input width (width x n_in_padded, tcnn) ->
    this library's (width x n_in | (n_hidden-1) x width x width | n_out_pad16 x width).  The padded input columns
    multiply the constant 1 in tcnn — a bias this library's networks do not have — so they are only dropped when they
    are all zero; otherwise the mismatch is reported instead of silently changing the function."""
    nop = (n_out + 15) // 16 * 16
    rest = (n_hidden - 1) * width * width + nop * width
    first = flat.numel() - rest
    if first <= 0 or first % width:
        raise RuntimeError(f"MLP parameter vector of {flat.numel()} values does not fit {n_in}->{width}x{n_hidden}->{n_out}")
    n_in_padded = first // width
    if n_in_padded < n_in:
        raise RuntimeError(f"MLP first layer has {n_in_padded} input columns, the module expects {n_in}")
    W0 = flat[:first].reshape(width, n_in_padded)
    if n_in_padded > n_in and float(W0[:, n_in:].abs().max()) != 0.0:
        raise RuntimeError(f"MLP first layer carries {n_in_padded - n_in} non-zero padded input columns (a tcnn bias column); "
                           "this library's networks are bias-free at the padded inputs")
    return torch.cat([W0[:, :n_in].reshape(-1), flat[first:]])
