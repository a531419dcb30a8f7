"""Import shim for `torch_scatter` (reference models/custom_functions.py:4: `from torch_scatter import segment_csr`).

The package is not installed in this image and cannot be fetched.  The reference's only use is RayMarcher.backward
(custom_functions.py:104-114), which is unreachable because the marcher is called under no_grad
(models/rendering.py:207-212); the shim still implements segment_csr(sum) faithfully so that the import — and that
dead backward, if anyone enables it — work.
"""
import torch


def segment_csr(src, indptr, out=None, reduce="sum"):
    """out[i] = reduce(src[indptr[i]:indptr[i+1]]) along dim 0 (torch_scatter semantics for 1-D indptr)."""
    if reduce not in ("sum", "add"):
        raise NotImplementedError("torch_scatter shim: only reduce='sum'")
    lengths = (indptr[1:] - indptr[:-1]).to(torch.int64)
    res = torch.segment_reduce(src, "sum", lengths=lengths, axis=0, unsafe=True)
    if out is not None:
        out.copy_(res)
        return out
    return res
