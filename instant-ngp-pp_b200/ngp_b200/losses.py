"""Loss terms that consume the hot path's outputs (reference losses.py:7-152).

Only DistortionLoss touches a kernel; the rest is elementwise torch on (N_rays, .) tensors and is
kept so that a training step produces exactly the upstream gradients the reference's step does
(SURVEY.md §3.1 "which gradients actually flow").
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .custom_functions import DistortionLoss  # noqa: F401  (re-exported: losses.py:32-58)


def compute_scale_and_shift(prediction, target, weight=None):
    """Least-squares (scale, shift) aligning prediction to target (losses.py:7-30), sync-free.
    weight (0/1 per element, optional): the solve runs over the elements with weight 1 — what the reference obtains by
    indexing `results['depth'][mask]`, `depth_2d[mask]` first (losses.py:128) — without a boolean-mask gather; a_11 is then
    the number of VALID elements, not numel()."""
    if weight is None:
        a11 = torch.tensor(float(prediction.numel()), device=prediction.device)
    else:
        prediction, target, a11 = prediction * weight, target * weight, weight.sum()
    a00, a01 = (prediction * prediction).sum(), prediction.sum()
    b0, b1 = (prediction * target).sum(), target.sum()
    det = a00 * a11 - a01 * a01
    ok = det != 0
    safe = torch.where(ok, det, torch.ones_like(det))
    return torch.where(ok, (a11 * b0 - a01 * b1) / safe, torch.zeros_like(det)), \
        torch.where(ok, (-a01 * b0 + a00 * b1) / safe, torch.zeros_like(det))


class ExponentialAnnealingWeight:
    def __init__(self, max, min, k):
        self.max, self.min, self.k = max, min, k

    def getWeight(self, Tcur):
        return max(self.min, self.max * math.exp(-Tcur * self.k))


class _BasicLossFn(torch.autograd.Function):
    """mean((rgb - target)**2) + mean(lambda_opa * -(o + 1e-10) * log(o + 1e-10)) as ONE kernel that also writes both gradients
    (csrc/losses.cu ngp_basic_loss); the backward only scales them by the upstream scalar."""

    @staticmethod
    def forward(ctx, rgb, target, opacity, lambda_opa):
        from . import _lib
        from ._lib import lib, ptr, check, stream
        _lib.require_device()
        rgb, target, opacity = rgb.contiguous(), target.contiguous(), opacity.contiguous()
        n = rgb.shape[0]
        loss = torch.zeros((), device=rgb.device)
        drgb = torch.empty_like(rgb); dopa = torch.empty_like(opacity)
        check(lib.ngp_basic_loss(ptr(rgb), ptr(target), ptr(opacity), n, float(lambda_opa), ptr(loss), ptr(drgb), ptr(dopa), stream()), "basic_loss")
        ctx.save_for_backward(drgb, dopa)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        drgb, dopa = ctx.saved_tensors
        return (drgb * g if ctx.needs_input_grad[0] else None), None, (dopa * g if ctx.needs_input_grad[2] else None), None


class NeRFLoss(nn.Module):
    """Per-ray loss dictionary; weights as in losses.py:75-83."""

    def __init__(self, lambda_opa=2e-4, lambda_distortion=3e-4):
        super().__init__()
        self.lambda_opa = lambda_opa
        self.lambda_distortion = lambda_distortion
        self.lambda_depth_mono = 1
        self.lambda_normal_mono = 1e-3
        self.lambda_normal_ref_rp = 1e-3
        self.lambda_normal_ref_ro = 1e-3
        self.lambda_sky = 1e-1
        self.lambda_semantic = 4e-2
        self.Annealing = ExponentialAnnealingWeight(max=1, min=6e-2, k=1e-3)
        self.CrossEntropyLoss = nn.CrossEntropyLoss(ignore_index=256)

    def total(self, results, target, **kwargs):
        """sum(v.mean() for v in self(results, target, **kwargs).values()) — train.py:310 — as a scalar tensor.  The photometric and
        opacity terms, which every configuration has, come from one fused kernel with their gradients; the others from forward()."""
        rgb, opa = results["rgb"], results["opacity"]
        fused = (rgb.is_cuda and rgb.dtype == torch.float32 and opa.dtype == torch.float32 and not kwargs.get("embed_msk", False)
                 and rgb.dim() == 2 and rgb.shape[1] == 3 and target["rgb"].shape == rgb.shape and target["rgb"].dtype == torch.float32)
        if not fused:
            return sum(v.mean() for v in self(results, target, **kwargs).values())
        loss = _BasicLossFn.apply(rgb, target["rgb"], opa, self.lambda_opa)
        rest = self(results, target, _skip_basic=True, **kwargs)
        for v in rest.values():
            loss = loss + v.mean()
        return loss

    def forward(self, results, target, _skip_basic=False, **kwargs):
        d = {}
        if _skip_basic:
            return self._other_terms(d, results, target, **kwargs)
        if kwargs.get("embed_msk", False):
            m = kwargs["mask"]
            d["r_ms"] = torch.mean(m ** 2) * self.Annealing.getWeight(kwargs["step"])
            d["rgb"] = (1 - m) * (results["rgb"] - target["rgb"]) ** 2
        else:
            d["rgb"] = (results["rgb"] - target["rgb"]) ** 2
        o = results["opacity"] + 1e-10
        d["opacity"] = self.lambda_opa * (-o * torch.log(o))
        return self._other_terms(d, results, target, **kwargs)

    def _other_terms(self, d, results, target, **kwargs):
        if self.lambda_distortion > 0:
            d["distortion"] = self.lambda_distortion * DistortionLoss.apply(
                results["ws"], results["deltas"], results["ts"], results["rays_a"])
        if kwargs.get("normal_ref", False):
            d["normal_ref_rp"] = self.lambda_normal_ref_rp * results["Rp"]
            d["normal_ref_ro"] = self.lambda_normal_ref_ro * results["Ro"]
        if kwargs.get("normal_mono", False):
            n_pred = F.normalize(results["normal_pred"], dim=-1)
            n_gt = F.normalize(target["normal"], dim=-1)
            d["normal_mono"] = self.lambda_normal_mono * (torch.abs(n_pred - n_gt) + 0.1 * (-(n_pred * n_gt)))
        if kwargs.get("semantic", False):
            d["CELoss"] = self.lambda_semantic * self.CrossEntropyLoss(results["semantic"], target["label"])
            sky = torch.where(target["label"] == 4, 1., 0.)
            d["sky_depth"] = self.lambda_sky * sky * torch.exp(-results["depth"])
        if kwargs.get("depth_mono", False):
            depth_2d = target["depth"] / 25
            w = (depth_2d > 0).float()
            pred = results["depth"].detach()
            scale, shift = compute_scale_and_shift(pred, depth_2d, weight=w)
            d["depth_mono"] = w * self.lambda_depth_mono * torch.exp(-pred / kwargs.get("scale", 1)) * \
                (scale * results["depth"] + shift - depth_2d) ** 2
        return d
