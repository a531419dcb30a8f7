"""Fused dense Adam (+ optional global-norm clip) over the flat parameter vectors.

Replaces torch.optim.Adam + Lightning's gradient_clip_val=50 on the reference's training path
(train.py:244-251, 435).  Same update rule as torch.optim.Adam (no weight decay,
no amsgrad); one kernel per parameter tensor reading p,g,m,v and writing p,m,v once.  The clip
coefficient (and the 1/world_size gradient averaging of data-parallel training) stay on the device.
"""
import torch

from . import _lib
from ._lib import lib, ptr, check, stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=None, grad_scale=1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.max_grad_norm = max_grad_norm
        self.grad_scale = grad_scale          # e.g. 1/world_size when gradients were summed across ranks
        self._scratch = None

    @torch.no_grad()
    def step(self, closure=None):
        _lib.require_device()
        ps = [(g, p) for g in self.param_groups for p in g["params"] if p.grad is not None]
        if not ps:
            return None
        dev = ps[0][1].device
        coef = None
        if self.max_grad_norm is not None or self.grad_scale != 1.0:
            if self._scratch is None or self._scratch.device != dev:
                self._scratch = torch.zeros(2, dtype=torch.float32, device=dev)
            sc = self._scratch
            if self.max_grad_norm is not None:
                sc.zero_()
                for _, p in ps:
                    check(lib.ngp_grad_sumsq(ptr(p.grad), p.numel(), ptr(sc), stream()), "grad_sumsq")
                # gradients are still un-averaged: clip on the averaged norm
                check(lib.ngp_clip_coef(ptr(sc), float(self.max_grad_norm) / self.grad_scale, sc.data_ptr() + 4, stream()),
                      "clip_coef")
                if self.grad_scale != 1.0:
                    sc[1].mul_(self.grad_scale)
            else:
                sc[1] = self.grad_scale
            coef = sc.data_ptr() + 4
        for g, p in ps:
            st = self.state[p]
            if not st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] += 1
            b1, b2 = g["betas"]
            check(lib.ngp_adam_step(ptr(p), ptr(p.grad), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(), float(g["lr"]),
                                    float(b1), float(b2), float(g["eps"]), int(st["step"]), coef, stream()), "adam_step")
        return None
