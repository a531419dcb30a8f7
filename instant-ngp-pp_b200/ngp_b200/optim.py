"""Fused Adam (+ optional global-norm clip) over the flat parameter vectors; dense, lazy and rank-sharded forms.

Replaces torch.optim.Adam + Lightning's gradient_clip_val=50 + the DDP gradient all-reduce in front of it on the
reference's training path (train.py:244-251, 431, 435).  Same update rule as torch.optim.Adam
(no weight decay, no amsgrad); one kernel per parameter tensor reading p,g,m,v and writing p,m,v once.  The clip coefficient
(and the 1/world_size gradient averaging of data-parallel training) stay on the device.  torch LR schedulers
(CosineAnnealingLR, train.py:249-251) drive it like any torch optimiser: `lr` is read from the param group every step.

  lazy=True        hash-table tensors (>= lazy_min_numel elements) skip entries whose gradient is exactly zero —
                   tiny-cuda-nn's rule for encoding parameters.  Opt-in (torch.optim.Adam, the reference's optimiser,
                   keeps moving untouched entries on their momentum).
  shard=(rank, world[, group])
                   ZeRO-1 style exchange for the big tensors (numel >= shard_min_numel, divisible by 4*world): the summed
                   gradient is REDUCE-SCATTERed, each rank runs Adam on its 1/world slice (exp_avg / exp_avg_sq exist for
                   that slice only), and the updated slices are ALL-GATHERed in place into the replicated parameter.
                   Same wire bytes as the all-reduce it replaces; optimiser traffic and state divided by world.  The
                   clip norm is assembled from the shards (one scalar all-reduce).  Tensors that are not sharded must
                   arrive already summed over ranks (Trainer.allreduce_grads skips the ones `is_sharded` names).
"""
import torch
import torch.distributed as dist

from . import _lib
from ._lib import lib, ptr, check, stream


def _cuda_update(p, g, m, v, lr, b1, b2, eps, step, coef, lazy):
    """One Adam update of the flat fp32 views through the CUDA kernel (coef: device pointer of the gradient scale or None)."""
    fn = lib.ngp_adam_step_lazy if lazy else lib.ngp_adam_step
    check(fn(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr), float(b1), float(b2), float(eps), int(step), coef, stream()),
          "adam_step")


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=None, grad_scale=1.0, lazy=False,
                 lazy_min_numel=1 << 20, shard=None, shard_min_numel=1 << 20, update_fn=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.max_grad_norm = max_grad_norm
        self.grad_scale = grad_scale          # e.g. 1/world_size when gradients were summed across ranks
        self.lazy, self.lazy_min_numel = lazy, lazy_min_numel
        self.rank, self.world, self.group = (shard[0], shard[1], shard[2] if len(shard) > 2 else None) if shard else (0, 1, None)
        self.shard_min_numel = shard_min_numel
        # update_fn(p, g, m, v, lr, b1, b2, eps, step, coef_tensor_or_None, lazy): the gloo tests of the sharding logic inject a
        # torch-op restatement here; the product path is the CUDA kernel and nothing else
        self._update = update_fn
        self._scratch = None

    def is_sharded(self, p):
        return self.world > 1 and p.numel() >= self.shard_min_numel and p.numel() % (4 * self.world) == 0 and p.is_contiguous()

    def _coef(self, ps, gshards, dev):
        """Device scalar the gradients are multiplied by: clip coefficient x grad_scale (None when it is 1)."""
        if self.max_grad_norm is None and self.grad_scale == 1.0:
            return None
        if self._scratch is None or self._scratch.device != dev:
            self._scratch = torch.zeros(2, dtype=torch.float32, device=dev)
        sc = self._scratch
        if self.max_grad_norm is None:
            sc[1] = self.grad_scale
            return sc
        sc.zero_()
        if self._update is None:
            for t in gshards:
                check(lib.ngp_grad_sumsq(ptr(t), t.numel(), ptr(sc), stream()), "grad_sumsq")
            if gshards:                                   # every rank holds a different slice: sum the partial norms
                dist.all_reduce(sc[:1], op=dist.ReduceOp.SUM, group=self.group)
            for _, p in ps:
                if not self.is_sharded(p):
                    check(lib.ngp_grad_sumsq(ptr(p.grad), p.numel(), ptr(sc), stream()), "grad_sumsq")
            # gradients are still un-averaged: clip on the averaged norm
            check(lib.ngp_clip_coef(ptr(sc), float(self.max_grad_norm) / self.grad_scale, sc.data_ptr() + 4, stream()), "clip_coef")
        else:
            if gshards:
                sc[0] = sum((t.double() ** 2).sum() for t in gshards).float()
                dist.all_reduce(sc[:1], op=dist.ReduceOp.SUM, group=self.group)
            sc[0] += sum((p.grad.double() ** 2).sum() for _, p in ps if not self.is_sharded(p))
            sc[1] = torch.clamp(float(self.max_grad_norm) / self.grad_scale / (sc[0].sqrt() + 1e-6), max=1.0)
        if self.grad_scale != 1.0:
            sc[1].mul_(self.grad_scale)
        return sc

    @torch.no_grad()
    def step(self, closure=None):
        if self._update is None:
            _lib.require_device()
        ps = [(g, p) for g in self.param_groups for p in g["params"] if p.grad is not None]
        if not ps:
            return None
        dev = ps[0][1].device
        # 1. reduce-scatter of the big tensors' gradients (issued together, waited for before the first use)
        shards, works = {}, []
        for _, p in ps:
            if self.is_sharded(p):
                n = p.numel() // self.world
                gs = torch.empty(n, dtype=p.grad.dtype, device=dev)
                works.append(dist.reduce_scatter_tensor(gs, p.grad.contiguous().view(-1), op=dist.ReduceOp.SUM, group=self.group,
                                                        async_op=True))
                shards[id(p)] = gs
        for w in works:
            w.wait()
        sc = self._coef(ps, list(shards.values()), dev)
        coef = None if sc is None else (sc.data_ptr() + 4 if self._update is None else sc[1])
        # 2. Adam: shard owners on their slice, everything else whole
        gathers = []
        for g, p in ps:
            st = self.state[p]
            sharded = id(p) in shards
            flat = p.view(-1) if sharded else p
            lo = self.rank * (p.numel() // self.world) if sharded else 0
            mine = flat[lo:lo + p.numel() // self.world] if sharded else p
            grad = shards[id(p)] if sharded else p.grad
            if not st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(mine, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(mine, memory_format=torch.preserve_format)
            st["step"] += 1
            b1, b2 = g["betas"]
            lazy = self.lazy and p.numel() >= self.lazy_min_numel
            (self._update or _cuda_update)(mine, grad, st["exp_avg"], st["exp_avg_sq"], g["lr"], b1, b2, g["eps"], st["step"], coef, lazy)
            if sharded:                                   # in place: rank r's slice already sits at offset r of the output
                gathers.append(dist.all_gather_into_tensor(flat, mine, group=self.group, async_op=True))
        for w in gathers:
            w.wait()
        return None
