"""One optimiser step of the hot path (reference train.py:268-345 `training_step` minus Lightning):
occupancy update every 16 steps -> render (AABB, march, field, composite) -> NeRFLoss -> backward
-> gradient all-reduce across ranks (rays are sharded, parameters replicated) -> Adam.

Multi-GPU follows the reference's data-parallel semantics (Lightning DDPPlugin, train.py:431): every
rank holds the full model and an R/N slice of the ray batch; per-rank mean losses are averaged by
averaging gradients.  The all-reduce is issued per parameter tensor on torch.distributed (NCCL over
NVLink on the GPU box, gloo in the CPU tests) — hash-table gradients first, since they dominate.
"""
import math

import torch
import torch.distributed as dist

from .losses import NeRFLoss
from .optim import FusedAdam
from .rendering import render


class Trainer:
    def __init__(self, model, lr=1e-2, eps=1e-15, update_interval=16, warmup_steps=256, density_threshold=0.01 * 1024 / 3 ** 0.5,
                 lambda_distortion=3e-4, lambda_opa=2e-4, render_kwargs=None, world_size=1, max_grad_norm=None,
                 extra_params=(), exchange="after", lazy_adam=False):
        self.model = model
        self.loss_fn = NeRFLoss(lambda_opa=lambda_opa, lambda_distortion=lambda_distortion)
        # extra_params: parameters outside the field that the step also trains (the appearance embedding, train.py:117-119)
        self.extra_params = [p for p in extra_params if p.numel() > 0]
        params = [p for p in model.parameters() if p.numel() > 0] + self.extra_params
        on_gpu = len(params) > 0 and params[0].is_cuda
        # train.py:244-251,435: Adam + global-norm clip, fused (device-resident clip coefficient, 1/world averaging)
        # Gradient exchange mode, see below ("sharded" lives in the optimiser: reduce-scatter -> Adam on 1/world -> all-gather)
        import os
        self._exchange = os.environ.get("NGP_DP_EXCHANGE", exchange)
        shard = None
        if world_size > 1 and on_gpu and self._exchange == "sharded":
            shard = (dist.get_rank(), world_size)
        self.opt = (FusedAdam(params, lr=lr, eps=eps, max_grad_norm=max_grad_norm, grad_scale=1.0 / world_size, shard=shard,
                              lazy=lazy_adam) if on_gpu
                    else torch.optim.Adam(params, lr=lr, eps=eps))
        self.fused = on_gpu
        self.update_interval, self.warmup_steps = update_interval, warmup_steps
        self.density_threshold = density_threshold
        self.render_kwargs = dict(render_kwargs or {})
        self.world_size = world_size
        self.step = 0
        self.last_samples = None
        self._loss_host = self._loss_event = None
        self._works, self._sinks = [], []
        # Gradient exchange mode.  "after" (default): per-tensor async all-reduce issued after backward(), waited for before Adam.
        # "overlap": table slices all-reduced under the remaining scatter launches (tcnn.GradSink) — measured SLOWER on B200 +
        # NVSwitch for the 44 MB table (profiles/r02b_dp_exchange_probe.txt: 7.22 vs 7.09 ms/step at N=2; the all-reduce is only
        # 0.04 ms exposed and competes with the L2-bound scatter when overlapped), kept for tables where the exchange is long.
        # "sharded": the tensors FusedAdam.is_sharded names (the hash tables) are not all-reduced at all — the optimiser
        # reduce-scatters their gradients, updates its 1/world slice (optimiser state and traffic / world) and all-gathers the
        # parameters in place; everything else goes through the all-reduce below.
        # "none": timing probe only.  NGP_DP_EXCHANGE overrides.
        if world_size > 1 and on_gpu and self._exchange == "overlap":
            self._install_grad_sinks()

    def _install_grad_sinks(self):
        """Overlapped gradient exchange for the hash table of the fused density path (> 99 % of the gradient bytes): the
        table gradient is scattered level range by level range into a persistent buffer that IS table.grad, and every
        finished range's all-reduce is issued at once on the process group's stream, under the next range's scatter
        (tcnn.GradSink).  What stays exposed is the last, smallest range (the dense coarse levels) and the MLP weights."""
        from . import tcnn
        enc, net = getattr(self.model, "xyz_encoder", None), getattr(self.model, "sigma_net", None)
        if enc is None or net is None or not getattr(self.model, "fused_density", False) or enc.params.dtype != torch.float32:
            return
        buf = torch.zeros_like(enc.params)
        sink = tcnn.GradSink(buf, enc.grid, lambda a, b: self._works.append(dist.all_reduce(buf[a:b], op=dist.ReduceOp.SUM, async_op=True)))
        tcnn.GRAD_SINKS[enc.params.data_ptr()] = sink
        self._sinks.append((enc.params, sink))

    def allreduce_grads(self):
        if self.world_size <= 1:
            return
        # largest tensors first: the hash tables are > 99 % of the bytes
        sunk = {id(p) for p, _ in self._sinks}           # already in flight, slice by slice, since the backward pass
        in_opt = getattr(self.opt, "is_sharded", lambda p: False)      # exchanged by the sharded optimiser itself
        ps = sorted((p for p in list(self.model.parameters()) + self.extra_params
                     if p.grad is not None and id(p) not in sunk and not in_opt(p)), key=lambda p: -p.numel())
        works = self._works + [dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, async_op=True) for p in ps]
        self._works = []
        for w in works:
            w.wait()
        if not self.fused:                     # the fused optimiser folds 1/world_size into its gradient scale
            for p in ps:
                p.grad.div_(self.world_size)

    def backward_and_exchange(self, loss):
        """loss.backward() + the data-parallel gradient sum; afterwards every p.grad holds the SUM over ranks (the optimiser
        folds in 1/world_size)."""
        self.opt.zero_grad(set_to_none=True)
        for p, sink in self._sinks:            # the scatter accumulates in place into table.grad (tcnn.GradSink)
            sink.buf.zero_()
            p.grad = sink.buf
        loss.backward()
        self.allreduce_grads()

    def train_step(self, rays_o, rays_d, rgb_gt, update_grid=True, target=None, host_loss=False, **step_kwargs):
        """-> (loss 0-dim tensor, results dict).  No host sync besides the marcher's sample count.
        target: further ground-truth tensors of the batch ('label', 'normal', 'depth': train.py:275-283);
        step_kwargs: per-step render arguments (embedding_a = the batch's appearance embeddings, train.py:285-288).
        host_loss=True returns the loss as a python float (what the reference's progress bar / logger reads every step,
        train.py:337-343): the value is copied to pinned host memory as soon as the forward pass has produced it and
        waited for only after backward + optimiser have been enqueued, so the read-back never drains the GPU."""
        m = self.model
        if update_grid and self.step % self.update_interval == 0:
            m.update_density_grid(self.density_threshold, warmup=self.step < self.warmup_steps)
        kw = {**self.render_kwargs, **step_kwargs} if step_kwargs else self.render_kwargs
        results = render(m, rays_o, rays_d, **kw)
        loss = self.loss_fn.total(results, {"rgb": rgb_gt, **(target or {})}, **kw)        # train.py:310 sum(lo.mean())
        if host_loss:
            if self._loss_host is None:
                self._loss_host = torch.empty((), dtype=torch.float32).pin_memory()
                self._loss_event = torch.cuda.Event()
            self._loss_host.copy_(loss.detach(), non_blocking=True)
            self._loss_event.record()
        self.backward_and_exchange(loss)
        self.opt.step()
        self.step += 1
        self.last_samples = results["total_samples"]
        if host_loss:
            self._loss_event.synchronize()
            return float(self._loss_host), results
        return loss.detach(), results


def psnr(pred, target):
    return -10.0 * torch.log10(torch.mean((pred - target) ** 2))
