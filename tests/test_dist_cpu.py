"""world_size-2 gloo test of the multi-GPU host logic (ray sharding + gradient all-reduce semantics of
ngp_b200.trainer.Trainer) on CPU: averaged gradients of two rank-local half batches must equal the
gradient of the union batch (the reference's Lightning DDP contract, train.py:431)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "instant-ngp-pp_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ngp_b200.trainer import Trainer
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    emb = torch.nn.Embedding(4, 6)                                # trained outside the field (appearance embedding, train.py:117-119)
    tr = Trainer(model, lr=1e-2, world_size=world, extra_params=emb.parameters())
    g = torch.Generator().manual_seed(1)
    x = torch.randn(64, 6, generator=g); y = torch.randn(64, 3, generator=g)
    img = torch.randint(4, (64,), generator=g)
    shard = slice(rank * 32, (rank + 1) * 32)                     # rays shard, parameters replicate
    loss = ((model(x[shard] + emb(img[shard])) - y[shard]) ** 2).mean()
    tr.opt.zero_grad(); loss.backward(); tr.allreduce_grads()
    every = list(model.parameters()) + list(emb.parameters())
    grads = torch.cat([p.grad.reshape(-1) for p in every])
    tr.opt.step()
    params = torch.cat([p.detach().reshape(-1) for p in every])
    if rank == 0:
        # union-batch reference on one process
        torch.manual_seed(0)
        ref = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
        remb = torch.nn.Embedding(4, 6)
        rall = list(ref.parameters()) + list(remb.parameters())
        opt = torch.optim.Adam(rall, lr=1e-2, eps=1e-15)
        ((ref(x + remb(img)) - y) ** 2).mean().backward()
        rg = torch.cat([p.grad.reshape(-1) for p in rall])
        opt.step()
        rp = torch.cat([p.detach().reshape(-1) for p in rall])
        torch.save(dict(g=grads, rg=rg, p=params, rp=rp), out)
    gathered = [torch.zeros_like(params) for _ in range(world)]
    dist.all_gather(gathered, params)
    assert all(torch.equal(gathered[0], t) for t in gathered)       # replicas stay bit-identical
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_union_batch(tmp_path):
    out = str(tmp_path / "r.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert torch.allclose(r["g"], r["rg"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(r["p"], r["rp"], rtol=1e-5, atol=1e-7)


# ---------------------------------------------------------------------------------------------------------------------
# Rank-sharded optimiser: reduce-scatter -> Adam on the 1/world slice -> all-gather, host logic
# on gloo with the update kernel replaced by a torch-op restatement (the product path is the CUDA kernel, test_dist_gpu).
def _torch_adam_update(p, g, m, v, lr, b1, b2, eps, step, coef, lazy):
    g = g * coef if coef is not None else g
    keep = (g == 0) if lazy else torch.zeros_like(g, dtype=torch.bool)
    m_new = b1 * m + (1 - b1) * g
    v_new = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p_new = p - (lr / bc1) * m_new / (v_new.sqrt() / bc2 ** 0.5 + eps)
    m.copy_(torch.where(keep, m, m_new)); v.copy_(torch.where(keep, v, v_new)); p.copy_(torch.where(keep, p, p_new))


def _sharded_worker(rank, world, port, out, clip):
    sys.path.insert(0, os.path.join(ROOT, "instant-ngp-pp_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ngp_b200.optim import FusedAdam
    torch.manual_seed(0)
    table = torch.nn.Parameter(torch.randn(4096) * 0.1)           # "hash table": sharded (>= shard_min_numel, divisible)
    small = torch.nn.Parameter(torch.randn(7, 3))                 # "MLP weights": replicated, all-reduced by the caller
    opt = FusedAdam([table, small], lr=1e-2, eps=1e-15, max_grad_norm=clip, grad_scale=1.0 / world, shard=(rank, world),
                    shard_min_numel=1024, update_fn=_torch_adam_update)
    ref_t, ref_s = table.detach().clone().requires_grad_(True), small.detach().clone().requires_grad_(True)
    ref = torch.optim.Adam([ref_t, ref_s], lr=1e-2, eps=1e-15)
    g = torch.Generator().manual_seed(1)
    assert opt.is_sharded(table) and not opt.is_sharded(small)
    for it in range(3):
        idx = torch.randint(4096, (64,), generator=g); y = torch.randn(64, generator=g); x = torch.randn(64, 7, generator=g)
        sh = slice(rank * 32, rank * 32 + 32)
        def loss_of(t, s, sl):
            return ((t[idx[sl]] * 3 + (x[sl] @ s).sum(-1) - y[sl]) ** 2).mean()
        opt.zero_grad()
        loss_of(table, small, sh).backward()
        dist.all_reduce(small.grad)                               # replicated tensors: the trainer's all-reduce
        opt.step()
        ref.zero_grad()
        loss_of(ref_t, ref_s, slice(0, 64)).backward()
        if clip is not None:
            torch.nn.utils.clip_grad_norm_([ref_t, ref_s], clip)
        ref.step()
    st = opt.state[table]
    assert st["exp_avg"].numel() == 4096 // world                 # optimiser state exists for the rank's slice only
    both = [torch.zeros(4096) for _ in range(world)]
    dist.all_gather(both, table.detach())
    assert torch.equal(both[0], both[1])                          # replicas identical after the in-place all-gather
    if rank == 0:
        torch.save(dict(t=table.detach(), rt=ref_t.detach(), s=small.detach(), rs=ref_s.detach()), out)
    dist.destroy_process_group()


@pytest.mark.parametrize("clip", [None, 0.05])
def test_two_rank_sharded_adam_equals_dense_adam_on_the_union_batch(tmp_path, clip):
    out = str(tmp_path / "s.pt")
    port = 31500 + os.getpid() % 2000 + (7 if clip else 0)
    mp.spawn(_sharded_worker, args=(2, port, out, clip), nprocs=2, join=True)
    r = torch.load(out)
    assert torch.allclose(r["t"], r["rt"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(r["s"], r["rs"], rtol=1e-5, atol=1e-7)
