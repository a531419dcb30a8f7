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
