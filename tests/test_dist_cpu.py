"""Host-side logic of the batch-sharded data parallelism (vit_torch_b200/dist.py) on CPU: world_size 2, gloo.

The block buckets are normally produced by BlockFn.backward on the GPU; here the bucket hook is driven by hand with the
same (flat buffer, params, alias views) protocol, plus the trailing reduction of non-bucket parameters."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, overlap, ret, arena=False, static=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vit_torch_b200 import functional as Fn
    from vit_torch_b200.dist import GradAllReducer
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4))   # identical replicas
    red = GradAllReducer(model, overlap=overlap)
    assert red.world == world and len(Fn.grad_bucket_hooks) == 1
    # per-rank data shard
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(6, 8, generator=g)
    loss = model(x).pow(2).mean()
    loss.backward()
    # emulate one block bucket: layer 0's grads live in a flat buffer whose views are the .grad tensors
    p0 = list(model[0].parameters())
    flat = torch.cat([p.grad.reshape(-1) for p in p0])
    if arena:   # the Trainer's gradient arena: bucket buffers are adjacent slices of one buffer -> one collective
        Fn.grad_arena = Fn.GradArena(1024, torch.device("cpu"))
        Fn.grad_arena.begin()
        a = Fn.grad_arena.take(flat.numel(), flat.device)
        a.copy_(flat)
        flat = a
    views, off = [], 0
    for p in p0:
        v = flat[off:off + p.numel()].view(p.shape)
        p.grad = v                      # what autograd does when it adopts the returned view
        views.append(v.view(v.shape))   # alias handed to the hook
        off += p.numel()
    if static:   # the graph Trainer's tail: no bucket notifications, the arena and the rest are reduced directly
        red.finish_static(Fn.grad_arena)
    else:
        for hook in Fn.grad_bucket_hooks:
            hook(flat, p0, views)
        red.finish()
    if arena:
        assert red.collectives == 2     # one over the arena, one for the parameters outside any bucket
        Fn.grad_arena = None
    ret[rank] = [p.grad.clone() for p in model.parameters()]
    red.close()
    assert len(Fn.grad_bucket_hooks) == 0
    dist.destroy_process_group()


def _run(overlap, arena=False, static=False):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), overlap, ret, arena, static), nprocs=world, join=True)
    # reference: average of the per-rank gradients computed in this process
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4))
    acc = None
    for rank in range(world):
        model.zero_grad()
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn(6, 8, generator=g)
        model(x).pow(2).mean().backward()
        gs = [p.grad.clone() for p in model.parameters()]
        acc = gs if acc is None else [a + b for a, b in zip(acc, gs)]
    want = [a / world for a in acc]
    for rank in range(world):
        for got, w in zip(ret[rank], want):
            assert torch.allclose(got, w, atol=1e-6), rank


def test_grad_allreduce_overlapped_gloo():
    _run(True)


def test_grad_allreduce_deferred_gloo():
    _run(False)


def test_grad_allreduce_deferred_arena_gloo():
    _run(False, arena=True)


def test_grad_allreduce_static_arena_gloo():
    _run(False, arena=True, static=True)
