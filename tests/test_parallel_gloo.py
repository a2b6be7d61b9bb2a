"""CPU, world_size 2, gloo: the data-parallel plumbing of dctn_b200.parallel.

The EPS kernels themselves need a GPU; what is checked here is the host logic around them: batch sharding,
parameter broadcast, gradient averaging through the gradient-ready hooks (must equal the single-process gradient of
the full batch), identical core-dropout masks on all ranks, and the sharded evaluation reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from dctn_b200.parallel import GradAllReducer, all_reduce_metrics, seed_core_dropout, shard_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(7)
    return torch.nn.Sequential(torch.nn.Linear(12, 8), torch.nn.Tanh(), torch.nn.Linear(8, 10)).double()


def _worker(rank, world, port, ret, overlap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(1, 8, 12, generator=g, dtype=torch.float64)  # (C, B, features): batch on dim 1
        y = torch.randint(0, 10, (8,), generator=g)
        model = _model()
        if rank == 1:  # replicas start different; GradAllReducer must broadcast rank 0's parameters
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)
        # overlap: parameters in the order their gradients become ready (last layer first)
        plist = list(model.parameters())
        reducer = GradAllReducer(plist[::-1] if overlap else plist, overlap=overlap)
        xs, ys = shard_batch(x, rank, world, dim=1), shard_batch(y, rank, world, dim=0)
        assert xs.shape == (1, 4, 12) and ys.shape == (4,)
        for step in range(2):   # second step: zero_grad() must restore a clean bucket
            reducer.zero_grad()
            assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(reducer.params, reducer._views))
            loss = F.cross_entropy(model(xs[0]), ys)
            loss.backward()
            reducer.wait()
        grads = [p.grad.clone() for p in model.parameters()]
        # a caller that detaches the gradients (optimizer.zero_grad(set_to_none=True)) and a parameter without a gradient
        # on ONE rank only: the fixed bucket layout keeps the ranks in step, the missing gradient counts as zero
        for p in model.parameters():
            p.grad = None
        out = model(xs[0])
        if rank == 0:
            F.cross_entropy(out, ys).backward()
        else:
            model[2].weight.requires_grad_(False)
            F.cross_entropy(out.detach() @ torch.eye(10, dtype=torch.float64) + model[2].bias, ys).backward()   # bias only
            model[2].weight.requires_grad_(True)
        reducer.wait()
        if rank == 0:
            ret["partial"] = [p.grad.clone() for p in model.parameters()]
        # identical dropout masks on every rank
        seed_core_dropout(123, 5, torch.device("cpu"))
        mask = torch.bernoulli(torch.full((16,), 0.5))
        gathered = [torch.zeros_like(mask) for _ in range(world)]
        dist.all_gather(gathered, mask)
        assert all(torch.equal(gathered[0], m) for m in gathered)
        mean_loss, acc = all_reduce_metrics(float(rank + 1) * 4, float(rank) * 2, 4.0, torch.device("cpu"))
        if rank == 0:
            ret["grads"] = grads
            ret["params"] = [p.detach().clone() for p in model.parameters()]
            ret["metrics"] = (mean_loss, acc)
        reducer.remove()
    finally:
        dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("overlap", [False, True])
def test_grad_allreduce_matches_single_process(overlap):
    world = 2
    port = _free_port()
    with mp.Manager() as manager:
        ret = manager.dict()
        mp.spawn(_worker, args=(world, port, ret, overlap), nprocs=world, join=True)
        grads, params, metrics, partial = ret["grads"], ret["params"], ret["metrics"], ret["partial"]
    # single-process reference: full batch, rank-0 parameters
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 8, 12, generator=g, dtype=torch.float64)
    y = torch.randint(0, 10, (8,), generator=g)
    model = _model()
    for p, q in zip(model.parameters(), params):
        assert torch.equal(p.detach(), q), "rank 0's parameters must have been broadcast unchanged"
    F.cross_entropy(model(x[0]), y).backward()
    for p, gr in zip(model.parameters(), grads):
        assert torch.allclose(p.grad, gr, rtol=1e-12, atol=1e-14)
    assert metrics == ((4 + 8) / 8.0, (0 + 2) / 8.0)
    # rank 1 contributed a gradient for the last bias only: every other averaged gradient is half of rank 0's
    model.zero_grad(set_to_none=True)
    F.cross_entropy(model(x[0, :4]), y[:4]).backward()
    for i, (p, gr) in enumerate(zip(model.parameters(), partial)):
        if i < 3:
            assert torch.allclose(p.grad / 2, gr, rtol=1e-12, atol=1e-14)


def test_shard_batch_requires_divisibility():
    import pytest

    with pytest.raises(AssertionError):
        shard_batch(torch.zeros(1, 7, 3), 0, 2)
