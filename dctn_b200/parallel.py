"""Data-parallel training of EPSesPlusLinear over the GPUs of one box.

The reference has no distributed code (multi-GPU = independent replica processes,
training_configurations/get_adequate_results_with_cifar10_rgb/lr_gridsearch.py:68-76).  The path shards over
the batch axis (dim 1 of the (C, B, H, W, Q) input): patches of different images are independent in the
forward and in the input gradient; only the parameter gradients (cores + linear, 7.5 MB at config 2) are sums
over the batch.  So: one process per GPU, parameters replicated, each rank runs the unchanged single-GPU
kernels on its B/world images, and ONE collective per step — an NCCL all-reduce over a flat bucket of all
parameter gradients, issued after the backward pass — averages the gradients (``F.cross_entropy`` averages over
the local batch, dctn/training.py:78).  No activation crosses GPUs.

Works with any torch.distributed backend (``nccl`` on GPUs, ``gloo`` in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor


def shard_batch(x: Tensor, rank: int, world_size: int, dim: int = 1) -> Tensor:
    """This rank's contiguous slice of the batch axis (dim 1 for (C, B, H, W, Q) inputs, dim 0 for labels).
    The batch must divide evenly so that the mean-of-means equals the global mean."""
    n = x.shape[dim]
    assert n % world_size == 0, f"global batch {n} not divisible by world size {world_size}"
    per = n // world_size
    return x.narrow(dim, rank * per, per)


class GradAllReducer:
    """Averages parameter gradients across ranks.  Call :meth:`wait` after ``backward()`` and before
    ``optimizer.step()``.

    Default (``overlap=False``): ONE all-reduce over a flat bucket holding every gradient (7.5 MB at config 2, ~20 us on
    NVLink 5) issued when the backward pass is complete.  The EPS kernels are persistent-style — one CTA per SM, all of the
    SM's shared memory and tensor memory — so a collective launched *during* backward cannot co-reside with them: its
    CTAs wait for SMs, then hold them while spinning on the peers, and every following one-wave kernel launch turns
    into two waves.  Measured at N=2: 27.2 ms/step with per-parameter all-reduces overlapped from the gradient hooks
    against 21.0 ms at N=1; the flat bucket after backward removes that.
    ``overlap=True`` keeps the hook-driven variant (one async all-reduce per parameter as soon as its gradient is
    accumulated) for models whose kernels leave SMs free."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                 broadcast_from: int = 0, overlap: bool = False):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = overlap
        self._pending = []
        self._handles = []
        self._flat: Optional[Tensor] = None
        if self.world > 1:
            for p in self.params:  # replicas start identical
                dist.broadcast(p.data, src=broadcast_from, group=group)
            if overlap:
                for p in self.params:
                    self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad_ready))

    def _on_grad_ready(self, p: torch.nn.Parameter) -> None:
        # pre-scale so that the SUM all-reduce yields the mean (cross_entropy averages over the local batch)
        p.grad.div_(self.world)
        work = dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending.append(work)

    def wait(self) -> None:
        if self.world == 1:
            return
        if self.overlap:
            for work in self._pending:
                work.wait()
            self._pending.clear()
            return
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return
        n = sum(g.numel() for g in grads)
        if self._flat is None or self._flat.numel() != n or self._flat.device != grads[0].device or self._flat.dtype != grads[0].dtype:
            self._flat = torch.empty(n, dtype=grads[0].dtype, device=grads[0].device)
        views = []
        off = 0
        for g in grads:
            views.append(self._flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(views, grads)
        self._flat.div_(self.world)   # SUM of pre-scaled gradients = mean over ranks
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
        torch._foreach_copy_(grads, views)

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()


def seed_core_dropout(seed: int, step: int, device: torch.device) -> None:
    """Core dropout draws ONE Bernoulli mask per step over the cores (dctn/eps_plus_linear.py:139-143), not per
    sample, so every rank must draw the same mask to stay equivalent to the single-GPU model: re-seed the
    device generator identically on all ranks before each forward."""
    gen_seed = (seed * 1000003 + step) % (2 ** 63)
    if device.type == "cuda":
        with torch.cuda.device(device):
            torch.cuda.manual_seed(gen_seed)
    else:
        torch.manual_seed(gen_seed)


def all_reduce_metrics(sum_loss: float, num_correct: float, num_samples: float, device: torch.device):
    """Evaluation (dctn/evaluation.py:7-22) shards the same way: a final all-reduce of the three sums."""
    t = torch.tensor([sum_loss, num_correct, num_samples], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t[0].item() / t[2].item(), t[1].item() / t[2].item()
