"""Data-parallel training of EPSesPlusLinear over the GPUs of one box.

The reference has no distributed code (multi-GPU = independent replica processes,
training_configurations/get_adequate_results_with_cifar10_rgb/lr_gridsearch.py:68-76).  The path shards over
the batch axis (dim 1 of the (C, B, H, W, Q) input): patches of different images are independent in the
forward and in the input gradient; only the parameter gradients (cores + linear, 7.5 MB at config 2) are sums
over the batch.  So: one process per GPU, parameters replicated, each rank runs the unchanged single-GPU
kernels on its B/world images, and the only collective is the average of the parameter gradients
(``F.cross_entropy`` averages over the local batch, dctn/training.py:78).  No activation crosses GPUs.

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


def _is_nccl(group) -> bool:
    try:
        return dist.get_backend(group) == "nccl"
    except Exception:
        return False


class GradAllReducer:
    """Averages parameter gradients across ranks.  Call :meth:`zero_grad` before ``backward()`` and :meth:`wait` after
    it, before ``optimizer.step()``.

    Layout: ONE flat bucket holds every gradient, in a FIXED layout over all parameters that require a gradient (the same
    on every rank, whatever subset of gradients a step produces), and every ``p.grad`` is a VIEW into it — autograd
    accumulates straight into the bucket, so the collective needs no gather / scatter copies.  A parameter that received
    no gradient in a step contributes zeros (its view was zeroed by :meth:`zero_grad`).

    ``overlap=False`` (default): one all-reduce (average) over the whole bucket when the backward pass is complete.
    ``overlap=True``: the bucket is cut in two at ``params[:early]`` / the rest, in the ORDER GRADIENTS BECOME READY
    (pass the parameters last-layer first: linear, last EPS core, ..., first EPS core).  The early part — everything but
    the first layer's core — is all-reduced asynchronously from the gradient hook of its last parameter while the first
    layer's core gradient is still being computed; :meth:`wait` reduces the remainder and joins.  Use it with a
    CTA-limited NCCL communicator (:func:`make_comm_group`): the EPS kernels occupy every SM with one CTA, so the
    collective's CTAs take over SMs as EPS CTAs retire, and a handful of them costs a few percent of the SMs for the
    ~50 us of the transfer instead of stalling whole waves."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                 broadcast_from: int = 0, overlap: bool = False, early: Optional[int] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        assert self.params, "no parameter requires a gradient"
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = overlap and self.world > 1
        self._handles = []
        self._work = []
        self._flat = None
        if self.world == 1:      # nothing to reduce: gradients stay ordinary tensors
            return
        first = self.params[0]
        assert all(p.dtype == first.dtype and p.device == first.device for p in self.params), \
            "all parameters must share dtype and device (one flat bucket)"
        self._sizes = [p.numel() for p in self.params]
        # 128-byte aligned slots so that every view satisfies the kernels' 16-byte alignment
        es = first.element_size()
        align = max(1, 128 // es)
        self._offsets, off = [], 0
        for n in self._sizes:
            self._offsets.append(off)
            off += (n + align - 1) // align * align
        self._flat = torch.zeros(off, dtype=first.dtype, device=first.device)
        self._views = [self._flat[o:o + n].view_as(p) for o, n, p in zip(self._offsets, self._sizes, self.params)]
        self._early = len(self.params) - 1 if early is None else early
        self._early = max(0, min(self._early, len(self.params)))
        self._cut = self._offsets[self._early] if self._early < len(self.params) else off
        self._avg = _is_nccl(group) if self.world > 1 else False
        if self.world > 1:
            for p in self.params:  # replicas start identical
                dist.broadcast(p.data, src=broadcast_from, group=group)
            if self.overlap and self._early > 0:
                trigger = self.params[self._early - 1]
                self._handles.append(trigger.register_post_accumulate_grad_hook(self._on_early_ready))
        self.zero_grad()

    # -- gradients live in the bucket
    def zero_grad(self) -> None:
        """Zeroes the bucket (one memset) and (re)attaches every ``p.grad`` to its view.  Replaces
        ``optimizer.zero_grad()``, whose default ``set_to_none=True`` would detach the gradients from the bucket."""
        if self.world == 1:
            for p in self.params:
                p.grad = None
            return
        self._flat.zero_()
        for p, v in zip(self.params, self._views):
            if p.grad is not v:
                p.grad = v

    def _attached(self) -> bool:
        return all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(self.params, self._views))

    def _reattach(self) -> None:
        """A caller used optimizer.zero_grad(set_to_none=True) (or autograd replaced a gradient tensor): copy what exists
        into the fixed layout — missing gradients stay zero — so every rank still reduces the same buffer."""
        self._flat.zero_()
        for p, v in zip(self.params, self._views):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            p.grad = v

    def _reduce(self, t: Tensor, async_op: bool):
        if self._avg:
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        t.div_(self.world)   # SUM of pre-scaled gradients = mean over ranks (gloo has no AVG)
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def _on_early_ready(self, _p: torch.nn.Parameter) -> None:
        # hooks fire in backward order: everything in params[:early] is accumulated when the last of them is
        if self._attached():
            self._work.append(self._reduce(self._flat[: self._cut], async_op=True))

    def wait(self) -> None:
        if self.world == 1:
            return
        if self.overlap:
            # ALWAYS the same two collectives in the same order on every rank ([:cut], then [cut:]), whether or not this
            # rank's hook fired (a rank whose trigger parameter got no gradient this step must not fall out of step)
            if not self._work:
                if not self._attached():
                    self._reattach()
                if self._cut > 0:
                    self._work.append(self._reduce(self._flat[: self._cut], async_op=True))
            if self._cut < self._flat.numel():
                self._work.append(self._reduce(self._flat[self._cut:], async_op=True))
            for w in self._work:
                w.wait()
            self._work.clear()
            return
        if not self._attached():
            self._reattach()
        self._reduce(self._flat, async_op=False)

    @property
    def flat_grad(self) -> Tensor:
        return self._flat

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()


def make_comm_group(max_ctas: int = 4):
    """A second NCCL communicator over all ranks whose kernels use at most ``max_ctas`` CTAs (ncclConfig_t.maxCTAs), for
    collectives that run WHILE the EPS kernels hold the SMs (GradAllReducer(overlap=True)).  Returns None when the
    backend is not NCCL."""
    if not dist.is_initialized() or dist.get_backend() != "nccl":
        return None
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.max_ctas = max_ctas
    opts.config.min_ctas = 1
    return dist.new_group(backend="nccl", pg_options=opts)


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
