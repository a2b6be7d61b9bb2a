"""One training step of the hot path — forward, loss, backward, gradient average, optimiser update
(dctn/training.py:77-83) — eager, or captured ONCE in a CUDA graph and replayed.

Why a graph: the step of a small model is a chain of ~35 short kernels (config 1: 8 us + 15 us of EPS kernels, the rest
is the linear layer, the loss, Adam) and the host needs ~0.6 ms to issue them; the GPU idles between launches.  A
captured step is one launch.  Everything the step allocates (outputs, workspaces, the saved intermediate, gradients)
comes from the graph's private pool, so replays reuse the same addresses and the per-call ``torch.empty`` disappears.
The EPS library launches on the stream it is handed and never synchronises, allocates or reads back, which is what makes
it capturable (include/dctn_b200.h, "Conventions").
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from .parallel import GradAllReducer


class TrainStep:
    """``step(x, y) -> loss`` (a device tensor; reading it is the caller's choice).

    ``graph=True`` captures the step for the shapes of ``example_x`` / ``example_y``: inputs are copied into static
    buffers (asynchronously, from pinned or device memory) and the captured graph is replayed; the returned loss is the
    graph's static output, overwritten by the next call.  The optimiser must be capturable (``capturable=True`` for
    torch.optim.Adam) — it steps on the device, no host read.  ``reducer`` averages the gradients over the data-parallel
    ranks inside the step (inside the graph too: NCCL collectives are capturable)."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, example_x: Tensor, example_y: Tensor,
                 loss_fn: Callable[[Tensor, Tensor], Tensor] = F.cross_entropy, reducer: Optional[GradAllReducer] = None,
                 graph: bool = False, warmup: int = 3):
        self.model, self.optimizer, self.loss_fn, self.reducer = model, optimizer, loss_fn, reducer
        self.graph = None
        if not graph:
            return
        if not example_x.is_cuda:
            raise RuntimeError("TrainStep(graph=True) needs CUDA example inputs")
        self._x = example_x.detach().clone()
        self._y = example_y.detach().clone()
        side = torch.cuda.Stream(device=example_x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # warm-up off the default stream, as graph capture requires
            for _ in range(max(1, warmup)):
                self._eager(self._x, self._y)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._loss = self._eager(self._x, self._y)

    def _eager(self, x: Tensor, y: Tensor) -> Tensor:
        if self.reducer is not None:
            self.reducer.zero_grad()
        else:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model(x), y)
        loss.backward()
        if self.reducer is not None:
            self.reducer.wait()
        self.optimizer.step()
        return loss

    def __call__(self, x: Tensor, y: Tensor) -> Tensor:
        if self.graph is None:
            return self._eager(x, y)
        self._x.copy_(x, non_blocking=True)
        self._y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self._loss
