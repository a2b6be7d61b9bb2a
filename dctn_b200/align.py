"""Patch alignment ("unfold") — API of dctn/align.py:11-46.

In the CUDA path these views never exist: the kernels do the same index arithmetic on the fly
(factor j = (dh*K + dw)*C + c reads x[c, b, h+dh, w+dw, :]).  The functions are kept because callers
of the reference use them directly (statistics, tests).
"""
from dataclasses import dataclass
from typing import Iterable, Sequence, Tuple, Union

import torch
from torch import Tensor

from . import _lib
from .pos2d import Pos2D


def align_with_positions(input: Union[Tensor, Sequence[Tensor]], positions: Tuple[Pos2D, ...]) -> Iterable[Tensor]:
    """For every position (outer loop) and channel (inner loop) yields the zero-copy view of that
    channel shifted by the position and cropped to the common valid size."""
    hs = [p.h for p in positions]
    ws = [p.w for p in positions]
    assert min(hs) == 0 and min(ws) == 0
    span_h, span_w = max(hs), max(ws)
    _, height, width, _ = input[0].shape
    out_h, out_w = height - span_h, width - span_w
    for p in positions:
        for channel in input:
            yield channel[:, p.h : p.h + out_h, p.w : p.w + out_w]


def align(input: Tensor, kernel_size: int) -> Iterable[Tensor]:
    """Window positions in row-major order (for kernel_size=3: 0 1 2 / 3 4 5 / 6 7 8)."""
    grid = tuple(Pos2D(i // kernel_size, i % kernel_size) for i in range(kernel_size * kernel_size))
    return align_with_positions(input, grid)


@dataclass(frozen=True)
class WindowsBatch:
    """All K x K windows of ``x`` (C, B, H, W, Q) seen as a batch of rank-one tensors with K*K*C factors of Q coordinates —
    what the reference's ``make_windows`` returns as a ``RankOneTensorsBatch`` (dctn/align.py:49-61,
    dctn/rank_one_tensor.py:14-110), with the same statistics API.  The reference materialises the (K*K*C, B, H', W', Q)
    stack of aligned views; here the windows are never formed: dctn_window_stats reduces two numbers per pixel and one
    product per window on the GPU (the rank-one identities: sum = product of factor sums, squared norm = product of
    factor squared norms)."""

    x: Tensor
    kernel_size: int

    @property
    def num_factors(self) -> int:
        return self.kernel_size ** 2 * self.x.shape[0]

    @property
    def num_coordinates_in_one_factor(self) -> int:
        return self.x.shape[4]

    @property
    def batch_shape(self) -> Tuple[int, ...]:
        _, B, H, W, _ = self.x.shape
        return (B, H - self.kernel_size + 1, W - self.kernel_size + 1)

    @property
    def ncoordinates(self) -> int:
        return self.num_coordinates_in_one_factor ** self.num_factors

    @property
    def ntensors(self) -> int:
        B, Ho, Wo = self.batch_shape
        return B * Ho * Wo

    def _stats(self) -> Tensor:
        """(sum of all elements of all windows, squared Frobenius norm of the whole batch) as a float64 pair."""
        cached = self.__dict__.get("_cache")
        if cached is None:
            x = self.x
            if not x.is_cuda:
                raise RuntimeError("dctn_b200.align.make_windows statistics run on CUDA tensors only (no CPU fallback)")
            dt = {torch.float32: _lib.F32, torch.float64: _lib.F64}[x.dtype]
            x = x.detach().contiguous()
            C, B, H, W, Q = x.shape
            lib = _lib.lib()
            cached = torch.zeros(2, dtype=torch.float64, device=x.device)
            with torch.cuda.device(x.device):
                ws = torch.empty(lib.dctn_window_stats_workspace_bytes(C, B, H, W), dtype=torch.uint8, device=x.device)
                rc = lib.dctn_window_stats(x.data_ptr(), C, B, H, W, Q, self.kernel_size, dt, cached.data_ptr(), ws.data_ptr(),
                                           ws.numel(), torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "dctn_window_stats")
            object.__setattr__(self, "_cache", cached)
        return cached

    def sum_over_batch(self) -> Tensor:
        return self._stats()[0]

    def squared_fro_norm_over_batch(self) -> Tensor:
        return self._stats()[1]

    def mean_over_batch(self) -> Tensor:
        return self.sum_over_batch() / (self.ntensors * self.ncoordinates)

    def var_over_batch(self, unbiased: bool = True) -> Tensor:
        """Same expression as dctn/rank_one_tensor.py:93-106."""
        total, mean = self.sum_over_batch(), self.mean_over_batch()
        nelement = self.ntensors * self.ncoordinates
        divisor = nelement - 1 if unbiased else nelement
        return self.squared_fro_norm_over_batch() / divisor - 2 * total / divisor * mean + nelement / divisor * mean ** 2

    def std_over_batch(self, unbiased: bool = True) -> Tensor:
        # the reference ignores `unbiased` here (dctn/rank_one_tensor.py:108-110 calls var_over_batch()): kept
        return self.var_over_batch() ** 0.5


def make_windows(x: Tensor, kernel_size: int) -> WindowsBatch:
    """`x`: (num_channels, batch_size, height, width, in_quantum_size) — dctn/align.py:49-61."""
    assert x.ndim == 5 and x.shape[2] >= kernel_size and x.shape[3] >= kernel_size
    return WindowsBatch(x, kernel_size)
