"""Patch alignment ("unfold") — API of dctn/align.py:11-46.

In the CUDA path these views never exist: the kernels do the same index arithmetic on the fly
(factor j = (dh*K + dw)*C + c reads x[c, b, h+dh, w+dw, :]).  The functions are kept because callers
of the reference use them directly (statistics, tests).
"""
from typing import Iterable, Sequence, Tuple, Union

from torch import Tensor

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
