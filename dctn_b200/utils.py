"""Initialisation descriptors and small helpers (API of dctn/utils.py:10-59)."""
from dataclasses import dataclass
from typing import Callable, Sequence, Union

import torch
from torch import Tensor


@torch.no_grad()
def transform_dataset(f: Callable[[Tensor], Tensor], x: Tensor, batch_size: int = 64) -> Tensor:
    """Applies an eps-like `f` to x (channel, sample, h, w, quantum) in slices along the sample axis and
    returns (1, sample, h', w', quantum')."""
    pieces = [f(piece) for piece in torch.split(x, batch_size, dim=1)]
    return torch.cat(pieces).unsqueeze(0)


def implies(x: bool, y: bool) -> bool:
    return (not x) or y


def xor(*args: bool) -> bool:
    acc = False
    for a in args:
        acc = acc != bool(a)
    return acc


def exactly_one_true(*args: bool) -> bool:
    assert all(isinstance(a, bool) for a in args)
    return sum(args) == 1


@dataclass(frozen=True)
class ZeroCenteredNormalInitialization:
    std: float


@dataclass(frozen=True)
class ZeroCenteredUniformInitialization:
    maximum: float


@dataclass(frozen=True)
class FromFileInitialization:
    path: str


OneTensorInitialization = Union[
    ZeroCenteredNormalInitialization, ZeroCenteredUniformInitialization, FromFileInitialization
]


def raise_exception(exception):
    raise exception


def id_assert_shape_matches(tensor: Tensor, shape: Sequence[int]) -> Tensor:
    assert tensor.shape == tuple(shape)
    return tensor
