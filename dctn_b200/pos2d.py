"""2-D positions inside a K x K window (API of dctn/pos2d.py:4-23)."""
from dataclasses import dataclass


@dataclass(frozen=True)
class Pos2D:
    h: int
    w: int


def pos_to_index(max_w: int, pos: Pos2D) -> int:
    """Row-major index of `pos` when the width coordinate runs over [0, max_w]."""
    assert pos.w <= max_w
    return pos.h * (max_w + 1) + pos.w


def index_to_pos(max_w: int, index: int) -> Pos2D:
    """Inverse of pos_to_index for the same max_w."""
    h, w = divmod(index, max_w + 1)
    return Pos2D(h, w)
