"""Static contraction-plan cache — API of dctn/contraction_path_cache.py:15-35.

The reference caches ``opt_einsum.contract_expression(..., optimize="auto-hq")`` objects keyed by the
operand shapes and subscripts.  The EPS hot path no longer goes through it (its plan is the per-shape
kernel plan inside libdctn_b200.so, see ``dctn_eps_plan_get``); what remains are the parameter-only
contractions (regulariser), for which a plan is a normalised ``torch.einsum`` sublist program.
All three call forms of the reference are accepted:
``contract("ij,jk->ik", a, b)``, ``contract(a, "ij", b, "jk", "ik")``, ``contract(a, (0, 1), b, (1, 2), (0, 2))``.
"""
from typing import Callable, Dict, Hashable, List, Tuple, Union

import torch
from torch import Tensor

from .singleton import Singleton

ContractArgs = Tuple[Union[Hashable, Tensor], ...]
ContractExpressionArgs = Tuple[Hashable, ...]


def tensors_to_shapes(*args) -> ContractExpressionArgs:
    return tuple(tuple(x.shape) if isinstance(x, Tensor) else (tuple(x) if isinstance(x, list) else x) for x in args)


def _compile(key: ContractExpressionArgs) -> Callable[..., Tensor]:
    """Turns the (shapes-in-place-of-tensors) argument tuple into an executable plan."""
    if isinstance(key[0], str):  # string form
        equation = key[0].replace(" ", "")
        lhs, rhs = equation.split("->") if "->" in equation else (equation, None)
        terms = [tuple(t) for t in lhs.split(",")]
        if rhs is None:
            flat = [s for t in terms for s in t]
            rhs = tuple(sorted(s for s in set(flat) if flat.count(s) == 1))
        out = tuple(rhs)
    else:  # interleaved form: shape, names, shape, names, ..., out names
        terms = [tuple(key[i]) for i in range(1, len(key) - 1, 2)]
        out = tuple(key[-1])
    symbols: List[Hashable] = []
    for t in terms + [out]:
        for s in t:
            if s not in symbols:
                symbols.append(s)
    assert len(symbols) <= 52, "torch.einsum supports at most 52 distinct indices"
    idx = {s: i for i, s in enumerate(symbols)}
    term_ids = [[idx[s] for s in t] for t in terms]
    out_ids = [idx[s] for s in out]

    def run(*operands: Tensor) -> Tensor:
        assert len(operands) == len(term_ids)
        flat = []
        for o, t in zip(operands, term_ids):
            flat += [o, t]
        return torch.einsum(*flat, out_ids)

    return run


class ContractionPathCache(metaclass=Singleton):
    def __init__(self):
        self.paths: Dict[ContractExpressionArgs, Callable[..., Tensor]] = {}

    def contract(self, *args) -> Tensor:
        key = tensors_to_shapes(*args)
        plan = self.paths.get(key)
        if plan is None:
            plan = self.paths[key] = _compile(key)
        return plan(*(x for x in args if isinstance(x, Tensor)))


def contract(*args) -> Tensor:
    return ContractionPathCache().contract(*args)
