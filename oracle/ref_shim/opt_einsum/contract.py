"""Implementation module of the opt_einsum shim (see package docstring)."""
from __future__ import annotations

from typing import Any, Hashable, List, Sequence, Tuple

import torch

__version__ = "0.0-shim"


def _parse(args: Tuple[Any, ...]):
    """Returns (operands, list of name-tuples, output name-tuple)."""
    if isinstance(args[0], str):
        subs = args[0].replace(" ", "")
        operands = list(args[1:])
        if "->" in subs:
            lhs, rhs = subs.split("->")
        else:
            lhs, rhs = subs, None
        terms = [tuple(t) for t in lhs.split(",")]
        if rhs is None:
            counts = {}
            for t in terms:
                for s in t:
                    counts[s] = counts.get(s, 0) + 1
            rhs = tuple(sorted(s for s, c in counts.items() if c == 1))
        else:
            rhs = tuple(rhs)
        assert len(terms) == len(operands)
        return operands, terms, rhs
    # interleaved
    rest = list(args)
    operands, terms = [], []
    while len(rest) >= 2:
        operands.append(rest.pop(0))
        terms.append(tuple(rest.pop(0)))
    assert len(rest) == 1, "interleaved form needs an explicit output sublist"
    return operands, terms, tuple(rest[0])


def _pairwise(ops: List[torch.Tensor], terms: List[Tuple[Hashable, ...]], keep: set):
    """einsum of the given operands, keeping the symbols in `keep` (order of first appearance)."""
    symbols: List[Hashable] = []
    for t in terms:
        for s in t:
            if s not in symbols:
                symbols.append(s)
    out = tuple(s for s in symbols if s in keep)
    idx = {s: i for i, s in enumerate(symbols)}
    assert len(symbols) <= 52
    flat: List[Any] = []
    for o, t in zip(ops, terms):
        flat += [o, [idx[s] for s in t]]
    flat.append([idx[s] for s in out])
    return torch.einsum(*flat), out


def _run(operands, terms, out, optimize):
    ops = list(operands)
    terms = [tuple(t) for t in terms]
    if isinstance(optimize, (list, tuple)) and len(optimize) > 0 and not isinstance(optimize, str):
        path = [tuple(p) for p in optimize]
    else:
        # any valid order gives the same value; fold from the left
        path = [(0, 1)] * (len(ops) - 1) if len(ops) > 1 else [(0,)]
    for step in path:
        picked = sorted(step, reverse=True)
        sel_ops = [ops[i] for i in sorted(step)]
        sel_terms = [terms[i] for i in sorted(step)]
        for i in picked:
            ops.pop(i)
            terms.pop(i)
        keep = set(out)
        for t in terms:
            keep |= set(t)
        res, res_term = _pairwise(sel_ops, sel_terms, keep)
        ops.append(res)
        terms.append(res_term)
    assert len(ops) == 1
    res, res_term = ops[0], tuple(terms[0])
    out = tuple(out)
    if res_term != out:
        if set(res_term) != set(out):  # leftover symbols to sum away
            res, res_term = _pairwise([res], [res_term], set(out))
        res = res.permute([res_term.index(s) for s in out])
    return res


def contract(*args, optimize="auto", **kwargs):
    operands, terms, out = _parse(args)
    return _run(operands, terms, out, optimize)


class ContractExpression:
    def __init__(self, terms, out, optimize):
        self.terms, self.out, self.optimize = terms, out, optimize

    def __call__(self, *operands, **kwargs):
        return _run(list(operands), self.terms, self.out, self.optimize)


def contract_expression(*args, optimize="auto", **kwargs):
    shapes_or_ops, terms, out = _parse(args)
    return ContractExpression(terms, out, optimize)


def contract_path(*args, **kwargs):  # only so torch's optional opt_einsum hook cannot crash
    operands, terms, out = _parse(args)
    n = len(operands)
    return [(0, 1)] * (n - 1) if n > 1 else [(0,)], None
