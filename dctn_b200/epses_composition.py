"""Stacked EPS layers — API of dctn/epses_composition.py.

``contract_with_input`` is the hot path (one fused CUDA forward per layer, inter-layer activation stays
(B, H', W', Q) in HBM: 5 MB at config 2, so layers are not fused with each other).  ``inner_product`` is the
composition-L2 regulariser: parameters only, batch independent, evaluated with cuBLAS-backed torch ops.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch
from torch import Tensor

from . import eps
from .utils import (
    FromFileInitialization,
    OneTensorInitialization,
    ZeroCenteredNormalInitialization,
    ZeroCenteredUniformInitialization,
    id_assert_shape_matches,
)


def inner_product(epses1: Sequence[Tensor], epses2: Sequence[Tensor]) -> Tensor:
    """<composition(epses1), composition(epses2)> without ever forming the composed tensors
    (dctn/epses_composition.py:21-58): the Gram matrix of the two first cores over their input dims is
    pushed through every input mode of the next core of epses1, then recurse."""
    epses1, epses2 = tuple(epses1), tuple(epses2)
    assert len(epses1) == len(epses2)
    for e1, e2 in zip(epses1, epses2):
        assert e1.shape == e2.shape
        assert eps.is_eps(e1)
    if len(epses1) == 1:
        return eps.inner_product(epses1[0], epses2[0])
    gram = eps.contract_on_input_dims(epses1[0], epses2[0])  # (out of a, out of k)
    nxt = epses1[1]
    # n-fold mode product: every input mode of `nxt` (size out_a) is mapped through gram to size out_k.
    # Each tensordot consumes the current leading mode and appends the mapped one, so after all input
    # modes are done the output mode sits in front.
    for _ in range(nxt.ndim - 1):
        nxt = torch.tensordot(nxt, gram, dims=([0], [0]))
    nxt = nxt.movedim(0, -1)
    assert eps.is_eps(nxt)
    return inner_product((nxt,) + epses1[2:], epses2[1:])


def specs_to_full_specs(epses_specs: Tuple[Tuple[int, int], ...], initial_in_size: int) -> Tuple[Dict[str, int], ...]:
    """(kernel_size, out_size) pairs -> full layer specs; every layer of a composition has one input
    channel and its in_size is the previous layer's out_size (dctn/epses_composition.py:61-76)."""
    full = []
    in_size = initial_in_size
    for kernel_size, out_size in epses_specs:
        full.append({"kernel_size": kernel_size, "in_num_channels": 1, "in_size": in_size, "out_size": out_size})
        in_size = out_size
    return tuple(full)


def make_epses_composition_unit_theoretical_output_std(
    epses_specs: Tuple[Tuple[int, int], ...], initial_in_size: int, device: torch.device, dtype: torch.dtype
) -> Tuple[Tensor, ...]:
    return tuple(
        eps.make_eps_unit_theoretical_output_std(**spec, device=device, dtype=dtype)
        for spec in specs_to_full_specs(epses_specs, initial_in_size)
    )


def make_epses_composition_unit_empirical_output_std(
    epses_specs: Tuple[Tuple[int, int], ...], input: Tensor, device: torch.device, dtype: torch.dtype, batch_size: int = 128
) -> Tuple[Tensor, ...]:
    """Layer-by-layer empirical-std initialisation: each new core is rescaled on the current representation
    of `input`, which is then pushed through it (dctn/epses_composition.py:91-105).  Forward-only consumer
    of the CUDA forward kernel."""
    cores = []
    for kernel_size, out_size in epses_specs:
        core, input = eps._empirical_std_core_and_output(kernel_size, out_size, input, device, dtype, batch_size)
        cores.append(core)
    return tuple(cores)


def _init_one(spec: Dict[str, int], init: OneTensorInitialization, device, dtype) -> Tensor:
    shape = eps.spec_to_shape(**spec)
    if isinstance(init, ZeroCenteredNormalInitialization):
        return torch.randn(shape, dtype=dtype).to(device) * init.std
    if isinstance(init, ZeroCenteredUniformInitialization):
        return torch.rand(shape, dtype=dtype).to(device) * (2 * init.maximum) - init.maximum
    if isinstance(init, FromFileInitialization):
        return id_assert_shape_matches(torch.load(init.path, device).to(dtype=dtype), shape)
    raise ValueError(f"unknown initialization {init!r}")


def make_epses_composition_manually_chosen_inializations(
    epses_specs: Tuple[Tuple[int, int], ...],
    initializations: Tuple[OneTensorInitialization, ...],
    initial_in_size: int,
    device: torch.device,
    dtype: torch.dtype,
) -> Tuple[Tensor, ...]:
    """(sic — the reference spells it 'inializations', dctn/epses_composition.py:108-130)."""
    assert len(epses_specs) == len(initializations)
    full = specs_to_full_specs(epses_specs, initial_in_size)
    return tuple(_init_one(spec, init, device, dtype) for spec, init in zip(full, initializations))


def contract_with_input(epses: Sequence[Tensor], input: Tensor) -> Tensor:
    """`input`: (channels, batch, height, width, quantum_in) -> (batch, new_height, new_width, quantum_out)
    (dctn/epses_composition.py:133-141).  Layers >= 2 see their input as one channel."""
    assert all(eps.is_eps(core) for core in epses)
    inter = input
    for core in epses[:-1]:
        inter = eps.eps(core, inter).unsqueeze(0)
    return eps.eps(epses[-1], inter)


def epswise_squared_fro_norm(epses: Sequence[Tensor]) -> Tensor:
    assert all(eps.is_eps(core) for core in epses)
    return sum(core.norm(p="fro") ** 2 for core in epses)
