"""EPSesPlusLinear — API and parameter layout of dctn/eps_plus_linear.py:52-159.

Parameters are ``epses.{i}`` (cores), ``linear.weight``, ``linear.bias`` and the buffer ``p`` (KEEP
probability of the core dropout), so reference ``state_dict`` files load unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from logging import getLogger
from typing import Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import eps, epses_composition
from .align import make_windows
from .eps import EPS  # noqa: F401  (re-exported like the reference)
from .utils import (  # noqa: F401
    FromFileInitialization,
    OneTensorInitialization,
    ZeroCenteredNormalInitialization,
    ZeroCenteredUniformInitialization,
)


@dataclass(frozen=True)
class UnitEmpiricalOutputStd:
    input: Tensor
    batch_size: int = 128


class UnitTheoreticalOutputStd:
    pass


@dataclass(frozen=True)
class ManuallyChosenInitialization:
    epses: Tuple[OneTensorInitialization, ...]
    linear_weight: OneTensorInitialization
    linear_bias: OneTensorInitialization


Initialization = Union[UnitEmpiricalOutputStd, UnitTheoreticalOutputStd, ManuallyChosenInitialization]


def _fill(param: Tensor, init: OneTensorInitialization) -> None:
    if isinstance(init, ZeroCenteredNormalInitialization):
        param.data.copy_(torch.randn_like(param) * init.std)
    elif isinstance(init, ZeroCenteredUniformInitialization):
        param.data.copy_(torch.rand_like(param) * (2 * init.maximum) - init.maximum)
    else:
        raise ValueError(f"linear layer initialization {init!r} must be normal or uniform")


class EPSesPlusLinear(nn.Module):
    def __init__(
        self,
        epses_specs: Tuple[Tuple[int, int], ...],
        initialization: Initialization,
        p: float,
        device: torch.device,
        dtype: torch.dtype,
        image_size: int = 28,
        Q_0: int = 2,
    ):
        """`p` is the probability of NOT dropping a component of a core (dctn/eps_plus_linear.py:53-63)."""
        assert 0.0 < p <= 1
        super().__init__()
        if isinstance(initialization, UnitEmpiricalOutputStd):
            assert initialization.input.shape[2] == image_size
            assert initialization.input.shape[3] == image_size
            cores = epses_composition.make_epses_composition_unit_empirical_output_std(
                epses_specs, initialization.input, device, dtype, initialization.batch_size
            )
        elif isinstance(initialization, UnitTheoreticalOutputStd):
            cores = epses_composition.make_epses_composition_unit_theoretical_output_std(epses_specs, Q_0, device, dtype)
        elif isinstance(initialization, ManuallyChosenInitialization):
            cores = epses_composition.make_epses_composition_manually_chosen_inializations(
                epses_specs, initialization.epses, Q_0, device, dtype
            )
        else:
            raise ValueError(f"{initialization=} is not {Initialization}")
        self.epses = nn.ParameterList(nn.Parameter(core) for core in cores)

        # every "valid" K x K layer shrinks the image by K-1
        side = image_size - sum(k - 1 for k, _ in epses_specs)
        self.linear = nn.Linear(side * side * eps.matrix_shape(self.epses[-1])[0], 10, bias=True).to(dtype)
        if isinstance(initialization, ManuallyChosenInitialization):
            _fill(self.linear.weight, initialization.linear_weight)
            _fill(self.linear.bias, initialization.linear_bias)
        else:
            logger = getLogger(f"{__name__}.EPSesPlusLinear.__init__")
            weight_std = self.linear.in_features ** -0.5 / 4.0
            _fill(self.linear.weight, ZeroCenteredNormalInitialization(weight_std))
            logger.info(f"Initialized linear.weight as randn * {weight_std:.30e}")
            bias_max = self.linear.in_features ** -0.5
            _fill(self.linear.bias, ZeroCenteredUniformInitialization(bias_max))
            logger.info(f"Initialized linear.bias from Uniform[{-bias_max:.30e}, {bias_max:.30e}]")
        self.linear.to(device)
        self.register_buffer("p", torch.tensor(p, device=device, dtype=dtype))
        self._p_host = float(p)

    def _keep_prob(self) -> float:
        """Host copy of the buffer ``p``.  The reference tests ``self.p < 1.0`` on the device buffer every forward
        (dctn/eps_plus_linear.py:139), a device-to-host read per step — and an illegal synchronisation while a CUDA graph
        is being captured (dctn_b200.train_step).  Read once, refreshed when a state_dict is loaded."""
        if self._p_host is None:
            self._p_host = float(self.p)
        return self._p_host

    def _load_from_state_dict(self, *args, **kwargs):
        self._p_host = None
        return super()._load_from_state_dict(*args, **kwargs)

    def forward(self, input: Tensor) -> Tensor:
        """Core dropout (train mode and p < 1): every core component is kept with probability p and the
        survivors are divided by p — ONE mask per step for the whole batch (dctn/eps_plus_linear.py:138-147);
        data-parallel ranks must draw it from identically seeded generators (dctn_b200.parallel)."""
        if self.training and self._keep_prob() < 1.0:
            cores = tuple(torch.bernoulli(self.p.expand_as(core)) * core / self.p for core in self.epses)
        else:
            cores = tuple(self.epses)
        inter = epses_composition.contract_with_input(cores, input)
        return self.linear(inter.flatten(start_dim=1))  # "b h w q -> b (h w q)"

    def epswise_l2_regularizer(self) -> Tensor:
        """Squared Frobenius norms of all cores and of linear.weight (bias excluded)."""
        return self.linear.weight.norm(p="fro") ** 2 + epses_composition.epswise_squared_fro_norm(self.epses)

    def epses_composition_l2_regularizer(self) -> Tensor:
        return self.linear.weight.norm(p="fro") ** 2 + epses_composition.inner_product(self.epses, self.epses)

    @torch.no_grad()
    def log_intermediate_reps_stats(self, x: Tensor, batch_size: int = 128) -> None:
        """Logs mean/std of every intermediate representation and of its K x K windows in eval mode
        (dctn/eps_plus_linear.py:161-196; same log lines)."""
        logger = getLogger(f"{__name__}.EPSesPlusLinear.log_intermediate_reps_stats")
        logger.info("Logging intermediate reps stats as if self.training == False")

        def log_one(t: Tensor, name: str) -> None:
            mu, sigma = t.mean(), t.std(unbiased=False)
            logger.info(f"{name}: μ={mu:.7e}, σ={sigma:.7e}, μ**2+σ**2={mu**2+sigma**2:.7e}, shape={tuple(t.shape)}")

        def log_windows(windows, name: str) -> None:
            mu, sigma = windows.mean_over_batch(), windows.std_over_batch(unbiased=False)
            logger.info(
                f"{name}: μ={mu:.7e}, σ={sigma:.7e}, μ**2+σ**2={mu**2+sigma**2:.7e}, "
                f"batch_shape={windows.batch_shape}, "
                f"num_factors={windows.num_factors}, "
                f"num_coordinates_in_one_factor={windows.num_coordinates_in_one_factor}"
            )

        for n, core in enumerate(self.epses):
            log_one(x, f"x_{n}")
            kernel_size = math.isqrt(core.ndim - 1)
            assert kernel_size ** 2 == core.ndim - 1
            log_windows(make_windows(x, kernel_size), f"w_{n}")
            x = eps.transform_in_slices(core, x, batch_size)
        x = x.squeeze(0).flatten(start_dim=1)
        log_one(x, f"x_{len(self.epses)}")
        log_one(F.linear(x, self.linear.weight), "output_of_linear_without_bias")
        log_one(self.linear(x), "output_of_linear_with_bias")
