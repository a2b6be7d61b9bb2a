"""CPU: host-side logic of the package — the C-ABI library loads and exports every declared symbol, plan
queries and argument validation work without a GPU, and the Python API keeps the reference's contracts
(shape asserts, parameter names, init statistics, regulariser known answers)."""
import ctypes
import os
import re
from itertools import product

import pytest
import torch

import dctn_b200
from dctn_b200 import _lib
from dctn_b200 import contraction_path_cache as cpc
from dctn_b200 import eps as eps_mod
from dctn_b200.eps import contract_on_input_dims, eps, eps_one_by_one
from dctn_b200.eps_plus_linear import (
    EPSesPlusLinear,
    ManuallyChosenInitialization,
    UnitTheoreticalOutputStd,
    ZeroCenteredNormalInitialization,
    ZeroCenteredUniformInitialization,
)
from dctn_b200.epses_composition import inner_product, specs_to_full_specs
from dctn_b200.pos2d import Pos2D, index_to_pos, pos_to_index

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "dctn_b200.h")).read()
    declared = set(re.findall(r"\b(dctn_[a-z0-9_]+)\s*\(", header))
    declared -= {"dctn_status", "dctn_plan"}
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in include/dctn_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert _lib.lib().dctn_version() == 100


def test_plan_queries_and_errors_without_gpu():
    lib = _lib.lib()
    plan = lib.dctn_eps_plan_get(1, 4, 2, 4, _lib.F32, _lib.VARIANT_AUTO)
    assert plan
    assert plan == lib.dctn_eps_plan_get(1, 4, 2, 4, _lib.F32, _lib.VARIANT_AUTO)  # cached
    desc = lib.dctn_eps_plan_describe(plan).decode()
    assert "n=16" in desc and "m=8" in desc and "D=65536" in desc
    assert lib.dctn_eps_workspace_bytes(plan, 512, 28, 28, _lib.WS_BACKWARD_CORE) > 0
    # K*K*C beyond the factor limit, and a core too large to exist
    assert not lib.dctn_eps_plan_get(4, 3, 2, 2, _lib.F32, 0)
    assert "factors" in _lib.last_error()
    assert not lib.dctn_eps_plan_get(1, 4, 4, 2, _lib.F32, 0)
    assert "too large" in _lib.last_error()
    assert not lib.dctn_eps_plan_get(1, 2, 2, 2, 7, 0)
    # image smaller than the kernel -> bad argument, null pointers -> bad argument
    assert lib.dctn_eps_forward(plan, None, None, None, 1, 3, 3, None, 0, None) == -1
    assert lib.dctn_eps_forward(plan, None, None, None, 1, 28, 28, None, 0, None) == -1
    assert lib.dctn_logmatmulexp_forward(None, None, None, 4, 4, 4, 0, None) == -1


def test_eps_shape_contract_and_no_cpu_fallback():
    x = torch.randn(1, 2, 5, 5, 2)
    good = torch.randn(2, 2, 2, 2, 3)
    with pytest.raises(AssertionError):  # dctn/eps.py:22
        eps(torch.randn(3, 3, 3, 3, 3), x)
    with pytest.raises(AssertionError):
        eps_one_by_one(torch.randn(2, 2, 2, 3), x)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        eps(good, x)
    from dctn_b200.logmatmulexp import logmatmulexp

    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        logmatmulexp(torch.randn(3, 4), torch.randn(4, 5))
    with pytest.raises(AssertionError):
        logmatmulexp(torch.randn(3, 4), torch.randn(5, 5))


def test_variant_switch():
    assert eps_mod.get_default_variant() in _lib.VARIANTS
    old = eps_mod.get_default_variant()
    eps_mod.set_default_variant("ffma")
    assert eps_mod.get_default_variant() == "ffma"
    eps_mod.set_default_variant(old)


def test_pos2d_round_trip():
    """reference tests/test_pos2d.py."""
    for max_w in (0, 3, 7):
        for i in range(3 * (max_w + 1)):
            assert pos_to_index(max_w, index_to_pos(max_w, i)) == i
    assert pos_to_index(3, Pos2D(2, 1)) == 9
    assert index_to_pos(3, 11) == Pos2D(2, 3)


def test_align_order_matches_reference_factor_order():
    from dctn_b200.align import align

    x = torch.arange(2 * 1 * 3 * 3 * 1, dtype=torch.float32).reshape(2, 1, 3, 3, 1)
    views = list(align(x, 2))
    assert len(views) == 8 and all(v.shape == (1, 2, 2, 1) for v in views)
    # factor j = (dh*K + dw)*C + c
    for dh, dw, c in product(range(2), range(2), range(2)):
        j = (dh * 2 + dw) * 2 + c
        assert torch.equal(views[j], x[c][:, dh : dh + 2, dw : dw + 2])


def test_contraction_path_cache_call_forms_and_singleton():
    """reference tests/test_contraction_path_cache.py."""
    cache = cpc.ContractionPathCache()
    a, b = torch.randn(3, 4), torch.randn(4, 5)
    ab0 = cache.contract("ij,jk->ijk", a, b)
    for ab in (
        cache.contract("ij,jk->ijk", a, b),
        cache.contract(a, "ij", b, "jk", "ijk"),
        cache.contract(a, (0, 1), b, (1, 2), (0, 1, 2)),
        cpc.contract(a, ("x", "y"), b, ("y", "z"), ("x", "y", "z")),
    ):
        assert torch.equal(ab, ab0)
    assert cpc.ContractionPathCache() is cache
    assert torch.allclose(cpc.contract("ij,jk", a, b), a @ b, atol=1e-6)


def test_contract_on_input_dims_known_answers():
    """reference tests/test_eps.py:64-73."""
    a = torch.einsum("oi,j->ijo", torch.eye(3), 2.0 * torch.ones(3))
    assert torch.allclose(contract_on_input_dims(a, a), 12.0 * torch.eye(3))
    a = torch.einsum("oi,j->ijo", 2.0 * torch.eye(4), torch.tensor([1.0, 2.0, 3.0, 4.0]))
    b = torch.einsum("pj,i->ijp", 3.0 * torch.eye(4), torch.ones(4))
    want = torch.einsum("o,p->op", 2.0 * torch.ones(4), torch.tensor([3.0, 6.0, 9.0, 12.0]))
    assert torch.allclose(contract_on_input_dims(a, b), want)


def test_composition_inner_product_known_answers():
    """reference tests/test_epses_composition.py:7-41."""
    a = torch.einsum("oi,j->ijo", torch.eye(3), torch.ones(3))
    assert torch.allclose(inner_product((a,), (a,)), torch.tensor(9.0))
    assert torch.allclose(inner_product((a, a), (a, a)), torch.tensor(3.0 ** 4))
    assert torch.allclose(inner_product((a, a, a), (a, a, a)), torch.tensor(3.0 ** 8))
    green = torch.einsum("oj,i->ijo", torch.eye(6)[:4], torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, 6.0]))
    black = torch.einsum("oi,j->ijo", torch.eye(4)[:3], torch.tensor([1.5, 0.0, 0.0, 0.0]))
    orange = torch.einsum("oi,j->ijo", torch.eye(6)[:4], torch.ones(6))
    red = torch.einsum("oi,j->ijo", torch.eye(4)[1:], torch.tensor([1.0, 0.0, 0.0, 1.0]))
    assert torch.allclose(inner_product((green, black), (orange, red)), torch.tensor((2 + 3 + 4) * 5 * 1.5))


def test_inner_product_matches_golden():
    from conftest import load_golden

    g = load_golden("inner_product_random")
    assert torch.allclose(inner_product((g["a1"], g["b1"]), (g["a2"], g["b2"])), g["ip_two"], rtol=1e-10, atol=1e-10)


def test_specs_to_full_specs():
    full = specs_to_full_specs(((4, 4), (3, 6)), 2)
    assert full == (
        {"kernel_size": 4, "in_num_channels": 1, "in_size": 2, "out_size": 4},
        {"kernel_size": 3, "in_num_channels": 1, "in_size": 4, "out_size": 6},
    )


def test_epses_plus_linear_manually_chosen_initialization():
    """reference tests/test_eps_plus_linear.py:13-36 (constructor API, parameter layout, init statistics)."""
    epses_specs = ((4, 4), (3, 4), (3, 6))
    initialization = ManuallyChosenInitialization(
        (
            ZeroCenteredNormalInitialization(0.1),
            ZeroCenteredUniformInitialization(77.0),
            ZeroCenteredNormalInitialization(10.0),
        ),
        ZeroCenteredUniformInitialization(500.0),
        ZeroCenteredNormalInitialization(1e-6),
    )
    for p, dtype in product((1e-3, 0.4, 0.6, 1.0), (torch.float32, torch.float64)):
        model = EPSesPlusLinear(epses_specs, initialization, p, torch.device("cpu"), dtype)
        assert 0.09 <= model.epses[0].std() <= 0.11
        assert -77.0 <= model.epses[1].min() <= -70.0
        assert 70.0 <= model.epses[1].max() <= 77
        assert 9.0 <= model.epses[2].std() <= 11.0
        assert -500.0 <= model.linear.weight.min() <= -460.0
        assert 460.0 <= model.linear.weight.max() <= 500.0
        assert 1e-9 <= model.linear.bias.std() <= 1e-3
        assert model.epses[0].dtype == dtype and model.linear.weight.dtype == dtype


def test_state_dict_layout_matches_reference():
    model = EPSesPlusLinear(((4, 4), (3, 6)), UnitTheoreticalOutputStd(), 1.0, torch.device("cpu"), torch.float32)
    assert list(model.state_dict().keys()) == ["p", "epses.0", "epses.1", "linear.weight", "linear.bias"]
    assert model.epses[0].shape == (2,) * 16 + (4,)
    assert model.epses[1].shape == (4,) * 9 + (6,)
    assert model.linear.in_features == 23 * 23 * 6


def test_install_as_dctn_alias():
    import sys

    saved = {k: v for k, v in sys.modules.items() if k == "dctn" or k.startswith("dctn.")}
    for k in saved:
        del sys.modules[k]
    try:
        dctn_b200.install_as_dctn()
        import dctn.eps as e  # noqa

        assert e is eps_mod
    finally:
        for k in [k for k in sys.modules if k == "dctn" or k.startswith("dctn.")]:
            del sys.modules[k]
        sys.modules.update(saved)
