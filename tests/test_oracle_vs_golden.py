"""CPU: pins the oracle (oracle/eps_oracle.py) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Tolerance: float64, rtol 1e-12 (the restatement follows the reference's own
contraction order, most cases agree bit for bit)."""
import os
from functools import reduce

import pytest
import torch

from conftest import CONVSBS_CASES, EPS_GOLDEN_CASES, LME_GOLDEN_CASES, WINDOW_STATS_CASES, load_golden
from oracle import eps_oracle as O

TOL = dict(rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", EPS_GOLDEN_CASES)
def test_eps_forward_and_grads(name):
    g = load_golden(name)
    for fn in (O.eps_4step, O.eps_one_by_one, O.eps_dense):
        assert torch.allclose(fn(g["core"], g["x"]), g["out"], **TOL), fn.__name__
    assert torch.allclose(g["out_one_by_one"], g["out"], **TOL)
    dcore, dx = O.eps_grads(g["core"], g["x"], g["gout"])
    assert torch.allclose(dcore, g["dcore"], **TOL)
    assert torch.allclose(dx, g["dx"], **TOL)


@pytest.mark.parametrize("name", CONVSBS_CASES)
def test_eps_matches_convsbs(name):
    """reference tests/test_conversion_of_convsbs_to_eps.py:13-56: eps(as_eps) == ConvSBS, outputs and input grads."""
    g = load_golden(name)
    assert torch.allclose(O.eps_4step(g["eps_tensor"], g["x"]), g["convsbs_out"], rtol=1e-9, atol=1e-11)
    _, dx = O.eps_grads(g["eps_tensor"], g["x"], g["gout"])
    assert torch.allclose(dx, g["convsbs_dx"], rtol=1e-9, atol=1e-11)


def test_composition_three_layers():
    g = load_golden("composition_3layers")
    cores = [g[k].clone().requires_grad_(True) for k in ("e1", "e2", "e3")]
    x = g["x"].clone().requires_grad_(True)
    out = O.contract_with_input(cores, x)
    assert torch.allclose(out, g["out"], **TOL)
    out.backward(g["gout"])
    for c, k in zip(cores, ("de1", "de2", "de3")):
        assert torch.allclose(c.grad, g[k], **TOL)
    assert torch.allclose(x.grad, g["dx"], **TOL)


@pytest.mark.parametrize("name,ncores", [("epl_cfg1_k2q2", 1), ("epl_two_layers", 2)])
def test_eps_plus_linear_logits_and_regularisers(name, ncores):
    g = load_golden(name)
    cores = [g[f"eps{i}"] for i in range(ncores)]
    assert torch.allclose(O.phi_cos_sin_squared(g["u"]), g["x"], **TOL)
    logits = O.eps_plus_linear_forward(cores, g["weight"], g["bias"], g["x"])
    assert torch.allclose(logits, g["logits"], **TOL)
    wn = (g["weight"] ** 2).sum()
    assert torch.allclose(wn + O.epswise_squared_fro_norm(cores), g["reg_epswise"], **TOL)
    assert torch.allclose(wn + O.composition_inner_product(cores, cores), g["reg_composition"], rtol=1e-10, atol=1e-12)


def test_inner_product_random_and_known_answers():
    g = load_golden("inner_product_random")
    assert torch.allclose(O.composition_inner_product((g["a1"],), (g["a2"],)), g["ip_single"], **TOL)
    assert torch.allclose(O.composition_inner_product((g["a1"], g["b1"]), (g["a2"], g["b2"])), g["ip_two"], rtol=1e-10, atol=1e-10)
    assert torch.allclose(O.contract_on_input_dims(g["a1"], g["a2"]), g["coid"], **TOL)
    # analytic answers of reference tests/test_epses_composition.py:7-13
    a = torch.einsum("oi,j->ijo", torch.eye(3), torch.ones(3))
    assert torch.allclose(O.composition_inner_product((a,), (a,)), torch.tensor(9.0))
    assert torch.allclose(O.composition_inner_product((a, a), (a, a)), torch.tensor(3.0 ** 4))
    assert torch.allclose(O.composition_inner_product((a, a, a), (a, a, a)), torch.tensor(3.0 ** 8))


@pytest.mark.parametrize("name", LME_GOLDEN_CASES)
def test_logmatmulexp(name):
    g = load_golden(name)
    assert torch.allclose(O.logmatmulexp(g["log_A"], g["log_B"]), g["out"], **TOL)
    assert torch.allclose(g["out_lowmem"], g["out"], **TOL)
    dA, dB = O.logmatmulexp_grads(g["log_A"], g["log_B"], g["gout"])
    assert torch.allclose(dA, g["dA"], **TOL) and torch.allclose(dB, g["dB"], **TOL)


def test_logmatmulexp_chain():
    g = load_golden("lme_chain6")
    mats = [g[f"m{i}"] for i in range(6)]
    assert torch.allclose(reduce(O.logmatmulexp, mats), g["out"], **TOL)


@pytest.mark.skipif(not os.path.isdir("/root/reference/dctn"), reason="reference tree only exists in the build container")
def test_oracle_against_live_reference():
    """Build container only: fresh random draw, oracle vs the reference imported through the shim."""
    from oracle.ref_import import import_reference

    ref = import_reference()
    torch.manual_seed(1234)
    x = torch.randn(1, 2, 6, 6, 3, dtype=torch.float64, requires_grad=True)
    core = torch.randn(*(3,) * 4, 4, dtype=torch.float64, requires_grad=True)
    out = ref.eps.eps(core, x)
    gout = torch.randn_like(out)
    out.backward(gout)
    assert torch.allclose(O.eps_4step(core.detach(), x.detach()), out, **TOL)
    dcore, dx = O.eps_grads(core, x, gout)
    assert torch.allclose(dcore, core.grad, **TOL) and torch.allclose(dx, x.grad, **TOL)


# ---------------------------------------------------------------- ConvSBS in log space, batched logmatmulexp
from conftest import CONVSBS_LOG_CASES, LME_BATCHED_CASES, convsbs_log_case  # noqa: E402


@pytest.mark.parametrize("name", CONVSBS_LOG_CASES)
def test_conv_sbs_oracle_linear_and_log(name):
    """The reference's ConvSBS.forward (dctn/conv_sbs.py:258-304) on positive cores/inputs, its log, and the
    gradients w.r.t. the LOG cores and LOG input."""
    g = load_golden(name)
    log_cores, positions, log_x = convsbs_log_case(g)
    lin = O.conv_sbs_forward([c.exp() for c in log_cores], positions, log_x.exp())
    assert torch.allclose(lin.log(), g["log_out"], rtol=1e-10, atol=1e-10)
    lc = [c.clone().requires_grad_(True) for c in log_cores]
    lx = log_x.clone().requires_grad_(True)
    out = O.conv_sbs_log_forward(lc, positions, lx)
    assert torch.allclose(out, g["log_out"], rtol=1e-10, atol=1e-10)
    out.backward(g["gout"])
    assert torch.allclose(lx.grad, g["dlog_x"], rtol=1e-8, atol=1e-10)
    for i, c in enumerate(lc):
        assert torch.allclose(c.grad, g[f"dlog_core{i}"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("name", LME_BATCHED_CASES)
def test_logmatmulexp_batched_oracle(name):
    g = load_golden(name)
    A = g["log_A"].clone().requires_grad_(True)
    B = g["log_B"].clone().requires_grad_(True)
    out = O.logmatmulexp_batched(A, B)
    assert torch.allclose(out, g["out"], rtol=1e-12, atol=1e-12)
    out.backward(g["gout"])
    assert torch.allclose(A.grad, g["dA"], rtol=1e-10, atol=1e-12) and torch.allclose(B.grad, g["dB"], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("name", WINDOW_STATS_CASES)
def test_window_stats(name):
    """make_windows + RankOneTensorsBatch statistics of the reference (tests/golden/make_golden_stats.py)."""
    g = load_golden(name)
    K = int(g["kernel_size"])
    total, sqn, ntensors, ncoord = O.window_stats(g["x"], K)
    assert ntensors == int(g["ntensors"]) and ncoord == int(g["ncoordinates"])
    assert torch.allclose(total, g["sum"], rtol=1e-12) and torch.allclose(sqn, g["sqnorm"], rtol=1e-12)
    mean, var_u = O.window_mean_var(g["x"], K, True)
    assert torch.allclose(mean, g["mean"], rtol=1e-12) and torch.allclose(var_u, g["var_unbiased"], rtol=1e-10)
    assert torch.allclose(O.window_mean_var(g["x"], K, False)[1], g["var_biased"], rtol=1e-10)
    # std_over_batch(unbiased=False) of the reference ignores its argument (dctn/rank_one_tensor.py:108-110)
    assert torch.allclose(var_u ** 0.5, g["std"], rtol=1e-10)


def test_empirical_std_init():
    """UnitEmpiricalOutputStd cores of the reference under a fixed seed, and unit output std per layer."""
    g = load_golden("stats_empirical_init")
    torch.manual_seed(int(g["seed"]))
    cores = O.empirical_std_cores(((2, 3), (2, 4)), g["x"], int(g["batch_size"]))
    assert torch.allclose(cores[0], g["core0"], rtol=1e-11, atol=1e-13)
    assert torch.allclose(cores[1], g["core1"], rtol=1e-11, atol=1e-13)
    assert torch.allclose(g["out_stds"], torch.ones(2, dtype=torch.float64), rtol=1e-10)


def test_intermediate_reps_log_numbers():
    """(mu, sigma) of every line log_intermediate_reps_stats writes (dctn/eps_plus_linear.py:161-196)."""
    g = load_golden("stats_log_lines")
    x, got = g["x"], []
    for core in (g["core0"], g["core1"]):
        got.append((x.mean(), x.std(unbiased=False)))
        K = 2
        mean, var = O.window_mean_var(x, K, True)
        got.append((mean, var ** 0.5))
        x = O.transform_in_slices(core, x, 4)
    flat = x.squeeze(0).flatten(start_dim=1)
    got.append((flat.mean(), flat.std(unbiased=False)))
    for t in (flat @ g["weight"].T, flat @ g["weight"].T + g["bias"]):
        got.append((t.mean(), t.std(unbiased=False)))
    mus = torch.stack([m for m, _ in got])
    sigmas = torch.stack([s for _, s in got])
    # the golden numbers were parsed from the log lines: 8 significant digits
    assert torch.allclose(mus, g["mus"], rtol=2e-7, atol=1e-12) and torch.allclose(sigmas, g["sigmas"], rtol=2e-7)
