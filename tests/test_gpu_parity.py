"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
(1) golden vectors produced by the unmodified reference, (2) the CPU oracle on seeded inputs, and
(3) size-independent identities at BASELINE.json's full sizes.

Tolerances (Frobenius-relative):
  float64 kernels vs float64 reference : 1e-11
  float32 kernels (FFMA and 3-pass TF32 tcgen05) vs float64 reference : 1e-5
"""
from functools import reduce

import pytest
import torch

from conftest import CONVSBS_CASES, EPS_GOLDEN_CASES, LME_GOLDEN_CASES, load_golden
from oracle import eps_oracle as O
from oracle.eps_oracle import rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-11, torch.float32: 1e-5}
DEV = "cuda:0"


def _eps_fwd_bwd(core64, x64, gout64, dtype, variant="auto"):
    from dctn_b200 import eps as E

    old = E.get_default_variant()
    E.set_default_variant(variant)
    try:
        core = core64.to(DEV, dtype).requires_grad_(True)
        x = x64.to(DEV, dtype).requires_grad_(True)
        out = E.eps(core, x)
        out.backward(gout64.to(DEV, dtype))
        torch.cuda.synchronize()
        return out.detach(), core.grad, x.grad
    finally:
        E.set_default_variant(old)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", EPS_GOLDEN_CASES)
def test_eps_golden(name, dtype):
    g = load_golden(name)
    out, dcore, dx = _eps_fwd_bwd(g["core"], g["x"], g["gout"], dtype)
    assert out.shape == g["out"].shape and out.dtype == dtype
    assert rel_err(out, g["out"]) <= TOL[dtype]
    assert rel_err(dcore, g["dcore"]) <= TOL[dtype]
    assert rel_err(dx, g["dx"]) <= TOL[dtype]


@pytest.mark.parametrize("name", EPS_GOLDEN_CASES[:3])
def test_eps_one_by_one_golden(name):
    """reference tests/test_eps.py:9-61 exercise eps_one_by_one."""
    from dctn_b200.eps import eps_one_by_one

    g = load_golden(name)
    out = eps_one_by_one(g["core"].to(DEV), g["x"].to(DEV))
    assert rel_err(out, g["out_one_by_one"]) <= TOL[torch.float64]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", CONVSBS_CASES)
def test_eps_matches_convsbs(name, dtype):
    """reference tests/test_conversion_of_convsbs_to_eps.py:13-56 (output and input-gradient vs ConvSBS)."""
    g = load_golden(name)
    out, _, dx = _eps_fwd_bwd(g["eps_tensor"], g["x"], g["gout"], dtype)
    tol = 1e-9 if dtype == torch.float64 else 1e-5
    assert rel_err(out, g["convsbs_out"]) <= tol
    assert rel_err(dx, g["convsbs_dx"]) <= tol


# (C, B, H, W, Q, K, O) — seeded random shapes checked against the oracle; includes the paper's layer
# shapes (K=4,Q=2,O=4 and K=3,Q=4,O=6), CIFAR-shaped (Q=23 -> O=24), ragged patch counts, B=1, H=K.
ORACLE_SHAPES = [
    (1, 3, 9, 8, 2, 4, 4),
    (1, 2, 7, 6, 4, 3, 6),
    (1, 2, 6, 5, 23, 2, 24),
    (1, 5, 28, 28, 2, 2, 2),
    (1, 1, 2, 2, 2, 2, 3),
    (1, 1, 3, 7, 3, 3, 2),
    (2, 3, 5, 4, 3, 2, 4),
    (1, 2, 8, 8, 4, 2, 23),
    (1, 130, 4, 4, 2, 2, 6),
    (1, 2, 5, 5, 12, 2, 24),
    (1, 3, 4, 6, 5, 1, 7),
]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("shape", ORACLE_SHAPES)
def test_eps_vs_oracle(shape, dtype):
    C, B, H, W, Q, K, Oq = shape
    g = torch.Generator().manual_seed(hash(shape) % (2 ** 31))
    n = K * K * C
    x = torch.randn(C, B, H, W, Q, dtype=torch.float64, generator=g) * (0.9 if n > 4 else 1.0)
    core = torch.randn(*(Q,) * n, Oq, dtype=torch.float64, generator=g) * Q ** (-n / 2)
    gout = torch.randn(B, H - K + 1, W - K + 1, Oq, dtype=torch.float64, generator=g)
    if dtype == torch.float32:  # compare on the float32-rounded inputs, in float64 arithmetic
        x, core, gout = x.float().double(), core.float().double(), gout.float().double()
    want = O.eps_4step(core, x)
    want_dcore, want_dx = O.eps_grads(core, x, gout)
    out, dcore, dx = _eps_fwd_bwd(core, x, gout, dtype)
    assert rel_err(out, want) <= TOL[dtype]
    assert rel_err(dcore, want_dcore) <= TOL[dtype]
    assert rel_err(dx, want_dx) <= TOL[dtype]


@pytest.mark.parametrize("variant", ["ffma"])
@pytest.mark.parametrize("shape", [(1, 3, 9, 8, 2, 4, 4), (1, 2, 7, 6, 4, 3, 6), (1, 5, 12, 12, 2, 2, 2)])
def test_eps_forced_variant_vs_oracle(shape, variant):
    C, B, H, W, Q, K, Oq = shape
    g = torch.Generator().manual_seed(7)
    n = K * K * C
    x = (torch.randn(C, B, H, W, Q, dtype=torch.float64, generator=g) * 0.9).float().double()
    core = (torch.randn(*(Q,) * n, Oq, dtype=torch.float64, generator=g) * Q ** (-n / 2)).float().double()
    gout = torch.randn(B, H - K + 1, W - K + 1, Oq, dtype=torch.float64, generator=g).float().double()
    want = O.eps_4step(core, x)
    want_dcore, want_dx = O.eps_grads(core, x, gout)
    out, dcore, dx = _eps_fwd_bwd(core, x, gout, torch.float32, variant)
    assert rel_err(out, want) <= 1e-5 and rel_err(dcore, want_dcore) <= 1e-5 and rel_err(dx, want_dx) <= 1e-5


def test_eps_non_contiguous_input_and_no_grad_paths():
    """transform_in_slices hands eps() non-contiguous slices when C > 1 (dctn/eps.py:136)."""
    from dctn_b200.eps import eps, transform_in_slices

    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 7, 5, 5, 2, dtype=torch.float64, generator=g)
    core = torch.randn(*(2,) * 8, 3, dtype=torch.float64, generator=g)
    want = torch.cat([O.eps_4step(core, s) for s in x.split(3, dim=1)]).unsqueeze(0)
    got = transform_in_slices(core.to(DEV), x.to(DEV), 3)
    assert got.shape == want.shape and rel_err(got, want) <= 1e-11
    xs = x.to(DEV)[:, 1:4]
    assert not xs.is_contiguous()
    assert rel_err(eps(core.to(DEV), xs), O.eps_4step(core, x[:, 1:4])) <= 1e-11
    # only the core requires grad (layer 1 of a model: input is data)
    c = core.to(DEV).requires_grad_(True)
    eps(c, x.to(DEV)).sum().backward()
    assert c.grad is not None


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_composition_three_layers(dtype):
    from dctn_b200.epses_composition import contract_with_input

    g = load_golden("composition_3layers")
    cores = [g[k].to(DEV, dtype).requires_grad_(True) for k in ("e1", "e2", "e3")]
    x = g["x"].to(DEV, dtype).requires_grad_(True)
    out = contract_with_input(cores, x)
    out.backward(g["gout"].to(DEV, dtype))
    assert rel_err(out, g["out"]) <= TOL[dtype]
    for c, k in zip(cores, ("de1", "de2", "de3")):
        assert rel_err(c.grad, g[k]) <= TOL[dtype]
    assert rel_err(x.grad, g["dx"]) <= TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name,specs,img", [("epl_cfg1_k2q2", ((2, 2),), 28), ("epl_two_layers", ((3, 4), (2, 6)), 10)])
def test_eps_plus_linear_logits_and_grads(name, specs, img, dtype):
    """Config-1-shaped model and a 2-layer model: logits, loss and every parameter gradient vs the reference."""
    from dctn_b200.eps_plus_linear import EPSesPlusLinear, UnitTheoreticalOutputStd

    g = load_golden(name)
    model = EPSesPlusLinear(specs, UnitTheoreticalOutputStd(), 1.0, torch.device(DEV), dtype, image_size=img)
    with torch.no_grad():
        for i in range(len(specs)):
            model.epses[i].copy_(g[f"eps{i}"])
        model.linear.weight.copy_(g["weight"])
        model.linear.bias.copy_(g["bias"])
    logits = model(g["x"].to(DEV, dtype))
    loss = torch.nn.functional.cross_entropy(logits, g["y"].to(DEV))
    loss.backward()
    assert rel_err(logits, g["logits"]) <= TOL[dtype]
    assert rel_err(loss, g["loss"]) <= TOL[dtype]
    for i in range(len(specs)):
        assert rel_err(model.epses[i].grad, g[f"deps{i}"]) <= TOL[dtype]
    assert rel_err(model.linear.weight.grad, g["dweight"]) <= TOL[dtype]
    assert rel_err(model.linear.bias.grad, g["dbias"]) <= TOL[dtype]
    assert rel_err(model.epswise_l2_regularizer(), g["reg_epswise"]) <= TOL[dtype]
    assert rel_err(model.epses_composition_l2_regularizer(), g["reg_composition"]) <= 10 * TOL[dtype]


def test_core_dropout_is_one_mask_per_step():
    from dctn_b200.eps_plus_linear import EPSesPlusLinear, UnitTheoreticalOutputStd

    model = EPSesPlusLinear(((2, 2),), UnitTheoreticalOutputStd(), 0.5, torch.device(DEV), torch.float32, image_size=6)
    x = torch.rand(1, 4, 6, 6, 2, device=DEV)
    model.train()
    torch.manual_seed(5)
    a = model(x)
    torch.manual_seed(5)
    b = model(x)
    assert torch.equal(a, b)  # same seed -> same mask -> identical (what the DP ranks rely on)
    model.eval()
    assert not torch.equal(model(x), a)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", LME_GOLDEN_CASES)
def test_logmatmulexp_golden(name, dtype):
    from dctn_b200.logmatmulexp import logmatmulexp, logmatmulexp_lowmem

    g = load_golden(name)
    A = g["log_A"].to(DEV, dtype).requires_grad_(True)
    B = g["log_B"].to(DEV, dtype).requires_grad_(True)
    if dtype == torch.float32:  # reference values on the float32-rounded inputs
        a64, b64 = A.detach().double().cpu(), B.detach().double().cpu()
        want = O.logmatmulexp(a64, b64)
        want_dA, want_dB = O.logmatmulexp_grads(a64, b64, g["gout"].float().double())
    else:
        want, want_dA, want_dB = g["out"], g["dA"], g["dB"]
    out = logmatmulexp(A, B)
    out.backward(g["gout"].to(DEV, dtype))
    tol = 1e-11 if dtype == torch.float64 else 1e-5
    assert rel_err(out, want) <= tol
    assert rel_err(A.grad, want_dA) <= tol and rel_err(B.grad, want_dB) <= tol
    assert rel_err(logmatmulexp_lowmem(A.detach(), B.detach()), want) <= tol
    assert torch.isfinite(out).all()


def test_logmatmulexp_chain_and_extremes():
    from dctn_b200.logmatmulexp import logmatmulexp

    g = load_golden("lme_chain6")
    mats = [g[f"m{i}"].to(DEV) for i in range(6)]
    mats[0].requires_grad_(True)
    out = reduce(logmatmulexp, mats)
    out.backward(torch.ones_like(out))
    assert rel_err(out, g["out"]) <= 1e-11
    assert rel_err(mats[0].grad, g["dm0"]) <= 1e-11
    # -inf entries (log of zero probabilities) must not produce NaNs
    A = torch.full((4, 5), float("-inf"), device=DEV, dtype=torch.float32)
    A[:, 0] = 0.0
    B = torch.randn(5, 3, device=DEV)
    out = logmatmulexp(A, B)
    assert torch.allclose(out, B[0].expand(4, 3), atol=1e-6)


# ---------------------------------------------------------------- batched logmatmulexp, ConvSBS in log space (8f-4)
from conftest import CONVSBS_LOG_CASES, LME_BATCHED_CASES, convsbs_log_case  # noqa: E402


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", LME_BATCHED_CASES)
def test_logmatmulexp_batched_golden(name, dtype):
    """One reference logmatmulexp call per batch element (dctn/logmatmulexp.py:5-14) vs the batched kernel."""
    from dctn_b200.logmatmulexp import logmatmulexp_batched

    g = load_golden(name)
    A = g["log_A"].to(DEV, dtype).requires_grad_(True)
    B = g["log_B"].to(DEV, dtype).requires_grad_(True)
    if dtype == torch.float32:  # reference values on the float32-rounded inputs
        a64, b64 = A.detach().double().cpu().requires_grad_(True), B.detach().double().cpu().requires_grad_(True)
        want = O.logmatmulexp_batched(a64, b64)
        want.backward(g["gout"].float().double())
        want, want_dA, want_dB = want.detach(), a64.grad, b64.grad
    else:
        want, want_dA, want_dB = g["out"], g["dA"], g["dB"]
    out = logmatmulexp_batched(A, B)
    out.backward(g["gout"].to(DEV, dtype))
    tol = TOL[dtype]
    assert out.shape == want.shape and out.dtype == dtype
    assert rel_err(out, want) <= tol
    assert rel_err(A.grad, want_dA) <= tol and rel_err(B.grad, want_dB) <= tol


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (1, 3, 5, 7), (5, 2, 40, 3), (1001, 4, 4, 4), (333, 12, 12, 12),
                                   (3, 70, 9, 33), (2, 1, 64, 1), (130, 8, 16, 8), (9, 24, 12, 36), (77, 5, 7, 20)])
def test_logmatmulexp_batched_ragged(shape, dtype):
    """Group tails, sizes that defeat the 128-bit staging, one element per CTA, leading dims, -inf entries."""
    from dctn_b200.logmatmulexp import logmatmulexp_batched

    NB, T, R, I = shape
    gen = torch.Generator().manual_seed(1000 + NB + T)
    A = (3 * torch.randn(NB, T, R, generator=gen, dtype=torch.float64)).to(dtype)
    B = (3 * torch.randn(NB, R, I, generator=gen, dtype=torch.float64)).to(dtype)
    if R > 1:
        A[:, :, 0] = float("-inf")  # log of a zero column: must not produce NaN
    gout = torch.randn(NB, T, I, generator=gen, dtype=torch.float64).to(dtype)
    a64, b64 = A.double().clone().requires_grad_(True), B.double().clone().requires_grad_(True)
    want = O.logmatmulexp_batched(a64, b64)
    want.backward(gout.double())
    Ad, Bd = A.to(DEV).requires_grad_(True), B.to(DEV).requires_grad_(True)
    out = logmatmulexp_batched(Ad, Bd)
    out.backward(gout.to(DEV))
    assert torch.isfinite(out).all() and torch.isfinite(Ad.grad).all() and torch.isfinite(Bd.grad).all()
    assert rel_err(out, want) <= TOL[dtype]
    assert rel_err(Ad.grad, a64.grad) <= TOL[dtype] and rel_err(Bd.grad, b64.grad) <= TOL[dtype]
    if NB > 1:  # extra leading dims are flattened
        out2 = logmatmulexp_batched(Ad.detach().reshape(1, NB, T, R), Bd.detach().reshape(1, NB, R, I))
        assert out2.shape == (1, NB, T, I) and torch.equal(out2[0], out.detach())


def test_logmatmulexp_batched_errors():
    from dctn_b200.logmatmulexp import logmatmulexp_batched

    with pytest.raises(RuntimeError):  # CPU tensors: no fallback
        logmatmulexp_batched(torch.zeros(2, 3, 3), torch.zeros(2, 3, 3))
    with pytest.raises(RuntimeError, match="shared memory"):  # one pair larger than a CTA's shared memory
        logmatmulexp_batched(torch.zeros(2, 200, 200, device=DEV), torch.zeros(2, 200, 200, device=DEV))
    with pytest.raises(AssertionError):
        logmatmulexp_batched(torch.zeros(2, 3, 4, device=DEV), torch.zeros(2, 3, 4, device=DEV))


def test_logmatmulexp_batched_full_size_identities():
    """Config 5 extension size: one 8x8 bond-matrix product per window of a 28x28 image with a 3x3 string, batch 2048."""
    from dctn_b200.logmatmulexp import logmatmulexp_batched

    NB, r = 2048 * 26 * 26, 8
    gen = torch.Generator(device=DEV).manual_seed(7)
    A = 2 * torch.randn(NB, r, r, device=DEV, generator=gen)
    B = 2 * torch.randn(NB, r, r, device=DEV, generator=gen)
    out = logmatmulexp_batched(A, B)
    assert torch.equal(out, logmatmulexp_batched(A, B))  # deterministic
    eye = torch.full((r, r), float("-inf"), device=DEV)
    eye.fill_diagonal_(0.0)
    assert torch.equal(logmatmulexp_batched(A, eye.expand(NB, r, r).contiguous()), A)  # log-identity is neutral
    shift = torch.randn(NB, 1, 1, device=DEV)
    assert torch.allclose(logmatmulexp_batched(A + shift, B), out + shift, rtol=0, atol=2e-5)  # shift equivariance
    # spot-check 4096 elements against the float64 oracle
    idx = torch.randint(0, NB, (4096,), device=DEV, generator=gen)
    want = O.logmatmulexp_batched(A[idx].double().cpu(), B[idx].double().cpu())
    assert rel_err(out[idx], want) <= TOL[torch.float32]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("shape", [(70001, 2, 64), (20000, 4, 24), (100000, 9, 5)])
def test_logmatmulexp_tall_product(shape, dtype):
    """One row per ConvSBS window: Theta beyond the 65535-block grid.y limit, dB as a chunked reduction over Theta."""
    from dctn_b200.logmatmulexp import logmatmulexp

    T, R, I = shape
    gen = torch.Generator().manual_seed(77)
    A = torch.randn(T, R, generator=gen, dtype=torch.float64).to(dtype)
    B = torch.randn(R, I, generator=gen, dtype=torch.float64).to(dtype)
    gout = torch.randn(T, I, generator=gen, dtype=torch.float64).to(dtype)
    a64, b64 = A.double().clone().requires_grad_(True), B.double().clone().requires_grad_(True)
    want = O.logmatmulexp(a64, b64)
    want.backward(gout.double())
    Ad, Bd = A.to(DEV).requires_grad_(True), B.to(DEV).requires_grad_(True)
    out = logmatmulexp(Ad, Bd)
    out.backward(gout.to(DEV))
    assert rel_err(out, want) <= TOL[dtype]
    assert rel_err(Ad.grad, a64.grad) <= TOL[dtype] and rel_err(Bd.grad, b64.grad) <= TOL[dtype]
    Bd.grad = None
    logmatmulexp(Ad.detach(), Bd).backward(gout.to(DEV))   # dB only
    assert rel_err(Bd.grad, b64.grad) <= TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", CONVSBS_LOG_CASES)
def test_conv_sbs_log_golden(name, dtype):
    """log of the reference's ConvSBS.forward (dctn/conv_sbs.py:258-304) and gradients w.r.t. log cores / log input."""
    from dctn_b200.conv_sbs_log import conv_sbs_log_forward
    from dctn_b200.pos2d import Pos2D

    g = load_golden(name)
    log_cores, positions, log_x = convsbs_log_case(g)
    lc = [c.to(DEV, dtype).requires_grad_(True) for c in log_cores]
    lx = log_x.to(DEV, dtype).requires_grad_(True)
    if dtype == torch.float32:
        lc64 = [c.detach().double().cpu().requires_grad_(True) for c in lc]
        lx64 = lx.detach().double().cpu().requires_grad_(True)
        want = O.conv_sbs_log_forward(lc64, positions, lx64)
        want.backward(g["gout"].float().double())
        want, want_dx, want_dc = want.detach(), lx64.grad, [c.grad for c in lc64]
    else:
        want, want_dx, want_dc = g["log_out"], g["dlog_x"], [g[f"dlog_core{i}"] for i in range(len(lc))]
    out = conv_sbs_log_forward(lc, tuple(Pos2D(*p) for p in positions), lx)
    out.backward(g["gout"].to(DEV, dtype))
    tol = TOL[dtype]
    assert out.shape == want.shape
    assert rel_err(out, want) <= tol
    assert rel_err(lx.grad, want_dx) <= tol
    for got, w in zip(lc, want_dc):
        assert rel_err(got.grad, w) <= tol


# ---------------------------------------------------------------- full BASELINE sizes, oracle-free identities
FULL_LAYERS = [
    # (name, B, H, W, Q, K, O)  — config 2: layer 1 and layer 2 at batch 512; config 1 at batch 128 (float64)
    ("cfg2_L1", 512, 28, 28, 2, 4, 4, torch.float32),
    ("cfg2_L2", 512, 25, 25, 4, 3, 6, torch.float32),
    ("cfg1", 128, 28, 28, 2, 2, 2, torch.float64),
]


@pytest.mark.parametrize("name,B,H,W,Q,K,Oq,dtype", FULL_LAYERS)
def test_full_size_identities(name, B, H, W, Q, K, Oq, dtype):
    """At full size the oracle is too slow; use exact algebraic identities of the contraction instead:
       * adjointness:   <gout, eps(core, x)> == <dcore, core>      (eps is linear in core)
       * homogeneity:   <dx, x> == n * <gout, eps(core, x)>        (eps is multilinear in the n factors)
       * linearity in core: eps(2.5*c1 - c2, x) == 2.5*eps(c1, x) - eps(c2, x)
       * batch independence: the first image's patches equal a B=1 run (checked against the oracle)."""
    from dctn_b200.eps import eps

    gen = torch.Generator().manual_seed(11)
    n = K * K
    x = (torch.rand(1, B, H, W, Q, generator=gen, dtype=torch.float64) * 1.2 + 0.2).to(DEV, dtype).requires_grad_(True)
    c1 = (torch.randn(*(Q,) * n, Oq, generator=gen, dtype=torch.float64) * Q ** (-n / 2)).to(DEV, dtype).requires_grad_(True)
    c2 = (torch.randn(*(Q,) * n, Oq, generator=gen, dtype=torch.float64) * Q ** (-n / 2)).to(DEV, dtype)
    out = eps(c1, x)
    gout = torch.randn(out.shape, generator=gen, dtype=torch.float64).to(DEV, dtype)
    out.backward(gout)
    inner = (gout.double() * out.detach().double()).sum()
    # the scalar <gout, out> is a sum of ~1e6 random-sign terms: measure errors on the Cauchy-Schwarz scale
    scale = gout.double().norm() * out.detach().double().norm()
    tol = 1e-10 if dtype == torch.float64 else 1e-5
    assert abs((c1.grad.double() * c1.detach().double()).sum() - inner) / scale <= tol
    assert abs((x.grad.double() * x.detach().double()).sum() - n * inner) / (n * scale) <= tol
    with torch.no_grad():
        lin = eps(2.5 * c1 - c2, x)
        assert rel_err(lin, 2.5 * out.detach().double() - eps(c2, x).double()) <= tol
    # first image against the oracle
    want0 = O.eps_4step(c1.detach().double().cpu(), x.detach().double().cpu()[:, :1])
    assert rel_err(out[:1], want0) <= (1e-11 if dtype == torch.float64 else 1e-5)
    # core-gradient of a single image against the oracle
    want_dc, want_dx = O.eps_grads(c1.detach().double().cpu(), x.detach().double().cpu()[:, :1], gout[:1].double().cpu())
    c3 = c1.detach().clone().requires_grad_(True)
    x3 = x.detach()[:, :1].clone().requires_grad_(True)
    eps(c3, x3).backward(gout[:1])
    assert rel_err(c3.grad, want_dc) <= (1e-11 if dtype == torch.float64 else 1e-5)
    assert rel_err(x3.grad, want_dx) <= (1e-11 if dtype == torch.float64 else 1e-5)
    # input gradient of the batch run restricted to image 0 equals the single-image run (patches independent)
    assert rel_err(x.grad[:, :1], x3.grad) <= (1e-11 if dtype == torch.float64 else 1e-5)


# ---------------------------------------------------------------- tcgen05 (tensor-core) kernels, called directly
def _raw_call(kind, variant, core, x, gout):
    """Calls one C-ABI entry point with an explicit kernel variant; returns the output tensor."""
    from dctn_b200 import _lib
    from dctn_b200 import eps as E

    C, K, Q, Oq = E._infer(core, x)
    _, B, H, W, _ = x.shape
    plan = E._plan(C, K, Q, Oq, x.dtype, _lib.VARIANTS[variant])
    lib = _lib.lib()
    ws = torch.empty(lib.dctn_eps_workspace_bytes(plan, B, H, W, kind), dtype=torch.uint8, device=x.device)
    st = torch.cuda.current_stream().cuda_stream
    if kind == _lib.WS_FORWARD:
        out = torch.empty(B, H - K + 1, W - K + 1, Oq, dtype=x.dtype, device=x.device)
        rc = lib.dctn_eps_forward(plan, x.data_ptr(), core.data_ptr(), out.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(), st)
    elif kind == _lib.WS_BACKWARD_CORE:
        out = torch.empty_like(core)
        rc = lib.dctn_eps_backward_core(plan, x.data_ptr(), gout.data_ptr(), out.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(), st)
    else:
        out = torch.empty_like(x)
        rc = lib.dctn_eps_backward_input(plan, x.data_ptr(), core.data_ptr(), gout.data_ptr(), out.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(), st)
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    return out


def _rand_layer(B, H, W, Q, K, Oq, seed):
    gen = torch.Generator().manual_seed(seed)
    n = K * K
    x = (torch.rand(1, B, H, W, Q, generator=gen, dtype=torch.float64) * 1.2 + 0.2).float()
    core = (torch.randn(*(Q,) * n, Oq, generator=gen, dtype=torch.float64) * Q ** (-n / 2)).float()
    gout = torch.randn(B, H - K + 1, W - K + 1, Oq, generator=gen, dtype=torch.float64).float()
    return x, core, gout


TC_SHAPES = [
    # (B, H, W, Q, K, O): P >= 4096 so that the tensor-core family accepts them; ragged A / N tiles included
    (8, 28, 28, 2, 4, 4),
    (8, 25, 25, 4, 3, 6),
    (9, 27, 26, 3, 3, 5),   # A = 243, N = 405: ragged row and column tiles
    (30, 16, 15, 2, 4, 3),   # N = 768, odd Q_out
]


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_tc_backward_core_vs_oracle(shape):
    from dctn_b200 import _lib

    B, H, W, Q, K, Oq = shape
    x, core, gout = _rand_layer(B, H, W, Q, K, Oq, seed=21)
    want, _ = O.eps_grads(core.double(), x.double(), gout.double())
    got3 = _raw_call(_lib.WS_BACKWARD_CORE, "tc3", core.to(DEV), x.to(DEV), gout.to(DEV))
    assert rel_err(got3, want) <= 1e-5, "3-pass TF32 must be fp32-accurate"
    goth = _raw_call(_lib.WS_BACKWARD_CORE, "tch3", core.to(DEV), x.to(DEV), gout.to(DEV))
    assert rel_err(goth, want) <= 1e-5, "3-pass split fp16 must be fp32-accurate"
    got1 = _raw_call(_lib.WS_BACKWARD_CORE, "tc1", core.to(DEV), x.to(DEV), gout.to(DEV))
    assert rel_err(got1, want) <= 5e-3, "single-pass TF32 (opt-in) tolerance"


@pytest.mark.parametrize("B,H,W,Q,K,Oq", [(512, 28, 28, 2, 4, 4), (512, 25, 25, 4, 3, 6)])
def test_tc_backward_core_full_size_vs_fp64(B, H, W, Q, K, Oq):
    """Long reduction (P = 320 000 / 270 848 patches): the tensor-core result against our own float64
    CUDA-core kernels (which are pinned to the oracle above) — checks the fp32 accumulation chain."""
    from dctn_b200 import _lib

    x, core, gout = _rand_layer(B, H, W, Q, K, Oq, seed=22)
    want = _raw_call(_lib.WS_BACKWARD_CORE, "ffma", core.double().to(DEV), x.double().to(DEV), gout.double().to(DEV))
    got = _raw_call(_lib.WS_BACKWARD_CORE, "tc3", core.to(DEV), x.to(DEV), gout.to(DEV))
    err = rel_err(got, want)
    print(f"tc3 dcore full-size rel err {err:.3e}")
    assert err <= 1e-5
    goth = _raw_call(_lib.WS_BACKWARD_CORE, "tch3", core.to(DEV), x.to(DEV), gout.to(DEV))
    print(f"tch3 dcore full-size rel err {rel_err(goth, want):.3e}")
    assert rel_err(goth, want) <= 1e-5
    ffma = _raw_call(_lib.WS_BACKWARD_CORE, "ffma", core.to(DEV), x.to(DEV), gout.to(DEV))
    print(f"ffma dcore full-size rel err {rel_err(ffma, want):.3e}")
    assert rel_err(ffma, want) <= 1e-5


def test_tch3_zero_and_sparse_inputs():
    """Range normalisation must not invent NaNs: all-zero factor vectors (scale exponent 0), all-zero gout rows, whole zero
    images and a zero core give exact zeros / the oracle's values."""
    from dctn_b200 import _lib

    B, H, W, Q, K, Oq = 8, 25, 25, 4, 3, 6
    x, core, gout = _rand_layer(B, H, W, Q, K, Oq, seed=35)
    gen = torch.Generator().manual_seed(36)
    x = x * (torch.rand(1, B, H, W, 1, generator=gen) > 0.3)       # 30 % of the pixels have an all-zero feature vector
    x[:, 3] = 0.0                                                    # one whole image is zero
    gout = gout * (torch.rand(B, H - K + 1, W - K + 1, 1, generator=gen) > 0.5)
    want = O.eps_4step(core.double(), x.double())
    want_dc, want_dx = O.eps_grads(core.double(), x.double(), gout.double())
    got = _raw_call(_lib.WS_FORWARD, "tch3", core.to(DEV), x.to(DEV), gout.to(DEV))
    assert torch.isfinite(got).all() and rel_err(got, want) <= 1e-5
    assert float(got[3].abs().max()) == 0.0
    got_dc = _raw_call(_lib.WS_BACKWARD_CORE, "tch3", core.to(DEV), x.to(DEV), gout.to(DEV))
    assert torch.isfinite(got_dc).all() and rel_err(got_dc, want_dc) <= 1e-5
    out_s, dx_s = _raw_train_call("tch3", core.to(DEV), x.to(DEV), gout.to(DEV))
    assert torch.isfinite(dx_s).all() and rel_err(dx_s, want_dx) <= 1e-5
    zero_core = torch.zeros_like(core).to(DEV)
    z = _raw_call(_lib.WS_FORWARD, "tch3", zero_core, x.to(DEV), gout.to(DEV))
    assert float(z.abs().max()) == 0.0
    zg = _raw_call(_lib.WS_BACKWARD_CORE, "tch3", core.to(DEV), x.to(DEV), torch.zeros_like(gout).to(DEV))
    assert float(zg.abs().max()) == 0.0


def test_tc_two_channel_layer_vs_oracle():
    """C = 2, K = 3, Q = 2: 18 factors per patch (A = 512, Bn = 512).  The first half has 9 factors, more than the fully
    fused input-gradient epilogue unrolls, so this runs the W path (first leave-one-out stage in the GEMM epilogue,
    second stage in loo_groups_kernel); also the only tensor-core test with interleaved channels in the factor order."""
    from dctn_b200 import _lib

    gen = torch.Generator().manual_seed(33)
    C, B, H, W, Q, K, Oq = 2, 8, 25, 25, 2, 3, 4
    n = K * K * C
    x = (torch.rand(C, B, H, W, Q, generator=gen, dtype=torch.float64) * 1.2 + 0.2).float()
    core = (torch.randn(*(Q,) * n, Oq, generator=gen, dtype=torch.float64) * Q ** (-n / 2)).float()
    gout = torch.randn(B, H - K + 1, W - K + 1, Oq, generator=gen, dtype=torch.float64).float()
    want = O.eps_4step(core.double(), x.double())
    want_dc, want_dx = O.eps_grads(core.double(), x.double(), gout.double())
    for variant in ("tch3", "tc3"):
        assert rel_err(_raw_call(_lib.WS_FORWARD, variant, core.to(DEV), x.to(DEV), gout.to(DEV)), want) <= 1e-5
        assert rel_err(_raw_call(_lib.WS_BACKWARD_CORE, variant, core.to(DEV), x.to(DEV), gout.to(DEV)), want_dc) <= 1e-5
        assert rel_err(_raw_call(_lib.WS_BACKWARD_INPUT, variant, core.to(DEV), x.to(DEV), gout.to(DEV)), want_dx) <= 1e-5
        out, dx = _raw_train_call(variant, core.to(DEV), x.to(DEV), gout.to(DEV))
        assert rel_err(out, want) <= 1e-5 and rel_err(dx, want_dx) <= 1e-5


@pytest.fixture
def generic_tc_kernels(monkeypatch):
    """Forces the generic table-lookup tcgen05 GEMM kernels instead of the register-table ones (eps_tc_fast.cu)."""
    monkeypatch.setenv("DCTN_B200_NO_FAST", "1")


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_tch3_generic_kernels_vs_oracle(shape, generic_tc_kernels):
    from dctn_b200 import _lib

    B, H, W, Q, K, Oq = shape
    x, core, gout = _rand_layer(B, H, W, Q, K, Oq, seed=29)
    want = O.eps_4step(core.double(), x.double())
    _, want_dx = O.eps_grads(core.double(), x.double(), gout.double())
    assert rel_err(_raw_call(_lib.WS_FORWARD, "tch3", core.to(DEV), x.to(DEV), gout.to(DEV)), want) <= 1e-5
    assert rel_err(_raw_call(_lib.WS_BACKWARD_INPUT, "tch3", core.to(DEV), x.to(DEV), gout.to(DEV)), want_dx) <= 1e-5
    out, dx = _raw_train_call("tch3", core.to(DEV), x.to(DEV), gout.to(DEV))
    assert rel_err(out, want) <= 1e-5 and rel_err(dx, want_dx) <= 1e-5


@pytest.mark.parametrize("variant,tol", [("tch3", 1e-5), ("tc3", 1e-5), ("tc1", 5e-3)])
@pytest.mark.parametrize("shape", TC_SHAPES + [(20, 16, 16, 8, 2, 24), (6, 30, 31, 8, 2, 5), (3, 40, 41, 16, 2, 3)])
def test_tc_forward_and_input_grad_vs_oracle(shape, variant, tol):
    """tcgen05 forward (fused KR2 epilogue) and input-gradient (two GEMMs + leave-one-out + gather) vs the oracle.
    The extra shapes exercise the register-table kernels with lo groups of 8 (Q=8, K=2) and 16 (Q=16, K=2) entries."""
    from dctn_b200 import _lib

    B, H, W, Q, K, Oq = shape
    x, core, gout = _rand_layer(B, H, W, Q, K, Oq, seed=23)
    want = O.eps_4step(core.double(), x.double())
    _, want_dx = O.eps_grads(core.double(), x.double(), gout.double())
    got = _raw_call(_lib.WS_FORWARD, variant, core.to(DEV), x.to(DEV), gout.to(DEV))
    assert rel_err(got, want) <= tol
    got_dx = _raw_call(_lib.WS_BACKWARD_INPUT, variant, core.to(DEV), x.to(DEV), gout.to(DEV))
    assert rel_err(got_dx, want_dx) <= tol


@pytest.mark.parametrize("B,H,W,Q,K,Oq", [(512, 28, 28, 2, 4, 4), (512, 25, 25, 4, 3, 6)])
def test_tc_forward_and_input_grad_full_size_vs_fp64(B, H, W, Q, K, Oq):
    """Full config-2 layer sizes: tensor-core forward / input-gradient against our own float64 CUDA-core kernels."""
    from dctn_b200 import _lib

    x, core, gout = _rand_layer(B, H, W, Q, K, Oq, seed=24)
    x64, c64, g64 = x.double().to(DEV), core.double().to(DEV), gout.double().to(DEV)
    want = _raw_call(_lib.WS_FORWARD, "ffma", c64, x64, g64)
    got = _raw_call(_lib.WS_FORWARD, "tc3", core.to(DEV), x.to(DEV), gout.to(DEV))
    print(f"tc3 forward full-size rel err {rel_err(got, want):.3e}")
    assert rel_err(got, want) <= 1e-5
    want_dx = _raw_call(_lib.WS_BACKWARD_INPUT, "ffma", c64, x64, g64)
    got_dx = _raw_call(_lib.WS_BACKWARD_INPUT, "tc3", core.to(DEV), x.to(DEV), gout.to(DEV))
    print(f"tc3 input-grad full-size rel err {rel_err(got_dx, want_dx):.3e}")
    assert rel_err(got_dx, want_dx) <= 1e-5
    out_s, dx_s = _raw_train_call("tc3", core.to(DEV), x.to(DEV), gout.to(DEV))
    print(f"tc3 saved-T path full-size rel err: forward {rel_err(out_s, want):.3e}, input-grad {rel_err(dx_s, want_dx):.3e}")
    assert torch.equal(out_s, got)
    assert rel_err(dx_s, want_dx) <= 1e-5
    # split-fp16 arithmetic (what AUTO uses): same tolerance
    got_h = _raw_call(_lib.WS_FORWARD, "tch3", core.to(DEV), x.to(DEV), gout.to(DEV))
    dx_h = _raw_call(_lib.WS_BACKWARD_INPUT, "tch3", core.to(DEV), x.to(DEV), gout.to(DEV))
    out_hs, dx_hs = _raw_train_call("tch3", core.to(DEV), x.to(DEV), gout.to(DEV))
    print(f"tch3 full-size rel err: forward {rel_err(got_h, want):.3e}, input-grad {rel_err(dx_h, want_dx):.3e}, saved-T input-grad {rel_err(dx_hs, want_dx):.3e}")
    assert rel_err(got_h, want) <= 1e-5 and rel_err(dx_h, want_dx) <= 1e-5 and rel_err(dx_hs, want_dx) <= 1e-5
    assert torch.equal(out_hs, got_h)


@pytest.mark.parametrize("core_scale", [1e-12, 1.0, 1e12])
def test_tch3_range_normalisation(core_scale):
    """fp16 has a 5-bit exponent: the split-fp16 kernels rescale every factor vector, the gout row and the core by exact
    powers of two.  Inputs whose magnitude varies over many decades from pixel to pixel (and a core far outside the fp16
    range) must come out as accurately as with well-scaled data."""
    from dctn_b200 import _lib

    B, H, W, Q, K, Oq = 8, 25, 25, 4, 3, 6
    x, core, gout = _rand_layer(B, H, W, Q, K, Oq, seed=27)
    gen = torch.Generator().manual_seed(28)
    # per-pixel magnitudes 10^U(-1.5, 1.5): a 3x3 patch product then spans ~27 decades between patches (fp16 covers 12)
    x = x * (10.0 ** (torch.rand(1, B, H, W, 1, generator=gen) * 3 - 1.5))
    gout = gout * (10.0 ** (torch.rand(B, H - K + 1, W - K + 1, 1, generator=gen) * 8 - 4))
    core = core * core_scale
    want = O.eps_4step(core.double(), x.double())
    _, want_dx = O.eps_grads(core.double(), x.double(), gout.double())
    got = _raw_call(_lib.WS_FORWARD, "tch3", core.to(DEV), x.to(DEV), gout.to(DEV))
    # row-wise check: every patch individually accurate (a Frobenius norm over all patches would hide the small ones)
    num = (got.double().cpu() - want).flatten(0, 2).norm(dim=1)
    den = want.flatten(0, 2).norm(dim=1)
    assert (num / den).max().item() <= 1e-5
    # core gradient: a sum over patches of very different magnitude — accurate relative to the whole sum
    want_dc, _ = O.eps_grads(core.double(), x.double(), gout.double())
    got_dc = _raw_call(_lib.WS_BACKWARD_CORE, "tch3", core.to(DEV), x.to(DEV), gout.to(DEV))
    assert rel_err(got_dc, want_dc) <= 1e-5
    got_dx = _raw_call(_lib.WS_BACKWARD_INPUT, "tch3", core.to(DEV), x.to(DEV), gout.to(DEV))
    out_s, dx_s = _raw_train_call("tch3", core.to(DEV), x.to(DEV), gout.to(DEV))
    assert torch.equal(out_s, got)
    for d in (got_dx, dx_s):
        # dx of a pixel sums over the patches containing it; compare per image
        num = (d.double().cpu() - want_dx).flatten(2).norm(dim=2)
        den = want_dx.flatten(2).norm(dim=2)
        assert (num / den).max().item() <= 1e-5


def _raw_train_call(variant, core, x, gout):
    """dctn_eps_forward_train + dctn_eps_backward_input_saved through the C ABI; returns (out, dx)."""
    from dctn_b200 import _lib
    from dctn_b200 import eps as E

    C, K, Q, Oq = E._infer(core, x)
    _, B, H, W, _ = x.shape
    plan = E._plan(C, K, Q, Oq, x.dtype, _lib.VARIANTS[variant])
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    nsave = lib.dctn_eps_saved_bytes(plan, B, H, W)
    assert nsave == x.shape[1] * (H - K + 1) * (W - K + 1) * Q ** (K * K * C // 2) * Oq * 4
    saved = torch.empty(nsave, dtype=torch.uint8, device=x.device)
    ws = torch.empty(lib.dctn_eps_workspace_bytes(plan, B, H, W, _lib.WS_FORWARD), dtype=torch.uint8, device=x.device)
    out = torch.empty(B, H - K + 1, W - K + 1, Oq, dtype=x.dtype, device=x.device)
    rc = lib.dctn_eps_forward_train(plan, x.data_ptr(), core.data_ptr(), out.data_ptr(), saved.data_ptr(), nsave, B, H, W,
                                    ws.data_ptr(), ws.numel(), st)
    assert rc == 0, _lib.last_error()
    ws2 = torch.empty(lib.dctn_eps_workspace_bytes(plan, B, H, W, _lib.WS_BACKWARD_INPUT_SAVED), dtype=torch.uint8, device=x.device)
    dx = torch.empty_like(x)
    rc = lib.dctn_eps_backward_input_saved(plan, x.data_ptr(), core.data_ptr(), gout.data_ptr(), saved.data_ptr(), nsave,
                                           dx.data_ptr(), B, H, W, ws2.data_ptr(), ws2.numel(), st)
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    return out, dx


@pytest.mark.parametrize("variant", ["tch3", "tc3"])
@pytest.mark.parametrize("shape", TC_SHAPES)
def test_tc_saved_intermediate_path_vs_oracle(shape, variant):
    """Training forward that keeps T + input gradient from the saved T (one GEMM instead of two) vs the oracle."""
    B, H, W, Q, K, Oq = shape
    x, core, gout = _rand_layer(B, H, W, Q, K, Oq, seed=25)
    want = O.eps_4step(core.double(), x.double())
    _, want_dx = O.eps_grads(core.double(), x.double(), gout.double())
    out, dx = _raw_train_call(variant, core.to(DEV), x.to(DEV), gout.to(DEV))
    assert rel_err(out, want) <= 1e-5
    assert rel_err(dx, want_dx) <= 1e-5


def test_saved_intermediate_entry_points_reject_other_families():
    """Families without a savable intermediate report 0 bytes and refuse the training entry points (no fallback)."""
    from dctn_b200 import _lib
    from dctn_b200 import eps as E

    lib = _lib.lib()
    plan = E._plan(1, 2, 2, 2, torch.float32, _lib.VARIANT_AUTO)      # direct family
    assert lib.dctn_eps_saved_bytes(plan, 8, 28, 28) == 0
    plan64 = E._plan(1, 3, 4, 6, torch.float64, _lib.VARIANT_AUTO)    # float64: CUDA-core family
    assert lib.dctn_eps_saved_bytes(plan64, 8, 25, 25) == 0
    buf = torch.empty(1 << 20, dtype=torch.uint8, device=DEV)
    rc = lib.dctn_eps_forward_train(plan, buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), buf.numel(), 8, 28, 28,
                                    buf.data_ptr(), buf.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == -2 and "no savable intermediate" in _lib.last_error()


def test_autograd_saved_and_recompute_paths_agree():
    """EpsFunction with and without the saved intermediate (save limit 0) gives the same gradients."""
    from dctn_b200 import eps as E

    x, core, gout = _rand_layer(8, 25, 25, 4, 3, 6, seed=26)
    res = []
    for limit in (16384, 0):
        E.set_save_limit_mb(limit)
        try:
            xd = x.to(DEV).requires_grad_(True)
            cd = core.to(DEV).requires_grad_(True)
            E.eps(cd, xd).backward(gout.to(DEV))
            res.append((xd.grad.clone(), cd.grad.clone()))
        finally:
            E.set_save_limit_mb(16384)
    _, want_dx = O.eps_grads(core.double(), x.double(), gout.double())
    assert rel_err(res[0][0], want_dx) <= 1e-5 and rel_err(res[1][0], want_dx) <= 1e-5
    assert rel_err(res[0][0], res[1][0]) <= 2e-6
    assert torch.equal(res[0][1], res[1][1])


def test_lopsided_split_for_mid_sized_core():
    """CIFAR (2, 6 -> 24), layer 2: the reference's split (dctn/eps.py:25-27) gives A = Bn = 36, too narrow for the 128-row
    tensor-core tiles; the plan takes m = 3 (A = 216, N = 144) so that both gradients run on tcgen05.  Same results."""
    from dctn_b200 import _lib
    from dctn_b200.eps import plan_description

    B, H, W, Q, K, Oq = 16, 20, 21, 6, 2, 24
    x, core, gout = _rand_layer(B, H, W, Q, K, Oq, seed=29)
    assert "split m=3 (A=216, Bn=6, N=144" in plan_description(core, x)
    assert "split m=2" in plan_description(core.double(), x.double())   # float64 keeps the reference's split
    want = O.eps_4step(core.double(), x.double())
    want_dcore, want_dx = O.eps_grads(core.double(), x.double(), gout.double())
    cd, xd, gd = core.to(DEV), x.to(DEV), gout.to(DEV)
    assert rel_err(_raw_call(_lib.WS_FORWARD, "auto", cd, xd, gd), want) <= 1e-5
    assert rel_err(_raw_call(_lib.WS_BACKWARD_CORE, "auto", cd, xd, gd), want_dcore) <= 1e-5
    assert rel_err(_raw_call(_lib.WS_BACKWARD_INPUT, "auto", cd, xd, gd), want_dx) <= 1e-5
    out, dcore, dx = _eps_fwd_bwd(core.double(), x.double(), gout.double(), torch.float32)   # autograd path (saved T or not)
    assert rel_err(out, want) <= 1e-5 and rel_err(dcore, want_dcore) <= 1e-5 and rel_err(dx, want_dx) <= 1e-5


# ---------------------------------------------------------------- direct (tiny-core) kernels and the host-buffer entry
DIRECT_SHAPES = [
    # (C, B, H, W, Q, K, O)
    (1, 7, 28, 28, 2, 2, 2),    # config-1 layer shape (warp-row kernel, compile-time Q_out)
    (1, 3, 9, 40, 2, 2, 5),     # W > 32: two column tiles; odd Q_out (runtime path)
    (1, 2, 6, 7, 3, 2, 6),      # Q = 3
    (1, 2, 8, 8, 4, 2, 23),     # CIFAR layer-1 shape (generic thread-per-patch kernel)
    (1, 2, 7, 6, 2, 3, 4),      # K = 3, Q = 2: 5 + 4 factors
    # larger patch counts: many chunks / slices of the tiled core gradient, many CTAs of the per-image input gradient
    (1, 40, 28, 28, 4, 2, 6),   # K = 2, Q = 4: 96 register tiles, 2 patch slices
    (1, 30, 20, 21, 6, 2, 4),   # Q = 6: 324 tiles = two tiles per thread; non-square image
    (1, 24, 14, 15, 2, 3, 6),   # K = 3, Q = 2, O = 6: 3072-element core (tiled kernel only); two-kernel input gradient
    (1, 33, 28, 28, 3, 2, 5),   # Q = 3: Bn = 9 padded to 12, odd Q_out
    (1, 64, 28, 28, 2, 2, 6),   # register-resident core gradient with 96 accumulators
    (1, 50, 12, 13, 5, 2, 3),   # Q = 5: A = Bn = 25 padded to 28
    (2, 9, 10, 11, 3, 1, 4),    # two channels, K = 1: channel planes in the per-image gather
    (1, 6, 16, 17, 6, 2, 24),   # CIFAR (2, 6 -> 24) layer: core gradient cut into 4 groups of column tiles
    (1, 5, 12, 12, 4, 2, 23),   # CIFAR layer 1 with Q_out = 23: odd Q_out across group boundaries
]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("shape", DIRECT_SHAPES)
def test_direct_kernels_vs_oracle(shape, dtype):
    from dctn_b200 import _lib

    C, B, H, W, Q, K, Oq = shape
    g = torch.Generator().manual_seed(31)
    n = K * K * C
    x = (torch.randn(C, B, H, W, Q, dtype=torch.float64, generator=g) * 0.9).to(dtype).double()
    core = (torch.randn(*(Q,) * n, Oq, dtype=torch.float64, generator=g) * Q ** (-n / 2)).to(dtype).double()
    gout = torch.randn(B, H - K + 1, W - K + 1, Oq, dtype=torch.float64, generator=g).to(dtype).double()
    want = O.eps_4step(core, x)
    want_dcore, want_dx = O.eps_grads(core, x, gout)
    xd, cd, gd = x.to(DEV, dtype), core.to(DEV, dtype), gout.to(DEV, dtype)
    tol = TOL[dtype]
    # the direct kernels cover only small cores (forward: transposed core in <= 36 KB of shared memory; backward:
    # everything of a patch in registers); outside that range the forced variant must report "unsupported" (never
    # silently fall back) and AUTO takes another family
    for kind, ref in ((_lib.WS_FORWARD, want), (_lib.WS_BACKWARD_CORE, want_dcore), (_lib.WS_BACKWARD_INPUT, want_dx)):
        try:
            got = _raw_call(kind, "direct", cd, xd, gd)
        except AssertionError as err:
            assert "does not support this shape" in str(err)
            got = _raw_call(kind, "auto", cd, xd, gd)
        assert rel_err(got, ref) <= tol


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("B,H,W,Oq", [(5, 28, 28, 2), (3, 9, 40, 5), (2, 7, 33, 6)])
def test_eps_from_pixels_vs_oracle(B, H, W, Oq, dtype):
    """dctn_eps_forward_from_pixels: the feature map of dctn/dataset_loading.py:33-36 fused into the K=2 forward kernel,
    against the oracle on phi(pixels); the core gradient too."""
    from dctn_b200 import eps as E

    gen = torch.Generator().manual_seed(41)
    u = torch.rand(B, H, W, generator=gen, dtype=torch.float64).to(dtype).double()
    core = (torch.randn(2, 2, 2, 2, Oq, generator=gen, dtype=torch.float64) * 0.25).to(dtype).double()
    gout = torch.randn(B, H - 1, W - 1, Oq, generator=gen, dtype=torch.float64).to(dtype).double()
    scale = 1.45646
    x = O.phi_cos_sin_squared(u, scale / 2)           # the oracle's nu convention: 2 * nu = scale
    want = O.eps_4step(core, x)
    want_dcore, _ = O.eps_grads(core, x, gout)
    cd = core.to(DEV, dtype).requires_grad_(True)
    out = E.eps_from_pixels(cd, u.to(DEV, dtype), scale)
    out.backward(gout.to(DEV, dtype))
    tol = TOL[dtype]
    assert rel_err(out, want) <= tol
    assert rel_err(cd.grad, want_dcore) <= tol
    assert rel_err(E.phi_sin_cos_squared(u.to(DEV, dtype), scale), x) <= tol


def test_eps_from_pixels_rejects_other_layers():
    from dctn_b200 import eps as E

    with pytest.raises(RuntimeError, match="K=2, C=1, Q_in=2"):
        E.eps_from_pixels(torch.randn(*(2,) * 9, 4, device=DEV), torch.rand(2, 8, 8, device=DEV))


def test_forward_host_entry():
    """dctn_eps_forward_host: host buffers in, host buffer out (copies inside), against the oracle."""
    import ctypes

    from dctn_b200 import _lib
    from dctn_b200 import eps as E

    g = torch.Generator().manual_seed(41)
    x = torch.randn(1, 4, 9, 8, 2, generator=g)
    core = torch.randn(*(2,) * 9, 4, generator=g) * 2 ** -4.5
    want = O.eps_4step(core.double(), x.double())
    plan = E._plan(1, 3, 2, 4, torch.float32, _lib.VARIANT_AUTO)
    lib = _lib.lib()
    nbytes = lib.dctn_eps_forward_host_device_bytes(plan, 4, 9, 8)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    out = torch.empty(4, 7, 6, 4)
    rc = lib.dctn_eps_forward_host(plan, x.data_ptr(), core.data_ptr(), out.data_ptr(), 4, 9, 8, scratch.data_ptr(), nbytes,
                                   torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    assert rel_err(out, want) <= 1e-5
    # too small a scratch buffer is reported, not overrun
    assert lib.dctn_eps_forward_host(plan, x.data_ptr(), core.data_ptr(), out.data_ptr(), 4, 9, 8, scratch.data_ptr(), 16,
                                     torch.cuda.current_stream().cuda_stream) == -3


@pytest.mark.parametrize("specs,Q0,img", [(((2, 6), (2, 24)), 4, 12), (((2, 12), (2, 24)), 4, 10)])
def test_cifar_shaped_model_vs_oracle(specs, Q0, img):
    """Config 4 shapes (Q_0 = 4, layers (2,6|12),(2,24)) at a reduced image size: logits and all gradients vs the oracle."""
    from dctn_b200.eps_plus_linear import EPSesPlusLinear, UnitTheoreticalOutputStd

    torch.manual_seed(51)
    model = EPSesPlusLinear(specs, UnitTheoreticalOutputStd(), 1.0, torch.device(DEV), torch.float32, image_size=img, Q_0=Q0)
    x = torch.randn(1, 5, img, img, Q0) * 0.8
    y = torch.randint(0, 10, (5,))
    logits = model(x.to(DEV))
    torch.nn.functional.cross_entropy(logits, y.to(DEV)).backward()
    cores = [c.detach().double().cpu().requires_grad_(True) for c in model.epses]
    w = model.linear.weight.detach().double().cpu().requires_grad_(True)
    b = model.linear.bias.detach().double().cpu().requires_grad_(True)
    ref = O.eps_plus_linear_forward(cores, w, b, x.double())
    torch.nn.functional.cross_entropy(ref, y).backward()
    assert rel_err(logits, ref) <= 1e-5
    for i, c in enumerate(cores):
        assert rel_err(model.epses[i].grad, c.grad) <= 1e-5
    assert rel_err(model.linear.weight.grad, w.grad) <= 1e-5


# (B, H, W, Q_out): the streaming K=2, Q_in=2 float32 kernels (csrc/eps_stream_k2q2.cu) — one warp per block of 9 or 4 output
# rows x 31 columns: full and ragged row blocks, several column tiles, odd Q_out, one and several output pairs, 2 x 2 images
STREAM_SHAPES = [(5, 28, 28, 2), (3, 28, 28, 3), (2, 28, 28, 4), (2, 28, 28, 5), (2, 28, 28, 6), (2, 28, 28, 8), (2, 28, 28, 1),
                 (3, 11, 40, 3), (2, 5, 70, 5), (4, 2, 2, 2), (2, 46, 33, 2), (2, 47, 64, 4), (1, 30, 32, 7), (7, 10, 19, 6)]


@pytest.mark.parametrize("shape", STREAM_SHAPES)
def test_stream_k2q2_vs_oracle(shape):
    """Forward (featurised and raw-pixel entry) and both gradients of the K=2, Q_in=2 float32 family against the oracle
    (dctn/eps.py:19-40); pixels outside [0, 1] exercise the period reduction of the fused feature map."""
    from dctn_b200 import eps as E

    B, H, W, Oq = shape
    g = torch.Generator().manual_seed(1000 + B + 7 * H + 31 * W + Oq)
    x = torch.randn(1, B, H, W, 2, dtype=torch.float64, generator=g).float().double()
    core = (torch.randn(2, 2, 2, 2, Oq, dtype=torch.float64, generator=g) * 0.25).float().double()
    gout = torch.randn(B, H - 1, W - 1, Oq, dtype=torch.float64, generator=g).float().double()
    want = O.eps_4step(core, x)
    want_dcore, want_dx = O.eps_grads(core, x, gout)
    out, dcore, dx = _eps_fwd_bwd(core, x, gout, torch.float32)
    assert out.shape == want.shape
    assert rel_err(out, want) <= 1e-5
    assert (out.double().cpu() - want).abs().max() <= 1e-5 * want.abs().max()     # element-wise: no patch is skipped
    assert rel_err(dcore, want_dcore) <= 1e-5
    assert rel_err(dx, want_dx) <= 1e-5
    u = (torch.rand(B, H, W, generator=g, dtype=torch.float64) * 5 - 2).float().double()
    u[0, 0, 0] = 0.0; u[0, -1, -1] = 1.0
    xp = O.phi_cos_sin_squared(u, 1.45646 / 2)
    wantp = O.eps_4step(core, xp)
    outp = E.eps_from_pixels(core.to(DEV, torch.float32), u.to(DEV, torch.float32), 1.45646)
    assert rel_err(outp, wantp) <= 1e-5
    assert (outp.double().cpu() - wantp).abs().max() <= 2e-5 * wantp.abs().max()
