"""GPU parity tests of every path bench.py times, against the CPU oracle (oracle/eps_oracle.py), plus the statistics /
empirical-std initialisation path against goldens of the unmodified reference.

Each model-level test asserts through dctn_eps_kernel_family() which kernel family served every call, so a silent
CUDA-core path cannot pass for the tcgen05 kernels.  Tolerance: float32 kernels vs the float64 oracle, Frobenius-relative
<= 1e-5 (BASELINE.json north_star); float64 kernels <= 1e-11.
"""
import logging
import re

import pytest
import torch
import torch.nn.functional as F

from conftest import WINDOW_STATS_CASES, load_golden
from oracle import eps_oracle as O
from oracle.eps_oracle import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model_vs_oracle(specs, Q0, img, B, seed, input_scale=1.0, positive=False):
    """Builds EPSesPlusLinear(specs), runs logits + cross-entropy backward on the GPU and on the oracle; returns
    (model, x on the device, {name: rel-err}, launches)."""
    from dctn_b200 import _lib
    from dctn_b200.eps_plus_linear import EPSesPlusLinear, UnitTheoreticalOutputStd

    torch.manual_seed(seed)
    model = EPSesPlusLinear(specs, UnitTheoreticalOutputStd(), 1.0, torch.device(DEV), torch.float32, image_size=img, Q_0=Q0)
    if positive:   # FashionMNIST-shaped phi features
        x = O.phi_cos_sin_squared(torch.rand(B, img, img, dtype=torch.float64), input_scale / 2).float()
    else:          # CIFAR-shaped: normalised channels plus the constant channel (dctn/dataset_loading.py:349-364)
        x = torch.randn(1, B, img, img, Q0) * input_scale
        x[..., -1] = input_scale
    y = torch.randint(0, 10, (B,))
    l0 = _lib.launch_count()
    logits = model(x.to(DEV))
    F.cross_entropy(logits, y.to(DEV)).backward()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    cores = [c.detach().double().cpu().requires_grad_(True) for c in model.epses]
    w = model.linear.weight.detach().double().cpu().requires_grad_(True)
    b = model.linear.bias.detach().double().cpu().requires_grad_(True)
    ref = O.eps_plus_linear_forward(cores, w, b, x.double())
    F.cross_entropy(ref, y).backward()
    errs = {"logits": rel_err(logits, ref), "dweight": rel_err(model.linear.weight.grad, w.grad), "dbias": rel_err(model.linear.bias.grad, b.grad)}
    for i, c in enumerate(cores):
        errs[f"deps{i}"] = rel_err(model.epses[i].grad, c.grad)
    return model, x.to(DEV), errs, launches


def _layer_families(model, x):
    """[(forward, backward_core, backward_input) family per layer] for the shapes this model/input produces."""
    from dctn_b200.eps import eps, kernel_families

    fams, inter = [], x
    with torch.no_grad():
        for core in model.epses:
            fams.append(kernel_families(core, inter))
            inter = eps(core, inter).unsqueeze(0)
    return fams


def test_cfg2_model_on_tcgen05_vs_oracle():
    """BASELINE config 2 specs ((4,4),(3,6)) at image 28: logits and every gradient; B = 8 puts both layers above the
    tcgen05 thresholds (P = 5000 and 4232 patches), so the kernels of the headline benchmark are the ones checked."""
    from dctn_b200 import _lib

    model, x, errs, launches = _model_vs_oracle(((4, 4), (3, 6)), 2, 28, 8, seed=101, input_scale=1.45646, positive=True)
    fams = _layer_families(model, x)
    assert fams[0]["forward"] == _lib.FAMILY_TCGEN05 and fams[0]["backward_core"] == _lib.FAMILY_TCGEN05, fams
    assert all(f == _lib.FAMILY_TCGEN05 for f in fams[1].values()), fams
    assert launches >= 10
    assert all(v <= 1e-5 for v in errs.values()), str(errs)


def _isolated_layer_errors(model, x, y):
    """Every layer of `model` run ALONE on the float64 oracle's own inputs and upstream gradients (rounded to float32):
    {(layer, 'out' | 'dcore' | 'dx'): Frobenius-relative error}.  This is the per-kernel accuracy, free of the errors the
    earlier layers feed in."""
    from dctn_b200.eps import eps

    B = x.shape[1]
    cores = [c.detach().double().cpu().requires_grad_(True) for c in model.epses]
    w = model.linear.weight.detach().double().cpu()
    b = model.linear.bias.detach().double().cpu()
    inters = [x.detach().double().cpu().requires_grad_(True)]
    for c in cores:
        nxt = O.eps_4step(c, inters[-1]).unsqueeze(0)
        nxt.retain_grad()
        inters.append(nxt)
    F.cross_entropy(inters[-1].squeeze(0).reshape(B, -1) @ w.T + b, y.cpu()).backward()
    errs = {}
    for li, c in enumerate(cores):
        c32, x32, g32 = c.detach().float(), inters[li].detach().float(), inters[li + 1].grad.squeeze(0).float()
        cd, xd = c32.to(DEV).requires_grad_(True), x32.to(DEV).requires_grad_(True)
        out = eps(cd, xd)
        out.backward(g32.to(DEV))
        want = O.eps_4step(c32.double(), x32.double())
        wdc, wdx = O.eps_grads(c32.double(), x32.double(), g32.double())
        errs[(li, "out")], errs[(li, "dcore")], errs[(li, "dx")] = rel_err(out, want), rel_err(cd.grad, wdc), rel_err(xd.grad, wdx)
    return errs


def test_three_layer_model_vs_oracle():
    """Config 4's 3-layer spec (4,4),(3,12),(2,24) (three_epses_on_fashionmnist.py:16) at image 28, B = 9:
    layer 2 is K=3, Q=4 -> 12 and layer 3 is K=2, Q=12 -> 24, all on the tcgen05 family.

    Tolerances.  Every layer IN ISOLATION (oracle inputs, oracle upstream gradients) must meet 1e-5 on output and both
    gradients.  End to end a stack of multilinear layers amplifies whatever error a layer makes: the output of layer 1
    enters layer 2 as K*K = 9 factors and layer 3 as 4 more, so a systematic relative error e1 of layer 1 arrives at
    the logits as up to 36 * e1 (the reference's own float32 arithmetic is at 3e-6 on this model for the same reason,
    tools/diag_layers.py).  The tensor core's truncating accumulator makes our per-layer errors (0.3 - 1.3e-6)
    systematic, so the end-to-end gradients get 5e-5; the logits — dominated here by the bias, the 3-layer output is
    1e-16 with this initialisation — keep 1e-5.  Upstream gradients reach 1e-20 and flush to exact zeros for some
    patches: the core gradient of layer 1 must survive that (regression test for the E_max poisoning bug)."""
    from dctn_b200 import _lib

    model, x, errs, _ = _model_vs_oracle(((4, 4), (3, 12), (2, 24)), 2, 28, 9, seed=102, input_scale=1.45646, positive=True)
    fams = _layer_families(model, x)
    assert all(f["forward"] == _lib.FAMILY_TCGEN05 and f["backward_core"] == _lib.FAMILY_TCGEN05 for f in fams), fams
    assert fams[1]["backward_input"] == _lib.FAMILY_TCGEN05 and fams[2]["backward_input"] == _lib.FAMILY_TCGEN05, fams
    assert errs["logits"] <= 1e-5 and errs["dbias"] <= 1e-5, str(errs)
    assert all(v <= 5e-5 for v in errs.values()), str(errs)
    torch.manual_seed(102 + 1)
    y = torch.randint(0, 10, (x.shape[1],))
    iso = _isolated_layer_errors(model, x, y)
    assert all(v <= 1e-5 for v in iso.values()), str(iso)


@pytest.mark.parametrize("Q1", [12, 23])
def test_cifar_models_on_tcgen05_vs_oracle(Q1):
    """Config 4 CIFAR-shaped models (2,Q1),(2,24) with Q_0 = 4 at the full 32x32 image, B = 5: layer 2 has P = 4500
    patches, non-power-of-two Q (the table-lookup tcgen05 GEMMs and the three-level generated operands of the core
    gradient).  Also runs the recompute path of the input gradient (no saved T) and compares the two."""
    from dctn_b200 import _lib
    from dctn_b200 import eps as E

    model, x, errs, _ = _model_vs_oracle(((2, Q1), (2, 24)), 4, 32, 5, seed=103 + Q1, input_scale=0.8)
    fams = _layer_families(model, x)
    assert all(f == _lib.FAMILY_TCGEN05 for f in fams[1].values()), fams
    assert all(v <= 1e-5 for v in errs.values()), str(errs)
    # the same step with and without the saved intermediate T (recompute path of the input gradient)
    model.zero_grad(set_to_none=True)
    old = E._save_limit_bytes
    E.set_save_limit_mb(0)
    try:
        torch.manual_seed(0)
        y = torch.randint(0, 10, (x.shape[1],), device=DEV)
        F.cross_entropy(model(x), y).backward()
        g0 = [p.grad.clone() for p in model.parameters()]
        model.zero_grad(set_to_none=True)
        E.set_save_limit_mb(old / (1 << 20))
        F.cross_entropy(model(x), y).backward()
        for a, b in zip(g0, [p.grad for p in model.parameters()]):
            assert rel_err(a, b) <= 1e-5
    finally:
        E.set_save_limit_mb(old / (1 << 20))


@pytest.mark.parametrize("shape", [(5, 32, 32, 12, 2, 24), (5, 32, 32, 23, 2, 24), (6, 31, 31, 6, 2, 24), (7, 28, 28, 3, 3, 5)])
def test_tcgen05_generic_q_layers_vs_oracle(shape):
    """Single layers with non-power-of-two Q through the public autograd entry at sizes the tcgen05 family accepts:
    output, core gradient, input gradient (saved-T path) against the oracle."""
    from dctn_b200 import _lib
    from dctn_b200.eps import eps, kernel_families

    B, H, W, Q, K, Oq = shape
    gen = torch.Generator().manual_seed(77 + Q)
    n = K * K
    x = (torch.randn(1, B, H, W, Q, generator=gen, dtype=torch.float64) * 0.8).float()
    core = (torch.randn(*(Q,) * n, Oq, generator=gen, dtype=torch.float64) * Q ** (-n / 2)).float()
    gout = torch.randn(B, H - K + 1, W - K + 1, Oq, generator=gen, dtype=torch.float64).float()
    c = core.to(DEV).requires_grad_(True)
    xd = x.to(DEV).requires_grad_(True)
    fams = kernel_families(c, xd)
    assert fams["backward_core"] == _lib.FAMILY_TCGEN05 and fams["backward_input"] == _lib.FAMILY_TCGEN05, fams
    out = eps(c, xd)
    out.backward(gout.to(DEV))
    want = O.eps_4step(core.double(), x.double())
    want_dc, want_dx = O.eps_grads(core.double(), x.double(), gout.double())
    assert rel_err(out, want) <= 1e-5
    assert rel_err(c.grad, want_dc) <= 1e-5
    assert rel_err(xd.grad, want_dx) <= 1e-5


@pytest.mark.parametrize("shape,force", [((5, 32, 32, 13, 2, 8), False), ((3, 30, 30, 5, 3, 4), False), ((6, 31, 31, 6, 2, 24), True),
                                         ((5, 32, 32, 23, 2, 24), True), ((7, 28, 28, 3, 3, 5), True)])
def test_first_half_leave_one_out_rows_kernel_vs_oracle(shape, force, monkeypatch):
    """Input gradient of generic-Q layers through the warp-per-patch first-half leave-one-out kernel (loo1_rows_kernel) in
    its register-block instantiations: Q = 13 (lo group of 13 -> 16 registers) and K = 3, Q = 5 (25 -> 32) select it by
    themselves (A >= 1024); DCTN_B200_LOO1=1 forces it for Q = 6 (8 registers), Q = 23 (24, nine K-segment slices summed
    in the kernel) and K = 3, Q = 3 (12, 27 rows)."""
    from dctn_b200 import _lib
    from dctn_b200.eps import eps, kernel_families

    if force:
        monkeypatch.setenv("DCTN_B200_LOO1", "1")
    B, H, W, Q, K, Oq = shape
    gen = torch.Generator().manual_seed(171 + Q)
    n = K * K
    x = (torch.randn(1, B, H, W, Q, generator=gen, dtype=torch.float64) * 0.8).float()
    core = (torch.randn(*(Q,) * n, Oq, generator=gen, dtype=torch.float64) * Q ** (-n / 2)).float()
    gout = torch.randn(B, H - K + 1, W - K + 1, Oq, generator=gen, dtype=torch.float64).float()
    c = core.to(DEV).requires_grad_(True)
    xd = x.to(DEV).requires_grad_(True)
    fams = kernel_families(c, xd)
    assert fams["backward_input"] == _lib.FAMILY_TCGEN05, fams
    out = eps(c, xd)
    out.backward(gout.to(DEV))
    want = O.eps_4step(core.double(), x.double())
    want_dc, want_dx = O.eps_grads(core.double(), x.double(), gout.double())
    assert rel_err(out, want) <= 1e-5
    assert rel_err(c.grad, want_dc) <= 1e-5
    assert rel_err(xd.grad, want_dx) <= 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_eps_module_forward_backward(dtype):
    """EPS(nn.Module) (dctn/eps.py:73-96): parameter `core`, He-style init std, forward = eps(core, input)."""
    from dctn_b200.eps import EPS

    torch.manual_seed(3)
    layer = EPS(kernel_size=2, in_num_channels=2, in_size=3, out_size=5)
    assert [n for n, _ in layer.named_parameters()] == ["core"]
    assert layer.core.shape == (3,) * 8 + (5,) and layer.matrix_shape == (5, 3 ** 8)
    assert abs(layer.core.std().item() / 3 ** -4 - 1) < 0.05
    layer = layer.to(DEV, dtype)
    x = torch.randn(2, 4, 6, 7, 3, dtype=torch.float64).to(dtype)
    xd = x.to(DEV).requires_grad_(True)
    out = layer(xd)
    gout = torch.randn(out.shape, dtype=torch.float64).to(dtype)
    out.backward(gout.to(DEV))
    core64 = layer.core.detach().double().cpu()
    want = O.eps_4step(core64, x.double())
    want_dc, want_dx = O.eps_grads(core64, x.double(), gout.double())
    tol = 1e-5 if dtype == torch.float32 else 1e-11
    assert out.shape == (4, 5, 6, 5)
    assert rel_err(out, want) <= tol and rel_err(layer.core.grad, want_dc) <= tol and rel_err(xd.grad, want_dx) <= tol


# ---------------------------------------------------------------- statistics / empirical-std initialisation
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", WINDOW_STATS_CASES)
def test_make_windows_stats_golden(name, dtype):
    """dctn_b200.align.make_windows(...) statistics against the reference's make_windows + RankOneTensorsBatch."""
    from dctn_b200.align import make_windows

    g = load_golden(name)
    w = make_windows(g["x"].to(DEV, dtype), int(g["kernel_size"]))
    tol = 1e-12 if dtype == torch.float64 else 2e-6
    assert tuple(w.batch_shape) == tuple(int(v) for v in g["batch_shape"])
    assert w.ntensors == int(g["ntensors"]) and w.ncoordinates == int(g["ncoordinates"])
    for got, key in ((w.sum_over_batch(), "sum"), (w.squared_fro_norm_over_batch(), "sqnorm"), (w.mean_over_batch(), "mean"),
                     (w.var_over_batch(True), "var_unbiased"), (w.var_over_batch(False), "var_biased"), (w.std_over_batch(unbiased=False), "std")):
        assert abs(got.item() - g[key].item()) <= tol * max(abs(g[key].item()), abs(g["sqnorm"].item()) / (w.ntensors * w.ncoordinates)), key


def test_forward_stats_entry_matches_torch():
    """dctn_eps_forward_stats: output identical to the plain forward, (sum, sum of squares) accumulated over slices."""
    from dctn_b200.eps import eps, transform_in_slices, transform_in_slices_with_stats

    gen = torch.Generator().manual_seed(9)
    x = torch.randn(2, 37, 9, 8, 2, generator=gen, dtype=torch.float64)
    core = torch.randn(*(2,) * 8, 5, generator=gen, dtype=torch.float64)
    for dtype, tol in ((torch.float64, 1e-13), (torch.float32, 1e-6)):
        xd, cd = x.to(DEV, dtype), core.to(DEV, dtype)
        out, stats = transform_in_slices_with_stats(cd, xd, 8)       # 5 slices, the last one ragged; odd slice offsets
        plain = transform_in_slices(cd, xd, 8)
        assert out.shape == plain.shape and torch.equal(out, plain)
        assert stats.dtype == torch.float64
        o64 = plain.double()
        assert abs(stats[0].item() - o64.sum().item()) <= tol * o64.abs().sum().item()
        assert abs(stats[1].item() - (o64 ** 2).sum().item()) <= tol * (o64 ** 2).sum().item()
    assert rel_err(plain, torch.cat([O.eps_4step(core, s) for s in x.split(8, dim=1)]).unsqueeze(0)) <= 1e-5
    del eps


def test_empirical_std_init_golden():
    """EPSesPlusLinear(..., UnitEmpiricalOutputStd(x, 4)) under the reference's seed: same cores as the reference
    (tests/golden/make_golden_stats.py) — the random draws happen on the CPU in the same order — and unit output std."""
    from dctn_b200.eps import transform_in_slices
    from dctn_b200.eps_plus_linear import EPSesPlusLinear, UnitEmpiricalOutputStd

    g = load_golden("stats_empirical_init")
    for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 1e-5)):
        torch.manual_seed(int(g["seed"]))
        x = g["x"].to(dtype)
        model = EPSesPlusLinear(((2, 3), (2, 4)), UnitEmpiricalOutputStd(x, int(g["batch_size"])), 1.0, torch.device(DEV), dtype, image_size=9)
        if dtype == torch.float64:     # float32 randn draws differ from the float64 golden draws: compare float64 only
            assert rel_err(model.epses[0], g["core0"]) <= tol and rel_err(model.epses[1], g["core1"]) <= tol
        reps = x.to(DEV)
        for core in model.epses:
            reps = transform_in_slices(core.detach(), reps, 4)
            assert abs(reps.double().std(unbiased=False).item() - 1.0) <= 10 * tol
        # the oracle on the same draw (float32: the oracle sees the float32 draws in float64 arithmetic)
        torch.manual_seed(int(g["seed"]))
        if dtype == torch.float64:
            want = O.empirical_std_cores(((2, 3), (2, 4)), x, int(g["batch_size"]))
            for got, w in zip(model.epses, want):
                assert rel_err(got, w) <= tol


def test_log_intermediate_reps_stats_golden():
    """Same log lines (names, order) and numbers as the reference's log_intermediate_reps_stats, window lines included."""
    from dctn_b200.eps_plus_linear import EPSesPlusLinear, UnitTheoreticalOutputStd

    g = load_golden("stats_log_lines")
    model = EPSesPlusLinear(((2, 3), (2, 4)), UnitTheoreticalOutputStd(), 1.0, torch.device(DEV), torch.float64, image_size=9)
    with torch.no_grad():
        model.epses[0].copy_(g["core0"]); model.epses[1].copy_(g["core1"])
        model.linear.weight.copy_(g["weight"]); model.linear.bias.copy_(g["bias"])
    records = []

    class Grab(logging.Handler):
        def emit(self, record):
            records.append(record.getMessage())

    lg = logging.getLogger("dctn_b200.eps_plus_linear")
    lg.setLevel(logging.INFO)
    handler = Grab()
    lg.addHandler(handler)
    try:
        model.log_intermediate_reps_stats(g["x"].to(DEV), batch_size=4)
    finally:
        lg.removeHandler(handler)
    assert len(records) == len(g["lines"])
    for ours, ref in zip(records, [str(s) for s in g["lines"]]):
        # identical text up to the last printed digits of the numbers
        assert re.sub(r"[-+]?\d\.\d+e[-+]\d+", "#", ours) == re.sub(r"[-+]?\d\.\d+e[-+]\d+", "#", ref), (ours, ref)
        a = [float(v) for v in re.findall(r"[-+]?\d\.\d+e[-+]\d+", ours)]
        b = [float(v) for v in re.findall(r"[-+]?\d\.\d+e[-+]\d+", ref)]
        assert len(a) == len(b) and all(abs(u - v) <= 3e-7 * max(abs(v), 1e-30) + 1e-13 for u, v in zip(a, b)), (ours, ref)


# ---------------------------------------------------------------- data-parallel equivalence on one GPU
def test_sharded_gradients_equal_full_batch():
    """What GradAllReducer computes at N ranks — the mean over ranks of the per-shard mean-loss gradients — equals the
    single-rank gradient of the concatenated batch (dctn/training.py:78 averages the loss over the local batch).
    Emulated on one GPU: 4 shards of 8 images against one batch of 32, config-2 specs on the tcgen05 kernels."""
    from dctn_b200.eps_plus_linear import EPSesPlusLinear, UnitTheoreticalOutputStd

    torch.manual_seed(12)
    model = EPSesPlusLinear(((4, 4), (3, 6)), UnitTheoreticalOutputStd(), 1.0, torch.device(DEV), torch.float32, image_size=28)
    x = O.phi_cos_sin_squared(torch.rand(32, 28, 28, dtype=torch.float64), 0.72823).float().to(DEV)
    y = torch.randint(0, 10, (32,)).to(DEV)
    F.cross_entropy(model(x), y).backward()
    full = [p.grad.clone() for p in model.parameters()]
    acc = [torch.zeros_like(p) for p in model.parameters()]
    for r in range(4):
        model.zero_grad(set_to_none=True)
        F.cross_entropy(model(x[:, 8 * r : 8 * r + 8]), y[8 * r : 8 * r + 8]).backward()
        for a, p in zip(acc, model.parameters()):
            a += p.grad / 4
    for a, f in zip(acc, full):
        assert rel_err(a, f) <= 1e-5
