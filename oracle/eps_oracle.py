"""CPU restatement of the reference EPS path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Every function cites the reference file:line it follows.  Plain PyTorch CPU ops only
(the reference's own arithmetic is PyTorch's; opt_einsum merely orders the pairwise
contractions), any dtype; parity tests use float64.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Sequence, Tuple

import torch
from torch import Tensor


# ----------------------------------------------------------------------------- align
def align_views(x: Tensor, kernel_size: int) -> List[Tensor]:
    """dctn/align.py:11-46.  Yields K*K*C views of shape (B, H', W', Q) in the order
    (dh, dw) row-major, then channel: factor j = (dh*K + dw)*C + c."""
    C, B, H, W, Q = x.shape
    K = kernel_size
    Ho, Wo = H - K + 1, W - K + 1
    views = []
    for dh in range(K):
        for dw in range(K):
            for c in range(C):
                views.append(x[c][:, dh : dh + Ho, dw : dw + Wo])
    return views


def infer_kernel_size(core: Tensor, x: Tensor) -> int:
    """dctn/eps.py:20-22."""
    C, B, H, W, Q = x.shape
    K = math.isqrt((core.ndim - 1) // C)
    assert core.shape[:-1] == tuple(Q for _ in range(K * K * C))
    return K


def khatri_rao(factors: Sequence[Tensor]) -> Tensor:
    """Row-wise Kronecker (Khatri-Rao) product of (P, Q) matrices, first factor slowest.
    This is what the reference's einsum over ("batch","height","width","in0".."in{k}")
    materialises for each half (dctn/eps.py:25-27)."""
    out = factors[0]
    for f in factors[1:]:
        out = (out.unsqueeze(2) * f.unsqueeze(1)).reshape(out.shape[0], -1)
    return out


# ----------------------------------------------------------------------------- eps
def eps_4step(core: Tensor, x: Tensor) -> Tensor:
    """dctn/eps.py:19-40 with its explicit 4-step contraction path:
    (1) Khatri-Rao of the first ceil(n/2) aligned factors, (2) Khatri-Rao of the rest,
    (3) GEMM of half-1 with the core, (4) batched dot with half-2."""
    K = infer_kernel_size(core, x)
    C, B, H, W, Q = x.shape
    Ho, Wo = H - K + 1, W - K + 1
    O = core.shape[-1]
    views = [v.reshape(-1, Q) for v in align_views(x, K)]
    n = len(views)
    m = math.ceil(n / 2)
    kr1 = khatri_rao(views[:m])  # (P, Q^m)
    if n - m > 0:
        kr2 = khatri_rao(views[m:])  # (P, Q^(n-m))
    else:
        kr2 = torch.ones(kr1.shape[0], 1, dtype=x.dtype)
    A, Bn = kr1.shape[1], kr2.shape[1]
    t = kr1 @ core.reshape(A, Bn * O)  # (P, Bn*O)
    out = torch.einsum("pb,pbo->po", kr2, t.reshape(-1, Bn, O))
    return out.reshape(B, Ho, Wo, O)


def eps_one_by_one(core: Tensor, x: Tensor) -> Tensor:
    """dctn/eps.py:43-63 — contract the core with one aligned factor at a time."""
    K = infer_kernel_size(core, x)
    inter = None
    for v in align_views(x, K):
        if inter is None:
            inter = torch.einsum("bhwi,i...->bhw...", v, core)
        else:
            inter = torch.einsum("bhwi,bhwi...->bhw...", v, inter)
    return inter


def eps_dense(core: Tensor, x: Tensor) -> Tensor:
    """Third, independent formulation: full Khatri-Rao (P x Q^n) @ core (Q^n x O).
    Small shapes only."""
    K = infer_kernel_size(core, x)
    C, B, H, W, Q = x.shape
    views = [v.reshape(-1, Q) for v in align_views(x, K)]
    kr = khatri_rao(views)
    O = core.shape[-1]
    return (kr @ core.reshape(-1, O)).reshape(B, H - K + 1, W - K + 1, O)


def eps_grads(core: Tensor, x: Tensor, gout: Tensor) -> Tuple[Tensor, Tensor]:
    """Gradients of eps_4step w.r.t. (core, x) by autograd on the restatement
    (the reference has no explicit backward: it is autograd through eps.py:31-40)."""
    core = core.detach().clone().requires_grad_(True)
    x = x.detach().clone().requires_grad_(True)
    out = eps_4step(core, x)
    out.backward(gout)
    return core.grad, x.grad


# ----------------------------------------------------------------------------- stacking
def contract_with_input(epses: Sequence[Tensor], x: Tensor) -> Tensor:
    """dctn/epses_composition.py:133-141."""
    inter = x
    for core in epses[:-1]:
        inter = eps_4step(core, inter).unsqueeze(0)  # "b h w q -> () b h w q"
    return eps_4step(epses[-1], inter)


def eps_plus_linear_forward(
    epses: Sequence[Tensor], weight: Tensor, bias: Tensor, x: Tensor
) -> Tensor:
    """dctn/eps_plus_linear.py:138-147 in eval mode / p == 1 (no core dropout)."""
    inter = contract_with_input(epses, x)
    flat = inter.reshape(inter.shape[0], -1)  # "b h w q -> b (h w q)"
    return flat @ weight.T + bias


def phi_cos_sin_squared(u: Tensor, nu: float = 1.0) -> Tensor:
    """dctn/dataset_loading.py:33-36 (order sin^2 then cos^2), scaled by nu
    (new_runner.py:358-361), returned as (1, N, H, W, 2)."""
    s = 2 * nu * torch.sin(u * math.pi / 2.0) ** 2
    c = 2 * nu * torch.cos(u * math.pi / 2.0) ** 2
    return torch.stack((s, c), dim=-1).unsqueeze(0)


# ----------------------------------------------------------------------------- statistics / empirical-std init
def window_stats(x: Tensor, kernel_size: int) -> Tuple[Tensor, Tensor, int, int]:
    """make_windows (dctn/align.py:49-61) + RankOneTensorsBatch (dctn/rank_one_tensor.py:53-98): for every K x K window
    seen as a rank-one tensor of K*K*C factors, returns (sum of all elements of all windows, squared Frobenius norm of
    the whole batch, number of windows, elements per window) via the rank-one identities (no expansion)."""
    views = align_views(x, kernel_size)                       # K*K*C views (B, H', W', Q)
    sums = torch.stack([v.sum(dim=-1) for v in views]).prod(dim=0)
    sqn = torch.stack([(v ** 2).sum(dim=-1) for v in views]).prod(dim=0)
    return sums.sum(), sqn.sum(), sums.numel(), x.shape[-1] ** len(views)


def window_mean_var(x: Tensor, kernel_size: int, unbiased: bool = True) -> Tuple[Tensor, Tensor]:
    """mean_over_batch / var_over_batch of dctn/rank_one_tensor.py:66-106."""
    total, sqn, ntensors, ncoord = window_stats(x, kernel_size)
    nelement = ntensors * ncoord
    mean = total / nelement
    divisor = nelement - 1 if unbiased else nelement
    return mean, sqn / divisor - 2 * total / divisor * mean + nelement / divisor * mean ** 2


def transform_in_slices(core: Tensor, x: Tensor, batch_size: int) -> Tensor:
    """dctn/eps.py:126-137."""
    return torch.cat([eps_4step(core, piece) for piece in x.split(batch_size, dim=1)]).unsqueeze(0)


def empirical_std_cores(specs: Sequence[Tuple[int, int]], x: Tensor, batch_size: int) -> List[Tensor]:
    """dctn/epses_composition.py:91-105 over dctn/eps.py:163-181: every layer's randn core (drawn from torch's global
    generator, in layer order, on the CPU — as the reference does) divided by the biased std of its output on the current
    representation of x, which is then pushed through the rescaled core."""
    cores = []
    for kernel_size, out_size in specs:
        C, _, _, _, Q = x.shape
        core = torch.randn(*(Q,) * (kernel_size ** 2 * C), out_size, dtype=x.dtype)
        core = core * transform_in_slices(core, x, batch_size).std(unbiased=False) ** -1
        x = transform_in_slices(core, x, batch_size)
        cores.append(core)
    return cores


# ----------------------------------------------------------------------------- regulariser
def contract_on_input_dims(a: Tensor, b: Tensor) -> Tensor:
    """dctn/eps.py:106-112."""
    return a.reshape(-1, a.shape[-1]).T @ b.reshape(-1, b.shape[-1])


def composition_inner_product(epses1: Sequence[Tensor], epses2: Sequence[Tensor]) -> Tensor:
    """dctn/epses_composition.py:21-58."""
    epses1, epses2 = tuple(epses1), tuple(epses2)
    assert len(epses1) == len(epses2)
    if len(epses1) == 1:
        return torch.dot(epses1[0].reshape(-1), epses2[0].reshape(-1))
    a, b = epses1[:2]
    k = epses2[0]
    xm = contract_on_input_dims(a, k)  # (out of a, out of k)
    new_d = b
    nin = b.ndim - 1
    for i in range(nin):  # n-fold mode product: in_i (size out_a) -> newin_i (size out_k)
        new_d = torch.tensordot(new_d, xm, dims=([0], [0]))  # contracted mode leaves, new mode appended
    # after nin rotations the layout is (out, newin0..newin{n-1}); move out to the back
    new_d = new_d.movedim(0, -1)
    return composition_inner_product((new_d,) + epses1[2:], epses2[1:])


def epswise_squared_fro_norm(epses: Sequence[Tensor]) -> Tensor:
    """dctn/epses_composition.py:144-146."""
    return sum((c.reshape(-1) ** 2).sum() for c in epses)


# ----------------------------------------------------------------------------- logmatmulexp
def logmatmulexp(log_A: Tensor, log_B: Tensor) -> Tensor:
    """dctn/logmatmulexp.py:5-14: logsumexp_r(log_A[t, r] + log_B[r, i])."""
    T, R = log_A.shape
    assert log_B.shape[0] == R
    return torch.logsumexp(log_A.unsqueeze(2) + log_B.unsqueeze(0), dim=1)


def logmatmulexp_grads(log_A: Tensor, log_B: Tensor, gout: Tensor) -> Tuple[Tensor, Tensor]:
    a = log_A.detach().clone().requires_grad_(True)
    b = log_B.detach().clone().requires_grad_(True)
    logmatmulexp(a, b).backward(gout)
    return a.grad, b.grad


def logmatmulexp_batched(log_A: Tensor, log_B: Tensor) -> Tensor:
    """One dctn/logmatmulexp.py:5-14 product per leading index: (NB, T, R) x (NB, R, I) -> (NB, T, I)."""
    return torch.logsumexp(log_A.unsqueeze(3) + log_B.unsqueeze(1), dim=2)


# ----------------------------------------------------------------------------- ConvSBS (linear and log space)
def conv_sbs_forward(cores: Sequence[Tensor], positions: Sequence[Tuple[int, int]], x: Tensor) -> Tensor:
    """dctn/conv_sbs.py:258-304 in LINEAR space.  cores[c]: (O, L, R, Q, ..., Q) (C input dims); x: (C, B, H, W, Q).
    Per core: bond matrices M_c[b,h,w,o,l,r] (:269-281); then the ring product over bonds, keeping every core's
    out_quantum dim, traced over the closing bond (:282-303); out dims flattened in core order (:304)."""
    C, B, H, W, Q = x.shape
    hs = [p[0] for p in positions]
    ws = [p[1] for p in positions]
    Ho, Wo = H - max(hs), W - max(ws)
    P = B * Ho * Wo
    T = None  # (P, L0, Otot, R_c)
    for core, (ph, pw) in zip(cores, positions):
        O, L, R = core.shape[:3]
        kr = khatri_rao([x[c][:, ph : ph + Ho, pw : pw + Wo].reshape(P, Q) for c in range(C)])  # (P, Q^C)
        M = torch.einsum("pi,olri->polr", kr, core.reshape(O, L, R, -1))
        T = M.permute(0, 2, 1, 3) if T is None else torch.einsum("pxyl,polr->pxyor", T, M).reshape(P, T.shape[1], -1, R)
    return torch.einsum("pxyx->py", T).reshape(B, Ho, Wo, -1)


def conv_sbs_log_forward(log_cores: Sequence[Tensor], positions: Sequence[Tuple[int, int]], log_x: Tensor) -> Tensor:
    """log(conv_sbs_forward(exp(log_cores), positions, exp(log_x))) computed with logsumexp only."""
    C, B, H, W, Q = log_x.shape
    hs = [p[0] for p in positions]
    ws = [p[1] for p in positions]
    Ho, Wo = H - max(hs), W - max(ws)
    P = B * Ho * Wo
    T = None  # (P, L0*Otot, R_c)
    L0 = log_cores[0].shape[1]
    for core, (ph, pw) in zip(log_cores, positions):
        O, L, R = core.shape[:3]
        kr = log_x[0][:, ph : ph + Ho, pw : pw + Wo].reshape(P, Q)
        for c in range(1, C):
            kr = (kr.unsqueeze(2) + log_x[c][:, ph : ph + Ho, pw : pw + Wo].reshape(P, 1, Q)).reshape(P, -1)
        M = logmatmulexp(kr, core.reshape(O, L, R, -1).permute(3, 1, 0, 2).reshape(-1, L * O * R)).reshape(P, L, O * R)
        T = M.reshape(P, L * O, R) if T is None else logmatmulexp_batched(T, M).reshape(P, -1, R)
    T = T.reshape(P, L0, -1, L0)
    return torch.logsumexp(torch.diagonal(T, dim1=1, dim2=3), dim=-1).reshape(B, Ho, Wo, -1)


# ----------------------------------------------------------------------------- helpers
def rel_err(a: Tensor, b: Tensor) -> float:
    """Frobenius-relative error ||a-b|| / ||b|| in float64."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = b.norm().item()
    num = (a - b).norm().item()
    return num / den if den > 0 else num
