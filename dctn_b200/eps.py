"""EPS operator — same API as the reference's dctn/eps.py, contraction done by CUDA kernels.

``eps(core, input)`` contracts, for every K x K patch of ``input`` (C, B, H, W, Q_in), the rank-one product
of the K*K*C per-pixel vectors with ``core`` ((Q_in,)*(K*K*C) + (Q_out,)) and returns (B, H-K+1, W-K+1, Q_out)
(reference: dctn/eps.py:19-40).  It is differentiable w.r.t. both arguments through
:class:`EpsFunction`, which saves only ``(core, input)`` and calls the C ABI of libdctn_b200.so
(include/dctn_b200.h): no Q^(K*K)-sized intermediate is ever kept for backward.
"""
from __future__ import annotations

import logging
import math
import os
from typing import Dict, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.float64: _lib.F64}
_default_variant = _lib.VARIANTS.get(os.environ.get("DCTN_B200_VARIANT", "auto").lower(), _lib.VARIANT_AUTO)
_plans: Dict[Tuple[int, int, int, int, int, int], int] = {}
# Training forward keeps the forward GEMM's per-patch intermediate T (P x Q_in^floor(n/2) * Q_out floats) for the
# input gradient when it is at most this many bytes (B200: 180 GB of HBM3e; config 2, B=512: 1.66 GB); larger
# layers recompute it.  0 disables saving.
_save_limit_bytes = int(float(os.environ.get("DCTN_B200_SAVE_LIMIT_MB", "16384")) * (1 << 20))


def set_save_limit_mb(mb: float) -> None:
    global _save_limit_bytes
    _save_limit_bytes = int(mb * (1 << 20))



def set_default_variant(name: str) -> None:
    """Select the kernel family for subsequent calls: 'auto' | 'ffma' | 'tc3' | 'tc1' | 'direct'."""
    global _default_variant
    _default_variant = _lib.VARIANTS[name.lower()]


def get_default_variant() -> str:
    return {v: k for k, v in _lib.VARIANTS.items()}[_default_variant]


def _plan(C: int, K: int, Q: int, O: int, dtype: torch.dtype, variant: int) -> int:
    key = (C, K, Q, O, _DTYPES[dtype], variant)
    handle = _plans.get(key)
    if handle is None:
        handle = _lib.lib().dctn_eps_plan_get(*key)
        if not handle:
            raise RuntimeError(f"no EPS kernel plan for C={C} K={K} Q_in={Q} Q_out={O} {dtype}: {_lib.last_error()}")
        _plans[key] = handle
    return handle


def plan_description(core: Tensor, input: Tensor) -> str:
    C, K, Q, O = _infer(core, input)
    return _lib.lib().dctn_eps_plan_describe(_plan(C, K, Q, O, input.dtype, _default_variant)).decode()


def kernel_families(core: Tensor, input: Tensor) -> Dict[str, int]:
    """Which kernel family (``_lib.FAMILY_*``) serves the forward, core-gradient and input-gradient calls of this
    layer at this input size under the current default variant (dctn_eps_kernel_family)."""
    C, K, Q, O = _infer(core, input)
    _, B, H, W, _ = input.shape
    plan = _plan(C, K, Q, O, input.dtype, _default_variant)
    kinds = {"forward": _lib.WS_FORWARD, "backward_core": _lib.WS_BACKWARD_CORE, "backward_input": _lib.WS_BACKWARD_INPUT}
    return {name: _lib.lib().dctn_eps_kernel_family(plan, B, H, W, kind) for name, kind in kinds.items()}


def _infer(core: Tensor, input: Tensor) -> Tuple[int, int, int, int]:
    """Shape contract of dctn/eps.py:20-23 (AssertionError on mismatch, like the reference)."""
    num_channels, batch_size, height, width, in_size = input.shape
    kernel_size = math.isqrt((core.ndim - 1) // num_channels)
    assert core.shape[:-1] == tuple(in_size for _ in range(kernel_size ** 2 * num_channels))
    return num_channels, kernel_size, in_size, core.shape[-1]


def _check_device(core: Tensor, input: Tensor) -> None:
    if not (input.is_cuda and core.is_cuda):
        raise RuntimeError(
            "dctn_b200.eps runs on CUDA tensors only (no CPU fallback); got "
            f"core on {core.device}, input on {input.device}"
        )
    if core.device != input.device:
        raise RuntimeError(f"core ({core.device}) and input ({input.device}) must be on the same device")
    if input.dtype not in _DTYPES or core.dtype != input.dtype:
        raise TypeError(f"eps supports float32/float64 with matching dtypes, got core {core.dtype}, input {input.dtype}")


def _dense(t: Tensor) -> Tensor:
    """Detached, contiguous and 16-byte aligned (the kernels use 128-bit accesses): a copy only when needed."""
    t = t.detach().contiguous()
    return t.clone() if t.data_ptr() % 16 else t


# Scratch arenas: ONE buffer per (device, stream), grown to the largest request and reused by every call.  The library's
# workspace is pure scratch — dead when the call's kernels have run — and calls on one stream execute in order, so
# sharing it is safe; a per-call torch.empty of hundreds of MB instead makes the caching allocator split and re-cut its
# large blocks next to the long-lived saved intermediates (seen as cudaMalloc / cudaFree stalls in the middle of a step).
_arenas: Dict[Tuple[int, int], Tensor] = {}


def _workspace(plan: int, B: int, H: int, W: int, kind: int, device) -> Tensor:
    nbytes = _lib.lib().dctn_eps_workspace_bytes(plan, B, H, W, kind)
    if torch.cuda.is_current_stream_capturing():
        # graph capture: the allocation must come from the graph's private pool (replays reuse the address)
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
    arena = _arenas.get(key)
    if arena is None or arena.numel() < nbytes:
        arena = None
        _arenas.pop(key, None)                    # release the old arena before allocating the larger one
        arena = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _arenas[key] = arena
    return arena


def release_workspaces() -> None:
    """Drops the scratch arenas (they are re-created on demand)."""
    _arenas.clear()


class EpsFunction(torch.autograd.Function):
    """autograd node of one EPS layer (forward: K1, backward: K2 core-gradient + K3 input-gradient).

    Saves ``(core, input)`` and — only when ``input`` needs a gradient, the kernel family has one, and it fits the
    save limit — the forward GEMM's intermediate ``T`` (dctn_eps_forward_train), which halves the tensor-core work
    of the input gradient.  The reference's autograd keeps that tensor and two larger ones (dctn/eps.py:31-40)."""

    @staticmethod
    def forward(ctx, core: Tensor, input: Tensor, variant: int) -> Tensor:
        C, K, Q, O = _infer(core, input)
        _, B, H, W, _ = input.shape
        core_c = _dense(core)
        x_c = _dense(input)
        plan = _plan(C, K, Q, O, input.dtype, variant)
        out = torch.empty((B, H - K + 1, W - K + 1, O), dtype=input.dtype, device=input.device)
        lib = _lib.lib()
        saved = None
        with torch.cuda.device(input.device):
            ws = _workspace(plan, B, H, W, _lib.WS_FORWARD, input.device)
            stream = torch.cuda.current_stream().cuda_stream
            nsave = lib.dctn_eps_saved_bytes(plan, B, H, W) if ctx.needs_input_grad[1] else 0
            if 0 < nsave <= _save_limit_bytes:
                # T is a convenience (it halves the tensor-core work of the input gradient), never a requirement: when
                # the device cannot hold it next to the workspaces, fall back to the recompute path
                try:
                    saved = torch.empty(nsave, dtype=torch.uint8, device=input.device)
                except torch.cuda.OutOfMemoryError:
                    saved = None
            if saved is not None:
                rc = lib.dctn_eps_forward_train(
                    plan, x_c.data_ptr(), core_c.data_ptr(), out.data_ptr(), saved.data_ptr(), saved.numel(), B, H, W,
                    ws.data_ptr(), ws.numel(), stream,
                )
                _lib.check(rc, "dctn_eps_forward_train")
            else:
                rc = lib.dctn_eps_forward(
                    plan, x_c.data_ptr(), core_c.data_ptr(), out.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(), stream
                )
                _lib.check(rc, "dctn_eps_forward")
        if saved is None:
            ctx.save_for_backward(core_c, x_c)
        else:
            ctx.save_for_backward(core_c, x_c, saved)
        ctx.plan = plan
        ctx.core_shape = core.shape
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout: Tensor):
        core_c, x_c, *rest = ctx.saved_tensors
        saved = rest[0] if rest else None
        _, B, H, W, _ = x_c.shape
        gout = _dense(gout)
        dcore = dx = None
        lib = _lib.lib()
        with torch.cuda.device(x_c.device):
            stream = torch.cuda.current_stream().cuda_stream
            if ctx.needs_input_grad[0]:
                dcore = torch.empty_like(core_c)
                ws = _workspace(ctx.plan, B, H, W, _lib.WS_BACKWARD_CORE, x_c.device)
                rc = lib.dctn_eps_backward_core(
                    ctx.plan, x_c.data_ptr(), gout.data_ptr(), dcore.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(), stream
                )
                _lib.check(rc, "dctn_eps_backward_core")
                dcore = dcore.view(ctx.core_shape)
            if ctx.needs_input_grad[1]:
                dx = torch.empty_like(x_c)
                if saved is not None:
                    ws = _workspace(ctx.plan, B, H, W, _lib.WS_BACKWARD_INPUT_SAVED, x_c.device)
                    rc = lib.dctn_eps_backward_input_saved(
                        ctx.plan, x_c.data_ptr(), core_c.data_ptr(), gout.data_ptr(), saved.data_ptr(), saved.numel(),
                        dx.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(), stream,
                    )
                    _lib.check(rc, "dctn_eps_backward_input_saved")
                else:
                    ws = _workspace(ctx.plan, B, H, W, _lib.WS_BACKWARD_INPUT, x_c.device)
                    rc = lib.dctn_eps_backward_input(
                        ctx.plan, x_c.data_ptr(), core_c.data_ptr(), gout.data_ptr(), dx.data_ptr(), B, H, W,
                        ws.data_ptr(), ws.numel(), stream,
                    )
                    _lib.check(rc, "dctn_eps_backward_input")
        return dcore, dx, None


def eps(core: Tensor, input: Tensor) -> Tensor:
    """Drop-in for dctn/eps.py:19-40."""
    _infer(core, input)  # AssertionError first, as in the reference
    _check_device(core, input)
    return EpsFunction.apply(core, input, _default_variant)


class EpsFromPixelsFunction(torch.autograd.Function):
    """First layer on RAW pixels: eps(core, phi(pixels)) with the feature map evaluated inside the kernel
    (dctn_eps_forward_from_pixels).  Differentiable w.r.t. ``core`` only — pixels are data; the core gradient is the
    ordinary one on phi(pixels)."""

    @staticmethod
    def forward(ctx, core: Tensor, pixels: Tensor, scale: float, variant: int) -> Tensor:
        B, H, W = pixels.shape
        K = math.isqrt(core.ndim - 1)
        assert core.shape[:-1] == (2,) * (K * K), "the fused feature map has two components (Q_in = 2, one channel)"
        core_c, pix_c = _dense(core), _dense(pixels)
        plan = _plan(1, K, 2, core.shape[-1], pixels.dtype, variant)
        out = torch.empty((B, H - K + 1, W - K + 1, core.shape[-1]), dtype=pixels.dtype, device=pixels.device)
        with torch.cuda.device(pixels.device):
            rc = _lib.lib().dctn_eps_forward_from_pixels(
                plan, pix_c.data_ptr(), float(scale), core_c.data_ptr(), out.data_ptr(), B, H, W,
                torch.cuda.current_stream().cuda_stream,
            )
        _lib.check(rc, "dctn_eps_forward_from_pixels")
        ctx.save_for_backward(pix_c)
        ctx.plan, ctx.scale, ctx.core_shape = plan, float(scale), core.shape
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout: Tensor):
        (pix_c,) = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        B, H, W = pix_c.shape
        x = phi_sin_cos_squared(pix_c, ctx.scale)
        gout = _dense(gout)
        dcore = torch.empty(ctx.core_shape, dtype=pix_c.dtype, device=pix_c.device)
        with torch.cuda.device(pix_c.device):
            ws = _workspace(ctx.plan, B, H, W, _lib.WS_BACKWARD_CORE, pix_c.device)
            rc = _lib.lib().dctn_eps_backward_core(
                ctx.plan, x.data_ptr(), gout.data_ptr(), dcore.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(),
                torch.cuda.current_stream().cuda_stream,
            )
        _lib.check(rc, "dctn_eps_backward_core")
        return dcore, None, None, None


def phi_sin_cos_squared(pixels: Tensor, scale: float) -> Tensor:
    """The reference's feature map (dctn/dataset_loading.py:33-36): (B, H, W) pixels in [0, 1] ->
    (1, B, H, W, 2) = scale * (sin^2(pi u / 2), cos^2(pi u / 2))."""
    return torch.stack((scale * torch.sin(pixels * (math.pi / 2)) ** 2, scale * torch.cos(pixels * (math.pi / 2)) ** 2), dim=-1)[None]


def eps_from_pixels(core: Tensor, pixels: Tensor, scale: float = 1.0) -> Tensor:
    """``eps(core, phi_sin_cos_squared(pixels, scale))`` with the feature map fused into the kernel: one float per pixel
    is read instead of two.  Additional entry point (the reference applies phi in its data loader); K = 2, one channel,
    Q_in = 2 layers with a small core only — other shapes raise."""
    if not (pixels.is_cuda and core.is_cuda):
        raise RuntimeError("dctn_b200.eps_from_pixels runs on CUDA tensors only (no CPU fallback)")
    if pixels.ndim != 3 or pixels.dtype not in _DTYPES or core.dtype != pixels.dtype:
        raise TypeError("eps_from_pixels: pixels must be (B, H, W) float32/float64 with the core's dtype")
    return EpsFromPixelsFunction.apply(core, pixels, scale, _default_variant)


def eps_one_by_one(core: Tensor, input: Tensor) -> Tensor:
    """Drop-in for dctn/eps.py:43-63.  The reference contracts one aligned factor at a time; the result is
    the same tensor, so this routes to the same fused kernel."""
    out = eps(core, input)
    num_channels, batch_size, height, width, in_size = input.shape
    kernel_size = math.isqrt((core.ndim - 1) // num_channels)
    assert out.shape == (batch_size, height - kernel_size + 1, width - kernel_size + 1, core.shape[-1])
    return out


def calc_eps_shape(kernel_size: int, in_num_channels: int, in_size: int, out_size: int) -> Tuple[int, ...]:
    """Shape an EPS core with these parameters must have (dctn/eps.py:66-70)."""
    return (in_size,) * (kernel_size ** 2 * in_num_channels) + (out_size,)


spec_to_shape = calc_eps_shape  # dctn/eps.py:184-187 is the same function under another name


def total_in_dim_size(kernel_size: int, in_num_channels: int, in_size: int) -> int:
    return in_size ** (in_num_channels * kernel_size ** 2)


def is_eps(a: Tensor) -> bool:
    """Whether `a` can plausibly be an EPS core judging by its shape (dctn/eps.py:115-117)."""
    return a.ndim >= 2 and all(s == a.shape[0] for s in a.shape[:-1])


def matrix_shape(eps_core: Tensor) -> Tuple[int, int]:
    """(out_size, total input size) — dctn/eps.py:99-103."""
    assert is_eps(eps_core)
    return eps_core.shape[-1], math.prod(eps_core.shape[:-1])


def contract_on_input_dims(a: Tensor, b: Tensor) -> Tensor:
    """(out dim of a, out dim of b) Gram matrix over all input dims (dctn/eps.py:106-112)."""
    assert is_eps(a) and is_eps(b)
    return a.reshape(-1, a.shape[-1]).T @ b.reshape(-1, b.shape[-1])


def inner_product(a: Tensor, b: Tensor) -> Tensor:
    """Frobenius inner product of two cores of equal shape (dctn/eps.py:120-123)."""
    assert a.shape == b.shape
    assert is_eps(a)
    return torch.dot(a.reshape(-1), b.reshape(-1))


@torch.no_grad()
def transform_in_slices(eps_core: Tensor, x: Tensor, batch_size: int) -> Tensor:
    """Forward-only transform of a whole dataset x (C, N, H, W, Q_in) in slices of `batch_size` samples;
    returns (1, N, H', W', Q_out) (dctn/eps.py:126-137).  Slices of a C>1 tensor are non-contiguous;
    EpsFunction makes them contiguous."""
    assert is_eps(eps_core)
    outs = [eps(eps_core, piece) for piece in x.split(batch_size, dim=1)]
    return torch.cat(outs).unsqueeze(0)


def make_eps_unit_theoretical_output_std(
    kernel_size: int, in_num_channels: int, in_size: int, out_size: int, device: torch.device, dtype: torch.dtype
) -> Tensor:
    """randn core scaled by D^-1/2 so that the forward pass preserves the std (dctn/eps.py:144-160)."""
    std = total_in_dim_size(kernel_size, in_num_channels, in_size) ** -0.5
    logging.getLogger(f"{__name__}.make_eps_unit_theoretical_output_std").info(
        f"Multiplying the output of randn by {std:.30e}"
    )
    shape = calc_eps_shape(kernel_size, in_num_channels, in_size, out_size)
    return std * torch.randn(*shape, dtype=dtype).to(device)


@torch.no_grad()
def transform_in_slices_with_stats(eps_core: Tensor, x: Tensor, batch_size: int) -> Tuple[Tensor, Tensor]:
    """transform_in_slices that also returns the float64 pair (sum, sum of squares) of the whole output, reduced on the
    GPU right after each slice's forward (dctn_eps_forward_stats): every slice writes straight into its part of ONE
    preallocated output, so there is no torch.cat and no separate pass for the statistics."""
    assert is_eps(eps_core)
    _check_device(eps_core, x)
    C, K, Q, O = _infer(eps_core, x)
    _, N, H, W, _ = x.shape
    core_c = _dense(eps_core)
    plan = _plan(C, K, Q, O, x.dtype, _default_variant)
    out = torch.empty((N, H - K + 1, W - K + 1, O), dtype=x.dtype, device=x.device)
    stats = torch.zeros(2, dtype=torch.float64, device=x.device)
    lib = _lib.lib()
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream().cuda_stream
        ws = None
        for b0 in range(0, N, batch_size):
            piece = _dense(x[:, b0 : b0 + batch_size])
            B = piece.shape[1]
            need = lib.dctn_eps_workspace_bytes(plan, B, H, W, _lib.WS_FORWARD_STATS)
            if ws is None or ws.numel() < need:
                ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            dst = out[b0 : b0 + B]
            aligned = dst.data_ptr() % 16 == 0        # the kernels store 128 bits at a time
            tmp = dst if aligned else torch.empty_like(dst)
            rc = lib.dctn_eps_forward_stats(
                plan, piece.data_ptr(), core_c.data_ptr(), tmp.data_ptr(), stats.data_ptr(), B, H, W,
                ws.data_ptr(), ws.numel(), stream,
            )
            _lib.check(rc, "dctn_eps_forward_stats")
            if not aligned:
                dst.copy_(tmp)
    return out.unsqueeze(0), stats


def _empirical_std_core_and_output(
    kernel_size: int, out_size: int, input: Tensor, device: torch.device, dtype: torch.dtype, batch_size: int
) -> Tuple[Tensor, Tensor]:
    """The core of make_eps_unit_empirical_output_std and its output on `input`.  The contraction is linear in the core, so
    the output of the rescaled core is the unscaled output times the same factor: ONE pass over the dataset instead of the
    reference's two (dctn/eps.py:176 and dctn/epses_composition.py:103)."""
    num_channels, dataset_size, height, width, in_size = input.shape
    core = torch.randn(*(in_size,) * (kernel_size ** 2 * num_channels), out_size, dtype=dtype).to(device)
    output, stats = transform_in_slices_with_stats(core, input.to(device, dtype), batch_size)
    n = output.numel()
    mean = stats[0] / n
    inv_std = ((stats[1] / n - mean * mean) ** -0.5).to(dtype)      # biased std, as output.std(unbiased=False)
    logger = logging.getLogger(f"{__name__}.make_eps_unit_empirical_output_std")
    logger.info(f"Multiplying the output of randn by {inv_std:.30e}")
    core *= inv_std
    output *= inv_std
    logger.info(f"Initialized an EPS with empirical std = {core.std(unbiased=False):.30e}")
    return core, output


def make_eps_unit_empirical_output_std(
    kernel_size: int, out_size: int, input: Tensor, device: torch.device, dtype: torch.dtype, batch_size: int
) -> Tensor:
    """randn core rescaled so that its output on `input` has unit (biased) std (dctn/eps.py:163-181)."""
    return _empirical_std_core_and_output(kernel_size, out_size, input, device, dtype, batch_size)[0]


class EPS(nn.Module):
    """One EPS layer holding its ``core`` parameter (dctn/eps.py:73-96)."""

    def __init__(self, kernel_size: int, in_num_channels: int, in_size: int, out_size: int):
        super().__init__()
        self.kernel_size = kernel_size
        self.in_num_channels = in_num_channels
        self.in_size = in_size
        self.out_size = out_size
        self.core = nn.Parameter(
            make_eps_unit_theoretical_output_std(
                kernel_size, in_num_channels, in_size, out_size, torch.device("cpu"), torch.float32
            )
        )

    @property
    def matrix_shape(self) -> Tuple[int, int]:
        return matrix_shape(self.core)

    def forward(self, input: Tensor) -> Tensor:
        return eps(self.core, input)
