"""Log-space matrix product — API of dctn/logmatmulexp.py:5-22.

``logmatmulexp(log_A, log_B)`` = log(exp(log_A) @ exp(log_B)) for 2-D inputs, computed by a max-shifted
online-logsumexp CUDA kernel.  Neither forward nor backward materialises the (Theta, R, I) tensor the
reference builds, so ``logmatmulexp_lowmem`` (checkpointing in the reference) is the same function.
"""
import torch
from torch import Tensor

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.float64: _lib.F64}


# dB of a TALL product (Theta = one row per ConvSBS window, R and I small) is a reduction over Theta that the 2-D backward
# kernel walks serially with R*I/256 CTAs.  Such products are cut into chunks of rows and handed to the batched kernel
# (one chunk per batch element, log_B repeated): it yields dA directly and one dB partial per chunk, summed in fixed order.
_TALL_THETA = 16384


def _tall_chunk(R: int, I: int, es: int) -> int:
    """Rows per chunk so that one chunk fits the batched backward kernel's shared memory (96 KiB)."""
    budget = 90 * 1024 // es - R * I
    return max(0, min(128, budget // (R + 2 * I)))


def _tall_backward(a: Tensor, b: Tensor, out: Tensor, gout: Tensor, need_dA: bool):
    theta, R = a.shape
    I = b.shape[1]
    tc = _tall_chunk(R, I, a.element_size())
    nchunk = (theta + tc - 1) // tc
    pad = nchunk * tc - theta

    def chunks(t, width):
        if pad:  # padding rows carry gout = 0: they add nothing to dB and their dA rows are dropped
            t = torch.cat([t, t.new_zeros(pad, width)])
        return t.reshape(nchunk, tc, width)

    a3, o3, g3 = chunks(a, R), chunks(out, I), chunks(gout, I)
    b3 = b.unsqueeze(0).expand(nchunk, R, I).contiguous()
    dA3 = torch.empty_like(a3) if need_dA else None
    dB3 = torch.empty_like(b3)
    with torch.cuda.device(a.device):
        rc = _lib.lib().dctn_logmatmulexp_batched_backward(
            a3.data_ptr(), b3.data_ptr(), o3.data_ptr(), g3.data_ptr(),
            dA3.data_ptr() if need_dA else None, dB3.data_ptr(), nchunk, tc, R, I, _DTYPES[a.dtype],
            torch.cuda.current_stream().cuda_stream,
        )
    _lib.check(rc, "dctn_logmatmulexp_batched_backward (tall product)")
    dA = dA3.reshape(-1, R)[:theta] if need_dA else None
    return dA, dB3.sum(dim=0)


class _LogMatMulExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_A: Tensor, log_B: Tensor) -> Tensor:
        theta, R = log_A.shape
        I = log_B.shape[1]
        a = log_A.detach().contiguous()
        b = log_B.detach().contiguous()
        out = torch.empty((theta, I), dtype=a.dtype, device=a.device)
        lib = _lib.lib()
        nws = lib.dctn_logmatmulexp_workspace_bytes(theta, R, I, _DTYPES[a.dtype])
        with torch.cuda.device(a.device):
            stream = torch.cuda.current_stream().cuda_stream
            if nws:   # one fused kernel: N^2 exponentials + a matrix product, per-element fallback inside (logmatmulexp_tile.cu)
                ws = torch.empty(nws, dtype=torch.uint8, device=a.device)
                rc = lib.dctn_logmatmulexp_forward_ws(a.data_ptr(), b.data_ptr(), out.data_ptr(), theta, R, I, _DTYPES[a.dtype],
                                                      ws.data_ptr(), ws.numel(), stream)
                _lib.check(rc, "dctn_logmatmulexp_forward_ws")
                ctx.save_for_backward(a, b, out, ws)
                return out
            rc = lib.dctn_logmatmulexp_forward(a.data_ptr(), b.data_ptr(), out.data_ptr(), theta, R, I, _DTYPES[a.dtype], stream)
        _lib.check(rc, "dctn_logmatmulexp_forward")
        ctx.save_for_backward(a, b, out)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout: Tensor):
        a, b, out, *rest = ctx.saved_tensors
        theta, R = a.shape
        I = b.shape[1]
        gout = gout.contiguous()
        if rest:
            ws = rest[0]
            dA = torch.empty_like(a) if ctx.needs_input_grad[0] else None
            dB = torch.empty_like(b) if ctx.needs_input_grad[1] else None
            with torch.cuda.device(a.device):
                rc = _lib.lib().dctn_logmatmulexp_backward_ws(
                    a.data_ptr(), b.data_ptr(), out.data_ptr(), gout.data_ptr(),
                    dA.data_ptr() if dA is not None else None, dB.data_ptr() if dB is not None else None,
                    theta, R, I, _DTYPES[a.dtype], ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream,
                )
            _lib.check(rc, "dctn_logmatmulexp_backward_ws")
            return dA, dB
        if theta >= _TALL_THETA and ctx.needs_input_grad[1] and _tall_chunk(R, I, a.element_size()) >= 16:
            return _tall_backward(a, b, out, gout, ctx.needs_input_grad[0])
        dA = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        dB = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(a.device):
            rc = _lib.lib().dctn_logmatmulexp_backward(
                a.data_ptr(), b.data_ptr(), out.data_ptr(), gout.data_ptr(),
                dA.data_ptr() if dA is not None else None, dB.data_ptr() if dB is not None else None,
                theta, R, I, _DTYPES[a.dtype], torch.cuda.current_stream().cuda_stream,
            )
        _lib.check(rc, "dctn_logmatmulexp_backward")
        return dA, dB


class _LogMatMulExpBatched(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_A: Tensor, log_B: Tensor) -> Tensor:
        nb, theta, R = log_A.shape
        I = log_B.shape[2]
        a = log_A.detach().contiguous()
        b = log_B.detach().contiguous()
        out = torch.empty((nb, theta, I), dtype=a.dtype, device=a.device)
        with torch.cuda.device(a.device):
            rc = _lib.lib().dctn_logmatmulexp_batched_forward(
                a.data_ptr(), b.data_ptr(), out.data_ptr(), nb, theta, R, I, _DTYPES[a.dtype],
                torch.cuda.current_stream().cuda_stream,
            )
        _lib.check(rc, "dctn_logmatmulexp_batched_forward")
        ctx.save_for_backward(a, b, out)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout: Tensor):
        a, b, out = ctx.saved_tensors
        nb, theta, R = a.shape
        I = b.shape[2]
        gout = gout.contiguous()
        dA = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        dB = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(a.device):
            rc = _lib.lib().dctn_logmatmulexp_batched_backward(
                a.data_ptr(), b.data_ptr(), out.data_ptr(), gout.data_ptr(),
                dA.data_ptr() if dA is not None else None, dB.data_ptr() if dB is not None else None,
                nb, theta, R, I, _DTYPES[a.dtype], torch.cuda.current_stream().cuda_stream,
            )
        _lib.check(rc, "dctn_logmatmulexp_batched_backward")
        return dA, dB


def logmatmulexp_batched(log_A: Tensor, log_B: Tensor, /) -> Tensor:
    """log_A: (..., Theta, R), log_B: (..., R, I) with equal leading dims -> (..., Theta, I), one log-space product
    per leading index.  ADDITIONAL entry (the reference's logmatmulexp is 2-D only, dctn/logmatmulexp.py:8-10): the
    ring step of a ConvSBS contraction in log space (dctn/conv_sbs.py:282-303 does it in linear space)."""
    assert log_A.ndim >= 3 and log_A.ndim == log_B.ndim
    assert log_A.shape[:-2] == log_B.shape[:-2] and log_A.shape[-1] == log_B.shape[-2]
    if not (log_A.is_cuda and log_B.is_cuda):
        raise RuntimeError("dctn_b200.logmatmulexp_batched runs on CUDA tensors only (no CPU fallback)")
    if log_A.dtype not in _DTYPES or log_A.dtype != log_B.dtype:
        raise TypeError(f"logmatmulexp_batched supports matching float32/float64 inputs, got {log_A.dtype} and {log_B.dtype}")
    lead = log_A.shape[:-2]
    out = _LogMatMulExpBatched.apply(log_A.reshape(-1, *log_A.shape[-2:]), log_B.reshape(-1, *log_B.shape[-2:]))
    return out.reshape(*lead, log_A.shape[-2], log_B.shape[-1])


def logmatmulexp(log_A: Tensor, log_B: Tensor, /) -> Tensor:
    """log_A: Theta x R, log_B: R x I -> (log_A.exp() @ log_B.exp()).log(), stable forward and backward."""
    theta, R = log_A.shape  # ValueError for non 2-D input, like the reference's unpacking
    I = log_B.shape[1]
    assert log_B.shape == (R, I)
    if not (log_A.is_cuda and log_B.is_cuda):
        raise RuntimeError("dctn_b200.logmatmulexp runs on CUDA tensors only (no CPU fallback)")
    if log_A.dtype not in _DTYPES or log_A.dtype != log_B.dtype:
        raise TypeError(f"logmatmulexp supports matching float32/float64 inputs, got {log_A.dtype} and {log_B.dtype}")
    if theta == 0 or I == 0 or R == 0:
        # empty operands, as the reference handles them: no rows / columns -> empty result; an empty sum is log 0 = -inf
        # (torch.logsumexp over an empty dimension); the expression keeps the result attached to the autograd graph
        return (log_A.sum(dim=1, keepdim=True) + log_B.sum(dim=0, keepdim=True)) * 0 + (float("-inf") if R == 0 else 0.0)
    return _LogMatMulExp.apply(log_A, log_B)


def logmatmulexp_lowmem(log_A: Tensor, log_B: Tensor, /) -> Tensor:
    """Same as logmatmulexp; the CUDA path never saves a (Theta, R, I)-shaped tensor."""
    return logmatmulexp(log_A, log_B)
