#!/usr/bin/env python
"""Measures the SIGNED relative bias of the tcgen05 kernels against the float64 oracle:
    alpha - 1,  alpha = <ours, ref> / <ref, ref>     (least-squares scale of ours onto ref)
next to the Frobenius-relative error, per call (forward, core gradient, input gradient), for layers of different
accumulation depth and for random-sign vs all-positive operands.  The tensor core truncates its fp32 accumulator toward
zero on every MMA, which shows as alpha < 1 growing with the number of accumulation steps (K / 16 per column tile).
    python tools/bias_probe.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dctn_b200 import eps as E
from oracle import eps_oracle as O
from oracle.eps_oracle import rel_err

dev = torch.device("cuda:0")
# (B, H, Q, K, O)
SHAPES = [(8, 28, 2, 4, 4), (8, 25, 4, 3, 6), (9, 25, 4, 3, 12), (5, 32, 12, 2, 24), (5, 32, 23, 2, 24), (30, 16, 2, 4, 3)]


def alpha(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return (a @ b / (b @ b)).item() - 1.0


for positive in (False, True):
    for (B, H, Q, K, Oq) in SHAPES:
        gen = torch.Generator().manual_seed(5)
        n = K * K
        x = (torch.rand(1, B, H, H, Q, generator=gen, dtype=torch.float64) * 1.2 + 0.2).float()
        if not positive:
            x = x * torch.sign(torch.randn(x.shape, generator=gen)).float()
        core = (torch.randn(*(Q,) * n, Oq, generator=gen, dtype=torch.float64) * Q ** (-n / 2)).float()
        gout = torch.randn(B, H - K + 1, H - K + 1, Oq, generator=gen, dtype=torch.float64).float()
        if positive:
            core, gout = core.abs(), gout.abs()
        c = core.to(dev).requires_grad_(True)
        xd = x.to(dev).requires_grad_(True)
        fam = E.kernel_families(c, xd)
        out = E.eps(c, xd)
        out.backward(gout.to(dev))
        want = O.eps_4step(core.double(), x.double())
        wdc, wdx = O.eps_grads(core.double(), x.double(), gout.double())
        m = (n + 1) // 2
        A, Bn = Q ** m, Q ** (n - m)
        print(f"{'pos ' if positive else 'rand'} K={K} Q={Q} O={Oq} P={B*(H-K+1)**2} fam={list(fam.values())} steps fwd {A/16:.0f} dx {Bn*Oq/16:.0f} | "
              f"fwd err {rel_err(out, want):.2e} bias {alpha(out, want):+.2e} | dcore err {rel_err(c.grad, wdc):.2e} bias {alpha(c.grad, wdc):+.2e} | "
              f"dx err {rel_err(xd.grad, wdx):.2e} bias {alpha(xd.grad, wdx):+.2e}", flush=True)
