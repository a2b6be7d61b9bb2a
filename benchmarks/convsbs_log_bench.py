#!/usr/bin/env python
"""Config 5, batched extension (BASELINE.json: "28x28 synthetic input, batch 2048"): the ring of
bond matrices of a 3x3 snake ConvSBS (dctn/conv_sbs.py:282-303 multiplies this ring in linear space) contracted in LOG
space — a chain of 9 r x r products per window, NB = 2048 * 26 * 26 windows, r in {4, 8, 12}, float32.

Reports, per r: forward and forward+backward time of `reduce(logmatmulexp_batched, 9 matrices)` (CUDA events), exponentials
per second against the MUFU rate, algorithmic HBM bytes per second, the same chain in the reference's formulation
(materialised broadcast sum + torch.logsumexp, dctn/logmatmulexp.py:5-14 applied per element) on the same GPU, and
optionally on the host CPU over a bounded sample of windows."""
import argparse
import json
import os
import sys
import time
from functools import reduce

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dctn_b200.logmatmulexp import logmatmulexp_batched  # noqa: E402


def ref_lme_batched(a, b):  # the reference formulation, one product per leading index
    return torch.logsumexp(a.unsqueeze(3) + b.unsqueeze(1), dim=2)


def time_cuda(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bonds", default="4,8,12")
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--cores", type=int, default=9)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--cpu-windows", type=int, default=0, help="also time the reference formulation on this many windows on the host")
    ap.add_argument("--json", default="")
    ap.add_argument("--full", action="store_true", help="also time conv_sbs_log_forward end to end (bond matrices from log x and log cores + ring) on a 3x3 snake string")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    NB, S = args.batch * 26 * 26, args.cores
    rows = []
    for r in map(int, args.bonds.split(",")):
        torch.manual_seed(0)
        mats = [torch.randn(NB, r, r, device=dev) for _ in range(S)]
        og = torch.ones(NB, r, r, device=dev)

        def fwd(f=logmatmulexp_batched):
            with torch.no_grad():
                return reduce(f, mats)

        def fwdbwd(f=logmatmulexp_batched):
            for m in mats:
                m.grad = None
            reduce(f, mats).backward(og)

        t_f = time_cuda(fwd, args.iters)
        t_rf = time_cuda(lambda: fwd(ref_lme_batched), max(1, args.iters // 2))
        for m in mats:
            m.requires_grad_(True)
        t_fb = time_cuda(fwdbwd, args.iters)
        try:
            t_rfb = time_cuda(lambda: fwdbwd(ref_lme_batched), max(1, args.iters // 2))
        except torch.OutOfMemoryError:  # autograd keeps 8 materialised (NB, r, r, r) tensors
            t_rfb = None
            torch.cuda.empty_cache()
        exps_f = (S - 1) * NB * r ** 3
        bytes_f = (S - 1) * NB * 3 * r * r * 4
        row = dict(bond=r, windows=NB, cores=S, dtype="float32", ours_fwd_ms=t_f, ours_fwdbwd_ms=t_fb,
                   ref_formulation_gpu_fwd_ms=t_rf, ref_formulation_gpu_fwdbwd_ms=t_rfb,
                   fwd_gexp_per_s=exps_f / t_f / 1e6, fwd_algorithmic_gb_per_s=bytes_f / t_f / 1e6,
                   fwdbwd_gexp_per_s=3 * exps_f / t_fb / 1e6, windows_per_s_fwdbwd=NB / t_fb * 1e3)
        if args.cpu_windows:
            n = args.cpu_windows
            cm = [m.detach()[:n].cpu().requires_grad_(True) for m in mats]
            reduce(ref_lme_batched, cm).backward(og[:n].cpu())
            t0 = time.perf_counter()
            reduce(ref_lme_batched, cm).backward(og[:n].cpu())
            dt = time.perf_counter() - t0
            row.update(ref_cpu_fwdbwd_windows_per_s=n / dt, cpu_sample_windows=n, cpu_threads=torch.get_num_threads())
        for m in mats:
            m.requires_grad_(False)
        rows.append(row)
        print(json.dumps(row), flush=True)
        del mats, og
        torch.cuda.empty_cache()
    if args.full:
        # BASELINE config 5 as worded: 28x28 synthetic input, batch 2048, through the public entry.  3x3 snake string, Q = 2,
        # C = 1, one output core (O = 4), closed ring; the linear-space reference formulation (dctn/conv_sbs.py:258-304
        # restated with einsum) on the same GPU beside it.
        from dctn_b200.conv_sbs_log import conv_sbs_log_forward
        from dctn_b200.pos2d import Pos2D

        snake = [(0, 0), (0, 1), (0, 2), (1, 2), (1, 1), (1, 0), (2, 0), (2, 1), (2, 2)]
        for r in map(int, args.bonds.split(",")):
            torch.manual_seed(1)
            outs = [1, 1, 1, 1, 4, 1, 1, 1, 1]
            log_cores = [(0.3 * torch.randn(o, r, r, 2, device=dev)).requires_grad_(True) for o in outs]
            log_x = (0.5 * torch.randn(1, args.batch, 28, 28, 2, device=dev)).requires_grad_(True)
            pos = tuple(Pos2D(*p) for p in snake)

            def lin_forward():
                x = log_x.exp()
                T = None
                P = args.batch * 26 * 26
                for core, (ph, pw) in zip(log_cores, snake):
                    M = torch.einsum("pi,olri->polr", x[0][:, ph:ph + 26, pw:pw + 26].reshape(P, 2), core.exp())
                    T = M.permute(0, 2, 1, 3) if T is None else torch.einsum("pxyl,polr->pxyor", T, M).reshape(P, T.shape[1], -1, M.shape[3])
                return torch.einsum("pxyx->py", T).log()

            def f_ours():
                with torch.no_grad():
                    return conv_sbs_log_forward(log_cores, pos, log_x)

            def fb_ours():
                for t in log_cores + [log_x]:
                    t.grad = None
                conv_sbs_log_forward(log_cores, pos, log_x).sum().backward()

            def f_lin():
                with torch.no_grad():
                    return lin_forward()

            def fb_lin():
                for t in log_cores + [log_x]:
                    t.grad = None
                lin_forward().sum().backward()

            err = (f_ours().reshape(-1, 4) - f_lin()).abs().max().item()
            row = dict(kind="conv_sbs_log_forward 3x3 snake", bond=r, batch=args.batch, image=28, windows=NB,
                       ours_fwd_ms=time_cuda(f_ours, args.iters), ours_fwdbwd_ms=time_cuda(fb_ours, args.iters),
                       linear_space_einsum_gpu_fwd_ms=time_cuda(f_lin, args.iters),
                       linear_space_einsum_gpu_fwdbwd_ms=time_cuda(fb_lin, args.iters), max_abs_diff_vs_linear_space=err)
            rows.append(row)
            print(json.dumps(row), flush=True)
            del log_cores, log_x
            torch.cuda.empty_cache()
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
