#!/usr/bin/env python
"""Config 5 (BASELINE.json): logmatmulexp chain — reduce(logmatmulexp, 6 N x N matrices), forward and forward+backward
with out_grad = ones, exactly the protocol of small_experiments/logmatmulexp_benchmark/benchmark.py:21-52.
Reference numbers (results.json): N=256 fp32 5.51 ms fwd / 11.08 ms fwd+bwd on its "Graphics Device", 303.8 / 517.4 ms on CPU.
Also times the reference formulation (materialised (N,N,N) tensor + torch.logsumexp) on the same GPU and on the host CPU."""
import argparse
import json
import os
import sys
import time
from functools import reduce

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dctn_b200.logmatmulexp import logmatmulexp  # noqa: E402


def ref_lme(a, b):  # dctn/logmatmulexp.py:5-14 restated (oracle formulation)
    return torch.logsumexp(a.unsqueeze(2) + b.unsqueeze(0), dim=1)


def bench(func, mats, out_grad, iters, cuda):
    def sync():
        if cuda:
            torch.cuda.synchronize()
    with torch.no_grad():
        reduce(func, mats)
    sync()
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(iters):
            reduce(func, mats)
    sync()
    fwd = (time.perf_counter() - t0) / iters
    mats[0].requires_grad_()
    reduce(func, mats).backward(out_grad)
    sync()
    t0 = time.perf_counter()
    for _ in range(iters):
        reduce(func, mats).backward(out_grad)
    sync()
    fb = (time.perf_counter() - t0) / iters
    mats[0].requires_grad_(False)
    return fwd * 1e3, fb * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="64,128,150,192,256,280,300")
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--json", default="")
    ap.add_argument("--cpu", action="store_true", help="also time the reference formulation on the host CPU (slow)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rows = []
    for dtype in (torch.float32, torch.float64):
        for N in map(int, args.sizes.split(",")):
            torch.manual_seed(0)
            mats = [torch.randn(N, N, dtype=dtype, device=dev) for _ in range(6)]
            og = torch.ones(N, N, dtype=dtype, device=dev)
            ours = bench(logmatmulexp, mats, og, args.iters, True)
            refg = bench(ref_lme, mats, og, max(3, args.iters // 5), True)
            row = dict(N=N, dtype=str(dtype), ours_fwd_ms=ours[0], ours_fwdbwd_ms=ours[1], ref_formulation_gpu_fwd_ms=refg[0],
                       ref_formulation_gpu_fwdbwd_ms=refg[1], exps_per_chain=5 * N ** 3)
            if args.cpu and N in (256, 300):
                cm = [m.cpu() for m in mats]
                refc = bench(ref_lme, cm, og.cpu(), 2, False)
                row.update(ref_cpu_fwd_ms=refc[0], ref_cpu_fwdbwd_ms=refc[1], cpu_threads=torch.get_num_threads())
            rows.append(row)
            print(row, flush=True)
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
