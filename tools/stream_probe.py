#!/usr/bin/env python
"""K=2, Q_in=2 streaming family at the config-3 size (B=4096, 28x28), steady-state device time per launch: NSET rotating
(x, out) buffer sets (NSET x 50 MB > the 126 MB L2, so every launch reads its input from HBM and its output is written
back by the launches that follow), all launches of a round captured in ONE CUDA graph (no host latency between them),
CUDA events around a graph replay.  An equal-bytes device copy is timed the same way.  (A single launch between two
events is quantised to 1.024 us and carries ~5.7 us of fixed cost on this box: a 50 MB copy reads 14.3 us that way.)
--once: one launch of each kernel after an L2 flush, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dctn_b200 import eps as E
dev = torch.device("cuda:0")
once = "--once" in sys.argv
B, NSET, REP = 4096, 8, 3

def graph_time(fns):
    """fns: one closure per buffer set; returns us per launch"""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for f in fns: f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(REP):
                for f in fns: f()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3 / (REP * len(fns))

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for Oq in ([2, 6] if once else [2, 3, 4, 5, 6]):
    torch.manual_seed(0)
    core = torch.randn(2, 2, 2, 2, Oq, device=dev)
    xs = [torch.rand(1, B, 28, 28, 2, device=dev) for _ in range(NSET)]
    us = [torch.rand(B, 28, 28, device=dev) for _ in range(NSET)]
    out = E.eps(core, xs[0])
    nb = (xs[0].numel() + out.numel()) * 4; nbp = (us[0].numel() + out.numel()) * 4
    srcs = [torch.empty(nb // 8, device=dev) for _ in range(NSET)]; dsts = [torch.empty_like(t) for t in srcs]
    if once:
        gout = torch.randn_like(out)
        c = core.clone().requires_grad_(True); x = xs[0].clone().requires_grad_(True)
        for f in (lambda: E.eps(core, xs[0]), lambda: E.eps_from_pixels(core, us[0], 1.45646), lambda: E.eps(c, x).backward(gout)):
            flush.zero_(); f()
        torch.cuda.synchronize(); continue
    with torch.no_grad():
        tf = graph_time([lambda x=x: E.eps(core, x) for x in xs])
        tp = graph_time([lambda u=u: E.eps_from_pixels(core, u, 1.45646) for u in us])
        tc = graph_time([lambda s=s, d=d: d.copy_(s) for s, d in zip(srcs, dsts)])
    print(f"O={Oq}: fwd {tf:6.2f} us ({nb / tf / 1e3:6.0f} GB/s = {nb / tf / 1e3 / 65.466:4.1f} % of 6546.6)   pixels {tp:6.2f} us "
          f"({nbp / tp / 1e3:6.0f} GB/s of its own {nbp / 1e6:.1f} MB = {nbp / tp / 1e3 / 65.466:4.1f} %)   copy of {nb / 1e6:.1f} MB {tc:6.2f} us ({nb / tc / 1e3:6.0f} GB/s)", flush=True)
