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
    from dctn_b200 import _lib
    lib = _lib.lib()
    plan = E._plan(1, 2, 2, Oq, torch.float32, _lib.VARIANTS["auto"])
    gouts = [torch.randn_like(out) for _ in range(NSET)]
    dcore = torch.empty_like(core); dxs = [torch.empty_like(x) for x in xs]
    wsb = max(lib.dctn_eps_workspace_bytes(plan, B, 28, 28, k) for k in (_lib.WS_BACKWARD_CORE, _lib.WS_BACKWARD_INPUT))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    def f_dcore(x, g):
        st = torch.cuda.current_stream().cuda_stream
        assert lib.dctn_eps_backward_core(plan, x.data_ptr(), g.data_ptr(), dcore.data_ptr(), B, 28, 28, ws.data_ptr(), ws.numel(), st) == 0
    def f_dx(x, g, d):
        st = torch.cuda.current_stream().cuda_stream
        assert lib.dctn_eps_backward_input(plan, x.data_ptr(), core.data_ptr(), g.data_ptr(), d.data_ptr(), B, 28, 28, ws.data_ptr(), ws.numel(), st) == 0
    tdc = graph_time([lambda x=x, g=g: f_dcore(x, g) for x, g in zip(xs, gouts)])
    tdx = graph_time([lambda x=x, g=g, d=d: f_dx(x, g, d) for x, g, d in zip(xs, gouts, dxs)])
    nbd = xs[0].numel() * 4 + out.numel() * 4; nbx = 2 * xs[0].numel() * 4 + out.numel() * 4
    print(f"O={Oq}: core gradient {tdc:6.2f} us ({nbd / tdc / 1e3:6.0f} GB/s = {nbd / tdc / 1e3 / 65.466:4.1f} %)   input gradient {tdx:6.2f} us ({nbx / tdx / 1e3:6.0f} GB/s = {nbx / tdx / 1e3 / 65.466:4.1f} %)", flush=True)
    with torch.no_grad():
        tf = graph_time([lambda x=x: E.eps(core, x) for x in xs])
        tp = graph_time([lambda u=u: E.eps_from_pixels(core, u, 1.45646) for u in us])
        tc = graph_time([lambda s=s, d=d: d.copy_(s) for s, d in zip(srcs, dsts)])
    print(f"O={Oq}: fwd {tf:6.2f} us ({nb / tf / 1e3:6.0f} GB/s = {nb / tf / 1e3 / 65.466:4.1f} % of 6546.6)   pixels {tp:6.2f} us "
          f"({nbp / tp / 1e3:6.0f} GB/s of its own {nbp / 1e6:.1f} MB = {nbp / tp / 1e3 / 65.466:4.1f} %)   copy of {nb / 1e6:.1f} MB {tc:6.2f} us ({nb / tc / 1e3:6.0f} GB/s)", flush=True)
