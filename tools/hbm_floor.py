#!/usr/bin/env python
"""Times an equal-bytes device copy next to the direct EPS forward (K=2,Q=2,O=2,B=4096) with the same event method."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dctn_b200 import eps as E
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.rand(1, B, 28, 28, 2, device=dev)
core = torch.randn(2, 2, 2, 2, 2, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
src = torch.empty(x.numel(), device=dev); dst = torch.empty_like(src)
def t(fn, n=20, fl=True):
    fn(); fn(); ts = []
    for _ in range(n):
        if fl: flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2] * 1e3
out = E.eps(core, x)
nbytes = x.numel() * 4 + out.numel() * 4
with torch.no_grad():
    for fl in (True,):
        te = t(lambda: E.eps(core, x), fl=fl); tc = t(lambda: dst.copy_(src), fl=fl)
        print(f"flush={fl}: eps fwd {te:7.1f} us ({nbytes / te / 1e3:7.1f} GB/s algorithmic)   copy of {2 * src.numel() * 4 / 1e6:.1f} MB {tc:7.1f} us ({2 * src.numel() * 4 / tc / 1e3:7.1f} GB/s)")
    # 10 back-to-back launches in one event pair (amortises launch latency)
    def many(): 
        for _ in range(10): E.eps(core, x)
    print(f"10 back-to-back eps fwd: {t(many, fl=False) / 10:7.1f} us each")
