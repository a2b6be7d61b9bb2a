#!/usr/bin/env python
"""Times individual EPS kernels (CUDA events, L2 flushed) for a layer shape and kernel variant.
   python tools/kbench.py --layers L1,L2 --kinds fwd,core,input --variants ffma,tc3 --batch 512"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dctn_b200 import _lib  # noqa: E402
from dctn_b200 import eps as E  # noqa: E402

LAYERS = {  # name: (H, Q, K, O)
    "L1": (28, 2, 4, 4), "L2": (25, 4, 3, 6), "cfg1": (28, 2, 2, 2), "c23": (31, 23, 2, 24), "c12": (31, 12, 2, 24), "c6": (31, 6, 2, 24),
    "k3q3": (28, 3, 3, 6), "k2q4": (28, 4, 2, 6), "k3q2": (28, 2, 3, 6), "k2q2o6": (28, 2, 2, 6), "k2q3": (28, 3, 2, 6),
}
KINDS = {"fwd": _lib.WS_FORWARD, "core": _lib.WS_BACKWARD_CORE, "input": _lib.WS_BACKWARD_INPUT}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", default="L1,L2")
    ap.add_argument("--kinds", default="fwd,core,input")
    ap.add_argument("--variants", default="auto")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--once", action="store_true", help="one call per (layer, kind, variant) and no timing: for ncu captures")
    ap.add_argument("--train", action="store_true", help="forward keeps T (dctn_eps_forward_train), input gradient uses it")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for lname in args.layers.split(","):
        H, Q, K, O = LAYERS[lname]
        B = args.batch
        n = K * K
        x = torch.rand(1, B, H, H, Q, device=dev) + 0.2
        core = torch.randn(*(Q,) * n, O, device=dev) * Q ** (-n / 2)
        Ho = H - K + 1
        out = torch.empty(B, Ho, Ho, O, device=dev)
        gout = torch.randn_like(out)
        dcore = torch.empty_like(core)
        dx = torch.empty_like(x)
        P, D = B * Ho * Ho, Q ** n
        for variant in args.variants.split(","):
            saved = ws3 = None
            plan = E._plan(1, K, Q, O, torch.float32, _lib.VARIANTS[variant])
            for kname in args.kinds.split(","):
                kind = KINDS[kname]
                ws = torch.empty(lib.dctn_eps_workspace_bytes(plan, B, H, H, kind), dtype=torch.uint8, device=dev)
                st = torch.cuda.current_stream().cuda_stream

                nsave = lib.dctn_eps_saved_bytes(plan, B, H, H) if args.train else 0
                if nsave and saved is None:
                    saved = torch.empty(nsave, dtype=torch.uint8, device=dev)
                    ws3 = torch.empty(lib.dctn_eps_workspace_bytes(plan, B, H, H, 3), dtype=torch.uint8, device=dev)

                def call():
                    if kind == 0:
                        if nsave:
                            return lib.dctn_eps_forward_train(plan, x.data_ptr(), core.data_ptr(), out.data_ptr(), saved.data_ptr(), nsave, B, H, H, ws.data_ptr(), ws.numel(), st)
                        return lib.dctn_eps_forward(plan, x.data_ptr(), core.data_ptr(), out.data_ptr(), B, H, H, ws.data_ptr(), ws.numel(), st)
                    if kind == 1:
                        return lib.dctn_eps_backward_core(plan, x.data_ptr(), gout.data_ptr(), dcore.data_ptr(), B, H, H, ws.data_ptr(), ws.numel(), st)
                    if nsave:
                        return lib.dctn_eps_backward_input_saved(plan, x.data_ptr(), core.data_ptr(), gout.data_ptr(), saved.data_ptr(), nsave, dx.data_ptr(), B, H, H, ws3.data_ptr(), ws3.numel(), st)
                    return lib.dctn_eps_backward_input(plan, x.data_ptr(), core.data_ptr(), gout.data_ptr(), dx.data_ptr(), B, H, H, ws.data_ptr(), ws.numel(), st)

                rc = call()
                if args.once:
                    torch.cuda.synchronize()
                    print(f"{lname} B={B} {kname} {variant}: rc={rc}", flush=True)
                    continue
                if rc != 0:
                    print(f"{lname} {kname} {variant}: unsupported ({_lib.last_error()[:80]})")
                    continue
                call()
                ts = []
                for _ in range(args.iters):
                    flush.zero_()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record(); call(); e.record(); e.synchronize()
                    ts.append(s.elapsed_time(e))
                ms = sum(ts) / len(ts)
                flops = 2.0 * P * D * O * (2 if kname == "input" and not nsave else 1)
                print(f"{lname} B={B} {kname:5s} {variant:5s}: {ms:8.3f} ms  {flops / ms / 1e9:8.2f} TFLOP/s (algorithmic)  ws={ws.numel() / 2**20:.0f} MiB", flush=True)


if __name__ == "__main__":
    main()
