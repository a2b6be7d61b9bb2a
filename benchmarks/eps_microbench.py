#!/usr/bin/env python
"""Config 3 (BASELINE.json): EPS layer microbenchmark — K in {2,3,4}, Q_in in {2,3,4}, Q_out in {2..6}, batch 4096 at
28x28, forward and forward+backward, against the roofline that binds each shape (small_experiments/eps2d_benchmark
shape and dctn/benchmark.py protocol: randn core and input, both requiring grad, fixed randn out_grad).

Infeasible corners are reported, not run: K=4 with Q_in=4 (core 4^16 x Q_out elements) and
K=4 with Q_in=3 (3^16-element core, > 0.4 PFLOP per forward at B=4096).  K=3,Q_in=4 (2.9-8.7 TFLOP per forward) runs at --big-batch (default 4096, like every other row).

    python benchmarks/eps_microbench.py [--batch 4096] [--qouts 2,6] [--json out.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dctn_b200 import _lib  # noqa: E402
from dctn_b200.eps import eps, eps_from_pixels, plan_description  # noqa: E402


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


def timeit(fn, flush, iters):
    fn(); fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def steady_time(fns, rep=3):
    """Device time per launch in steady state: `fns` are the same call on ROTATING buffer sets (together larger than the
    126 MB L2, so every launch reads its input from HBM and its output is written back while the next ones run), all
    captured in one CUDA graph (no host latency between launches), CUDA events around a replay.  A single launch between
    two events is quantised to 1.024 us and carries ~5.7 us of fixed cost on these boxes (a 50 MB device copy reads
    14.3 us that way, 9.7 us this way), which is most of a 10 us kernel."""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for f in fns:
            f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(rep):
                for f in fns:
                    f()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] / (rep * len(fns))


def steady_rows(core, x, gout, K, Q, O, B):
    """forward / core gradient / input gradient of one layer through the C ABI on rotating buffers (ms per call)"""
    lib = _lib.lib()
    nbytes = 4 * (2 * x.numel() + gout.numel())
    nset = max(3, int(400e6 // nbytes) + 1)
    xs = [torch.randn_like(x) for _ in range(nset)]
    gs = [torch.randn_like(gout) for _ in range(nset)]
    outs = [torch.empty_like(gout) for _ in range(nset)]
    dxs = [torch.empty_like(x) for _ in range(nset)]
    dcore = torch.empty_like(core)
    from dctn_b200.eps import _plan
    plan = _plan(1, K, Q, O, torch.float32, _lib.VARIANTS["auto"])
    wsb = max(lib.dctn_eps_workspace_bytes(plan, B, 28, 28, k) for k in (_lib.WS_FORWARD, _lib.WS_BACKWARD_CORE, _lib.WS_BACKWARD_INPUT))
    ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=x.device)

    def stream():
        return torch.cuda.current_stream().cuda_stream

    def f_fwd(xi, oi):
        assert lib.dctn_eps_forward(plan, xi.data_ptr(), core.data_ptr(), oi.data_ptr(), B, 28, 28, ws.data_ptr(), ws.numel(), stream()) == 0

    def f_dcore(xi, gi):
        assert lib.dctn_eps_backward_core(plan, xi.data_ptr(), gi.data_ptr(), dcore.data_ptr(), B, 28, 28, ws.data_ptr(), ws.numel(), stream()) == 0

    def f_dx(xi, gi, di):
        assert lib.dctn_eps_backward_input(plan, xi.data_ptr(), core.data_ptr(), gi.data_ptr(), di.data_ptr(), B, 28, 28, ws.data_ptr(), ws.numel(), stream()) == 0

    t_f = steady_time([lambda a=a, b=b: f_fwd(a, b) for a, b in zip(xs, outs)])
    t_c = steady_time([lambda a=a, b=b: f_dcore(a, b) for a, b in zip(xs, gs)])
    t_x = steady_time([lambda a=a, b=b, c=c: f_dx(a, b, c) for a, b, c in zip(xs, gs, dxs)])
    return t_f, t_c, t_x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--big-batch", type=int, default=4096, help="batch for K=3,Q=4 (8.7 TFLOP per forward at 4096: the largest feasible point of the grid)")
    ap.add_argument("--ks", default="2,3,4")
    ap.add_argument("--qs", default="2,3,4")
    ap.add_argument("--qouts", default="2,3,4,5,6")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    hbm, bf16, src = peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for K in map(int, args.ks.split(",")):
        for Q in map(int, args.qs.split(",")):
            for O in map(int, args.qouts.split(",")):
                D = Q ** (K * K)
                if K == 4 and Q >= 3:
                    rows.append(dict(K=K, Q=Q, O=O, skipped="infeasible: core has %d^16 x %d elements" % (Q, O)))
                    continue
                B = args.big_batch if D >= 262144 else args.batch
                torch.manual_seed(0)
                core = (torch.randn(*(Q,) * (K * K), O, device=dev) * Q ** (-(K * K) / 2)).requires_grad_(True)
                x = torch.randn(1, B, 28, 28, Q, device=dev).requires_grad_(True)
                out = eps(core, x)
                gout = torch.randn_like(out)
                P = out.numel() // O
                flops = 2.0 * P * D * O

                def fwd():
                    with torch.no_grad():
                        eps(core, x)

                def fwdbwd():
                    core.grad = None; x.grad = None
                    eps(core, x).backward(gout)

                l0 = _lib.launch_count()
                t_f = timeit(fwd, flush, args.iters)
                t_fb = timeit(fwdbwd, flush, args.iters)
                assert _lib.launch_count() > l0
                fbytes = 4.0 * (x.numel() + out.numel() + core.numel())
                ai = flops / fbytes
                row = dict(K=K, Q=Q, O=O, B=B, D=D, P=P, fwd_ms=t_f, fwdbwd_ms=t_fb, fwd_tflops=flops / t_f / 1e9,
                           fwdbwd_tflops=4 * flops / t_fb / 1e9, fwd_gbs=fbytes / t_f / 1e6, ai_flop_per_byte=ai,
                           bound="hbm" if ai < 11.4 else ("ffma" if ai < 128 else "tensor"),
                           fwd_frac_hbm=fbytes / t_f / 1e6 / hbm, fwd_frac_bf16=flops / t_f / 1e9 / bf16,
                           imgs_per_s_fwdbwd=B / t_fb * 1e3, patches_per_s_fwdbwd=P / t_fb * 1e3,
                           plan=plan_description(core, x))
                if K == 2 and D * O <= 4096 and core.dtype == torch.float32:
                    # HBM- / CUDA-core-bound rows: microsecond kernels, timed in steady state on rotating buffers
                    s_f, s_c, s_x = steady_rows(core.detach(), x.detach(), gout, K, Q, O, B)
                    cbytes = 4.0 * (x.numel() + out.numel() + core.numel())
                    xbytes = 4.0 * (2 * x.numel() + out.numel() + core.numel())
                    row.update(steady=dict(fwd_ms=s_f, fwd_gbs=fbytes / s_f / 1e6, fwd_frac_hbm=fbytes / s_f / 1e6 / hbm,
                                           dcore_ms=s_c, dcore_gbs=cbytes / s_c / 1e6, dcore_frac_hbm=cbytes / s_c / 1e6 / hbm,
                                           dx_ms=s_x, dx_gbs=xbytes / s_x / 1e6, dx_frac_hbm=xbytes / s_x / 1e6 / hbm,
                                           method="rotating buffers > L2, one CUDA graph per round, events around a replay"))
                    print(f"K={K} Q={Q} O={O} B={B}: steady state  fwd {s_f * 1e3:6.1f} us ({fbytes / s_f / 1e6 / hbm * 100:4.1f} % HBM)  "
                          f"core grad {s_c * 1e3:6.1f} us ({cbytes / s_c / 1e6 / hbm * 100:4.1f} %)  input grad {s_x * 1e3:6.1f} us "
                          f"({xbytes / s_x / 1e6 / hbm * 100:4.1f} %)", flush=True)
                if K == 2 and Q == 2:
                    # feature map fused into the forward (raw pixels in): 1 float per pixel instead of 2 crosses HBM
                    u = torch.rand(B, 28, 28, device=dev)

                    def fwd_pix():
                        with torch.no_grad():
                            eps_from_pixels(core, u, 1.45646)

                    t_p = timeit(fwd_pix, flush, args.iters)
                    pbytes = 4.0 * (u.numel() + out.numel() + core.numel())
                    row.update(pixels_fwd_ms=t_p, pixels_fwd_gbs=pbytes / t_p / 1e6, pixels_fwd_frac_hbm=pbytes / t_p / 1e6 / hbm,
                               pixels_fwd_gbs_as_featurised=fbytes / t_p / 1e6)
                    print(f"K={K} Q={Q} O={O} B={B}: fwd from raw pixels {t_p:8.3f} ms ({row['pixels_fwd_gbs']:7.1f} GB/s of its own "
                          f"{pbytes / 1e6:.1f} MB; {fbytes / t_p / 1e6:7.1f} GB/s counted on the featurised input)", flush=True)
                rows.append(row)
                print(f"K={K} Q={Q} O={O} B={B}: fwd {t_f:8.3f} ms ({row['fwd_tflops']:7.2f} TF/s, {row['fwd_gbs']:7.1f} GB/s, "
                      f"{row['bound']}-bound)  fwd+bwd {t_fb:8.3f} ms ({row['fwdbwd_tflops']:7.2f} TF/s)", flush=True)
                del core, x, out, gout
    if args.json:
        json.dump(dict(peaks=dict(hbm_gbs=hbm, bf16_tflops=bf16, source=src), rows=rows), open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
