#!/usr/bin/env python
"""Per-layer diagnosis of a stacked model against the float64 oracle: every layer is run in isolation on the ORACLE's
inputs / upstream gradients (rounded to float32), so a failing layer is not masked by its neighbours.
    python tools/diag_layers.py "((4,4),(3,12),(2,24))" 2 28 9 102
"""
import ast
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from dctn_b200 import eps as E
from dctn_b200.eps_plus_linear import EPSesPlusLinear, UnitTheoreticalOutputStd
from oracle import eps_oracle as O
from oracle.eps_oracle import rel_err

specs = ast.literal_eval(sys.argv[1]) if len(sys.argv) > 1 else ((4, 4), (3, 12), (2, 24))
Q0, img, B, seed = (int(v) for v in (sys.argv[2:6] + ["2", "28", "9", "102"][len(sys.argv) - 2:])) if len(sys.argv) > 2 else (2, 28, 9, 102)
dev = torch.device("cuda:0")
torch.manual_seed(seed)
model = EPSesPlusLinear(specs, UnitTheoreticalOutputStd(), 1.0, dev, torch.float32, image_size=img, Q_0=Q0)
if Q0 == 2:
    x = O.phi_cos_sin_squared(torch.rand(B, img, img, dtype=torch.float64), 1.45646 / 2).float()
else:
    x = torch.randn(1, B, img, img, Q0) * 0.8
    x[..., -1] = 0.8
y = torch.randint(0, 10, (B,))
cores = [c.detach().double().cpu().requires_grad_(True) for c in model.epses]
w = model.linear.weight.detach().double().cpu().requires_grad_(True)
b = model.linear.bias.detach().double().cpu().requires_grad_(True)
inters = [x.double().requires_grad_(True)]
for c in cores:
    out = O.eps_4step(c, inters[-1])
    out.retain_grad()
    nxt = out.unsqueeze(0)
    nxt.retain_grad()
    inters.append(nxt)
logits = inters[-1].squeeze(0).reshape(B, -1) @ w.T + b
F.cross_entropy(logits, y).backward()
for li, c in enumerate(cores):
    xin = inters[li].detach()
    gout = inters[li + 1].grad.squeeze(0)
    print(f"--- layer {li}: K,O={specs[li]} |x|max={xin.abs().max():.3e} |out|={inters[li+1].detach().norm():.3e} |gout|={gout.norm():.3e} "
          f"|dcore|={c.grad.norm():.3e} |dx|={inters[li].grad.norm():.3e}")
    for variant in ("auto", "ffma"):
        E.set_default_variant(variant)
        try:
            cd = c.detach().float().to(dev).requires_grad_(True)
            xd = xin.float().to(dev).requires_grad_(True)
            fam = E.kernel_families(cd, xd)
            o = E.eps(cd, xd)
            o.backward(gout.float().to(dev))
            # oracle on the float32-rounded inputs
            c32, x32, g32 = c.detach().float().double(), xin.float().double(), gout.float().double()
            want = O.eps_4step(c32, x32)
            wdc, wdx = O.eps_grads(c32, x32, g32)
            print(f"   {variant:5s} fam={fam} out {rel_err(o, want):.2e} dcore {rel_err(cd.grad, wdc):.2e} (|ours|={cd.grad.double().norm():.3e}) "
                  f"dx {rel_err(xd.grad, wdx):.2e} (|ours|={xd.grad.double().norm():.3e})")
        finally:
            E.set_default_variant("auto")
# whole model
logits_gpu = model(x.to(dev))
F.cross_entropy(logits_gpu, y.to(dev)).backward()
print("model: logits", rel_err(logits_gpu, logits), [f"deps{i} {rel_err(model.epses[i].grad, c.grad):.2e}" for i, c in enumerate(cores)],
      "dweight", rel_err(model.linear.weight.grad, w.grad))
# what the reference's own float32 arithmetic gives on the CPU (same graph in float32)
c32 = [c.detach().float().requires_grad_(True) for c in cores]
w32, b32 = w.detach().float().requires_grad_(True), b.detach().float().requires_grad_(True)
F.cross_entropy(O.eps_plus_linear_forward(c32, w32, b32, x), y).backward()
print("float32 CPU restatement vs float64:", [f"deps{i} {rel_err(c32[i].grad, c.grad):.2e}" for i, c in enumerate(cores)], "dweight", rel_err(w32.grad, w.grad))
