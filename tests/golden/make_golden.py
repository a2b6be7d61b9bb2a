"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz (float64 unless stated).  Every case stores the seeded inputs and the
reference's own outputs / gradients, so the parity tests never need the reference at run time.
Cases mirror the reference's tests for this path plus the pins the
reference lacks (core gradient, multi-layer composition, EPSesPlusLinear logits, logmatmulexp).
"""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402

dctn = import_reference()
from dctn.eps import eps, eps_one_by_one, contract_on_input_dims  # noqa: E402
from dctn.epses_composition import contract_with_input, inner_product  # noqa: E402
from dctn.eps_plus_linear import EPSesPlusLinear, UnitTheoreticalOutputStd  # noqa: E402
from dctn.logmatmulexp import logmatmulexp, logmatmulexp_lowmem  # noqa: E402
from dctn.conv_sbs import ConvSBS  # noqa: E402
from dctn.conv_sbs_spec import SBSSpecCore, SBSSpecString  # noqa: E402
from dctn.pos2d import Pos2D  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
F64 = torch.float64


def save(name, **arrays):
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **conv)
    print(name, {k: v.shape for k, v in conv.items()})


def eps_case(name, C, B, H, W, Q, K, O, seed, dtype=F64, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = (scale * torch.randn(C, B, H, W, Q, dtype=dtype, generator=g)).requires_grad_(True)
    core = torch.randn(*(Q,) * (K * K * C), O, dtype=dtype, generator=g).requires_grad_(True)
    out = eps(core, x)
    gout = torch.randn(out.shape, dtype=dtype, generator=g)
    out.backward(gout)
    with torch.no_grad():
        out_obo = eps_one_by_one(core, x)
    save(name, x=x, core=core, out=out, out_one_by_one=out_obo, gout=gout, dcore=core.grad, dx=x.grad)


# --- reference tests/test_eps.py:9-26 shape (C=2, K=2 factor-order pin) and :29-61 (K=3 window order)
eps_case("eps_c2_k2_single_pixel", C=2, B=3, H=2, W=2, Q=2, K=2, O=4, seed=1)
eps_case("eps_c1_k3_two_pixels", C=1, B=1, H=4, W=3, Q=2, K=3, O=4, seed=2)
# --- further eps cases: ragged sizes, Q=3, C=3 with K=1, non-square images, K=4
eps_case("eps_c1_k2_q2", C=1, B=5, H=7, W=9, Q=2, K=2, O=2, seed=3)
eps_case("eps_c1_k2_q3", C=1, B=3, H=6, W=5, Q=3, K=2, O=5, seed=4)
eps_case("eps_c2_k2_q2_4x5", C=2, B=3, H=4, W=5, Q=2, K=2, O=24, seed=5)
eps_case("eps_c3_k1_q2", C=3, B=2, H=3, W=4, Q=2, K=1, O=2, seed=6)
# (K=3,Q=4 has a 4^9-element core: 12.6 MB per float64 array — too large for a fixture; that shape is
# covered by the CUDA-vs-oracle parity tests, the oracle being pinned on the cases below.)
eps_case("eps_c1_k4_q2", C=1, B=2, H=6, W=7, Q=2, K=4, O=2, seed=8)
eps_case("eps_c1_k2_q6", C=1, B=2, H=5, W=5, Q=6, K=2, O=24, seed=9, scale=0.5)
eps_case("eps_c1_k3_q3", C=1, B=1, H=4, W=5, Q=3, K=3, O=3, seed=10, scale=0.7)

# --- reference tests/test_conversion_of_convsbs_to_eps.py:13-56 — eps() output AND input-grad vs ConvSBS
cores = (
    SBSSpecCore(Pos2D(0, 0), 1),
    SBSSpecCore(Pos2D(0, 1), 3),
    SBSSpecCore(Pos2D(1, 0), 2),
    SBSSpecCore(Pos2D(1, 1), 4),
)
perms = list(itertools.permutations(cores))
for idx in (0, 7, 23):
    torch.manual_seed(100 + idx)
    spec = SBSSpecString(perms[idx], (3, 4, 5, 6), 2, 2)
    convsbs = ConvSBS(spec).double()
    with torch.no_grad():
        eps_tensor = convsbs.as_eps()
    x = torch.randn(2, 3, 4, 5, 2, dtype=F64, requires_grad=True)
    sbs_out = convsbs(x)
    gout = torch.randn_like(sbs_out)
    sbs_out.backward(gout)
    save(f"convsbs_as_eps_perm{idx}", eps_tensor=eps_tensor, x=x, convsbs_out=sbs_out, gout=gout,
         convsbs_dx=x.grad)

# --- stacked composition (epses_composition.py:133-141) with all gradients
g = torch.Generator().manual_seed(20)
x = torch.randn(1, 3, 9, 8, 2, dtype=F64, generator=g).requires_grad_(True)
e1 = (0.5 * torch.randn(*(2,) * 9, 4, dtype=F64, generator=g)).requires_grad_(True)  # K=3
e2 = (0.5 * torch.randn(*(4,) * 4, 3, dtype=F64, generator=g)).requires_grad_(True)  # K=2
e3 = (0.5 * torch.randn(*(3,) * 4, 5, dtype=F64, generator=g)).requires_grad_(True)  # K=2
out = contract_with_input((e1, e2, e3), x)
gout = torch.randn(out.shape, dtype=F64, generator=g)
out.backward(gout)
save("composition_3layers", x=x, e1=e1, e2=e2, e3=e3, out=out, gout=gout, de1=e1.grad, de2=e2.grad,
     de3=e3.grad, dx=x.grad)

# --- EPSesPlusLinear logits + parameter grads (eps_plus_linear.py:138-147), config-1-shaped and 2-layer
for name, specs, img, B, seed in (("epl_cfg1_k2q2", ((2, 2),), 28, 6, 30), ("epl_two_layers", ((3, 4), (2, 6)), 10, 4, 31)):
    torch.manual_seed(seed)
    model = EPSesPlusLinear(specs, UnitTheoreticalOutputStd(), 1.0, torch.device("cpu"), F64, image_size=img)
    u = torch.rand(B, img, img, dtype=F64)
    x = torch.stack((2 * torch.sin(u * np.pi / 2) ** 2, 2 * torch.cos(u * np.pi / 2) ** 2), dim=-1)[None]
    y = torch.randint(0, 10, (B,))
    logits = model(x)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    arrays = dict(u=u, x=x, y=y, logits=logits, loss=loss, weight=model.linear.weight, bias=model.linear.bias,
                  dweight=model.linear.weight.grad, dbias=model.linear.bias.grad,
                  reg_epswise=model.epswise_l2_regularizer(), reg_composition=model.epses_composition_l2_regularizer())
    for i, c in enumerate(model.epses):
        arrays[f"eps{i}"] = c
        arrays[f"deps{i}"] = c.grad
    save(name, **arrays)

# --- regulariser known answers (tests/test_eps.py:64-73, tests/test_epses_composition.py:7-41) on random data
g = torch.Generator().manual_seed(40)
a1 = torch.randn(3, 3, 3, 3, 4, dtype=F64, generator=g)
b1 = torch.randn(4, 4, 4, 4, 2, dtype=F64, generator=g)
a2 = torch.randn(3, 3, 3, 3, 4, dtype=F64, generator=g)
b2 = torch.randn(4, 4, 4, 4, 2, dtype=F64, generator=g)
save("inner_product_random", a1=a1, b1=b1, a2=a2, b2=b2, ip_single=inner_product((a1,), (a2,)),
     ip_two=inner_product((a1, b1), (a2, b2)), coid=contract_on_input_dims(a1, a2))

# --- logmatmulexp (logmatmulexp.py:5-22): no reference test exists; pin against the function itself
for name, T, R, I, scale, seed in (("lme_small", 5, 7, 3, 1.0, 50), ("lme_64", 64, 64, 64, 1.0, 51),
                                   ("lme_ragged", 33, 130, 17, 3.0, 52), ("lme_scale150", 20, 24, 28, 150.0, 53)):
    g = torch.Generator().manual_seed(seed)
    A = (scale * torch.randn(T, R, dtype=F64, generator=g)).requires_grad_(True)
    Bm = (scale * torch.randn(R, I, dtype=F64, generator=g)).requires_grad_(True)
    out = logmatmulexp(A, Bm)
    gout = torch.randn(out.shape, dtype=F64, generator=g)
    out.backward(gout)
    with torch.no_grad():
        out_low = logmatmulexp_lowmem(A, Bm)
    save(name, log_A=A, log_B=Bm, out=out, out_lowmem=out_low, gout=gout, dA=A.grad, dB=Bm.grad)
# chain of 6 as in small_experiments/logmatmulexp_benchmark/benchmark.py:21-52
g = torch.Generator().manual_seed(54)
mats = [torch.randn(24, 24, dtype=F64, generator=g) for _ in range(6)]
mats[0].requires_grad_(True)
from functools import reduce  # noqa: E402
out = reduce(logmatmulexp, mats)
out.backward(torch.ones_like(out))
save("lme_chain6", **{f"m{i}": m for i, m in enumerate(mats)}, out=out, dm0=mats[0].grad)
print("done")
