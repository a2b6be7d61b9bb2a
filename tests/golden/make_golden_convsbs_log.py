"""Golden vectors for the log-space ConvSBS path, from the UNMODIFIED reference.

    python tests/golden/make_golden_convsbs_log.py          (build container only)

Each case runs the reference's ConvSBS.forward (dctn/conv_sbs.py:258-304, LINEAR space) on entrywise positive cores
and inputs given by their logs, takes the log of its output and back-propagates a seeded cotangent to the LOG cores
and the LOG input (chain rule through exp, done by autograd).  Also pins a batch of reference logmatmulexp calls for
the batched entry.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402

dctn = import_reference()
from dctn.conv_sbs import ConvSBS  # noqa: E402
from dctn.conv_sbs_spec import SBSSpecCore, SBSSpecString  # noqa: E402
from dctn.logmatmulexp import logmatmulexp  # noqa: E402
from dctn.pos2d import Pos2D  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
F64 = torch.float64


def save(name, **arrays):
    conv = {k: np.asarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **conv)
    print(name, {k: v.shape for k, v in conv.items()})


def convsbs_log_case(name, positions, outs, bonds, C, Q, B, H, W, seed, core_scale=0.4, x_scale=0.6):
    g = torch.Generator().manual_seed(seed)
    spec = SBSSpecString(tuple(SBSSpecCore(Pos2D(*p), o) for p, o in zip(positions, outs)), tuple(bonds), C, Q)
    model = ConvSBS(spec).double()
    log_cores = [(core_scale * torch.randn(*shape.as_tuple(), dtype=F64, generator=g)).requires_grad_(True)
                 for shape in spec.shapes]
    log_x = (x_scale * torch.randn(C, B, H, W, Q, dtype=F64, generator=g)).requires_grad_(True)
    # the module's own forward with its parameters REPLACED by exp(log_cores) (kept in the autograd graph)
    del model._modules["cores"]
    model.__dict__["cores"] = [lc.exp() for lc in log_cores]
    out = ConvSBS.forward(model, log_x.exp())
    assert (out > 0).all()
    log_out = out.log()
    gout = torch.randn(log_out.shape, dtype=F64, generator=g)
    log_out.backward(gout)
    arrays = dict(log_x=log_x, log_out=log_out, gout=gout, dlog_x=log_x.grad,
                  positions=np.asarray(positions), outs=np.asarray(outs), bonds=np.asarray(bonds))
    for i, lc in enumerate(log_cores):
        arrays[f"log_core{i}"] = lc
        arrays[f"dlog_core{i}"] = lc.grad
    save(name, **arrays)


# the 2x2 string of reference tests/test_conversion_of_convsbs_to_eps.py:13-29 (C=2, bonds 3,4,5,6, outs 1,3,2,4)
convsbs_log_case("convsbs_log_2x2_ring", [(0, 0), (0, 1), (1, 0), (1, 1)], [1, 3, 2, 4], [3, 4, 5, 6], 2, 2, 3, 4, 5, 200)
# same cores visited in another order (permutation 7 of the reference test)
convsbs_log_case("convsbs_log_2x2_perm", [(0, 1), (0, 0), (1, 1), (1, 0)], [3, 1, 4, 2], [2, 5, 3, 4], 2, 2, 2, 5, 4, 201)
# 3x3 snake, one output core, uniform bond 4, closed ring (bond_sizes[0] = 4), C=1
snake = [(0, 0), (0, 1), (0, 2), (1, 2), (1, 1), (1, 0), (2, 0), (2, 1), (2, 2)]
convsbs_log_case("convsbs_log_3x3_snake_ring", snake, [1, 1, 1, 1, 3, 1, 1, 1, 1], [4] * 9, 1, 2, 2, 6, 7, 202)
# open string (bond_sizes[0] = 1: "can't work with a tensor ring" initialisations use this), Q=3
convsbs_log_case("convsbs_log_3x3_snake_open", snake, [1, 1, 1, 1, 1, 1, 1, 1, 5], [1] + [3] * 8, 1, 3, 2, 5, 5, 203,
                 core_scale=0.3)

# batch of independent reference logmatmulexp calls (the reference function is 2-D only: one call per element)
for name, NB, T, R, I, scale, seed in (("lme_batched_small", 7, 3, 5, 4, 1.0, 210), ("lme_batched_r8", 37, 8, 8, 8, 2.0, 211),
                                       ("lme_batched_scale150", 5, 6, 4, 12, 150.0, 212)):
    g = torch.Generator().manual_seed(seed)
    A = (scale * torch.randn(NB, T, R, dtype=F64, generator=g)).requires_grad_(True)
    Bm = (scale * torch.randn(NB, R, I, dtype=F64, generator=g)).requires_grad_(True)
    out = torch.stack([logmatmulexp(A[p], Bm[p]) for p in range(NB)])
    gout = torch.randn(out.shape, dtype=F64, generator=g)
    out.backward(gout)
    save(name, log_A=A, log_B=Bm, out=out, gout=gout, dA=A.grad, dB=Bm.grad)
print("done")
