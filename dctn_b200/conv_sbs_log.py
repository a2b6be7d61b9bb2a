"""ConvSBS forward in log space — SURVEY.md section 8f-4, an ADDITIONAL entry.

The reference contracts a ConvSBS (a string of bond-connected cores laid over a window, dctn/conv_sbs.py:258-304)
in linear space: per core, the window's input vectors are contracted with the core into one bond matrix per
(image, window, out_quantum) (:269-281); then the ring of bond matrices is multiplied and traced (:282-303).  The
reference's ``logmatmulexp`` (dctn/logmatmulexp.py) was written for exactly this product but never wired in
(SURVEY.md section 9.1).  Here both steps run in log space for entrywise POSITIVE cores and inputs, given as logs:

* per core:   ``log M_c[p, l, o, r] = logsumexp_i(log x_c[p, i] + log core_c[o, l, r, i])`` — the 2-D
  ``logmatmulexp`` kernel on ``(P, Q^C) x (Q^C, L*O*R)``;
* the ring:   ``T <- logmatmulexp_batched(T[p], log M_c[p])`` core by core (one small product per window), out_quantum
  dims accumulating in the row index, and a final log-trace over the closing bond.

Same argument meaning as ``ConvSBS.forward``: input ``(C, B, H, W, Q)`` (or a tuple of channels), cores shaped
``(out_quantum, bond_left, bond_right, Q, ..., Q)`` (dctn/conv_sbs_spec.py:24-27), output
``(B, H', W', prod out_quantum)`` — the LOG of what ``ConvSBS.forward`` returns.
"""
from typing import Sequence, Tuple, Union

import torch
from torch import Tensor

from .align import align_with_positions
from .logmatmulexp import logmatmulexp, logmatmulexp_batched
from .pos2d import Pos2D


def log_bond_matrices(log_core: Tensor, log_channels: Sequence[Tensor]) -> Tensor:
    """One core's step (dctn/conv_sbs.py:269-281) in log space.  log_core: (O, L, R, Q, ..., Q); log_channels: C views
    of shape (B, H', W', Q).  Returns (P, L, O*R) — bond_left leading so the ring step is a plain row-major product."""
    O, L, R = log_core.shape[:3]
    C = len(log_channels)
    assert log_core.ndim == 3 + C
    P = log_channels[0].numel() // log_channels[0].shape[-1]
    kr = log_channels[0].reshape(P, -1)
    for ch in log_channels[1:]:  # log of the rank-one product of the channels' vectors, first channel slowest
        kr = (kr.unsqueeze(2) + ch.reshape(P, 1, -1)).reshape(P, -1)
    mat = log_core.permute(*range(3, 3 + C), 1, 0, 2).reshape(-1, L * O * R)
    return logmatmulexp(kr.contiguous(), mat.contiguous()).reshape(P, L, O * R)


def conv_sbs_log_forward(
    log_cores: Sequence[Tensor], positions: Tuple[Pos2D, ...], log_input: Union[Tensor, Tuple[Tensor, ...]]
) -> Tensor:
    """log(ConvSBS(spec with these positions, cores = exp(log_cores)).forward(exp(log_input)))."""
    assert len(log_cores) == len(positions)
    num_channels = len(log_input)
    batch_size, height, width, _ = log_input[0].shape
    aligned = list(align_with_positions(log_input, tuple(positions)))
    out_h, out_w = aligned[0].shape[1:3]
    P = batch_size * out_h * out_w
    L0 = log_cores[0].shape[1]
    assert log_cores[-1].shape[2] == L0, "the last core's right bond closes the ring on the first core's left bond"
    T = None  # (P, L0 * prod(O so far), R_c)
    for c, core in enumerate(log_cores):
        O, L, R = core.shape[:3]
        M = log_bond_matrices(core, aligned[c * num_channels : (c + 1) * num_channels])  # (P, L, O*R)
        if T is None:
            T = M.reshape(P, L * O, R)
        else:
            assert T.shape[2] == L, "bond sizes of neighbouring cores must match"
            T = logmatmulexp_batched(T, M).reshape(P, -1, R)
    T = T.reshape(P, L0, -1, L0)
    out = torch.logsumexp(torch.diagonal(T, dim1=1, dim2=3), dim=-1)  # log trace over the closing bond
    return out.reshape(batch_size, out_h, out_w, -1)
