"""CPU oracle for the EPS contraction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  The product path (``dctn_b200``) never imports this package
and fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``eps_oracle`` is a float64/float32 PyTorch-CPU restatement of
the reference algorithm (dctn/eps.py:19-63, dctn/align.py:11-46,
dctn/epses_composition.py:21-58,133-141, dctn/eps_plus_linear.py:138-147,
dctn/logmatmulexp.py:5-14).  It is checked (tests/test_oracle_vs_golden.py) against
golden vectors produced by importing the UNMODIFIED reference from /root/reference
through the ``ref_shim`` stand-ins for its two missing third-party imports
(``opt_einsum`` — un-vendored, no pinned version, only orders pairwise
torch.einsum calls — and ``more_itertools``); the generating script is
tests/golden/make_golden.py.  ``logmatmulexp`` has no test in the reference at all;
its golden vectors also come from running the reference function itself.
"""
