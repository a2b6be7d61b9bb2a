"""Minimal stand-in for the third-party ``opt_einsum`` package (TEST INFRASTRUCTURE).

The reference imports opt_einsum (dctn/eps.py:13, dctn/contraction_path_cache.py:3) but the
package is absent from this image and there is no network.  opt_einsum does no arithmetic of
its own: it orders pairwise ``torch.einsum`` calls.  This shim reproduces that behaviour for
the call forms the reference uses so that the UNMODIFIED reference can be imported from
/root/reference to generate golden vectors (tests/golden/make_golden.py):

* ``contract(subscripts, *operands)`` — string form, arbitrary unicode symbols
  (reference tests/test_eps.py:14 uses digits and a Greek theta);
* ``contract(op0, names0, op1, names1, ..., out_names)`` — interleaved form with
  arbitrary hashable names (dctn/eps.py:31-40);
* ``optimize=`` an explicit path (list of index tuples, opt_einsum semantics: operands
  are popped from the current list, the contraction result is appended) — honoured so
  that the intermediates are the reference's own (dctn/eps.py:25-30);
* ``contract_expression`` with shapes in place of tensors
  (dctn/contraction_path_cache.py:26-31).
"""
from .contract import (  # noqa: F401  (rebinds the name `contract` to the function, as the real package does)
    ContractExpression,
    contract,
    contract_expression,
    contract_path,
)

__version__ = "0.0-shim"
