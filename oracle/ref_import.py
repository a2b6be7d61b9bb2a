"""Import the UNMODIFIED reference (``dctn``) from /root/reference (TEST INFRASTRUCTURE).

Only usable in the build container: /root/reference does not exist on the GPU box, so nothing
that runs there (``-m gpu`` tests, smoke(), bench.py) may call this.  Used by
tests/golden/make_golden.py and by the optional CPU tests that re-check the oracle against
the live reference when it is present.
"""
import os
import sys

REFERENCE_ROOT = "/root/reference"
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dctn"))


def import_reference():
    """Returns the reference's ``dctn`` package (with eps, epses_composition, eps_plus_linear,
    logmatmulexp, conv_sbs loaded)."""
    import torch  # noqa: F401  (must be imported BEFORE the shim is importable)

    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import dctn  # the reference package
    import dctn.eps, dctn.epses_composition, dctn.eps_plus_linear, dctn.logmatmulexp  # noqa
    import dctn.conv_sbs, dctn.conv_sbs_spec, dctn.pos2d, dctn.contraction_path_cache  # noqa

    assert os.path.abspath(dctn.__file__).startswith(REFERENCE_ROOT)
    return dctn
