"""dctn_b200 — B200-native (sm_100a) implementation of dctn's EPS contraction hot path.

Drop-in for the reference modules of the same names (``dctn.eps``, ``dctn.epses_composition``,
``dctn.eps_plus_linear``, ``dctn.logmatmulexp``, ``dctn.contraction_path_cache``, ``dctn.align``,
``dctn.pos2d``, ``dctn.utils``): same function names, argument meaning, parameter layout and
error behaviour, with the contraction running in hand-written CUDA kernels behind a C-ABI library
(``libdctn_b200.so``, header ``include/dctn_b200.h``).  There is no CPU fallback: tensors must live on
a CUDA device and the library must be built (``python -c "import __graft_entry__ as g; g.build()"``).
"""
from . import _lib  # noqa: F401  (does not load the .so until first use)

__all__ = [
    "align",
    "contraction_path_cache",
    "conv_sbs_log",
    "eps",
    "eps_plus_linear",
    "epses_composition",
    "logmatmulexp",
    "pos2d",
    "utils",
]
__version__ = "0.1.0"


def install_as_dctn() -> None:
    """Make ``import dctn.eps`` etc. resolve to this package (drop-in switch for reference users)."""
    import importlib
    import sys

    pkg = sys.modules[__name__]
    sys.modules.setdefault("dctn", pkg)
    for name in __all__:
        sys.modules.setdefault("dctn." + name, importlib.import_module(__name__ + "." + name))
