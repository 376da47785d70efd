"""Locate the CUDA binding whether ``src`` is imported as a sub-package of
``multi-modal-uncertainty_b200`` or as a top-level package (the reference's own layout, with the
package directory on ``sys.path``)."""
try:
    from .. import _lib, ops  # noqa: F401
except ImportError:  # pragma: no cover - top-level ``src`` layout
    import importlib
    import os
    import sys

    _root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if _root not in sys.path:
        sys.path.insert(0, _root)
    _pkg = importlib.import_module("multi-modal-uncertainty_b200")
    _lib, ops = _pkg._lib, _pkg.ops
