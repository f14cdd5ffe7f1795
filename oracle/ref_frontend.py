"""
Loads the reference's own Python frontend (tracing, fusing, bounds, GraphConverter pre-passes, MLIR
emission, clear evaluation) from /root/reference on top of this backend's stand-ins for the absent
wheel (`concrete.compiler`, `concrete.lang`, `mlir`).  TEST INFRASTRUCTURE: used in the development
container to generate tests/golden/ and to prove the drop-in; /root/reference does not exist on
the GPU box, nothing under `-m gpu` may import this module.

numpy >= 2 shims (the reference targets numpy 1.x):
  * `np.round_` was removed (concrete/numpy/tracing/tracer.py:288)
  * `table % (2 << 63)` overflows int64 under NEP 50 (concrete/numpy/mlir/node_converter.py:1162):
    tables are handed over as an ndarray subclass whose `%` falls back to Python integers.
"""
import os
import sys
import warnings

import numpy as np

REFERENCE = os.environ.get("CNP_REFERENCE_PATH", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _ModSafe(np.ndarray):
    def __mod__(self, other):
        if isinstance(other, int) and other > np.iinfo(np.int64).max:
            return np.asarray(self).astype(object) % other
        return np.asarray(self) % other


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "concrete", "numpy"))


_loaded = None


def load():
    """Returns (cnp, connx) = the reference's `concrete.numpy` and `concrete.onnx` modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE}")
    if not hasattr(np, "round_"):
        np.round_ = np.round
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from concrete_numpy_b200.mlir_text import install_shims

    install_shims()
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import concrete.numpy as cnp
        import concrete.onnx as connx
        from concrete.numpy.mlir import node_converter

    from concrete_numpy_b200.mlir_text import bind_to_parent_package

    bind_to_parent_package()
    original = node_converter.construct_deduplicated_tables

    def construct_tables_nep50(node, preds):
        return [(np.asarray(t).view(_ModSafe), idx) for t, idx in original(node, preds)]

    node_converter.construct_deduplicated_tables = construct_tables_nep50
    _loaded = (cnp, connx)
    return _loaded
