"""Loader for the in-tree native pieces.  There is no Python / CPU fallback: if the compiled extension is
missing or cannot be loaded, importing this module raises, loudly."""
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmedpy_b200_gc.so")

try:
    from . import _mgc  # noqa: F401  (pybind11 binding, built by medpy_b200.build)
except ImportError as exc:  # pragma: no cover - exercised only on a broken install
    raise ImportError(
        "medpy_b200: the native CUDA extension (medpy_b200/_mgc*.so + lib/libmedpy_b200_gc.so) is not built or "
        "cannot be loaded (%s). Build it in-tree with `python -m medpy_b200.build`; there is no CPU fallback." % exc
    ) from exc

Graph = _mgc.Graph
ABI_VERSION = _mgc.ABI_VERSION
SOURCE = _mgc.SOURCE
SINK = _mgc.SINK
