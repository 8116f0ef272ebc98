"""The two ``medpy.filter`` functions bin/medpy_graphcut_label.py uses around the region cut (medpy/filter/label.py:31-105);
everything else of ``medpy.filter`` is outside this repository's scope."""
from medpy_b200.relabel import relabel, relabel_map  # noqa: F401

__all__ = ["relabel", "relabel_map"]
