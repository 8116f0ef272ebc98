"""``medpy.graphcut`` -> ``medpy_b200.graphcut`` (voxel path only; see medpy_b200/graphcut/__init__.py)."""
import sys

import medpy_b200.graphcut as _gc
from medpy_b200.graphcut import *  # noqa: F401,F403
from medpy_b200.graphcut import GCGraph, energy_voxel, graph_from_voxels, maxflow, split_marker  # noqa: F401
from medpy_b200.graphcut import generate, graph, wrapper  # noqa: F401

for _name in ("energy_voxel", "maxflow", "generate", "graph", "wrapper"):
    sys.modules[__name__ + "." + _name] = getattr(_gc, _name)
