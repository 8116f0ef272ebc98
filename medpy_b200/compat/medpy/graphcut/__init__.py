"""``medpy.graphcut`` -> ``medpy_b200.graphcut`` (see medpy_b200/graphcut/__init__.py for what is covered)."""
import sys

import medpy_b200.graphcut as _gc
from medpy_b200.graphcut import *  # noqa: F401,F403
from medpy_b200.graphcut import (GCGraph, energy_label, energy_voxel, graph_from_labels, graph_from_voxels,  # noqa: F401
                                 maxflow, split_marker)
from medpy_b200.graphcut import generate, graph, wrapper, write  # noqa: F401

for _name in ("energy_voxel", "energy_label", "maxflow", "generate", "graph", "wrapper", "write"):
    sys.modules[__name__ + "." + _name] = getattr(_gc, _name)
