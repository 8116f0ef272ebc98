"""``medpy_b200.graphcut`` -- B200-native drop-in for the voxel half of ``medpy.graphcut``.

Exports the names the reference package exports for this path (medpy/graphcut/__init__.py:186-222):
``graph_from_voxels``, the ``energy_voxel`` module, ``GCGraph``, ``split_marker`` and the ``maxflow`` module
with ``GraphDouble`` / ``GraphFloat`` / ``GraphInt``.  The label/region path (``graph_from_labels``,
``energy_label``, ``graphcut_split`` ...) is outside this path's scope.
"""
from . import energy_voxel, maxflow
from .generate import graph_from_voxels
from .graph import GCGraph
from .maxflow import GraphDouble, GraphFloat, GraphInt
from .wrapper import split_marker

__all__ = ["graph_from_voxels", "energy_voxel", "GCGraph", "GraphDouble", "GraphFloat", "GraphInt",
           "split_marker", "maxflow"]
