"""``medpy_b200.graphcut`` -- B200-native drop-in for ``medpy.graphcut``'s graph construction and min-cut.

Exports the names the reference package exports (medpy/graphcut/__init__.py:186-222): ``graph_from_voxels`` and the
``energy_voxel`` module (the voxel path), ``graph_from_labels`` and the ``energy_label`` module (the
region path, §8 row f3), ``GCGraph``, ``split_marker`` and the ``maxflow`` module with ``GraphDouble`` /
``GraphFloat`` / ``GraphInt`` (general sparse graphs: row f4), the plain ``Graph`` record and ``graph_to_dimacs``.  ``graphcut_stawiaski`` / ``graphcut_split`` /
``graphcut_subprocesses`` (wrapper.py:72-329) are provided with the reference's splitting rules; their jobs run back to
back on the GPU instead of in a process pool.
"""
from . import energy_label, energy_voxel, maxflow
from .generate import graph_from_labels, graph_from_voxels, label_cut_mask
from .graph import GCGraph, Graph
from .maxflow import GraphDouble, GraphFloat, GraphInt
from .wrapper import graphcut_split, graphcut_stawiaski, graphcut_subprocesses, split_marker
from .write import graph_to_dimacs

__all__ = ["graph_from_voxels", "graph_from_labels", "label_cut_mask", "energy_voxel", "energy_label", "GCGraph", "Graph",
           "graph_to_dimacs", "GraphDouble", "GraphFloat", "GraphInt", "split_marker", "graphcut_split", "graphcut_stawiaski",
           "graphcut_subprocesses", "maxflow"]
