#!/usr/bin/env python
"""Host-side timing of each public-API call of one end-to-end step at 512^3 from pinned host arrays."""
import os, sys, time
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medpy_b200 import synthetic
import medpy_b200.graphcut as gc
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vol = synthetic.two_blob_volume((size,) * 3, seed=0)
def pin(a):
    t = torch.from_numpy(numpy.ascontiguousarray(a)).pin_memory(); return t, t.numpy()
keep = [pin(vol["image"]), pin(vol["prob"]), pin(vol["fg"].view(numpy.uint8)), pin(vol["bg"].view(numpy.uint8))]
img, prob, fg, bg = (k[1] for k in keep)
fg = fg.view(numpy.bool_); bg = bg.view(numpy.bool_)
from medpy_b200.graphcut.graph import GCGraph
from medpy_b200.graphcut.generate import voxel_edge_count
for rep in range(4):
    T = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    graph = GCGraph(fg.size, voxel_edge_count(fg.shape), shape=fg.shape); graph.get_graph()._nat()
    T["create"] = time.perf_counter() - t0; t0 = time.perf_counter()
    gc.energy_voxel.regional_probability_map(graph, (prob, vol["alpha"]))
    T["regional"] = time.perf_counter() - t0; t0 = time.perf_counter()
    gc.energy_voxel.boundary_difference_exponential(graph, (img, vol["sigma"], False))
    T["boundary"] = time.perf_counter() - t0; t0 = time.perf_counter()
    graph._add_markers(fg, bg)
    T["markers"] = time.perf_counter() - t0; t0 = time.perf_counter()
    g = graph.get_graph(); e = g.maxflow()
    T["maxflow"] = time.perf_counter() - t0; t0 = time.perf_counter()
    m = g.get_mask()
    T["get_mask"] = time.perf_counter() - t0; t0 = time.perf_counter()
    del g, graph
    torch.cuda.synchronize()
    T["destroy"] = time.perf_counter() - t0
    print({k: round(1e3 * v, 2) for k, v in T.items()}, "total", round(1e3 * sum(T.values()), 2), flush=True)
