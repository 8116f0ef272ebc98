"""Probe of the fused build kernel's image staging variants (TMA with clamped / negative box coordinates, plain loads)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
mode = sys.argv[1]
os.environ["MEDPY_GC_BUILD_TMA"] = mode
from medpy_b200 import synthetic
import medpy_b200.graphcut as gc
from oracle import energy_terms as et, solvers
shape = tuple(int(x) for x in sys.argv[2].split("x")) if len(sys.argv) > 2 else (24, 28, 32)
vol = synthetic.two_blob_volume(shape, seed=0)
g = gc.graph_from_voxels(vol["fg"], vol["bg"], regional_term=gc.energy_voxel.regional_probability_map,
                         regional_term_args=(vol["prob"], vol["alpha"]),
                         boundary_term=gc.energy_voxel.boundary_difference_exponential,
                         boundary_term_args=(vol["image"], vol["sigma"], False))
flow = g.maxflow()
mask = g.get_mask()
prob = et.build_problem(vol["fg"], vol["bg"], regional=(vol["prob"], vol["alpha"]),
                        boundary=("difference_exponential", vol["image"], vol["sigma"], False))
oflow, omask, _ = solvers.solve_port(prob)
print(json.dumps(dict(mode=mode, shape=shape, flow=flow, oflow=oflow, mask_equal=bool(numpy.array_equal(mask, omask)), stats=g.stats())))
