"""medpy_b200 -- B200-native voxel graph-cut path behind MedPy's ``medpy.graphcut`` API.

Sub-packages / modules:
  medpy_b200.graphcut   host-side mirror of medpy.graphcut (graph_from_voxels, energy_voxel.*, GCGraph)
  medpy_b200.csrc       hand-written sm_100a CUDA + the C-ABI (include/medpy_b200_graphcut.h)
  medpy_b200._lib       loader for the in-tree C-ABI shared library (fails loudly when it is missing)
  medpy_b200.synthetic  deterministic synthetic workloads (bench + tests)
"""
__version__ = "0.1.0"
