"""CPU tests (no GPU): host-side mirror of the reference API -- validation, error conventions, arity checks --
and the C-ABI library: it loads and exports every symbol include/medpy_b200_graphcut.h declares."""
import ctypes
import os
import re

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    from medpy_b200 import build
    build.build_all()
    return build.LIB


def test_cabi_exports_every_declared_symbol():
    lib_path = _ensure_built()
    header = open(os.path.join(ROOT, "include", "medpy_b200_graphcut.h")).read()
    declared = sorted(set(re.findall(r"\b(mgc_[a-z_]+)\s*\(", header)))
    assert len(declared) >= 25
    lib = ctypes.CDLL(lib_path)
    missing = [name for name in declared if not hasattr(lib, name)]
    assert not missing, missing
    lib.mgc_abi_version.restype = ctypes.c_int
    assert lib.mgc_abi_version() == 3


def test_pybind_module_imports_and_matches_abi():
    _ensure_built()
    from medpy_b200 import _lib
    assert _lib.ABI_VERSION == 3 and _lib.SOURCE == 0 and _lib.SINK == 1


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must fail loudly, not fall back to anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from medpy_b200.graphcut import graph_from_voxels, energy_voxel
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        graph_from_voxels(numpy.ones((3, 3)), numpy.zeros((3, 3)),
                          boundary_term=energy_voxel.boundary_difference_exponential,
                          boundary_term_args=(numpy.zeros((3, 3)), 1.0, False))


def test_gcgraph_error_conventions():
    """Mirrors the reference's tests/graphcut_/graph.py:28-85 (setter contract), device-free."""
    from medpy_b200.graphcut import GCGraph
    nodes, edges = 10, 20
    graph = GCGraph(nodes, edges)
    graph.set_source_nodes(list(range(0, nodes)))
    with pytest.raises(ValueError):
        graph.set_source_nodes([-1])
    with pytest.raises(ValueError):
        graph.set_source_nodes([nodes])
    graph.set_sink_nodes(list(range(0, nodes)))
    with pytest.raises(ValueError):
        graph.set_sink_nodes([-1])
    with pytest.raises(ValueError):
        graph.set_sink_nodes([nodes])
    graph.set_nweight(0, nodes - 1, 1, 2)
    graph.set_nweight(nodes - 1, 0, 0.5, 1.5)
    for bad in [(-1, 0, 1, 1), (0, nodes, 1, 1), (0, 0, 1, 1), (0, nodes - 1, 0, 0), (0, nodes - 1, -1, -2),
                (0, nodes - 1, -0.5, -1.5)]:
        with pytest.raises(ValueError):
            graph.set_nweight(*bad)
    graph.set_nweights({(0, nodes - 1): (1, 2)})
    with pytest.raises(ValueError):
        graph.set_nweights({(0, 0): (1, 1)})
    graph.set_tweight(0, 1, 2)
    graph.set_tweight(nodes - 1, 0.5, 1.5)
    graph.set_tweight(0, -1, -2)
    graph.set_tweight(0, 0, 0)
    with pytest.raises(ValueError):
        graph.set_tweight(-1, 1, 1)
    with pytest.raises(ValueError):
        graph.set_tweight(nodes, 1, 1)
    graph.set_tweights({0: (1, 2)})
    with pytest.raises(ValueError):
        graph.set_tweights({nodes: (1, 1)})
    assert graph.get_node_count() == nodes
    assert graph.get_edge_count() == edges
    assert graph.get_nodes() == list(range(nodes))
    # a non-chain edge was accepted above: the graph moved to the general sparse backend,
    # which -- like everything else -- has no CPU solver
    assert graph.get_graph().is_sparse
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            graph.get_graph().maxflow()
    # a graph with a lattice shape still refuses edges between non-neighbours
    lattice = GCGraph(6, 7, shape=(2, 3))
    lattice.set_nweight(0, 5, 1, 1)
    with pytest.raises(NotImplementedError):
        lattice.get_graph().maxflow()


def test_term_callables_keep_reference_signature():
    """graph_from_voxels checks len(getfullargspec(f)[0]) == 2 (generate.py:135-146)."""
    import inspect
    from medpy_b200.graphcut import energy_voxel, graph_from_voxels
    for name in energy_voxel.__all__:
        assert len(inspect.getfullargspec(getattr(energy_voxel, name))[0]) == 2, name
    z = numpy.zeros((3, 3))
    with pytest.raises(AttributeError):
        graph_from_voxels(z, z, boundary_term=lambda a: None)
    with pytest.raises(AttributeError):
        graph_from_voxels(z, z, regional_term=lambda a, b, c: None)
    with pytest.raises(AttributeError):
        graph_from_voxels(z, z, boundary_term=3)
    # wrong tuple arity -> ValueError from unpacking, like the reference (energy_voxel.py:95,170)
    with pytest.raises(ValueError):
        graph_from_voxels(z, z, boundary_term=energy_voxel.boundary_difference_linear,
                          boundary_term_args=(z, 1.0, False))


def test_edge_count_matches_reference_formula():
    from medpy_b200.graphcut.generate import voxel_edge_count
    assert voxel_edge_count((64, 64, 64)) == 774144
    assert voxel_edge_count((256, 256, 256)) == 50135040
    assert voxel_edge_count((256, 256, 128, 4)) == 125304832
    assert voxel_edge_count((4, 1, 4)) == 24
    assert voxel_edge_count((2, 3, 5)) == 15 + 20 + 24


def test_split_marker():
    from medpy_b200.graphcut import split_marker
    m = numpy.asarray([[0, 1, 2], [2, 3, 1]])
    fg, bg = split_marker(m)
    assert fg.dtype == numpy.bool_ and bg.dtype == numpy.bool_
    assert fg.tolist() == [[False, True, False], [False, False, True]]
    assert bg.tolist() == [[False, False, True], [True, False, False]]


def test_element_staging_batches_in_call_order():
    from medpy_b200.graphcut.maxflow import GraphDouble
    g = GraphDouble(6, 5, shape=(2, 3))
    g.add_tweights(0, 5, 0)
    g.add_tweights(1, 0, 3)
    g.add_tweights(0, 0, 2)  # same node again -> new batch
    g._close_tweight_batch()
    assert [op[0] for op in g._pending] == ["tw", "tw"]
    assert g._pending[0][1][0] == 5 and g._pending[0][2][1] == 3 and g._pending[1][2][0] == 2
    g.sum_edge(0, 1, 1.0, 2.0)
    g.sum_edge(1, 0, 0.5, 0.25)   # reversed orientation accumulates onto the same pair
    g.sum_edge(0, 3, 7.0, 8.0)
    assert g._st_nw[1][0][0] == 1.25 and g._st_nw[1][1][0] == 2.5
    assert g._st_nw[0][0][0] == 7.0 and g._st_nw[0][1][0] == 8.0
    assert g._axis_of(2, 3) is None  # end of a row: not neighbours
    assert GraphDouble.termtype.SINK == 1 and GraphDouble.termtype.SOURCE == 0
