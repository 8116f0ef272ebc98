"""The Python glue of the region-graph path, end to end on the CPU: the checks of tests/test_gpu_labels.py are run with
the two native classes replaced by oracle-backed test doubles (tests/fake_native.py).  What this covers: energy_label,
graph_from_labels, GCGraph's bulk and element-wise setters, the sparse GraphDouble staging and journal migration,
label_cut_mask.  What it cannot cover: the CUDA kernels and the C ABI (that is what ``-m gpu`` is for)."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import solvers  # noqa: E402

import fake_native  # noqa: E402
import test_gpu_labels as T  # noqa: E402



@pytest.fixture(autouse=True)
def fake_native_classes(monkeypatch):
    from medpy_b200 import _lib
    monkeypatch.setattr(_lib._mgc, "LabelImage", fake_native.FakeLabelImage)
    monkeypatch.setattr(_lib._mgc, "SparseGraph", fake_native.FakeSparseGraph)
    yield


@pytest.mark.parametrize("nm", T.NAMES)
def test_stawiaski_and_means_glue(nm):
    T.test_stawiaski_edges_vs_reference(nm)
    T.test_difference_of_means_edges_vs_reference(nm)


@pytest.mark.parametrize("nm", T.FULL)
def test_directed_atlas_and_whole_cut_glue(nm):
    T.test_directed_edges_vs_reference(nm)
    T.test_atlas_tweights_vs_reference(nm)
    for tag in ("cut_stawiaski", "cut_means", "cut_directed_atlas"):
        T.test_graph_from_labels_whole_cut_vs_reference(nm, tag)


def test_label_image_checks_glue():
    T.test_label_image_checks()


@pytest.mark.parametrize("seed", [0, 1, 3, 4])
def test_general_sparse_graph_glue(seed):
    T.test_general_sparse_graph_vs_reference_bk(seed)


def test_sparse_fixture_glue():
    T.test_sparse_graph_reference_fixture_and_reset()


def test_wrapper_functions_glue():
    T.test_wrapper_functions_vs_reference()


def test_set_tweights_all_on_general_graphs():
    """GCGraph.set_tweights_all (graph.py:532-552) on a sparse graph, and on a shape-less graph BEFORE it turns sparse."""
    import numpy
    from medpy_b200.graphcut import GCGraph
    from oracle import solvers as S
    tw = numpy.asarray([[5.0, 0.0], [0.0, 1.0], [0.0, 0.0], [0.0, 4.0]])
    edges = [(0, 2, 3.0, 1.0), (2, 3, 2.0, 2.0), (0, 1, 0.5, 0.5)]
    want = S.solve_sparse(4, [e[0] for e in edges], [e[1] for e in edges], [e[2] for e in edges], [e[3] for e in edges],
                          [(numpy.arange(4), tw[:, 0], tw[:, 1])])
    for sparse_first in (True, False):
        g = GCGraph(4, 3, sparse=True if sparse_first else None)
        g.set_tweights_all(tw)                      # journaled when the graph is still a chain
        for a, b, c, d in edges:
            g.set_nweight(a, b, c, d)               # (0, 2) is no chain neighbour -> sparse backend
        assert g.get_graph().is_sparse
        assert g.get_graph().maxflow() == want[0]
        assert numpy.array_equal(g.get_graph().get_mask(), want[1])
