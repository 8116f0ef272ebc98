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
