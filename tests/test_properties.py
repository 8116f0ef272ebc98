"""Property-based checks (hypothesis) of host logic and oracle pieces that have an exact specification:
relabel / relabel_map against the reference's own loops, sum_edge accumulation semantics, the BK restatement for general
graphs against max-flow == min-cut duality (and against the real solver where it is built)."""
import os
import sys

import numpy
import pytest
from hypothesis import given, settings, strategies as st

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import energy_label_terms as elt  # noqa: E402
from oracle import solvers  # noqa: E402

from medpy_b200.relabel import relabel, relabel_map  # noqa: E402


def _reference_relabel(label_image, start=1):
    """The reference's algorithm (medpy/filter/label.py:95-104) restated as the obvious loop."""
    flat = numpy.asarray(label_image).ravel().copy()
    mapping = {}
    for k, v in enumerate(flat.tolist()):
        if v not in mapping:
            mapping[v] = start
            start += 1
        flat[k] = mapping[v]
    return flat.reshape(numpy.asarray(label_image).shape)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(-5, 12), min_size=1, max_size=60), st.integers(0, 5))
def test_relabel_is_first_appearance_order(values, start):
    a = numpy.asarray(values, dtype=numpy.int64).reshape(1, -1)
    assert numpy.array_equal(relabel(a, start), _reference_relabel(a, start))
    out = relabel(a, start)
    assert sorted(set(out.ravel().tolist())) == list(range(start, start + len(set(values))))


@settings(max_examples=40, deadline=None)
@given(st.lists(st.integers(1, 6), min_size=1, max_size=40))
def test_relabel_map_applies_the_mapping(values):
    a = numpy.asarray(values)
    mapping = [0, 10, 20, 30, 40, 50, 60]
    assert relabel_map(a, mapping).tolist() == [mapping[v] for v in values]
    assert relabel_map(a, {v: -v for v in range(1, 7)}).tolist() == [-v for v in values]


graphs = st.integers(2, 12).flatmap(lambda n: st.tuples(
    st.just(n),
    st.lists(st.tuples(st.integers(0, n - 1), st.integers(0, n - 1), st.integers(0, 9), st.integers(0, 9)), min_size=0, max_size=40),
    st.lists(st.tuples(st.integers(0, n - 1), st.integers(-3, 9), st.integers(-3, 9)), min_size=1, max_size=20)))


def _cut_capacity(n, lo, hi, c_lh, c_hl, tr, const, mask):
    """Capacity of the cut a mask describes (mask 1 = source side), SURVEY.md App. A.6."""
    e = const
    for v in range(n):
        if mask[v] == 0 and tr[v] > 0:
            e += tr[v]
        if mask[v] == 1 and tr[v] < 0:
            e -= tr[v]
    for a, b, x, y in zip(lo, hi, c_lh, c_hl):
        if mask[a] == 1 and mask[b] == 0:
            e += x
        if mask[b] == 1 and mask[a] == 0:
            e += y
    return e


@settings(max_examples=120, deadline=None)
@given(graphs)
def test_sparse_bk_restatement_duality_and_reference(g):
    n, edges, tws = g
    edges = [e for e in edges if e[0] != e[1]]
    i = numpy.asarray([e[0] for e in edges], dtype=numpy.int64)
    j = numpy.asarray([e[1] for e in edges], dtype=numpy.int64)
    cap = numpy.asarray([e[2] for e in edges], dtype=float)
    rev = numpy.asarray([e[3] for e in edges], dtype=float)
    tw = [(numpy.asarray([t[0]]), numpy.asarray([float(t[1])]), numpy.asarray([float(t[2])])) for t in tws]
    flow, mask, _ = solvers.solve_sparse_port(n, i, j, cap, rev, tw)
    lo, hi, c_lh, c_hl = elt.merge_edges(i, j, cap, rev) if i.size else (numpy.zeros(0, int),) * 2 + (numpy.zeros(0),) * 2
    tr, const = elt.add_tweights_replay(n, tw)
    # integer capacities: the returned flow IS the capacity of the returned cut, and no cut is cheaper than a cut
    # obtained by flipping a single node (local optimality; global optimality is the solver comparison below)
    assert flow == _cut_capacity(n, lo, hi, c_lh, c_hl, tr, const, mask)
    for v in range(n):
        m2 = mask.copy()
        m2[v] ^= 1
        assert _cut_capacity(n, lo, hi, c_lh, c_hl, tr, const, m2) >= flow
    if solvers.have_ref():
        rflow, rmask, _ = solvers.solve_sparse_ref(n, i, j, cap, rev, tw)
        assert rflow == flow and numpy.array_equal(rmask, mask)
