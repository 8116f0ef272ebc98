"""CPU tests of the round-2 host logic: graph_from_voxels records its whole-lattice terms and hands them to the native side
in ONE build_voxel_graph call (regional term, boundary term, markers in the reference's order, generate.py:159-172);
anything that does not fit that pattern falls back to one call per term in call order.  The native class is replaced by the
oracle-backed test double (tests/fake_native.py) wrapped in a call recorder -- no GPU, no product path through the oracle."""
import numpy
import pytest


class _Recorder:
    """FakeGraph with a log of the native calls it receives."""

    def __init__(self, shape, device=-1):
        import fake_native
        self.inner = fake_native.FakeGraph(shape, device)
        self.calls = []

    def __getattr__(self, name):
        attr = getattr(self.inner, name)
        if callable(attr) and not name.startswith("_"):
            def wrapped(*a, **k):
                self.calls.append(name)
                return attr(*a, **k)
            return wrapped
        return attr


@pytest.fixture()
def fake(monkeypatch):
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from medpy_b200 import _lib
    made = []

    def factory(shape, device=-1):
        r = _Recorder(shape, device)
        made.append(r)
        return r
    monkeypatch.setattr(_lib, "Graph", factory)
    return made


def _volume(shape=(6, 7, 8), seed=0):
    from medpy_b200 import synthetic
    return synthetic.two_blob_volume(shape, seed=seed)


def _oracle(vol, regional=True):
    from oracle import energy_terms as et, solvers
    prob = et.build_problem(vol["fg"], vol["bg"], regional=(vol["prob"], vol["alpha"]) if regional else None,
                            boundary=("difference_exponential", vol["image"], vol["sigma"], False))
    return solvers.solve_port(prob)


def test_graph_from_voxels_issues_one_build_call(fake):
    import medpy_b200.graphcut as gc
    vol = _volume()
    g = gc.graph_from_voxels(vol["fg"], vol["bg"], regional_term=gc.energy_voxel.regional_probability_map,
                             regional_term_args=(vol["prob"], vol["alpha"]),
                             boundary_term=gc.energy_voxel.boundary_difference_exponential,
                             boundary_term_args=(vol["image"], vol["sigma"], False))
    calls = [c for c in fake[0].calls if c not in ("set_option", "check_deferred")]
    assert calls == ["build_voxel_graph"], calls
    flow, mask, _ = _oracle(vol)
    assert g.maxflow() == flow
    assert numpy.array_equal(g.get_mask(), mask)


def test_boundary_only_and_no_terms(fake):
    import medpy_b200.graphcut as gc
    vol = _volume(seed=1)
    g = gc.graph_from_voxels(vol["fg"], vol["bg"], boundary_term=gc.energy_voxel.boundary_difference_exponential,
                             boundary_term_args=(vol["image"], vol["sigma"], False))
    assert [c for c in fake[0].calls if c not in ("set_option", "check_deferred")] == ["build_voxel_graph"]
    flow, mask, _ = _oracle(vol, regional=False)
    assert g.maxflow() == flow and numpy.array_equal(g.get_mask(), mask)
    # no term at all: only the markers, applied directly
    g2 = gc.graph_from_voxels(vol["fg"], vol["bg"])
    assert "add_markers" in fake[1].calls and "build_voxel_graph" not in fake[1].calls
    g2.maxflow()


def test_element_wise_calls_between_terms_keep_the_reference_order(fake):
    """A user-written boundary term (set_nweight per edge) after a built-in regional term: the collected regional term is
    committed first, then the staged n-weights, then the markers -- the order graph_from_voxels applies them in."""
    import medpy_b200.graphcut as gc
    vol = _volume((4, 5, 6), seed=2)

    def my_boundary(graph, args):
        (img,) = args
        n = img.size
        for p in range(n - 1):
            if (p + 1) % img.shape[-1]:
                graph.set_nweight(p, p + 1, 2.0, 2.0)

    g = gc.graph_from_voxels(vol["fg"], vol["bg"], regional_term=gc.energy_voxel.regional_probability_map,
                             regional_term_args=(vol["prob"], vol["alpha"]), boundary_term=my_boundary,
                             boundary_term_args=(vol["image"],))
    calls = [c for c in fake[0].calls if c not in ("set_option", "check_deferred")]
    assert calls == ["build_voxel_graph", "add_nweights_dense", "add_markers"], calls
    # same graph assembled by hand on the oracle
    from oracle import energy_terms as et, solvers
    prob = et.build_problem(vol["fg"], vol["bg"], regional=(vol["prob"], vol["alpha"]))
    n = vol["image"].size
    w = numpy.zeros(n)
    w[[p for p in range(n - 1) if (p + 1) % vol["image"].shape[-1]]] = 2.0
    zero = numpy.zeros(n)
    prob["wf"] = [zero, zero, w]
    prob["wb"] = [zero, zero, w]
    flow, mask, _ = solvers.solve_port(prob)
    assert abs(g.maxflow() - flow) <= 1e-12 * max(1.0, abs(flow))
    assert numpy.array_equal(g.get_mask(), mask)


def test_terms_outside_graph_from_voxels_are_immediate_and_reset_clears_collection(fake):
    from medpy_b200.graphcut.maxflow import GraphDouble
    vol = _volume((4, 4, 4), seed=3)
    g = GraphDouble(64, 0, shape=(4, 4, 4))
    g.add_regional_probability(vol["prob"], vol["alpha"], True)          # no collection outside graph_from_voxels
    assert fake[0].calls[-1] == "add_regional_probability"
    g.defer_weight_check(True)                                          # what graph_from_voxels switches on ...
    g.reset()
    g.add_boundary(1, vol["image"], vol["sigma"], None, float("nan"))   # ... collected
    assert "add_boundary" not in fake[0].calls and "build_voxel_graph" not in fake[0].calls
    g.reset()                                                           # dropped with the reset
    g.add_markers(vol["fg"], vol["bg"])
    assert "build_voxel_graph" not in fake[0].calls and fake[0].calls[-1] == "add_markers"
    # a second boundary term cannot join the collection: the first is committed, the second runs on its own
    g.reset()
    g.add_boundary(1, vol["image"], vol["sigma"], None, float("nan"))
    g.add_boundary(2, vol["image"], vol["sigma"], None, float("nan"))
    tail = [c for c in fake[0].calls if c in ("build_voxel_graph", "add_boundary")]
    assert tail[-2:] == ["build_voxel_graph", "add_boundary"], tail
