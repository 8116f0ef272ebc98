"""GPU parity tests (``-m gpu``) of the region-graph path: the device-built region adjacency
graph and the sparse push-relabel behind ``graph_from_labels`` / ``energy_label`` / ``GCGraph`` against

(1) golden vectors recorded from the unmodified reference (tests/golden/golden_labels_v1.npz: every set_nweight /
    set_tweight call of the four label terms, and whole graph_from_labels -> maxflow -> what_segment runs),
(2) the numpy oracle (oracle/energy_label_terms.py) and the real reference BK (oracle/_ref, when built) on seeded
    inputs, and the host emulation's random graphs.

Tolerances: region masks, adjacency, means-based weights, float32-gradient Stawiaski weights and atlas t-links are
bit-exact; weights from float64 / integer gradients (and all of the directed term) are within 4 ulp per contribution
(the reference squares with libm's pow, which is not always the correctly rounded product the device forms); energies
within 1e-9 relative, exact for integer capacities.
"""
import os
import sys

import numpy
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import energy_label_terms as elt  # noqa: E402
from oracle import solvers  # noqa: E402

pytestmark = pytest.mark.gpu
os.environ.setdefault("MEDPY_GC_SPARSE_TIMEOUT", "30")    # test graphs are tiny: fail fast instead of spinning

G = numpy.load(os.path.join(HERE, "golden", "golden_labels_v1.npz"))
NAMES = [str(n) for n in G["names"]]
FULL = [n for n in NAMES if n + "/directed" in G.files]


def _gc():
    import medpy_b200.graphcut as gc
    return gc


def _label(nm):
    lab = G[nm + "/label"]
    return numpy.asfortranarray(lab) if bool(G[nm + "/label_forder"]) else lab


class Recorder:
    """Someone else's graph object: receives one set_nweight / set_tweight per region pair / region."""

    def __init__(self):
        self.n = {}
        self.t = []

    def set_nweight(self, a, b, w1, w2):
        assert (a, b) not in self.n
        self.n[(a, b)] = (w1, w2)

    def set_tweight(self, node, ws, wk):
        self.t.append((node, ws, wk))


def _merged(calls):
    """golden call list (i, j, there, back) -> {(lo, hi): (cap lo->hi, cap hi->lo)} with the reference's += order."""
    lo, hi, a, b = elt.merge_edges(calls[:, 0].astype(numpy.int64), calls[:, 1].astype(numpy.int64), calls[:, 2], calls[:, 3])
    return {(int(x), int(y)): (u, v) for x, y, u, v in zip(lo, hi, a, b)}


def _assert_edges(got, want, exact, n_contrib=None):
    assert set(got) == set(want)
    for key, (u, v) in want.items():
        gu, gv = got[key]
        if exact:
            assert gu == u and gv == v, (key, gu, u, gv, v)
        else:
            assert gu == pytest.approx(u, rel=1e-14, abs=1e-320) and gv == pytest.approx(v, rel=1e-14, abs=1e-320), key


@pytest.mark.parametrize("nm", NAMES)
def test_stawiaski_edges_vs_reference(nm):
    el = _gc().energy_label
    r = Recorder()
    el.boundary_stawiaski(r, _label(nm), G[nm + "/image"])
    exact = G[nm + "/image"].dtype == numpy.float32 or not G[nm + "/image"].any()
    _assert_edges(r.n, _merged(G[nm + "/stawiaski"]), exact)


@pytest.mark.parametrize("nm", NAMES)
def test_difference_of_means_edges_vs_reference(nm):
    el = _gc().energy_label
    r = Recorder()
    el.boundary_difference_of_means(r, _label(nm), G[nm + "/image"])
    _assert_edges(r.n, _merged(G[nm + "/means"]), True)


@pytest.mark.parametrize("nm", FULL)
def test_directed_edges_vs_reference(nm):
    el = _gc().energy_label
    r = Recorder()
    el.boundary_stawiaski_directed(r, _label(nm), (G[nm + "/image"], float(G[nm + "/directedness"])))
    _assert_edges(r.n, _merged(G[nm + "/directed"]), False)


@pytest.mark.parametrize("nm", FULL)
def test_atlas_tweights_vs_reference(nm):
    el = _gc().energy_label
    r = Recorder()
    el.regional_atlas(r, _label(nm), (G[nm + "/prob"], float(G[nm + "/alpha"])))
    got = numpy.asarray(r.t, dtype=numpy.float64)
    assert numpy.array_equal(got.view(numpy.uint64), G[nm + "/atlas"].view(numpy.uint64))


@pytest.mark.parametrize("tag", ["cut_stawiaski", "cut_means", "cut_directed_atlas"])
@pytest.mark.parametrize("nm", FULL)
def test_graph_from_labels_whole_cut_vs_reference(nm, tag):
    gc = _gc()
    el = gc.energy_label
    lab, img = _label(nm), G[nm + "/image"]
    if tag == "cut_stawiaski":
        kw = dict(boundary_term=el.boundary_stawiaski, boundary_term_args=img)
    elif tag == "cut_means":
        kw = dict(boundary_term=el.boundary_difference_of_means, boundary_term_args=img)
    else:
        kw = dict(boundary_term=el.boundary_stawiaski_directed, boundary_term_args=(img, float(G[nm + "/directedness"])),
                  regional_term=el.regional_atlas, regional_term_args=(G[nm + "/prob"], float(G[nm + "/alpha"])))
    g = gc.graph_from_labels(lab, G[nm + "/fg"], G[nm + "/bg"], **kw)
    flow = g.maxflow()
    want_mask = G[nm + "/" + tag + "_mask"]
    assert numpy.array_equal(g.get_mask(), want_mask)
    assert flow == pytest.approx(float(G[nm + "/" + tag + "_flow"]), rel=1e-9, abs=1e-300)
    # the reference's read-out loop (bin/medpy_graphcut_label.py:139-145) and the voxel mask
    seg = [0 if g.termtype.SINK == g.what_segment(v) else 1 for v in range(int(lab.max()))]
    assert seg == want_mask.tolist()
    vox = gc.label_cut_mask(g)
    assert vox.shape == lab.shape and numpy.array_equal(vox, want_mask[numpy.asarray(lab) - 1])
    assert g.maxflow() == flow          # idempotent


def test_label_image_checks():
    gc = _gc()
    el = gc.energy_label
    for bad in ([[1, 4, 8], [1, 3, 10], [1, 3, 10]], [[2, 3, 4], [2, 3, 4], [2, 3, 4]]):   # tests/graphcut_/energy_label.py:106-122
        for term, args in ((el.boundary_stawiaski, None), (el.boundary_difference_of_means, None),
                           (el.boundary_stawiaski_directed, (None, None)), (el.regional_atlas, (None, None))):
            with pytest.raises(AttributeError):
                term(None, numpy.asarray(bad), args)
        with pytest.raises(AttributeError):
            gc.graph_from_labels(numpy.asarray(bad), numpy.zeros((3, 3), bool), numpy.zeros((3, 3), bool))
    lab = numpy.asarray([[1, 2], [1, 2]])
    with pytest.raises(ValueError):     # no foreground marker: max() of an empty sequence in the reference (graph.py:334)
        gc.graph_from_labels(lab, numpy.zeros((2, 2), bool), numpy.ones((2, 2), bool),
                             boundary_term=el.boundary_stawiaski, boundary_term_args=numpy.zeros((2, 2)))
    with pytest.raises(ValueError):     # gradient of another shape
        el.boundary_stawiaski(Recorder(), lab, numpy.zeros((3, 2)))
    with pytest.raises(ValueError):     # size-1 axis: numpy.vectorize on an empty slice (energy_label.py:325-328)
        el.boundary_stawiaski_directed(Recorder(), numpy.asarray([[1, 2, 3]]), (numpy.zeros((1, 3)), -0.1))


def supervoxel_volume(shape, cell, seed):
    """Jittered block labels 1..K, a two-blob image and its gradient magnitude (float32)."""
    from medpy_b200 import synthetic
    rng = numpy.random.default_rng(seed)
    grids = numpy.meshgrid(*[numpy.arange(s) for s in shape], indexing="ij")
    jit = [g + rng.integers(-1, 2, size=shape) for g in grids]
    blocks = [numpy.clip(j, 0, s - 1) // cell for j, s in zip(jit, shape)]
    nb = [-(-s // cell) for s in shape]
    lab = numpy.zeros(shape, numpy.int64)
    for b, n in zip(blocks, nb):
        lab = lab * n + b
    _, inv = numpy.unique(lab, return_inverse=True)
    lab = (inv + 1).reshape(shape).astype(numpy.int32)
    img = synthetic.two_blob_volume(shape, seed, with_prob=False)["image"]
    grad = numpy.sqrt(sum(numpy.gradient(img.astype(numpy.float64))[d] ** 2 for d in range(len(shape)))).astype(numpy.float32)
    return lab, img.astype(numpy.float32), grad


def test_supervoxel_volume_vs_oracle_and_bk():
    """48^3 volume, ~1700 regions: device RAG == oracle restatement bit for bit (float32 gradient); cut == real BK."""
    gc = _gc()
    el = gc.energy_label
    shape = (48, 48, 48)
    lab, img, grad = supervoxel_volume(shape, 4, 3)
    r = Recorder()
    el.boundary_stawiaski(r, lab, grad)
    oi, oj, ow, owr = elt.stawiaski_calls(lab, grad)
    lo, hi, a, b = elt.merge_edges(oi, oj, ow, owr)
    want = {(int(x), int(y)): (u, v) for x, y, u, v in zip(lo, hi, a, b)}
    _assert_edges(r.n, want, True)
    # markers: a ball inside the first blob, the volume's faces
    zz, yy, xx = numpy.meshgrid(*[numpy.arange(s) for s in shape], indexing="ij")
    fg = (zz - 14) ** 2 + (yy - 14) ** 2 + (xx - 14) ** 2 <= 9
    bg = numpy.zeros(shape, bool)
    bg[0], bg[-1], bg[:, 0], bg[:, -1], bg[:, :, 0], bg[:, :, -1] = True, True, True, True, True, True
    g = gc.graph_from_labels(lab, fg, bg, boundary_term=el.boundary_stawiaski, boundary_term_args=grad)
    flow = g.maxflow()
    mask = g.get_mask()
    assert 0 < int(mask.sum()) < mask.size
    n = int(lab.max())
    fgr, bgr = elt.marker_regions(lab, fg), elt.marker_regions(lab, bg)
    tw = [(fgr, numpy.full(fgr.size, 65535.0), numpy.zeros(fgr.size)), (bgr, numpy.zeros(bgr.size), numpy.full(bgr.size, 65535.0))]
    rflow, rmask, _ = solvers.solve_sparse(n, oi, oj, ow, owr, tw)
    assert numpy.array_equal(mask, rmask)
    assert flow == pytest.approx(rflow, rel=1e-9)
    # means term on the same volume: device means and adjacency against the oracle
    r = Recorder()
    el.boundary_difference_of_means(r, lab, img)
    mi, mj, mw, _ = elt.difference_of_means_calls(lab, img)
    _assert_edges(r.n, {(int(x), int(y)): (u, u) for x, y, u in zip(mi, mj, mw)}, True)


def test_wrapper_functions_vs_reference():
    """graphcut_stawiaski on the whole volume and graphcut_split over 8 overlapping sub-volumes (wrapper.py:72-329)."""
    gc = _gc()
    from medpy_b200.errors import ArgumentError
    lab, grad, fg, bg = G["split/label"], G["split/gradient"], G["split/fg"], G["split/bg"]
    whole = gc.graphcut_stawiaski(lab, grad, fg, bg)
    assert whole.dtype == numpy.bool_ and numpy.array_equal(whole, G["split/whole_mask"].astype(bool))
    assert numpy.array_equal(gc.graphcut_stawiaski((lab, grad, fg, bg)), whole)          # the tuple form the pool used
    split = gc.graphcut_split(gc.graphcut_stawiaski, lab, grad, fg, bg, 10, 3, 2)
    assert numpy.array_equal(split, G["split/split_mask"].astype(bool))
    with pytest.raises(ArgumentError):
        gc.graphcut_split(gc.graphcut_stawiaski, lab, grad, fg, bg, 9, 3)
    with pytest.raises(ArgumentError):
        gc.graphcut_split(gc.graphcut_stawiaski, lab, grad, fg, bg, 10, 10)
    with pytest.raises(ArgumentError):
        gc.graphcut_split(gc.graphcut_stawiaski, lab, grad[1:], fg, bg, 10, 3)
    with pytest.raises(ArgumentError):
        gc.graphcut_subprocesses(gc.graphcut_stawiaski, [], -1)


def random_graph(rng, n, m, integer):
    i = rng.integers(0, n, size=m)
    j = rng.integers(0, n, size=m)
    keep = i != j
    i, j = i[keep], j[keep]
    if integer:
        cap = rng.integers(1, 20, size=i.size).astype(float)
        rev = rng.integers(1, 20, size=i.size).astype(float)
        src = rng.integers(0, 30, size=n).astype(float)
        snk = rng.integers(0, 30, size=n).astype(float)
    else:
        cap = rng.uniform(1e-3, 2.0, size=i.size)
        rev = rng.uniform(1e-3, 2.0, size=i.size)
        src = rng.uniform(0, 3.0, size=n)
        snk = rng.uniform(0, 3.0, size=n)
    fg = rng.choice(n, size=max(1, n // 20), replace=False)
    bg = rng.choice(n, size=max(1, n // 20), replace=False)
    return i, j, cap, rev, src, snk, fg, bg


@pytest.mark.parametrize("seed", range(10))
def test_general_sparse_graph_vs_reference_bk(seed):
    """GCGraph used the way tests/graphcut_/graph.py uses it: arbitrary node pairs, element-wise and bulk setters."""
    gc = _gc()
    rng = numpy.random.default_rng(100 + seed)
    n = int(rng.integers(3, 3000 if seed >= 6 else 300))
    m = int(rng.integers(1, 6 * n))
    integer = seed % 2 == 0
    i, j, cap, rev, src, snk, fg, bg = random_graph(rng, n, m, integer)
    i, j = numpy.append(i, 0), numpy.append(j, n - 1)          # at least one pair that is no chain neighbour
    cap, rev = numpy.append(cap, 1.0), numpy.append(rev, 2.0)
    graph = gc.GCGraph(n, m)
    if seed % 3 == 0:                      # element-wise, like the reference's callers
        for v in range(n):
            graph.set_tweight(v, src[v], snk[v])
        for a, b, c, d in zip(i.tolist(), j.tolist(), cap.tolist(), rev.tolist()):
            graph.set_nweight(a, b, c, d)
    else:
        graph.set_tweights_bulk(numpy.arange(n), src, snk)
        graph.set_nweights_bulk(i, j, cap, rev)
    graph.set_source_nodes(fg)
    graph.set_sink_nodes(bg)
    g = graph.get_graph()
    assert g.is_sparse
    tw = [(numpy.arange(n), src, snk), (fg, numpy.full(fg.size, 65535.0), numpy.zeros(fg.size)),
          (bg, numpy.zeros(bg.size), numpy.full(bg.size, 65535.0))]
    rflow, rmask, _ = solvers.solve_sparse(n, i, j, cap, rev, tw)
    flow = g.maxflow()
    assert numpy.array_equal(g.get_mask(), rmask)
    if integer:
        assert flow == rflow
    else:
        assert flow == pytest.approx(rflow, rel=1e-9)
    # getters return the assembled values (sum_edge accumulation in call order)
    lo, hi, a, b = elt.merge_edges(i, j, cap, rev)
    k = int(rng.integers(0, lo.size))
    assert g.get_edge(int(lo[k]), int(hi[k])) == a[k] and g.get_edge(int(hi[k]), int(lo[k])) == b[k]
    assert g.get_arc_num() == 2 * lo.size and g.get_node_num() == n
    tr, _ = elt.add_tweights_replay(n, tw)
    assert g.get_trcap(int(fg[0])) == tr[int(fg[0])]


def test_sparse_graph_reference_fixture_and_reset():
    """The diamond of lib/maxflow/src/sum_edge_test.py:20-37 (flows 2, then 4 after doubling the edges) + reset."""
    from medpy_b200.graphcut.maxflow import GraphDouble
    g = GraphDouble(4, 4, sparse=True)
    g.add_tweights(0, 99, 0)
    g.add_tweights(3, 0, 99)
    for a, b in ((0, 1), (0, 2), (1, 3), (2, 3)):
        g.sum_edge(a, b, 1, 0)
    assert g.maxflow() == 2.0
    for a, b in ((0, 1), (0, 2), (1, 3), (2, 3)):
        g.sum_edge(a, b, 1, 0)
    assert g.maxflow() == 4.0
    assert [int(g.what_segment(v)) for v in range(4)] == [0, 0, 0, 1]
    g.reset()
    g.add_tweights(0, 5, 0)
    g.add_tweights(1, 0, 3)
    g.sum_edge(0, 1, 2, 2)
    assert g.maxflow() == 2.0 and g.get_mask().tolist() == [1, 0, 1, 1]


def test_sparse_c_abi_direct_ctypes():
    """The sparse entry points straight through the C ABI."""
    import ctypes
    from medpy_b200 import build
    lib = ctypes.CDLL(build.LIB)
    h = ctypes.c_void_p()
    assert lib.mgc_sparse_create(ctypes.c_int64(3), -1, ctypes.byref(h)) == 0
    i = (ctypes.c_int32 * 2)(0, 1)
    j = (ctypes.c_int32 * 2)(1, 2)
    cap = (ctypes.c_double * 2)(3.0, 1.0)
    rev = (ctypes.c_double * 2)(0.5, 0.5)
    assert lib.mgc_sparse_sum_edges(h, ctypes.c_int64(2), i, j, cap, rev) == 0
    nodes = (ctypes.c_int32 * 2)(0, 2)
    src = (ctypes.c_double * 2)(10.0, 0.0)
    snk = (ctypes.c_double * 2)(0.0, 10.0)
    assert lib.mgc_sparse_add_tweights(h, ctypes.c_int64(2), nodes, src, snk) == 0
    e = ctypes.c_double()
    assert lib.mgc_sparse_maxflow(h, ctypes.byref(e)) == 0 and e.value == 1.0
    mask = (ctypes.c_uint8 * 3)()
    assert lib.mgc_sparse_get_mask(h, mask) == 0 and list(mask) == [1, 1, 0]
    bad = (ctypes.c_int32 * 1)(7)
    assert lib.mgc_sparse_sum_edges(h, ctypes.c_int64(1), bad, j, cap, rev) == -1      # MGC_E_ARG
    lib.mgc_sparse_last_error.restype = ctypes.c_char_p
    assert b"Invalid node id" in lib.mgc_sparse_last_error(h)
    lib.mgc_sparse_destroy(h)
