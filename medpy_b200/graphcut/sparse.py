"""General sparse graphs behind the ``GraphDouble`` API.

The reference's ``GraphDouble`` (lib/maxflow/src/wrapper.cpp:63-83) is a general graph; the voxel path only ever fills it
with lattice edges, which is why ``medpy_b200.graphcut.maxflow.GraphDouble`` stores a dense lattice.  Graphs that are
not lattices -- the region adjacency graph of ``graph_from_labels`` (generate.py:177-338) and graphs users assemble
edge by edge (tests/graphcut_/graph.py:47) -- are held by ``SparseGraphDouble``: calls are staged in host lists and
cross the C ABI in bulk (``mgc_sparse_sum_edges`` / ``mgc_sparse_add_tweights``, which apply them in order with the
reference's accumulation semantics); ``maxflow()`` runs a CSR push-relabel on the device (csrc/gc_sparse.cuh).

No CPU solver: the first call that needs a result creates the native graph and raises ``RuntimeError`` without a CUDA
device or built extension.
"""
import numpy

from .maxflow import _termtype

__all__ = ["SparseGraphDouble"]


class SparseGraphDouble:
    """``GraphDouble(node_num_max, edge_num_max)`` for arbitrary node pairs."""

    termtype = _termtype

    def __init__(self, node_num_max, edge_num_max=0, device=-1):
        self._n = int(node_num_max)
        if self._n < 1:
            raise ValueError("a graph needs at least one node")
        self._edges = int(edge_num_max)
        self._device = device
        self._native = None
        self._ops = []            # staged calls in order: ("e", i, j, cap, rev) / ("t", nodes|None, src, snk) arrays
        self._e = ([], [], [], [])
        self._t = ([], [], [])
        self._mask = None

    # ------------------------------------------------------------------ staging
    def _nat(self):
        if self._native is None:
            from .. import _lib  # raises ImportError loudly when the extension is not built
            self._native = _lib._mgc.SparseGraph(self._n, self._device)
        return self._native

    def _close_edges(self):
        if self._e[0]:
            self._ops.append(("e", numpy.asarray(self._e[0], dtype=numpy.int32), numpy.asarray(self._e[1], dtype=numpy.int32),
                              numpy.asarray(self._e[2], dtype=numpy.float64), numpy.asarray(self._e[3], dtype=numpy.float64)))
            self._e = ([], [], [], [])

    def _close_tweights(self):
        if self._t[0]:
            self._ops.append(("t", numpy.asarray(self._t[0], dtype=numpy.int32), numpy.asarray(self._t[1], dtype=numpy.float64),
                              numpy.asarray(self._t[2], dtype=numpy.float64)))
            self._t = ([], [], [])

    def _flush(self):
        # edges and t-links are independent state in the reference (graph.h:415-480), so the two staged groups may be
        # applied one after the other; inside a group the call order is kept
        self._close_edges()
        self._close_tweights()
        ops, self._ops = self._ops, []
        for op in ops:
            if op[0] == "e":
                self._nat().sum_edges(op[1], op[2], op[3], op[4])
            else:
                self._nat().add_tweights(op[1], op[2], op[3])

    def _check_node(self, i):
        if i < 0 or i >= self._n:
            raise ValueError("Invalid node id of {}. Valid values are 0 to {}.".format(i, self._n - 1))

    # ------------------------------------------------------------------ reference GraphDouble API
    def add_node(self, num=1):
        """graph.h:388-413: nodes exist from construction; returns the id of the first one."""
        return 0

    def add_tweights(self, i, cap_source, cap_sink):
        """graph.h:415-425."""
        i = int(i)
        self._check_node(i)
        self._t[0].append(i)
        self._t[1].append(float(cap_source))
        self._t[2].append(float(cap_sink))
        self._mask = None

    def stage_tweights_many(self, ids, cap_source, cap_sink):
        """add_tweights(v, cap_source, cap_sink) for every v in ids, in order (ids already range-checked)."""
        ids = numpy.asarray(ids, dtype=numpy.int32).ravel()
        self.add_tweights_bulk(ids, numpy.full(ids.size, float(cap_source)), numpy.full(ids.size, float(cap_sink)))

    def add_tweights_bulk(self, nodes, src, snk):
        """One add_tweights call per entry, in array order; ``nodes`` None means 0..len-1."""
        src = numpy.ascontiguousarray(src, dtype=numpy.float64).ravel()
        snk = numpy.ascontiguousarray(snk, dtype=numpy.float64).ravel()
        if nodes is not None:
            nodes = numpy.ascontiguousarray(nodes, dtype=numpy.int32).ravel()
            if nodes.size and (nodes.min() < 0 or nodes.max() >= self._n):
                raise ValueError("Invalid node id of {} or {}. Valid values are 0 to {}.".format(nodes.max(), nodes.min(), self._n - 1))
        elif src.size > self._n:
            raise ValueError("Invalid node id of {}. Valid values are 0 to {}.".format(src.size - 1, self._n - 1))
        self._close_tweights()
        self._ops.append(("t", nodes, src, snk))
        self._mask = None

    def sum_edge(self, i, j, cap, rev_cap):
        """graph.h:456-480: creates the arc pair on the first call for (i, j), accumulates afterwards."""
        i, j = int(i), int(j)
        if i < 0 or j < 0 or i >= self._n or j >= self._n or i == j:
            raise ValueError("invalid node ids ({}, {})".format(i, j))
        self._e[0].append(i)
        self._e[1].append(j)
        self._e[2].append(float(cap))
        self._e[3].append(float(rev_cap))
        self._mask = None

    add_edge = sum_edge  # graph.h:427-454: parallel arcs carry the summed capacity

    def sum_edges_bulk(self, i, j, cap, rev_cap):
        """One sum_edge call per entry, in array order."""
        i = numpy.ascontiguousarray(i, dtype=numpy.int32).ravel()
        j = numpy.ascontiguousarray(j, dtype=numpy.int32).ravel()
        cap = numpy.ascontiguousarray(cap, dtype=numpy.float64).ravel()
        rev_cap = numpy.ascontiguousarray(rev_cap, dtype=numpy.float64).ravel()
        if not (i.size == j.size == cap.size == rev_cap.size):
            raise ValueError("edge arrays differ in length")
        if i.size and (min(i.min(), j.min()) < 0 or max(i.max(), j.max()) >= self._n or (i == j).any()):
            raise ValueError("invalid node ids in the edge arrays")
        self._close_edges()
        self._ops.append(("e", i, j, cap, rev_cap))
        self._mask = None

    def maxflow(self):
        """Graph::maxflow (maxflow.cpp:471-604): min-cut energy including the add_tweights constants."""
        self._flush()
        return self._nat().maxflow()

    def get_mask(self):
        """uint8[n]: 0 where what_segment == SINK else 1 (the loop of bin/medpy_graphcut_label.py:139-145 in bulk)."""
        if self._mask is None:
            self.maxflow()
            self._mask = self._nat().get_mask()
        return self._mask

    def what_segment(self, i, default_segm=None):
        """graph.h:560-571."""
        i = int(i)
        self._check_node(i)
        return _termtype.SOURCE if self.get_mask()[i] else _termtype.SINK

    def reset(self):
        self._ops = []
        self._e = ([], [], [], [])
        self._t = ([], [], [])
        self._mask = None
        if self._native is not None:
            self._native.reset()

    def get_edge(self, i, j):
        self._flush()
        return self._nat().get_edge(int(i), int(j))

    def get_trcap(self, i):
        self._flush()
        return self._nat().get_trcap(int(i))

    def get_node_num(self):
        return self._n

    def get_arc_num(self):
        self._flush()
        return self._nat().get_arc_num()

    def stats(self):
        return self._nat().stats()
