"""TEST DOUBLES of the two native classes the region-graph host layer talks to (``_mgc.LabelImage``, ``_mgc.SparseGraph``),
built on the oracle: they let the CPU suite drive energy_label / graph_from_labels / GCGraph / the CLIs end to end --
every line of the Python glue -- where no GPU exists.  They are NOT a fallback: only tests install them."""
import numpy

from oracle import energy_label_terms as elt
from oracle import solvers

LABELS_ADJACENCY, LABELS_STAWIASKI, LABELS_STAWIASKI_DIRECTED = 0, 1, 2
SUM_BINCOUNT, SUM_PAIRWISE = 0, 1


class FakeLabelImage:
    def __init__(self, label_image, device=-1):
        lab = numpy.asarray(label_image)
        assert lab.dtype == numpy.int32, "the native class only takes int32"
        self.lab = lab
        self.k = elt.check_label_image(lab)     # AttributeError like MGC_E_LABELS
        self.shape = list(lab.shape)

    def region_count(self):
        return self.k

    def boundary(self, kind, values=None, directedness=0.0):
        if kind == LABELS_ADJACENCY:
            lo, hi = _adjacency(self.lab)
            z = numpy.zeros(lo.size)
            return (lo - 1).astype(numpy.int32), (hi - 1).astype(numpy.int32), z, z.copy()
        values = numpy.asarray(values)
        assert values.shape == self.lab.shape
        if kind == LABELS_STAWIASKI:
            calls = elt.stawiaski_calls(self.lab, values)
        else:
            calls = elt.stawiaski_directed_calls(self.lab, values, directedness)
        lo, hi, a, b = elt.merge_edges(*calls)
        return lo.astype(numpy.int32), hi.astype(numpy.int32), a, b

    def region_sums(self, values, mode):
        values = numpy.asarray(values)
        flat = self.lab.ravel()
        counts = numpy.bincount(flat, minlength=self.k + 1)[1:].astype(numpy.int64)
        if mode == SUM_BINCOUNT or values.dtype.kind != "f":
            sums = numpy.zeros(self.k + 1)
            numpy.add.at(sums, flat, values.ravel().astype(numpy.float64))
            return sums[1:], counts
        order = numpy.argsort(flat, kind="stable")
        bounds = numpy.searchsorted(flat[order], numpy.arange(1, self.k + 2))
        v = values.ravel()[order]
        return numpy.asarray([float(numpy.sum(v[bounds[r]:bounds[r + 1]])) for r in range(self.k)]), counts

    def region_flags(self, markers):
        m = numpy.asarray(markers).astype(bool)
        flags = numpy.zeros(self.k, numpy.uint8)
        flags[numpy.unique(self.lab[m] - 1)] = 1
        return flags

    def apply(self, per_region):
        return numpy.asarray(per_region, dtype=numpy.uint8)[self.lab - 1]


def _adjacency(lab):
    keys = []
    for dim in range(lab.ndim):
        a = [slice(None)] * lab.ndim
        b = [slice(None)] * lab.ndim
        a[dim], b[dim] = slice(None, -1), slice(1, None)
        kf, kt = lab[tuple(a)].ravel().astype(numpy.int64), lab[tuple(b)].ravel().astype(numpy.int64)
        valid = kf != kt
        keys.append(numpy.minimum(kf, kt)[valid] * (1 << 32) + numpy.maximum(kf, kt)[valid])
    keys = numpy.unique(numpy.concatenate(keys))
    return keys >> 32, keys & 0xffffffff


class FakeSparseGraph:
    def __init__(self, n_nodes, device=-1):
        self.n = int(n_nodes)
        self.reset()

    def reset(self):
        self.e = [numpy.zeros(0, numpy.int64), numpy.zeros(0, numpy.int64), numpy.zeros(0), numpy.zeros(0)]
        self.tw = []
        self.result = None

    def sum_edges(self, i, j, cap, rev):
        i, j = numpy.asarray(i, dtype=numpy.int64), numpy.asarray(j, dtype=numpy.int64)
        if i.size and (min(i.min(), j.min()) < 0 or max(i.max(), j.max()) >= self.n or (i == j).any()):
            raise ValueError("Invalid node id")
        for k, a in enumerate((i, j, numpy.asarray(cap, dtype=float), numpy.asarray(rev, dtype=float))):
            self.e[k] = numpy.concatenate([self.e[k], a])
        self.result = None

    def add_tweights(self, nodes, src, snk):
        src = numpy.asarray(src, dtype=float)
        nodes = numpy.arange(src.size) if nodes is None else numpy.asarray(nodes, dtype=numpy.int64)
        self.tw.append((nodes, src, numpy.asarray(snk, dtype=float)))
        self.result = None

    def _solve(self):
        if self.result is None:
            flow, mask, _ = solvers.solve_sparse(self.n, self.e[0], self.e[1], self.e[2], self.e[3], self.tw)
            self.result = (flow, mask)
        return self.result

    def maxflow(self):
        return self._solve()[0]

    def get_mask(self):
        return self._solve()[1].copy()

    def what_segment(self, i):
        return 0 if self._solve()[1][int(i)] else 1

    def get_edge(self, i, j):
        lo, hi, a, b = elt.merge_edges(*self.e)
        for x, y, u, v in zip(lo, hi, a, b):
            if (x, y) == (min(i, j), max(i, j)):
                return u if i < j else v
        return 0.0

    def get_trcap(self, i):
        return elt.add_tweights_replay(self.n, self.tw)[0][int(i)]

    def get_node_num(self):
        return self.n

    def get_arc_num(self):
        return 2 * elt.merge_edges(*self.e)[0].size

    def stats(self):
        return dict(n_nodes=self.n, global_relabels=0, push_sweeps=0, kernel_launches=0, ms_solve=0.0)


# ---- lattice double (``_mgc.Graph``) ------------------------------------------------------------------------------------
from oracle import energy_terms as et  # noqa: E402

_KINDS = ["difference_linear", "difference_exponential", "difference_division", "difference_power",
          "maximum_linear", "maximum_exponential", "maximum_division", "maximum_power"]


class FakeGraph:
    """The dense-lattice native class on top of oracle/energy_terms.py + oracle/bk_lattice.c."""

    def __init__(self, shape, device=-1):
        self.shape = [int(s) for s in shape]
        self.n = int(numpy.prod(self.shape))
        self.reset()

    def reset(self):
        self.wf = [numpy.zeros(self.n) for _ in self.shape]
        self.wb = [numpy.zeros(self.n) for _ in self.shape]
        self.tr = numpy.zeros(self.n)
        self.flow = 0.0
        self.result = None

    def set_option(self, option, value):
        pass

    def check_deferred(self):
        pass

    def set_stream(self, s):
        pass

    def synchronize(self):
        pass

    def add_regional_probability(self, prob, alpha, compute_f32):
        prob = numpy.asarray(prob)
        assert bool(compute_f32) == (prob.dtype == numpy.float32)
        src, snk = et.regional_probability_tweights(prob, alpha)
        self.flow = et.add_tweights_pass(self.tr, self.flow, src, snk)
        self.result = None

    def add_tweights_dense(self, src, snk):
        self.flow = et.add_tweights_pass(self.tr, self.flow, numpy.asarray(src, dtype=float).ravel(), numpy.asarray(snk, dtype=float).ravel())
        self.result = None

    def add_markers(self, fg, bg):
        if fg is not None and numpy.asarray(fg).any():
            self.flow = et.add_tweights_pass(self.tr, self.flow, 65535.0, 0.0, where=numpy.asarray(fg).ravel().astype(bool))
        if bg is not None and numpy.asarray(bg).any():
            self.flow = et.add_tweights_pass(self.tr, self.flow, 0.0, 65535.0, where=numpy.asarray(bg).ravel().astype(bool))
        self.result = None

    def add_boundary(self, kind, image, sigma, spacing, norm):
        image = numpy.asarray(image)
        w = et.boundary_weights(_KINDS[kind], image, sigma, spacing if spacing else False)
        for arr in w:
            if (arr <= 0).any():
                raise ValueError("Negative or zero weights are not allowed.")
        for d, full in enumerate(et.dense_axis_arrays(tuple(self.shape), w)):
            self.wf[d] += full
            self.wb[d] += full
        self.result = None

    def build_voxel_graph(self, prob, alpha, compute_f32, kind, image, sigma, spacing, norm, fg, bg):
        """mgc_build_voxel_graph: the three terms in the reference's order."""
        if prob is not None:
            self.add_regional_probability(prob, alpha, compute_f32)
        if kind >= 0:
            self.add_boundary(kind, image, sigma, spacing, norm)
        self.add_markers(fg, bg)

    def can_fuse(self):
        return False

    def add_nweights_dense(self, axis, fwd, bwd):
        fwd, bwd = numpy.asarray(fwd, dtype=float).ravel(), numpy.asarray(bwd, dtype=float).ravel()
        if (fwd < 0).any() or (bwd < 0).any():
            raise ValueError("Negative or zero weights are not allowed.")
        stride = int(numpy.prod(self.shape[axis + 1:]))
        last = (numpy.arange(self.n) // stride) % self.shape[axis] == self.shape[axis] - 1
        self.wf[axis] += numpy.where(last, 0.0, fwd)
        self.wb[axis] += numpy.where(last, 0.0, bwd)
        self.result = None

    def _solve(self):
        if self.result is None:
            prob = dict(shape=tuple(self.shape), wf=self.wf, wb=self.wb, tr=self.tr.copy(), flow_const=self.flow)
            flow, mask, _ = solvers.solve_port(prob)
            self.result = (flow, mask)
        return self.result

    def maxflow(self):
        return self._solve()[0]

    def get_mask(self):
        return self._solve()[1].copy()

    def what_segment(self, i):
        return 0 if self._solve()[1].flat[int(i)] else 1

    def _axis(self, i, j):
        lo, d = min(i, j), abs(i - j)
        for axis in range(len(self.shape)):
            stride = int(numpy.prod(self.shape[axis + 1:]))
            if d == stride and self.shape[axis] > 1 and (lo // stride) % self.shape[axis] < self.shape[axis] - 1:
                return axis, lo
        return None, lo

    def get_edge(self, i, j):
        axis, lo = self._axis(int(i), int(j))
        if axis is None:
            return 0.0
        return float(self.wf[axis][lo] if i < j else self.wb[axis][lo])

    def get_trcap(self, i):
        return float(self.tr[int(i)])

    def get_node_num(self):
        return self.n

    def get_arc_num(self):
        return 2 * sum((self.n // s) * (s - 1) for s in self.shape if s > 1)

    def stats(self):
        return dict(n_voxels=self.n, kernel_launches=0, push_sweeps=0, global_relabels=0, flow_const=self.flow)
