"""TEST DOUBLE (CPU): a numpy stand-in for the device slab handle, implementing the ``slab_*`` stepping protocol of
include/medpy_b200_graphcut.h so that medpy_b200.distributed (partitioning, border messages, distributed global
relabel, termination) can be exercised over gloo without a GPU.  It is deliberately simple and slow; it is NOT a
product path (the product has no CPU solver) and lives only under tests/.  Weights come from the oracle."""
import numpy

from oracle import energy_terms as et

HINF = 0x3FFFFFFF
NAMES = {v: k for k, v in {
    "difference_linear": 0, "difference_exponential": 1, "difference_division": 2, "difference_power": 3,
    "maximum_linear": 4, "maximum_exponential": 5, "maximum_division": 6, "maximum_power": 7}.items()}


class FakeSlab:
    def __init__(self, shape, z0, z1):
        self.gshape = tuple(shape)
        self.z0, self.z1 = z0, z1
        self.glo, self.ghi = z0 > 0, z1 < shape[0]
        self.shape = ((z1 - z0) + int(self.glo) + int(self.ghi),) + tuple(shape[1:])
        self.own0 = int(self.glo)
        self.own1 = self.own0 + (z1 - z0)
        self.n = int(numpy.prod(self.shape))
        self.nd = len(self.shape)
        self.reset()

    # ---- terms -------------------------------------------------------------------------------------------
    def reset(self):
        self.cap = numpy.zeros((2 * self.nd,) + self.shape)   # [2*axis + (0:-,1:+)]
        self.tr = numpy.zeros(self.shape)
        self.flow_const = 0.0
        self.owned = numpy.zeros(self.shape, bool)
        self.owned[self.own0:self.own1] = True

    def _np(self, a):
        return a.numpy() if hasattr(a, "numpy") and not isinstance(a, numpy.ndarray) else numpy.asarray(a)

    def _tw(self, src, snk, where=None):
        tr = self.tr.ravel()
        before = 0.0
        # constants only over owned voxels: replay on a copy restricted to them
        own = self.owned.ravel() if where is None else (self.owned.ravel() & where)
        tr_own = tr.copy()
        self.flow_const = et.add_tweights_pass(tr_own, self.flow_const, src, snk, where=own)
        et.add_tweights_pass(tr, before, src, snk, where=where)
        self.tr = tr.reshape(self.shape)

    def add_regional_probability(self, prob, alpha, compute_f32):
        src, snk = et.regional_probability_tweights(self._np(prob), alpha)
        self._tw(src, snk)

    def add_boundary(self, kind, image, sigma, spacing, norm):
        img = self._np(image)
        w = et.boundary_weights(NAMES[int(kind)], img, sigma, tuple(spacing) if spacing else False)
        for d in range(self.nd):
            lo = [slice(None)] * self.nd
            hi = [slice(None)] * self.nd
            lo[d] = slice(0, -1)
            hi[d] = slice(1, None)
            self.cap[2 * d + 1][tuple(lo)] += w[d]
            self.cap[2 * d][tuple(hi)] += w[d]

    def add_markers(self, fg, bg):
        if fg is not None:
            self._tw(65535.0, 0.0, where=self._np(fg).astype(bool).ravel())
        if bg is not None:
            self._tw(0.0, 65535.0, where=self._np(bg).astype(bool).ravel())

    def build_voxel_graph(self, prob, alpha, compute_f32, kind, image, sigma, spacing, norm, fg, bg):
        """mgc_build_voxel_graph: regional term, boundary term, markers, in the reference's order."""
        if prob is not None:
            self.add_regional_probability(prob, alpha, compute_f32)
        if kind >= 0:
            self.add_boundary(kind, image, sigma, spacing, norm)
        self.add_markers(fg, bg)

    # ---- stepping ------------------------------------------------------------------------------------------
    def slab_plane_elems(self):
        return int(numpy.prod(self.shape[1:]))

    def slab_begin(self):
        out = self.cap.sum(axis=0)
        self.excess = numpy.where(self.tr > 0, numpy.minimum(self.tr, out), 0.0) * self.owned
        self.absorbed = numpy.zeros(self.shape)
        self.height = numpy.full(self.shape, HINF, dtype=numpy.int64)

    def _sinkres(self):
        return numpy.maximum(-self.tr, 0) - self.absorbed

    def _shift(self, a, d, sgn, fill):
        """value of the neighbour in direction (d, sgn) for every voxel"""
        out = numpy.full(a.shape, fill, dtype=a.dtype)
        src = [slice(None)] * self.nd
        dst = [slice(None)] * self.nd
        if sgn > 0:
            src[d] = slice(1, None); dst[d] = slice(0, -1)
        else:
            src[d] = slice(0, -1); dst[d] = slice(1, None)
        out[tuple(dst)] = a[tuple(src)]
        return out

    def slab_relabel_begin(self):
        self.height = numpy.where(self.owned & (self._sinkres() > 0), 1, HINF).astype(numpy.int64)

    def slab_relabel_relax(self, want_changed=False):
        any_change = False
        while True:
            best = self.height.copy()
            for d in range(self.nd):
                for s, k in ((-1, 2 * d), (1, 2 * d + 1)):
                    hn = self._shift(self.height, d, s, HINF)
                    cand = numpy.where((self.cap[k] > 0) & (hn < HINF), hn + 1, HINF)
                    best = numpy.minimum(best, cand)
            best = numpy.where(self.owned, best, self.height)
            if (best == self.height).all():
                break
            self.height = best
            any_change = True
        return int(any_change)

    def slab_count_active(self):
        return int(((self.excess > 0) & (self.height < HINF) & self.owned).sum())

    def slab_push(self, n):
        for _ in range(8 * n):
            act = numpy.argwhere((self.excess > 0) & (self.height < HINF) & self.owned)
            if act.size == 0:
                return
            h_snap = self.height.copy()
            for idx in map(tuple, act):
                e = self.excess[idx]
                r = self._sinkres()[idx]
                if r > 0:
                    d_ = min(e, r)
                    self.absorbed[idx] = max(-self.tr[idx], 0) if d_ == r else self.absorbed[idx] + d_
                    e -= d_
                    self.excess[idx] -= d_
                minh = HINF
                for d in range(self.nd):
                    for s, k in ((-1, 2 * d), (1, 2 * d + 1)):
                        c = self.cap[k][idx]
                        if c <= 0:
                            continue
                        nb = list(idx); nb[d] += s; nb = tuple(nb)
                        hw = h_snap[nb]
                        if hw < h_snap[idx] and e > 0:
                            dd = min(e, c)
                            self.cap[k][idx] -= dd
                            self.cap[k ^ 1][nb] += dd
                            self.excess[nb] += dd
                            self.excess[idx] -= dd
                            e -= dd
                        if self.cap[k][idx] > 0:
                            minh = min(minh, hw)
                if e > 0:
                    self.height[idx] = HINF if minh >= HINF else minh + 1

    def slab_pack(self, h_lo, f_lo, h_hi, f_hi):
        P = self.slab_plane_elems()
        if self.glo and not isinstance(h_lo, int):
            h_lo.numpy()[:] = self.height[self.own0].ravel()[:P]
            f_lo.numpy()[:] = self.excess[self.own0 - 1].ravel()
            self.excess[self.own0 - 1] = 0
        if self.ghi and not isinstance(h_hi, int):
            h_hi.numpy()[:] = self.height[self.own1 - 1].ravel()
            f_hi.numpy()[:] = self.excess[self.own1].ravel()
            self.excess[self.own1] = 0

    def slab_count_active_dev(self, out):
        out.numpy()[0] = self.slab_count_active()

    def slab_unpack(self, h_lo, f_lo, h_hi, f_hi, changed_out=0):
        changed = 0
        psh = self.shape[1:]
        if self.glo and not isinstance(h_lo, int):
            hn = h_lo.numpy().reshape(psh).astype(numpy.int64)
            if (self.height[self.own0 - 1] != hn).any():
                changed = 1
            self.height[self.own0 - 1] = hn
            f = f_lo.numpy().reshape(psh)
            self.excess[self.own0] += f
            self.cap[0][self.own0] += f
        if self.ghi and not isinstance(h_hi, int):
            hn = h_hi.numpy().reshape(psh).astype(numpy.int64)
            if (self.height[self.own1] != hn).any():
                changed = 1
            self.height[self.own1] = hn
            f = f_hi.numpy().reshape(psh)
            self.excess[self.own1 - 1] += f
            self.cap[1][self.own1 - 1] += f
        if changed and not isinstance(changed_out, int):
            changed_out.numpy()[0] = 1
        return changed

    def slab_finish(self):
        self._mask = (self.height[self.own0:self.own1] >= HINF).astype(numpy.uint8)
        return float(self.flow_const + self.absorbed[self.own0:self.own1].sum())

    def get_mask(self):
        return self._mask
