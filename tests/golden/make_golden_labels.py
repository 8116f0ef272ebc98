#!/usr/bin/env python
"""Generate tests/golden/golden_labels_v1.npz by running the UNMODIFIED reference label path.

Run in the build container only (needs /root/reference and `make -C oracle pyshim`):

    python tests/golden/make_golden_labels.py

For every case we record the inputs, the exact sequence of ``set_nweight`` / ``set_tweight`` calls the reference's
energy_label terms issue (medpy/graphcut/energy_label.py, observed through a recording stand-in like the reference's own
GCGraphTest, tests/graphcut_/energy_label.py:189-210), and -- for whole ``graph_from_labels`` runs
(generate.py:177-338) -- the flow ``maxflow()`` returns and ``what_segment`` of every region.

``graph_from_labels`` calls ``inspect.getargspec`` (generate.py:280,286), which Python 3.11 removed; the generator
aliases it to ``inspect.getfullargspec`` in THIS process (same first field) so that the function runs at all; nothing
under /root/reference is touched.  The light-to-dark branch of boundary_stawiaski_directed (directedness >= 0) raises
TypeError in the reference (energy_label.py:281: a fifth parameter nobody passes), so only directedness < 0 has
golden vectors.  ``graphcut_split`` (wrapper.py:72-225) needs two more crutches, again in this process only: an empty
``SimpleITK`` module (medpy.filter imports medpy.io at module level) and ``functools.reduce`` under the name the
function expects (wrapper.py:145 calls the Python-2 builtin ``reduce``).
"""
import inspect
import os
import sys
import warnings

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as base  # noqa: E402


class Recorder:
    """Stands where a GCGraph stands; keeps the calls."""

    def __init__(self):
        self.n = []
        self.t = []

    def set_nweight(self, a, b, w1, w2):
        self.n.append((int(a), int(b), float(w1), float(w2)))

    def set_tweight(self, node, ws, wk):
        self.t.append((int(node), float(ws), float(wk)))


def supervoxels(shape, cells, seed):
    """Voronoi-like consecutive labels 1..K (every label guaranteed to occur)."""
    rng = numpy.random.default_rng(seed)
    pts = numpy.stack([rng.uniform(0, s, size=cells) for s in shape], axis=1)
    grids = numpy.stack(numpy.meshgrid(*[numpy.arange(s) for s in shape], indexing="ij"), axis=-1).reshape(-1, len(shape))
    d = ((grids[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    lab = d.argmin(1)
    _, inv = numpy.unique(lab, return_inverse=True)
    return (inv + 1).reshape(shape)


def cases():
    out = []
    # the reference's own fixtures (tests/graphcut_/energy_label.py:41-73,131-160)
    out.append(dict(name="ref_3d", label=numpy.asarray([[[1, 1], [1, 1]], [[1, 2], [2, 2]], [[2, 2], [2, 2]]]),
                    image=numpy.zeros((3, 2, 2), dtype=int)))
    out.append(dict(name="ref_zero_edge", label=numpy.asarray([[1, 2, 3], [1, 2, 4]]),
                    image=numpy.asarray([[0.0, 0.0, 0.0], [0.0, 0.0, sys.float_info.max]])))
    out.append(dict(name="ref_int_gradient", label=numpy.asarray([[1, 3, 4], [1, 2, 5], [1, 2, 5]]),
                    image=numpy.zeros((3, 3), dtype=int)))
    out.append(dict(name="ref_forder", label=numpy.asfortranarray(numpy.asarray([[1, 3, 4], [1, 2, 5], [1, 2, 5]])),
                    image=numpy.zeros((3, 3), order="C")))
    # generated: 2-D / 3-D / 4-D, float32 / float64 / int16 images, first pair of an axis on a border
    k = 0
    for shape, cells in [((9, 11), 7), ((6, 7, 8), 12), ((12, 10, 9), 40), ((4, 5, 3, 4), 9)]:
        for dt in (numpy.float32, numpy.float64, numpy.int16):
            k += 1
            lab = supervoxels(shape, cells, 40 + k)
            if k % 2 == 0:     # make element 0 of axis 0 a border pair (the numpy.vectorize double call)
                lab = lab.copy()
                lab.flat[0] = lab[(1,) + (0,) * (len(shape) - 1)] % lab.max() + 1
                _, inv = numpy.unique(lab, return_inverse=True)
                lab = (inv + 1).reshape(shape)
            rng = numpy.random.default_rng(900 + k)
            img = (rng.normal(0, 30, size=shape) + 10 * lab).astype(dt)
            if dt is numpy.int16:
                img = numpy.round(img).astype(dt)
            prob = rng.uniform(0, 1, size=shape).astype(numpy.float32 if dt is numpy.float32 else numpy.float64)
            fg = numpy.zeros(shape, bool)
            bg = numpy.zeros(shape, bool)
            fg[tuple(s // 2 for s in shape)] = True
            fg[tuple(s // 2 - 1 for s in shape)] = True
            bg[(0,) * len(shape)] = True
            bg[tuple(s - 1 for s in shape)] = True
            out.append(dict(name="gen_%s_%s" % ("x".join(map(str, shape)), numpy.dtype(dt).name), label=lab, image=img,
                            prob=prob, alpha=0.05 + 0.01 * k, directedness=-0.0002 * k, fg=fg, bg=bg))
    return out


def main():
    base.import_reference()
    if not hasattr(inspect, "getargspec"):
        inspect.getargspec = inspect.getfullargspec
    from medpy.graphcut import energy_label as el
    from medpy.graphcut import graph_from_labels
    store = {}
    names = []
    for c in cases():
        nm = c["name"]
        names.append(nm)
        lab, img = c["label"], c["image"]
        store[nm + "/label"] = numpy.ascontiguousarray(lab)
        store[nm + "/label_forder"] = numpy.asarray(bool(numpy.asarray(lab).flags["F_CONTIGUOUS"] and lab.ndim > 1))
        store[nm + "/image"] = numpy.ascontiguousarray(img)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = Recorder(); el.boundary_stawiaski(r, lab, img)
            store[nm + "/stawiaski"] = numpy.asarray(r.n, dtype=numpy.float64).reshape(-1, 4)
            r = Recorder(); el.boundary_difference_of_means(r, lab, img)
            store[nm + "/means"] = numpy.asarray(sorted(r.n), dtype=numpy.float64).reshape(-1, 4)
            if "directedness" in c:
                r = Recorder(); el.boundary_stawiaski_directed(r, lab, (img, c["directedness"]))
                store[nm + "/directed"] = numpy.asarray(r.n, dtype=numpy.float64).reshape(-1, 4)
                store[nm + "/directedness"] = numpy.asarray(c["directedness"])
                r = Recorder(); el.regional_atlas(r, lab, (c["prob"], c["alpha"]))
                store[nm + "/atlas"] = numpy.asarray(r.t, dtype=numpy.float64).reshape(-1, 3)
                store[nm + "/prob"] = c["prob"]
                store[nm + "/alpha"] = numpy.asarray(c["alpha"])
                store[nm + "/fg"] = c["fg"]
                store[nm + "/bg"] = c["bg"]
                # whole runs through graph_from_labels -> maxflow -> what_segment (bin/medpy_graphcut_label.py:128-146)
                for tag, kw in (("cut_stawiaski", dict(boundary_term=el.boundary_stawiaski, boundary_term_args=img)),
                                ("cut_means", dict(boundary_term=el.boundary_difference_of_means, boundary_term_args=img)),
                                ("cut_directed_atlas", dict(boundary_term=el.boundary_stawiaski_directed,
                                                            boundary_term_args=(img, c["directedness"]),
                                                            regional_term=el.regional_atlas,
                                                            regional_term_args=(c["prob"], c["alpha"])))):
                    g = graph_from_labels(lab, c["fg"], c["bg"], **kw)
                    flow = g.maxflow()
                    nreg = int(lab.max())
                    seg = numpy.asarray([0 if g.what_segment(v) == g.termtype.SINK else 1 for v in range(nreg)], dtype=numpy.uint8)
                    store[nm + "/" + tag + "_flow"] = numpy.asarray(flow)
                    store[nm + "/" + tag + "_mask"] = seg
    # graphcut_split / graphcut_stawiaski (wrapper.py:72-329): 8 overlapping sub-volumes, every one holding both markers
    import types
    sys.modules.setdefault("SimpleITK", types.ModuleType("SimpleITK"))   # medpy.filter -> medpy.io imports it at module level; unused here
    import functools
    import medpy.graphcut.wrapper as ref_wrapper
    if not hasattr(ref_wrapper, "reduce"):
        ref_wrapper.reduce = functools.reduce      # wrapper.py:145 uses the Python-2 builtin: NameError on Python 3 otherwise
    from medpy.graphcut.wrapper import graphcut_split, graphcut_stawiaski
    shape = (24, 22, 20)
    lab = supervoxels(shape, 60, 77)
    rng = numpy.random.default_rng(78)
    grad = numpy.abs(rng.normal(0, 20, size=shape) + 40 * (lab % 3 == 0)).astype(numpy.float32)
    fg = numpy.zeros(shape, bool)
    bg = numpy.zeros(shape, bool)
    for corner in numpy.ndindex(2, 2, 2):
        origin = [c * (s // 2) for c, s in zip(corner, shape)]
        fg[tuple(b + s // 4 for b, s in zip(origin, shape))] = True
        bg[tuple(b + 1 for b in origin)] = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        whole = graphcut_stawiaski((lab.copy(), grad, fg, bg))     # the only calling form that works in the reference
        split = graphcut_split(graphcut_stawiaski, lab.copy(), grad, fg, bg, 10, 3, 2)
    store["split/label"] = lab
    store["split/gradient"] = grad
    store["split/fg"] = fg
    store["split/bg"] = bg
    store["split/whole_mask"] = numpy.asarray(whole, dtype=numpy.uint8)
    store["split/split_mask"] = numpy.asarray(split, dtype=numpy.uint8)
    store["names"] = numpy.asarray(names)
    numpy.savez_compressed(os.path.join(HERE, "golden_labels_v1.npz"), **store)
    print("wrote", len(names), "cases")


if __name__ == "__main__":
    main()
