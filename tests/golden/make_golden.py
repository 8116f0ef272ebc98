#!/usr/bin/env python
"""Generate tests/golden/golden_v1.npz by running the UNMODIFIED reference package.

Run in the build container only (needs /root/reference and `make -C oracle pyshim`):

    python tests/golden/make_golden.py

The reference's Python layer (medpy/graphcut/{generate,energy_voxel,graph}.py) is imported from
/root/reference as is; its compiled ``medpy.graphcut.maxflow`` extension (Boost.Python, not buildable
here) is replaced by oracle/_ref/maxflow*.so, a pybind11 binding of the reference's own
Graph<double,double,double> compiled from /root/reference/lib/maxflow/src.  For every case we record
the inputs and what the reference returns: the pre-solve n-link weights (get_edge) and t-links
(get_trcap), the flow returned by maxflow() and the mask read out exactly the way
bin/medpy_graphcut_voxel.py:177-181 does.  The GPU box has no /root/reference, so these vectors are
what the `-m gpu` parity tests compare against there.
"""
import glob
import importlib.util
import json
import os
import sys
import warnings

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def import_reference():
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "maxflow*.so"))
    if not so:
        raise SystemExit("run `make -C oracle pyshim` first")
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    spec = importlib.util.spec_from_file_location("medpy.graphcut.maxflow", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["medpy.graphcut.maxflow"] = mod
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import medpy.graphcut as gc  # noqa
    return gc


def two_blob(shape, seed, integer=False):
    """The synthetic volume of SURVEY.md §8d scaled to `shape` (any ndim)."""
    rng = numpy.random.default_rng(seed)
    grids = numpy.meshgrid(*[numpy.arange(s) for s in shape], indexing="ij")
    inside = numpy.zeros(shape, bool)
    for c in (0.3, 0.7):
        r2 = sum(((g - c * s) / (0.18 * s + 1e-9)) ** 2 for g, s in zip(grids, shape))
        inside |= r2 <= 1.0
    img = 100.0 * inside + rng.normal(0, 10, size=shape)
    if integer:
        img = numpy.round(img)
    return img, inside


def markers(shape):
    grids = numpy.meshgrid(*[numpy.arange(s) for s in shape], indexing="ij")
    fg = numpy.zeros(shape, bool)
    for c in (0.3, 0.7):
        r2 = sum(((g - c * s) / (0.09 * s + 0.5)) ** 2 for g, s in zip(grids, shape))
        fg |= r2 <= 1.0
    bg = numpy.zeros(shape, bool)
    for d, s in enumerate(shape):
        if s < 3:
            continue
        sl = [slice(None)] * len(shape)
        sl[d] = 0
        bg[tuple(sl)] = True
        sl[d] = s - 1
        bg[tuple(sl)] = True
    return fg, bg


TERMS = ["difference_linear", "difference_exponential", "difference_division", "difference_power",
         "maximum_linear", "maximum_exponential", "maximum_division", "maximum_power"]


def cases():
    out = []
    # --- the reference's own fixtures -------------------------------------------------------
    image = numpy.asarray([[0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 1, 1], [0, 0, 1, 1]], dtype=float)
    gradient = numpy.asarray([[0, 0, 0, 0], [0, 1, 1, 1], [0, 1, 0, 0], [0, 1, 0, 0]], dtype=float)
    fgm = numpy.zeros((4, 4), int); fgm[3, 3] = 1
    bgm = numpy.zeros((4, 4), int); bgm[0, 0] = 1
    sig = {"exponential": 1.0, "division": 0.5, "power": 2.0}
    for t in TERMS:  # tests/graphcut_/energy_voxel.py:69-103
        base, fn = t.split("_")
        img = image if base == "difference" else gradient
        out.append(dict(name="ref_fixture_%s" % t, fg=fgm, bg=bgm, boundary=t, image=img,
                        sigma=sig.get(fn), spacing=False, expect_mask=image.astype(bool)))
    out.append(dict(name="ref_fixture_regional", fg=fgm, bg=bgm, prob=image / 2.0, alpha=1.0,
                    expect_mask=image.astype(bool)))  # energy_voxel.py:105-107
    sp_img = numpy.zeros((5, 5)); sp_img[1:, 2] = 2
    sp_fg = numpy.zeros((5, 5), bool); sp_fg[4, 2] = True
    sp_bg = numpy.zeros((5, 5), bool); sp_bg[0, 0] = sp_bg[0, 4] = True
    out.append(dict(name="ref_fixture_spacing", fg=sp_fg, bg=sp_bg, boundary="difference_division",
                    image=sp_img, sigma=1.0, spacing=(1.0, 5.0), expect_mask=sp_img.astype(bool)))  # :110-149
    vimg = numpy.asarray([[[1, 0, 1, 2, 3], [1, 0, 1, 4, 3], [0, 1, 1, 6, 4]]] * 2)
    vfg = numpy.zeros((2, 3, 5), int); vfg[:, 2, 0] = 1
    vbg = numpy.zeros((2, 3, 5), int); vbg[:, 0, 4] = 1
    vexp = numpy.asarray([[[1, 1, 1, 0, 0]] * 3] * 2, bool)
    out.append(dict(name="ref_fixture_cut3d", fg=vfg, bg=vbg, boundary="difference_linear", image=vimg,
                    sigma=None, spacing=False, expect_mask=vexp, expect_flow=3.0))  # tests/graphcut_/cut.py:32-50
    neg = numpy.asarray([[-1, 1, -4], [2, -7, 3], [-2.3, 3, -7]], dtype=float)  # energy_voxel.py:152-154
    fg3 = numpy.zeros((3, 3), int); fg3[2, 2] = 1
    bg3 = numpy.zeros((3, 3), int); bg3[0, 0] = 1
    for t in TERMS:
        fn = t.split("_")[1]
        out.append(dict(name="ref_negative_%s" % t, fg=fg3, bg=bg3, boundary=t, image=neg,
                        sigma=None if fn == "linear" else 1.0, spacing=False))
    # --- generated cases (gaps listed in SURVEY.md §4) ---------------------------------------
    seed = 100
    for shape in [(6, 7, 8), (9, 5, 12)]:
        for dt in (numpy.float32, numpy.float64):
            for t in TERMS:
                seed += 1
                img, _ = two_blob(shape, seed)
                if t.startswith("maximum"):
                    img = numpy.abs(numpy.gradient(img)[0]) * (1 if seed % 2 else -1)
                fn = t.split("_")[1]
                fg, bg = markers(shape)
                out.append(dict(name="gen_%s_%s_%s" % ("x".join(map(str, shape)), numpy.dtype(dt).name, t),
                                fg=fg, bg=bg, boundary=t, image=img.astype(dt),
                                sigma={"linear": None, "exponential": 15.0, "division": 7.5, "power": 1.5}[fn],
                                spacing=(1.0, 0.5, 2.5) if seed % 3 == 0 else False))
    # regional + boundary, overlapping markers, fp32 probabilities (config-3 shape in small)
    for k, shape in enumerate([(8, 8, 8), (5, 11, 7)]):
        img, _ = two_blob(shape, 500 + k)
        img = img.astype(numpy.float32)
        prob = (1.0 / (1.0 + numpy.exp(-(img - 50.0) / 15.0))).astype(numpy.float32)
        fg, bg = markers(shape)
        bg2 = bg.copy(); bg2[tuple(s // 3 for s in shape)] = True
        fg2 = fg.copy(); fg2[tuple(s // 3 for s in shape)] = True  # fg AND bg on one voxel
        out.append(dict(name="gen_regional_exp_%d" % k, fg=fg2, bg=bg2, boundary="difference_exponential",
                        image=img, sigma=14.0, spacing=False, prob=prob, alpha=0.1))
    # 2-D, 1-D and 4-D lattices
    img, _ = two_blob((17, 23), 600)
    fg, bg = markers((17, 23))
    out.append(dict(name="gen_2d_exp", fg=fg, bg=bg, boundary="difference_exponential",
                    image=img.astype(numpy.float32), sigma=12.0, spacing=(2.0, 1.0)))
    img1 = numpy.asarray([0, 1, 0, 2, 9, 10, 9, 11, 10, 2, 1, 0], dtype=numpy.float32)
    fg1 = numpy.zeros(12, bool); fg1[5] = True
    bg1 = numpy.zeros(12, bool); bg1[0] = bg1[11] = True
    out.append(dict(name="gen_1d_div", fg=fg1, bg=bg1, boundary="difference_division", image=img1,
                    sigma=2.0, spacing=False))
    shape4 = (6, 5, 4, 4)
    img4 = numpy.stack([two_blob(shape4[:3], 700 + c)[0] / (c + 1) for c in range(4)], axis=-1)
    fg4, bg4 = markers(shape4[:3])
    fg4 = numpy.repeat(fg4[..., None], 4, axis=-1)
    bg4 = numpy.repeat(bg4[..., None], 4, axis=-1)
    out.append(dict(name="gen_4d_max_exp", fg=fg4, bg=bg4, boundary="maximum_exponential",
                    image=img4.astype(numpy.float32), sigma=40.0, spacing=False))
    # Fortran-ordered input as medpy.io.load returns it (io/load.py:125-127): node ids stay C-order
    img, _ = two_blob((7, 6, 5), 800)
    fg, bg = markers((7, 6, 5))
    out.append(dict(name="gen_forder_exp", fg=numpy.asfortranarray(fg), bg=numpy.asfortranarray(bg),
                    boundary="difference_exponential", image=numpy.asfortranarray(img.astype(numpy.float32)),
                    sigma=15.0, spacing=False))
    # integer-valued image and int16 dtype -> integer-exact path through difference_linear
    img, _ = two_blob((6, 6, 6), 900, integer=True)
    fg, bg = markers((6, 6, 6))
    out.append(dict(name="gen_int16_division", fg=fg, bg=bg, boundary="difference_division",
                    image=img.astype(numpy.int16), sigma=4.0, spacing=False))
    return out


def run_case(gc, c):
    from medpy.graphcut import energy_voxel as ev, graph_from_voxels
    kw = {}
    if c.get("boundary"):
        fn = getattr(ev, "boundary_" + c["boundary"])
        if c["boundary"].endswith("linear"):
            kw.update(boundary_term=fn, boundary_term_args=(c["image"], c["spacing"]))
        else:
            kw.update(boundary_term=fn, boundary_term_args=(c["image"], c["sigma"], c["spacing"]))
    if c.get("prob") is not None:
        kw.update(regional_term=ev.regional_probability_map, regional_term_args=(c["prob"], c["alpha"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g = graph_from_voxels(c["fg"], c["bg"], **kw)
    shape = numpy.asarray(c["fg"]).shape
    n = int(numpy.prod(shape))
    nd = len(shape)
    # pre-solve capacities and t-links, as the reference's graph holds them
    w = numpy.zeros((nd, n))
    stride = n
    for d in range(nd):
        stride //= shape[d]
        for p in range(n):
            if (p % (stride * shape[d])) // stride < shape[d] - 1:
                w[d, p] = g.get_edge(p, p + stride)
    tr = numpy.asarray([g.get_trcap(p) for p in range(n)])
    flow = g.maxflow()
    mask = numpy.zeros(n, dtype=numpy.bool_)
    for idx in range(n):  # bin/medpy_graphcut_voxel.py:177-181
        mask[idx] = 0 if g.termtype.SINK == g.what_segment(idx) else 1
    return w, tr, flow, mask.reshape(shape)


def main():
    gc = import_reference()
    arrays = {}
    meta = []
    for c in cases():
        w, tr, flow, mask = run_case(gc, c)
        name = c["name"]
        if "expect_mask" in c:
            assert (mask == c["expect_mask"]).all(), name
        if "expect_flow" in c:
            assert flow == c["expect_flow"], (name, flow)
        m = dict(name=name, boundary=c.get("boundary"), sigma=c.get("sigma"),
                 spacing=list(c["spacing"]) if c.get("spacing") else False,
                 alpha=c.get("alpha"), flow_hex=float(flow).hex(),
                 forder=bool(numpy.asarray(c["fg"]).flags.f_contiguous and numpy.asarray(c["fg"]).ndim > 1))
        meta.append(m)
        arrays[name + "/fg"] = numpy.ascontiguousarray(numpy.asarray(c["fg"]).astype(numpy.uint8))
        arrays[name + "/bg"] = numpy.ascontiguousarray(numpy.asarray(c["bg"]).astype(numpy.uint8))
        if c.get("image") is not None:
            arrays[name + "/image"] = numpy.ascontiguousarray(c["image"])
        if c.get("prob") is not None:
            arrays[name + "/prob"] = numpy.ascontiguousarray(c["prob"])
        arrays[name + "/w"] = w
        arrays[name + "/tr"] = tr
        arrays[name + "/mask"] = mask.astype(numpy.uint8)
    arrays["__meta__"] = numpy.frombuffer(json.dumps(meta).encode(), dtype=numpy.uint8)
    path = os.path.join(HERE, "golden_v1.npz")
    numpy.savez_compressed(path, **arrays)
    print("wrote", path, "cases:", len(meta), "bytes:", os.path.getsize(path))


if __name__ == "__main__":
    main()
