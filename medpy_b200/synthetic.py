"""Deterministic synthetic workloads for the voxel graph-cut path (SURVEY.md §8d / BASELINE.md §3).

All generators use ``numpy.random.default_rng(seed)``; arrays are C-ordered with the LAST axis fastest.
Nothing here touches the GPU; bench.py and the tests feed these arrays through the public API.
"""
import numpy


def _shell_planes(shape, a0, b0):
    """shell_mask(shape)[a0:b0] without building the full volume."""
    sub = (b0 - a0,) + tuple(shape[1:])
    bg = numpy.zeros(sub, dtype=numpy.bool_)
    for d, s in enumerate(shape):
        if s < 3:
            continue
        if d == 0:
            if a0 == 0:
                bg[0] = True
            if b0 == s:
                bg[-1] = True
            continue
        sl = [slice(None)] * len(shape)
        sl[d] = 0
        bg[tuple(sl)] = True
        sl[d] = s - 1
        bg[tuple(sl)] = True
    return bg


def _ball_mask(shape, centres, radius_frac, min_radius=0.0, planes=None):
    """Union of ellipsoids: centre c*extent, semi-axis radius_frac*extent (+min_radius) per axis.
    Built plane by plane along axis 0 to keep the temporary small at 512^3; `planes` restricts axis 0."""
    nd = len(shape)
    a0, b0 = (0, shape[0]) if planes is None else planes
    out = numpy.zeros((b0 - a0,) + tuple(shape[1:]), dtype=numpy.bool_)
    axes = [numpy.arange(s, dtype=numpy.float32) for s in shape]
    for c in centres:
        terms = []
        for d in range(nd):
            r = radius_frac * shape[d] + min_radius
            t = ((axes[d] - c * shape[d]) / max(r, 1e-6)) ** 2
            terms.append(t.astype(numpy.float32))
        # separable sum via broadcasting, one axis-0 plane at a time
        rest = terms[1]
        for d in range(2, nd):
            rest = rest[..., None] + terms[d]
        for i in range(a0, b0):
            if terms[0][i] <= 1.0:
                out[i - a0] |= (rest + terms[0][i]) <= 1.0
    return out


def shell_mask(shape):
    """1-voxel shell on every face whose axis has extent >= 3 (background seeds)."""
    bg = numpy.zeros(shape, dtype=numpy.bool_)
    for d, s in enumerate(shape):
        if s < 3:
            continue
        sl = [slice(None)] * len(shape)
        sl[d] = 0
        bg[tuple(sl)] = True
        sl[d] = s - 1
        bg[tuple(sl)] = True
    return bg


def neighbour_difference_partials(image, next_plane=None):
    """Per axis-0 plane i: the float64 sum of (I_p - I_q)^2 over the neighbour pairs whose lower voxel lies in plane
    i (pairs inside the plane along the other axes + pairs towards plane i+1; the last plane pairs with `next_plane`
    when given).  Summing these with math.fsum gives a partition-independent total: the z-slab ranks each compute the
    partials of their own planes and obtain bit-identical sigma."""
    out = numpy.zeros(image.shape[0], dtype=numpy.float64)
    for i in range(image.shape[0]):
        p = image[i].astype(numpy.float64)
        acc = 0.0
        for d in range(p.ndim):
            if p.shape[d] < 2:
                continue
            lo = [slice(None)] * p.ndim
            hi = [slice(None)] * p.ndim
            lo[d] = slice(0, -1)
            hi[d] = slice(1, None)
            diff = p[tuple(lo)] - p[tuple(hi)]
            acc += float(numpy.sum(diff * diff))
        nxt = image[i + 1] if i + 1 < image.shape[0] else next_plane
        if nxt is not None:
            diff = p - numpy.asarray(nxt, dtype=numpy.float64)
            acc += float(numpy.sum(diff * diff))
        out[i] = acc
    return out


def neighbour_pair_count(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return sum((n // int(s)) * (int(s) - 1) for s in shape if int(s) > 1)


def rms_neighbour_difference(image):
    """The harness' "sigma=auto": sqrt(mean over all lattice neighbour pairs of (I_p - I_q)^2), float64.
    (The reference has no automatic sigma -- bin/medpy_graphcut_voxel.py:203-205 takes a float.)"""
    import math
    if image.ndim == 1:
        image = image[None]
        parts = neighbour_difference_partials(image)
        return float(math.sqrt(math.fsum(parts) / max(neighbour_pair_count(image.shape[1:]), 1)))
    parts = neighbour_difference_partials(image)
    return float(math.sqrt(math.fsum(parts) / max(neighbour_pair_count(image.shape), 1)))


def two_blob_volume(shape, seed=0, contrast=100.0, noise=10.0, integer=False, with_prob=True, planes=None):
    """Two-blob volume + markers.

    image  : float32, 100*[inside either ball] + N(0, 10^2); balls at 0.3 and 0.7 of the extent,
             radius 0.18 of the extent
    fg     : balls of half that radius around the same centres
    bg     : 1-voxel shell on all faces
    prob   : sigmoid((image-50)/15) float32 (regional_probability_map input, alpha 0.1)
    sigma  : RMS neighbour difference ("sigma=auto")

    The noise of axis-0 plane i comes from its own generator ``default_rng([seed, i])``, so a rank of the z-slab
    multi-GPU path can build just its planes: ``planes=(a, b)`` returns the arrays restricted to planes [a, b)
    (identical to slicing the full volume) and ``sigma`` is then None (it needs the whole volume).
    """
    shape = tuple(int(s) for s in shape)
    a0, b0 = (0, shape[0]) if planes is None else (int(planes[0]), int(planes[1]))
    inside = _ball_mask(shape, (0.3, 0.7), 0.18, planes=(a0, b0))
    image = numpy.empty((b0 - a0,) + shape[1:], dtype=numpy.float32)
    for i in range(a0, b0):
        rng = numpy.random.default_rng([seed, i])
        image[i - a0] = rng.normal(0.0, noise, size=shape[1:]).astype(numpy.float32)
    image += numpy.float32(contrast) * inside
    if integer:
        numpy.round(image, out=image)
    fg = _ball_mask(shape, (0.3, 0.7), 0.09, min_radius=0.5, planes=(a0, b0))
    bg = _shell_planes(shape, a0, b0)
    out = dict(image=image, fg=fg, bg=bg, inside=inside,
               sigma=rms_neighbour_difference(image) if planes is None else None)
    if with_prob:
        prob = image.astype(numpy.float32)
        prob -= numpy.float32(50.0)
        prob /= numpy.float32(15.0)
        numpy.negative(prob, out=prob)
        numpy.exp(prob, out=prob)
        prob += numpy.float32(1.0)
        numpy.reciprocal(prob, out=prob)
        out["prob"] = prob
        out["alpha"] = 0.1
    return out


def multispectral_volume(shape=(256, 256, 128, 4), seed=0):
    """Config 4: 4-D multi-spectral fp32 volume, channel axis last; channel k = two-blob geometry with
    contrast 100/(k+1), seed+k; markers replicated in every channel."""
    shape = tuple(int(s) for s in shape)
    sp, nc = shape[:-1], shape[-1]
    image = numpy.empty(shape, dtype=numpy.float32)
    for k in range(nc):
        v = two_blob_volume(sp, seed=seed + k, contrast=100.0 / (k + 1), with_prob=False)
        image[..., k] = v["image"]
    fg = numpy.repeat(v["fg"][..., None], nc, axis=-1)
    bg = numpy.repeat(v["bg"][..., None], nc, axis=-1)
    return dict(image=image, fg=fg, bg=bg, sigma=rms_neighbour_difference(image))


def integer_weight_boundary_term(image):
    """The integer parity set of SURVEY.md §8d: a custom 2-argument boundary term (API-legal,
    generate.py:79-86) with integer weights w = 1 + (255 - min(|dI|, 255)).  Returns the callable to pass
    as ``boundary_term`` with ``boundary_term_args=(image,)``; it drives ``graph.set_nweights_dense`` when
    the graph offers it (ours) and ``graph.set_nweight`` edge by edge otherwise (the reference's GCGraph)."""
    def term(graph, args):
        (img,) = args
        img = numpy.asarray(img).astype(numpy.float64)
        n = img.size
        stride = n
        for d in range(img.ndim):
            stride //= img.shape[d]
            lo = [slice(None)] * img.ndim
            hi = [slice(None)] * img.ndim
            lo[d] = slice(0, -1)
            hi[d] = slice(1, None)
            w = 1.0 + (255.0 - numpy.minimum(numpy.abs(img[tuple(lo)] - img[tuple(hi)]), 255.0))
            if hasattr(graph, "set_nweights_dense"):
                graph.set_nweights_dense(d, w, w)
            else:
                full = numpy.zeros(img.shape)
                full[tuple(lo)] = w
                flat = full.ravel()
                keys = numpy.flatnonzero(full.ravel() > 0)
                for p in keys:
                    graph.set_nweight(int(p), int(p + stride), float(flat[p]), float(flat[p]))
    return term
