"""Deterministic synthetic workloads for the voxel graph-cut path (SURVEY.md §8d / BASELINE.md §3).

All generators use ``numpy.random.default_rng(seed)``; arrays are C-ordered with the LAST axis fastest.
Nothing here touches the GPU; bench.py and the tests feed these arrays through the public API.
"""
import numpy


def _ball_mask(shape, centres, radius_frac, min_radius=0.0):
    """Union of ellipsoids: centre c*extent, semi-axis radius_frac*extent (+min_radius) per axis.
    Built plane by plane along axis 0 to keep the temporary small at 512^3."""
    nd = len(shape)
    out = numpy.zeros(shape, dtype=numpy.bool_)
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
        for i in range(shape[0]):
            if terms[0][i] <= 1.0:
                out[i] |= (rest + terms[0][i]) <= 1.0
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


def rms_neighbour_difference(image):
    """The harness' "sigma=auto": sqrt(mean over all lattice neighbour pairs of (I_p - I_q)^2), float64.
    (The reference has no automatic sigma -- bin/medpy_graphcut_voxel.py:203-205 takes a float.)"""
    acc = 0.0
    cnt = 0
    for d in range(image.ndim):
        if image.shape[d] < 2:
            continue
        lo = [slice(None)] * image.ndim
        hi = [slice(None)] * image.ndim
        lo[d] = slice(0, -1)
        hi[d] = slice(1, None)
        # chunk along axis 0 to bound temporaries
        a = image[tuple(lo)]
        b = image[tuple(hi)]
        step = max(1, (1 << 24) // max(1, int(numpy.prod(a.shape[1:]))))
        for i in range(0, a.shape[0], step):
            diff = a[i:i + step].astype(numpy.float64) - b[i:i + step].astype(numpy.float64)
            acc += float(numpy.dot(diff.ravel(), diff.ravel()))
            cnt += diff.size
    return float(numpy.sqrt(acc / max(cnt, 1)))


def two_blob_volume(shape, seed=0, contrast=100.0, noise=10.0, integer=False, with_prob=True):
    """Two-blob volume + markers.

    image  : float32, 100*[inside either ball] + N(0, 10^2); balls at 0.3 and 0.7 of the extent,
             radius 0.18 of the extent
    fg     : balls of half that radius around the same centres
    bg     : 1-voxel shell on all faces
    prob   : sigmoid((image-50)/15) float32 (regional_probability_map input, alpha 0.1)
    sigma  : RMS neighbour difference ("sigma=auto")
    """
    shape = tuple(int(s) for s in shape)
    rng = numpy.random.default_rng(seed)
    inside = _ball_mask(shape, (0.3, 0.7), 0.18)
    image = numpy.empty(shape, dtype=numpy.float32)
    # generate plane-wise so the float64 normal temporaries stay small
    for i in range(shape[0]):
        image[i] = rng.normal(0.0, noise, size=shape[1:]).astype(numpy.float32)
    image += numpy.float32(contrast) * inside
    if integer:
        numpy.round(image, out=image)
    fg = _ball_mask(shape, (0.3, 0.7), 0.09, min_radius=0.5)
    bg = shell_mask(shape)
    out = dict(image=image, fg=fg, bg=bg, inside=inside, sigma=rms_neighbour_difference(image))
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
