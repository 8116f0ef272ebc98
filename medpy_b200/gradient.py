"""Gradient-magnitude pre-step of the ``boundary_maximum_*`` terms.

``bin/medpy_gradient.py:79-85`` computes ``scipy.ndimage.generic_gradient_magnitude(image, prewitt, output=float32)``
and writes the result as the "gradient image" that ``medpy_graphcut_voxel.py --boundary max_*`` consumes.  This module
does the same on the GPU (kernel ``k_gradient_magnitude``, csrc/gc_gradient.cuh), bit for bit, with no CPU fallback."""
import numpy


def gradient_magnitude_prewitt(image, device=-1):
    """float32 Prewitt gradient magnitude of an n-D image (1 <= n <= 4), mode 'reflect', identical to
    ``scipy.ndimage.generic_gradient_magnitude(image, scipy.ndimage.prewitt, output=numpy.float32)``."""
    from . import _lib
    image = numpy.asarray(image)
    if image.dtype == numpy.bool_:
        image = image.view(numpy.uint8)
    if image.dtype.type not in (numpy.float32, numpy.float64, numpy.uint8, numpy.int16, numpy.int32):
        image = image.astype(numpy.float64)      # SciPy widens every line to double anyway
    image = numpy.ascontiguousarray(image)
    return _lib._mgc.gradient_magnitude_prewitt(image, device)
