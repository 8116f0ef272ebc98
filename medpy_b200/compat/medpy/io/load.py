"""``load(path) -> (array in x,y,z[,c] order, Header)`` (reference contract: medpy/io/load.py:34-129)."""
import os

import numpy

from ..core import ImageLoadingError, Logger
from . import _metaimage, _nifti
from .header import Header


def load(image):
    logger = Logger.getInstance()
    logger.info("Loading image {}...".format(image))
    if not os.path.exists(image):
        raise ImageLoadingError("The supplied image {} does not exist.".format(image))
    ext = os.path.splitext(image)[1].lower()
    try:
        if ext == ".npy":
            zyx = numpy.load(image)
            spacing, offset, nchan = [1.0] * zyx.ndim, [0.0] * zyx.ndim, 1
        elif ext in (".mha", ".mhd"):
            zyx, spacing, offset, nchan = _metaimage.read(image)
        elif ext == ".nii" or image.lower().endswith(".nii.gz"):
            zyx, spacing, offset, nchan = _nifti.read(image)
        else:
            raise ImageLoadingError("Only .npy, uncompressed MetaImage (.mha/.mhd) and NIfTI-1 (.nii/.nii.gz) are supported without SimpleITK.")
    except ImageLoadingError:
        raise
    except Exception as e:  # noqa: BLE001
        raise ImageLoadingError("Failed to read image {}: {}".format(image, e)) from e
    # z,y,x[,c] storage -> x,y,z[,c] view (no copy; Fortran-strided like the reference's arr.T, load.py:125-127)
    if nchan > 1 and zyx.ndim == 4:
        arr = numpy.moveaxis(zyx, -1, 0).T  # (c, z, y, x) -> (x, y, z, c), as load.py:122-125
    else:
        arr = zyx.T
    return arr, Header(spacing=spacing, offset=offset)
