"""``save(arr, filename, hdr=False, force=True)`` (reference contract: medpy/io/save.py:33-124)."""
import os

import numpy

from ..core import ImageSavingError, Logger
from . import _metaimage, _nifti


def save(arr, filename, hdr=False, force=True, use_compression=False):
    logger = Logger.getInstance()
    logger.info("Saving image as {}...".format(filename))
    if not force and os.path.exists(filename):
        raise ImageSavingError("The target file {} already exists.".format(filename))
    arr = numpy.asarray(arr)
    if arr.ndim == 4:
        arr = numpy.moveaxis(arr, -1, 0)
    zyx = arr.T  # x,y,z -> z,y,x (save.py:111-113)
    if zyx.dtype == numpy.bool_:
        zyx = zyx.astype(numpy.uint8)
    ext = os.path.splitext(filename)[1].lower()
    try:
        if ext == ".npy":
            numpy.save(filename, numpy.ascontiguousarray(zyx))
        elif ext in (".mha", ".mhd"):
            spacing = hdr.get_voxel_spacing() if hdr else None
            offset = hdr.get_offset() if hdr else None
            _metaimage.write(filename, zyx, spacing, offset)
        elif ext == ".nii" or filename.lower().endswith(".nii.gz"):
            spacing = hdr.get_voxel_spacing() if hdr else None
            offset = hdr.get_offset() if hdr else None
            _nifti.write(filename, zyx, spacing, offset)
        else:
            raise ImageSavingError("Only .npy, MetaImage (.mha/.mhd) and NIfTI-1 (.nii/.nii.gz) are supported without SimpleITK.")
    except ImageSavingError:
        raise
    except Exception as e:  # noqa: BLE001
        raise ImageSavingError("Failed to write image {}: {}".format(filename, e)) from e
