"""The two label-map helpers that surround a region cut in the reference (medpy/filter/label.py): ``relabel`` (:76-105)
before ``graph_from_labels`` and ``relabel_map`` (:31-73) after ``maxflow()``.  Fresh vectorised numpy implementations
with the reference's results."""
import numpy

from .errors import ArgumentError

__all__ = ["relabel", "relabel_map"]


def relabel(label_image, start=1):
    """Consecutive ids from ``start`` in order of first appearance (C order), like label.py:76-105."""
    label_image = numpy.asarray(label_image)
    flat = label_image.ravel()
    uniq, first, inverse = numpy.unique(flat, return_index=True, return_inverse=True)
    rank = numpy.empty(uniq.size, dtype=numpy.int64)
    rank[numpy.argsort(first, kind="stable")] = numpy.arange(uniq.size)
    return (rank[inverse] + start).astype(label_image.dtype).reshape(label_image.shape)


def relabel_map(label_image, mapping, key=lambda x, y: x[y]):
    """New id ``key(mapping, old id)`` for every voxel (label.py:31-73); ArgumentError for ids the mapping lacks."""
    label_image = numpy.array(label_image)
    uniq, inverse = numpy.unique(label_image.ravel(), return_inverse=True)
    new = numpy.empty(uniq.size, dtype=label_image.dtype)
    for k, x in enumerate(uniq.tolist()):
        try:
            new[k] = key(mapping, x)
        except Exception as e:
            raise ArgumentError("No conversion for region id {} found in the supplied mapping. Error: {}".format(x, e))
    return new[inverse].reshape(label_image.shape)
