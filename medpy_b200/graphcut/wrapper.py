"""Host-side mirror of the part of ``medpy/graphcut/wrapper.py`` that is on the voxel path."""
import numpy


def split_marker(marker, fg_id=1, bg_id=2):
    """Marker image -> (foreground, background) boolean arrays; by default label 1 marks foreground and
    label 2 background (same contract as the reference's medpy/graphcut/wrapper.py:39-69)."""
    labels = numpy.asarray(marker)
    return numpy.equal(labels, fg_id), numpy.equal(labels, bg_id)
