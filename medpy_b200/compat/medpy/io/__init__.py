"""SimpleITK-free ``medpy.io`` for the voxel graph-cut CLI: ``load``, ``save``, ``header`` / ``Header`` for
``.npy`` and uncompressed ITK MetaImage (``.mha`` / ``.mhd`` + ``.raw``).

Conventions follow the reference (medpy/io/load.py:116-127, save.py:107-124): ``load`` returns the array in
``x, y, z[, c]`` order as a *transposed view* of the file's z,y,x storage -- i.e. Fortran-strided, exactly what
the reference hands to ``graph_from_voxels`` -- plus a header carrying voxel spacing / offset; ``save`` writes
bool as uint8 and transposes back."""
from . import header
from .header import Header
from .load import load
from .save import save

__all__ = ["load", "save", "header", "Header"]
