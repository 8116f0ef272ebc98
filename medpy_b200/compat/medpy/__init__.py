"""Import shim: put ``medpy_b200/compat`` FIRST on PYTHONPATH and ``import medpy.graphcut`` resolves to the
B200-native voxel graph-cut path (medpy_b200.graphcut), so scripts written against MedPy -- in particular the
reference's ``bin/medpy_graphcut_voxel.py`` -- run unchanged.  Only what that script imports is provided:
``medpy.graphcut`` (voxel half), ``medpy.core`` (Logger, ArgumentError) and a small SimpleITK-free ``medpy.io``
(load / save / header for .npy and uncompressed MetaImage).  Everything else of MedPy is out of scope."""
__version__ = "0.5.2+b200"
