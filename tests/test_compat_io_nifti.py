"""CPU tests of the SimpleITK-free NIfTI-1 reader / writer behind the import shim's medpy.io: the
formats MedPy users actually have must be readable by the CLIs that run on the B200 path."""
import gzip
import os
import struct
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "medpy_b200", "compat"))


def _io():
    from medpy.io import load, save, header
    return load, save, header


@pytest.mark.parametrize("dtype", ["uint8", "int16", "int32", "float32", "float64", "uint16", "bool"])
@pytest.mark.parametrize("ext", [".nii", ".nii.gz"])
def test_round_trip_keeps_values_layout_and_geometry(tmp_path, dtype, ext):
    load, save, header = _io()
    rng = numpy.random.default_rng(0)
    arr = (rng.normal(0, 50, (5, 6, 7)) > 0) if dtype == "bool" else rng.normal(0, 50, (5, 6, 7)).astype(dtype)   # x, y, z
    path = str(tmp_path / ("v" + ext))
    hdr = header.Header(spacing=(0.5, 0.75, 2.0), offset=(10.0, -20.0, 30.0))
    save(arr, path, hdr)
    back, h2 = load(path)
    assert back.shape == arr.shape
    assert numpy.array_equal(back, arr.astype(numpy.uint8) if dtype == "bool" else arr)
    assert back.flags.f_contiguous                                   # x,y,z view of z,y,x storage, like the reference's arr.T
    assert h2.get_voxel_spacing() == pytest.approx((0.5, 0.75, 2.0))
    assert h2.get_offset() == pytest.approx((10.0, -20.0, 30.0))


def test_on_disk_layout_is_nifti(tmp_path):
    """Header fields and voxel order as the specification says: x fastest, dims in dim[1..3], magic n+1."""
    _, save, _ = _io()
    arr = numpy.arange(2 * 3 * 4, dtype=numpy.int16).reshape(2, 3, 4)         # x, y, z
    path = str(tmp_path / "a.nii")
    save(arr, path)
    raw = open(path, "rb").read()
    assert struct.unpack("<i", raw[:4])[0] == 348 and raw[344:348] == b"n+1\x00"
    assert struct.unpack("<8h", raw[40:56])[:4] == (3, 2, 3, 4)
    assert struct.unpack("<h", raw[70:72])[0] == 4 and struct.unpack("<h", raw[72:74])[0] == 16
    data = numpy.frombuffer(raw, "<i2", offset=352)
    assert numpy.array_equal(data, arr.ravel(order="F"))


def test_big_endian_scaled_and_4d_files(tmp_path):
    load, _, _ = _io()
    x, y, z, t = 3, 4, 5, 2
    vals = numpy.arange(x * y * z * t, dtype=">i2")
    hdr = bytearray(348)
    struct.pack_into(">i", hdr, 0, 348)
    struct.pack_into(">8h", hdr, 40, 4, x, y, z, t, 1, 1, 1)
    struct.pack_into(">h", hdr, 70, 4)
    struct.pack_into(">h", hdr, 72, 16)
    struct.pack_into(">8f", hdr, 76, 1.0, 1.5, 2.5, 3.5, 1.0, 1.0, 1.0, 1.0)
    struct.pack_into(">f", hdr, 108, 352.0)
    struct.pack_into(">2f", hdr, 112, 2.0, 1.0)                        # scl_slope, scl_inter
    struct.pack_into(">2h", hdr, 252, 0, 1)                            # sform only
    struct.pack_into(">4f", hdr, 280, -1.5, 0, 0, -7.0)
    struct.pack_into(">4f", hdr, 296, 0, -2.5, 0, 8.0)
    struct.pack_into(">4f", hdr, 312, 0, 0, 3.5, 9.0)
    hdr[344:348] = b"n+1\x00"
    path = str(tmp_path / "be.nii.gz")
    with gzip.open(path, "wb") as fh:
        fh.write(bytes(hdr) + b"\0\0\0\0" + vals.tobytes())
    arr, h = load(path)
    assert arr.shape == (x, y, z, t)
    expect = (numpy.arange(x * y * z * t).reshape((x, y, z, t), order="F") * 2.0 + 1.0)
    assert numpy.array_equal(arr, expect)
    assert h.get_voxel_spacing()[:3] == pytest.approx((1.5, 2.5, 3.5))
    assert h.get_offset()[:3] == pytest.approx((7.0, -8.0, 9.0))        # RAS -> LPS


def test_rejects_what_it_cannot_read(tmp_path):
    load, save, _ = _io()
    from medpy.core import ImageLoadingError, ImageSavingError
    p = tmp_path / "bad.nii"
    p.write_bytes(b"\0" * 400)
    with pytest.raises(ImageLoadingError):
        load(str(p))
    with pytest.raises(ImageSavingError):
        save(numpy.zeros((2, 2, 2), dtype=numpy.complex64), str(tmp_path / "c.nii"))
