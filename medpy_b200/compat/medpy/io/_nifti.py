"""NIfTI-1 single-file images (.nii, .nii.gz) without SimpleITK / nibabel: enough of the format for the graph-cut CLIs
(scalar 1-D..4-D volumes, the ten numeric datatypes, either byte order, scl_slope / scl_inter, spacing and origin).

Layout facts used (NIfTI-1 specification, nifti1.h): 348-byte header, `dim[0]` = number of axes, voxel data in Fortran order
(x fastest) starting at `vox_offset`; `pixdim[1..]` = spacing; origin from the qform (`qoffset_*`) or the sform (`srow_*[3]`).
NIfTI coordinates are RAS, ITK / SimpleITK (what medpy.io.load reports, io/load.py:116-127) are LPS: x and y of the origin
change sign; an identity LPS direction is the quaternion (0, 0, 1) with srow = diag(-sx, -sy, sz).
"""
import gzip
import struct

import numpy

_DTYPES = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4", 1024: "i8", 1280: "u8"}
_CODES = {numpy.dtype(v).str[1:]: k for k, v in _DTYPES.items()}


def _open(path, mode):
    return gzip.open(path, mode) if str(path).lower().endswith(".gz") else open(path, mode)


def read(path):
    """-> (array in z,y,x[,t] C order, spacing (x,y,z[,t]), offset in ITK's LPS convention (x,y,z[,t]), 1)."""
    with _open(path, "rb") as fh:
        raw = fh.read()
    if len(raw) < 348:
        raise ValueError("not a NIfTI-1 file (shorter than its 348-byte header)")
    order = "<"
    if struct.unpack("<i", raw[:4])[0] != 348:
        order = ">"
        if struct.unpack(">i", raw[:4])[0] != 348:
            raise ValueError("not a NIfTI-1 file (sizeof_hdr != 348)")
    if raw[344:347] not in (b"n+1", b"ni1"):
        raise ValueError("not a NIfTI-1 file (magic)")
    if raw[344:347] == b"ni1":
        raise ValueError("two-file NIfTI (.hdr/.img) is not supported; use a single .nii file")
    dim = struct.unpack(order + "8h", raw[40:56])
    ndim = dim[0]
    if not 1 <= ndim <= 4 or any(d < 1 for d in dim[1:ndim + 1]):
        raise ValueError("unsupported NIfTI dimensionality %r" % (dim,))
    datatype = struct.unpack(order + "h", raw[70:72])[0]
    if datatype not in _DTYPES:
        raise ValueError("unsupported NIfTI datatype %d" % datatype)
    pixdim = struct.unpack(order + "8f", raw[76:108])
    vox_offset = int(struct.unpack(order + "f", raw[108:112])[0])
    slope, inter = struct.unpack(order + "2f", raw[112:120])
    qform_code, sform_code = struct.unpack(order + "2h", raw[252:256])
    qoff = struct.unpack(order + "3f", raw[268:280])
    srow = [struct.unpack(order + "4f", raw[280 + 16 * i:296 + 16 * i]) for i in range(3)]
    shape_xyz = dim[1:ndim + 1]
    dt = numpy.dtype(order + _DTYPES[datatype])
    count = int(numpy.prod(shape_xyz))
    if len(raw) < max(vox_offset, 352) + count * dt.itemsize:
        raise ValueError("truncated NIfTI file")
    data = numpy.frombuffer(raw, dtype=dt, count=count, offset=max(vox_offset, 352))
    data = data.reshape(tuple(reversed(shape_xyz)))                     # Fortran order on disk == C order over (t,)z,y,x
    data = data.astype(dt.newbyteorder("="), copy=True) if order == ">" else data.copy()
    if slope not in (0.0,) and not (slope == 1.0 and inter == 0.0) and numpy.isfinite(slope):
        data = data.astype(numpy.float64) * float(slope) + float(inter)
    spacing = [abs(float(p)) if p else 1.0 for p in pixdim[1:ndim + 1]]
    if qform_code > 0:
        origin_ras = list(qoff)
    elif sform_code > 0:
        origin_ras = [srow[0][3], srow[1][3], srow[2][3]]
    else:
        origin_ras = [0.0, 0.0, 0.0]
    origin_lps = [-origin_ras[0], -origin_ras[1], origin_ras[2]] + [0.0] * max(0, ndim - 3)
    return data, spacing, origin_lps[:ndim], 1


def write(path, zyx, spacing=None, offset=None):
    """`zyx`: array in (t,)z,y,x C order (x fastest = NIfTI's own order); spacing / offset in x,y,z[,t] (LPS origin)."""
    zyx = numpy.ascontiguousarray(zyx)
    if zyx.dtype == numpy.bool_:
        zyx = zyx.astype(numpy.uint8)
    key = zyx.dtype.newbyteorder("=").str[1:] if zyx.dtype.byteorder != "|" else zyx.dtype.str[1:]
    if key not in _CODES:
        raise ValueError("dtype %s cannot be stored in a NIfTI-1 file" % zyx.dtype)
    zyx = zyx.astype(zyx.dtype.newbyteorder("<"), copy=False)
    ndim = zyx.ndim
    if not 1 <= ndim <= 4:
        raise ValueError("NIfTI writer supports 1 to 4 dimensions")
    shape_xyz = tuple(reversed(zyx.shape))
    spacing = [float(s) for s in spacing] if spacing else [1.0] * ndim
    offset = [float(o) for o in offset] if offset else [0.0] * ndim
    spacing = (spacing + [1.0] * 4)[:max(ndim, 3)]
    offset = (offset + [0.0] * 3)[:3]
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, ndim, *(list(shape_xyz) + [1] * (7 - ndim)))
    struct.pack_into("<h", hdr, 70, _CODES[key])
    struct.pack_into("<h", hdr, 72, zyx.dtype.itemsize * 8)
    struct.pack_into("<8f", hdr, 76, 1.0, *((spacing + [1.0] * 7)[:7]))
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, 1.0, 0.0)
    hdr[123] = 2                                                         # xyzt_units: millimetres
    struct.pack_into("<2h", hdr, 252, 1, 1)                              # qform_code, sform_code = scanner
    struct.pack_into("<3f", hdr, 256, 0.0, 0.0, 1.0)                     # identity LPS direction as a RAS quaternion
    struct.pack_into("<3f", hdr, 268, -offset[0], -offset[1], offset[2])
    struct.pack_into("<4f", hdr, 280, -spacing[0], 0.0, 0.0, -offset[0])
    struct.pack_into("<4f", hdr, 296, 0.0, -spacing[1], 0.0, -offset[1])
    struct.pack_into("<4f", hdr, 312, 0.0, 0.0, spacing[2], offset[2])
    hdr[344:348] = b"n+1\x00"
    with _open(path, "wb") as fh:
        fh.write(bytes(hdr))
        fh.write(b"\x00" * 4)
        fh.write(zyx.tobytes())
