"""Reader / writer for uncompressed ITK MetaImage files (text header + raw little-endian voxels, x fastest)."""
import os

import numpy

_TYPES = {
    "MET_UCHAR": numpy.uint8, "MET_CHAR": numpy.int8, "MET_USHORT": numpy.uint16, "MET_SHORT": numpy.int16,
    "MET_UINT": numpy.uint32, "MET_INT": numpy.int32, "MET_FLOAT": numpy.float32, "MET_DOUBLE": numpy.float64,
}
_NAMES = {numpy.dtype(v).name: k for k, v in _TYPES.items()}


def read(path):
    with open(path, "rb") as f:
        fields = {}
        while True:
            line = f.readline()
            if not line:
                raise ValueError("MetaImage header of {} has no ElementDataFile".format(path))
            key, _, val = line.decode("ascii", "replace").partition("=")
            key, val = key.strip(), val.strip()
            fields[key] = val
            if key == "ElementDataFile":
                break
        if fields.get("CompressedData", "False").lower() == "true":
            raise ValueError("compressed MetaImage files are not supported by this reader")
        dims = [int(v) for v in fields["DimSize"].split()]
        dtype = numpy.dtype(_TYPES[fields["ElementType"]])
        if fields.get("BinaryDataByteOrderMSB", fields.get("ElementByteOrderMSB", "False")).lower() == "true":
            dtype = dtype.newbyteorder(">")
        nchan = int(fields.get("ElementNumberOfChannels", "1"))
        count = int(numpy.prod(dims)) * nchan
        if fields["ElementDataFile"] == "LOCAL":
            data = numpy.fromfile(f, dtype=dtype, count=count)
        else:
            data = numpy.fromfile(os.path.join(os.path.dirname(path), fields["ElementDataFile"]), dtype=dtype, count=count)
    if data.size != count:
        raise ValueError("MetaImage {} is truncated".format(path))
    shape = list(reversed(dims)) + ([nchan] if nchan > 1 else [])   # z, y, x[, c]
    spacing = [float(v) for v in fields.get("ElementSpacing", " ".join(["1"] * len(dims))).split()]
    offset = [float(v) for v in fields.get("Offset", fields.get("Position", " ".join(["0"] * len(dims)))).split()]
    return data.astype(dtype.newbyteorder("="), copy=False).reshape(shape), spacing, offset, nchan


def write(path, zyx, spacing, offset):
    dims = list(reversed(zyx.shape))
    name = _NAMES.get(zyx.dtype.name)
    if name is None:
        raise ValueError("dtype {} cannot be stored in a MetaImage".format(zyx.dtype))
    spacing = list(spacing) if spacing else [1.0] * len(dims)
    offset = list(offset) if offset else [0.0] * len(dims)
    raw = os.path.splitext(path)[0] + ".raw"
    local = not path.lower().endswith(".mhd")
    hdr = ["ObjectType = Image", "NDims = {}".format(len(dims)), "BinaryData = True", "BinaryDataByteOrderMSB = False",
           "CompressedData = False", "Offset = " + " ".join(repr(float(o)) for o in offset[: len(dims)]),
           "ElementSpacing = " + " ".join(repr(float(s)) for s in spacing[: len(dims)]),
           "DimSize = " + " ".join(str(d) for d in dims), "ElementType = " + name,
           "ElementDataFile = " + ("LOCAL" if local else os.path.basename(raw))]
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode("ascii"))
        if local:
            numpy.ascontiguousarray(zyx).tofile(f)
    if not local:
        numpy.ascontiguousarray(zyx).tofile(raw)
