"""Header object and accessor functions (reference surface: medpy/io/header.py:33-166, 168-387)."""
import warnings


class Header:
    def __init__(self, spacing=None, offset=None, direction=None, meta=None):
        self.spacing = tuple(float(s) for s in spacing) if spacing is not None else None
        self.offset = tuple(float(o) for o in offset) if offset is not None else None
        self.direction = direction
        self.meta = dict(meta or {})

    def get_voxel_spacing(self):
        return self.spacing

    def get_offset(self):
        return self.offset

    def get_direction(self):
        return self.direction

    def set_voxel_spacing(self, spacing):
        self.spacing = tuple(float(s) for s in spacing)

    def set_offset(self, offset):
        self.offset = tuple(float(o) for o in offset)


def get_voxel_spacing(hdr):
    return hdr.get_voxel_spacing()


def get_pixel_spacing(hdr):
    warnings.warn("get_pixel_spacing() is depreciated, use get_voxel_spacing() instead", category=DeprecationWarning)
    return get_voxel_spacing(hdr)


def get_offset(hdr):
    return hdr.get_offset()


def set_voxel_spacing(hdr, spacing):
    hdr.set_voxel_spacing(spacing)


def set_pixel_spacing(hdr, spacing):
    set_voxel_spacing(hdr, spacing)


def set_offset(hdr, offset):
    hdr.set_offset(offset)


def copy_meta_data(hdr_to, hdr_from):
    hdr_to.spacing, hdr_to.offset, hdr_to.direction = hdr_from.spacing, hdr_from.offset, hdr_from.direction
    return hdr_to
