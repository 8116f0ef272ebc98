"""Minimal stand-ins for the two names the voxel CLI takes from ``medpy.core``
(reference: medpy/core/logger.py:35-148, medpy/core/exceptions.py:31)."""
import logging
import sys


from medpy_b200.errors import ArgumentError  # noqa: F401,E402  (one class for the host layer and the shim)


class ImageLoadingError(Exception):
    """Raised when an image cannot be read."""


class ImageSavingError(Exception):
    """Raised when an image cannot be written."""


class Logger(logging.Logger):
    """Process-wide logger writing to stdout, WARNING by default; ``Logger.getInstance()`` returns the singleton."""

    _instance = None

    def __init__(self, name="MedPyLogger", level=logging.WARNING):
        super().__init__(name, level)
        handler = logging.StreamHandler(sys.stdout)
        handler.setFormatter(logging.Formatter("%(asctime)s %(levelname)s: %(message)s"))
        self.addHandler(handler)

    @classmethod
    def getInstance(cls):
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance
