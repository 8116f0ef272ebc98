"""Exception types shared by the host layer and the ``medpy`` import shim (reference: medpy/core/exceptions.py)."""


class ArgumentError(Exception):
    """Raised for invalid arguments (medpy/core/exceptions.py:31)."""
