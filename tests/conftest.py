"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path."""
import json
import os
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


def _cuda_device_present():
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a CUDA device skips the gpu-marked tests instead of failing them
    (the product has no CPU fallback, so they cannot run there); `-m gpu` on the B200 box runs all of them."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: gpu tests run on the B200 box (-m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """tests/golden/golden_v1.npz -- outputs of the unmodified reference (tests/golden/make_golden.py)."""

    def __init__(self):
        self.z = numpy.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
        self.meta = json.loads(bytes(self.z["__meta__"]).decode())

    def names(self):
        return [m["name"] for m in self.meta]

    def case(self, name):
        m = next(m for m in self.meta if m["name"] == name)
        c = dict(m)
        for k in ("fg", "bg", "image", "prob", "w", "tr", "mask"):
            key = name + "/" + k
            c[k] = self.z[key] if key in self.z.files else None
        if m.get("forder"):
            for k in ("fg", "bg", "image"):
                c[k] = numpy.asfortranarray(c[k])
        c["flow"] = float.fromhex(m["flow_hex"])
        return c


_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = Golden()
    return _golden


@pytest.fixture(scope="session")
def golden_set():
    return golden()
