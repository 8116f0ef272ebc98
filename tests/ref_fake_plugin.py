"""pytest plugin (TEST INFRASTRUCTURE): installs the oracle-backed doubles of tests/fake_native.py in place of the native
classes (dense lattice, label image, sparse graph), so that the reference's OWN test files can be run against the import shim on a machine without a GPU
(tests/test_reference_suite.py).  Never loaded by the product or by the `-m gpu` tests."""


def pytest_configure(config):
    import fake_native
    from medpy_b200 import _lib
    _lib._mgc.LabelImage = fake_native.FakeLabelImage
    _lib._mgc.SparseGraph = fake_native.FakeSparseGraph
    _lib._mgc.Graph = fake_native.FakeGraph
    _lib.Graph = fake_native.FakeGraph
