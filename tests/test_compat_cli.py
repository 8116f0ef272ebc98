"""The reference CLI against the import shim (CPU part) and our own CLI on the GPU."""
import os
import subprocess
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "medpy_b200", "compat")
REF_CLI = "/root/reference/bin/medpy_graphcut_voxel.py"


def _write_case(tmp_path, shape=(12, 10, 9), ext=".mha"):
    sys.path.insert(0, COMPAT)
    try:
        from medpy.io import save, Header
    finally:
        sys.path.remove(COMPAT)
    from medpy_b200 import synthetic
    vol = synthetic.two_blob_volume(shape, seed=2, with_prob=False)
    # MedPy's convention: arrays are x,y,z -> transpose our z,y,x volume
    img_xyz = numpy.ascontiguousarray(vol["image"]).T
    markers = (vol["fg"].astype(numpy.uint8) + 2 * vol["bg"].astype(numpy.uint8)).T
    hdr = Header(spacing=(1.0, 1.0, 2.0), offset=(0.0, 0.0, 0.0))
    ip, mp = str(tmp_path / ("img" + ext)), str(tmp_path / ("markers" + ext))
    save(img_xyz, ip, hdr, True)
    save(markers, mp, hdr, True)
    return vol, ip, mp


def test_metaimage_roundtrip_and_xyz_order(tmp_path):
    sys.path.insert(0, COMPAT)
    try:
        import importlib
        import medpy.io as mio
        importlib.reload(mio)
        a = numpy.arange(2 * 3 * 4, dtype=numpy.float32).reshape(2, 3, 4)  # x, y, z
        p = str(tmp_path / "a.mha")
        mio.save(a, p, mio.Header(spacing=(1, 2, 3), offset=(0, 0, 0)), True)
        b, hdr = mio.load(p)
        assert b.shape == (2, 3, 4) and numpy.array_equal(a, b)
        assert not b.flags.c_contiguous and b.flags.f_contiguous      # transposed view, like the reference's load()
        assert mio.header.get_voxel_spacing(hdr) == (1.0, 2.0, 3.0)
        with pytest.warns(DeprecationWarning):
            assert mio.header.get_pixel_spacing(hdr) == (1.0, 2.0, 3.0)
    finally:
        sys.path.remove(COMPAT)
        for k in [k for k in sys.modules if k == "medpy" or k.startswith("medpy.")]:
            del sys.modules[k]


@pytest.mark.skipif(not os.path.exists(REF_CLI), reason="reference tree not present")
def test_reference_cli_runs_unchanged_up_to_the_device(tmp_path):
    """bin/medpy_graphcut_voxel.py, executed in place and unmodified, imports medpy.* from the shim, parses its
    arguments, loads the images, splits the markers and reaches graph_from_voxels; without a GPU that is where the
    product refuses (no CPU fallback); with one it must finish and write the mask."""
    import torch
    vol, ip, mp = _write_case(tmp_path)
    out = str(tmp_path / "out.mha")
    env = dict(os.environ, PYTHONPATH=COMPAT + os.pathsep + ROOT, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, REF_CLI, "15.0", ip, mp, out, "--boundary", "diff_exp", "-f"],
                       env=env, capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr[-2000:]
        assert os.path.exists(out)
    else:
        assert r.returncode != 0
        assert "no CPU path" in r.stderr or "CUDA" in r.stderr, r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("ext", [".mha", ".nii.gz"])
def test_own_cli_matches_oracle(tmp_path, ext):
    from oracle import energy_terms as et, solvers
    vol, ip, mp = _write_case(tmp_path, shape=(20, 16, 18), ext=ext)
    out = str(tmp_path / ("out" + ext))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "medpy_b200", "cli", "graphcut_voxel.py"), "15.0", ip, mp, out,
                        "--boundary", "diff_exp", "-s", "-f"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    sys.path.insert(0, COMPAT)
    try:
        from medpy.io import load
        mask_xyz, _ = load(out)
    finally:
        sys.path.remove(COMPAT)
    # spacing (1,1,2) is in x,y,z order = the CLI's array order; our volume is z,y,x -> reversed
    prob = et.build_problem(vol["fg"], vol["bg"], boundary=("difference_exponential", vol["image"], 15.0, (2.0, 1.0, 1.0)))
    oflow, omask, _ = solvers.solve_port(prob)
    assert numpy.array_equal(mask_xyz.T.astype(numpy.uint8), omask)


REF_LABEL_CLI = "/root/reference/bin/medpy_graphcut_label.py"


def _write_label_case(tmp_path, shape=(12, 10, 9)):
    sys.path.insert(0, COMPAT)
    try:
        from medpy.io import save, Header
    finally:
        sys.path.remove(COMPAT)
    from medpy_b200 import synthetic
    vol = synthetic.two_blob_volume(shape, seed=5, with_prob=False)
    grids = numpy.meshgrid(*[numpy.arange(s) for s in shape], indexing="ij")
    regions = sum((g // 3) * m for g, m in zip(grids, (100, 10, 1))) + 7          # arbitrary ids: the CLI relabels
    grad = numpy.sqrt(sum(d ** 2 for d in numpy.gradient(vol["image"].astype(numpy.float64)))).astype(numpy.float32)
    markers = (vol["fg"].astype(numpy.uint8) + 2 * vol["bg"].astype(numpy.uint8))
    hdr = Header(spacing=(1.0, 1.0, 1.0), offset=(0.0, 0.0, 0.0))
    paths = {}
    for name, arr in (("grad", grad), ("regions", regions.astype(numpy.int32)), ("markers", markers)):
        paths[name] = str(tmp_path / (name + ".mha"))
        save(numpy.ascontiguousarray(arr).T, paths[name], hdr, True)
    return vol, regions, grad, paths


@pytest.mark.skipif(not os.path.exists(REF_LABEL_CLI), reason="reference tree not present")
def test_reference_label_cli_runs_unchanged_up_to_the_device(tmp_path):
    """bin/medpy_graphcut_label.py in place and unmodified: medpy.filter.relabel, medpy.graphcut.graph_from_labels,
    energy_label.boundary_stawiaski, what_segment per region and medpy.filter.relabel_map all come from the shim."""
    import torch
    vol, regions, grad, paths = _write_label_case(tmp_path)
    out = str(tmp_path / "out.mha")
    env = dict(os.environ, PYTHONPATH=COMPAT + os.pathsep + ROOT, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, REF_LABEL_CLI, paths["grad"], paths["regions"], paths["markers"], out, "-f"],
                       env=env, capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr[-2000:]
        assert os.path.exists(out)
    else:
        assert r.returncode != 0
        assert "no CPU" in r.stderr or "CUDA" in r.stderr, r.stderr[-2000:]


@pytest.mark.gpu
def test_own_label_cli_matches_oracle(tmp_path):
    from oracle import energy_label_terms as elt, solvers
    vol, regions, grad, paths = _write_label_case(tmp_path, shape=(18, 15, 12))
    out = str(tmp_path / "out.mha")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "medpy_b200", "cli", "graphcut_label.py"), paths["grad"],
                        paths["regions"], paths["markers"], out, "-f"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    sys.path.insert(0, COMPAT)
    try:
        from medpy.io import load
        from medpy.filter import relabel
        mask_xyz, _ = load(out)
        lab_xyz = relabel(numpy.ascontiguousarray(regions.astype(numpy.int32)).T)     # what the CLI cut: x,y,z order
    finally:
        sys.path.remove(COMPAT)
    g_xyz, fg_xyz, bg_xyz = grad.T, vol["fg"].T, vol["bg"].T
    i, j, w, wr = elt.stawiaski_calls(lab_xyz, g_xyz)
    fgr, bgr = elt.marker_regions(lab_xyz, fg_xyz), elt.marker_regions(lab_xyz, bg_xyz)
    tw = [(fgr, numpy.full(fgr.size, 65535.0), numpy.zeros(fgr.size)), (bgr, numpy.zeros(bgr.size), numpy.full(bgr.size, 65535.0))]
    _, rmask, _ = solvers.solve_sparse(int(lab_xyz.max()), i, j, w, wr, tw)
    assert numpy.array_equal(numpy.asarray(mask_xyz).astype(numpy.uint8), rmask[lab_xyz - 1])


OTHER_LABEL_CLIS = [
    ("medpy_graphcut_label_w_regional.py", lambda p, out: [p["grad"], p["regions"], p["markers"], out, "--regional", "atlas",
                                                            "--radditional", p["atlas"], "--alpha", "0.2", "-f"]),
    ("medpy_graphcut_label_wsplit.py", lambda p, out: [p["grad"], p["regions"], p["markers"], out, "-f"]),
]
# (bin/medpy_graphcut_label_bgreduced.py cannot run on numpy >= 1.13 at all: it subtracts boolean arrays in its own
#  pre-processing, :217, long before it reaches the graph cut.)


@pytest.mark.skipif(not os.path.exists(REF_LABEL_CLI), reason="reference tree not present")
@pytest.mark.parametrize("script,argv", OTHER_LABEL_CLIS, ids=[c[0] for c in OTHER_LABEL_CLIS])
def test_other_reference_label_clis_run_unchanged_up_to_the_device(tmp_path, script, argv):
    """The remaining region-cut scripts of the reference, in place and unmodified: regional_atlas + boundary term
    (w_regional), graphcut_split(graphcut_stawiaski, ...) (wsplit).
    Everything they import resolves in the shim; without a GPU they stop where the product refuses to run on a CPU."""
    import torch
    sys.path.insert(0, COMPAT)
    try:
        from medpy.io import save, Header
    finally:
        sys.path.remove(COMPAT)
    vol, regions, grad, paths = _write_label_case(tmp_path, shape=(24, 22, 20))
    atlas = (1.0 / (1.0 + numpy.exp(-(vol["image"].astype(numpy.float64) - 50.0) / 15.0))).astype(numpy.float32)
    paths["atlas"] = str(tmp_path / "atlas.mha")
    save(numpy.ascontiguousarray(atlas).T, paths["atlas"], Header(spacing=(1.0, 1.0, 1.0), offset=(0.0, 0.0, 0.0)), True)
    out = str(tmp_path / "out.mha")
    env = dict(os.environ, PYTHONPATH=COMPAT + os.pathsep + ROOT, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(REF_LABEL_CLI), script)] + argv(paths, out),
                       env=env, capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr[-2000:]
        assert os.path.exists(out)
    else:
        assert r.returncode != 0
        assert "no CPU" in r.stderr or "CUDA" in r.stderr, r.stderr[-3000:]
