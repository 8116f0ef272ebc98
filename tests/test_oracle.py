"""CPU tests: the oracle (oracle/energy_terms.py + oracle/bk_lattice.c) against the golden vectors
produced by the unmodified reference, and against the real reference solver when oracle/_ref exists."""
import os

import numpy
import pytest

from conftest import golden
from oracle import energy_terms as et
from oracle import solvers

NAMES = golden().names()


def _problem(c):
    boundary = None
    if c["boundary"]:
        boundary = (c["boundary"], c["image"], c["sigma"], tuple(c["spacing"]) if c["spacing"] else False)
    regional = (c["prob"], c["alpha"]) if c["prob"] is not None else None
    return et.build_problem(c["fg"], c["bg"], regional=regional, boundary=boundary)


@pytest.mark.parametrize("name", NAMES)
def test_terms_bit_exact_vs_reference(name):
    """n-link weights and net t-links equal the reference's get_edge / get_trcap bit for bit."""
    c = golden().case(name)
    with numpy.errstate(all="ignore"):
        prob = _problem(c)
    nd = len(prob["shape"])
    for d in range(nd):
        got = prob["wf"][d] if prob["wf"] is not None else numpy.zeros(prob["tr"].size)
        assert numpy.array_equal(got, c["w"][d], equal_nan=True), "axis %d weights differ" % d
    assert numpy.array_equal(prob["tr"], c["tr"])


@pytest.mark.parametrize("name", NAMES)
def test_port_flow_and_mask_vs_reference(name):
    """BK restatement: flow bit-identical, mask identical to the reference run."""
    c = golden().case(name)
    with numpy.errstate(all="ignore"):
        prob = _problem(c)
    flow, mask, _ = solvers.solve_port(prob)
    assert numpy.array_equal(mask, c["mask"].reshape(mask.shape))
    if numpy.isnan(c["flow"]):
        assert numpy.isnan(flow)
    else:
        assert flow == c["flow"], (flow.hex(), c["flow"].hex())
        if not numpy.isnan(prob["wf"][0]).any() if prob["wf"] is not None else True:
            e = et.cut_energy(prob["shape"], prob["wf"], prob["wb"], prob["tr"], prob["flow_const"], mask)
            assert abs(e - flow) <= 1e-12 * max(1.0, abs(flow))


def test_reference_known_answers():
    """The reference's own KATs survive the trip: cut.py maxflow == 3, fixtures' masks."""
    c = golden().case("ref_fixture_cut3d")
    assert c["flow"] == 3.0
    expect = numpy.asarray([[[1, 1, 1, 0, 0]] * 3] * 2, dtype=numpy.uint8)
    assert numpy.array_equal(c["mask"], expect)
    flow, mask, _ = solvers.solve_port(_problem(c))
    assert flow == 3.0 and numpy.array_equal(mask, expect)


@pytest.mark.skipif(not solvers.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_equals_real_reference_on_random_lattices():
    rng = numpy.random.default_rng(7)
    for _ in range(120):
        nd = int(rng.integers(1, 5))
        shape = tuple(int(x) for x in rng.integers(1, 7, size=nd))
        n = int(numpy.prod(shape))
        wf = [rng.integers(0, 3, size=n).astype(float) for _ in range(nd)]
        wb = [rng.integers(0, 3, size=n).astype(float) for _ in range(nd)]
        tr = rng.integers(-3, 4, size=n).astype(float)
        prob = dict(shape=shape, wf=wf, wb=wb, tr=tr, flow_const=0.0, fg=numpy.zeros(n, numpy.uint8),
                    bg=numpy.zeros(n, numpy.uint8), src=numpy.maximum(tr, 0), snk=numpy.maximum(-tr, 0))
        f1, m1, _ = solvers.solve_port(prob)
        f2, m2, _ = solvers.solve_ref(prob)
        assert f1 == f2 and numpy.array_equal(m1, m2)


@pytest.mark.skipif(not solvers.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_equals_real_reference_float_32cubed():
    from medpy_b200 import synthetic
    vol = synthetic.two_blob_volume((32, 32, 32), seed=0)
    prob = et.build_problem(vol["fg"], vol["bg"], regional=(vol["prob"], 0.1),
                            boundary=("difference_exponential", vol["image"], vol["sigma"], False))
    f1, m1, _ = solvers.solve_port(prob)
    f2, m2, _ = solvers.solve_ref(prob)
    assert f1 == f2 and numpy.array_equal(m1, m2)
    assert 0 < m1.sum() < m1.size


@pytest.mark.skipif(not (os.path.isdir("/root/reference/medpy") and solvers.have_ref()), reason="reference tree / pyshim not present")
def test_voxel_oracle_fuzzed_against_the_live_reference():
    """120 random lattices (1-D..4-D, the eight boundary terms, five image dtypes, random sigma / spacing / regional term /
    overlapping markers) through the reference's own graph_from_voxels: weights, t-links, flow and mask equal the
    restatement's bit for bit; weight <= 0 ValueErrors occur in the same cases
    (tests/golden/fuzz_voxels_against_reference.py)."""
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fuzz_voxels_against_reference.py")
    r = subprocess.run([sys.executable, script, "120", "11"], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert r.returncode == 0 and "ok 120" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
