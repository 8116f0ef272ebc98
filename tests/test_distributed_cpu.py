"""CPU tests of the multi-GPU host logic: two gloo ranks run medpy_b200.distributed over a numpy test double of the
device slab handle (tests/fake_slab.py) and must reproduce the single-process oracle: identical mask, energy
within 1e-9, for a volume whose object straddles the slab border (so flow and labels really cross ranks)."""
import os
import socket
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fake_slab import FakeSlab
    from medpy_b200 import distributed as md
    vol = _volume(case)
    kw = dict(image=vol["image"], boundary=vol["boundary"], sigma=vol["sigma"], handle_factory=FakeSlab)
    if case == "regional":
        kw.update(prob=vol["prob"], alpha=0.1)
    energy, mask = md.graphcut_slab(vol["fg"], vol["bg"], **kw)
    if rank == 0:
        numpy.savez(out, energy=energy, mask=mask)
    dist.destroy_process_group()


def _volume(case):
    rng = numpy.random.default_rng(3)
    shape = (10, 7, 6)
    z, y, x = numpy.meshgrid(*[numpy.arange(s) for s in shape], indexing="ij")
    inside = ((z - 4.5) ** 2 / 9 + (y - 3) ** 2 / 4 + (x - 2.5) ** 2 / 4) <= 1.0   # ellipsoid across the z=5 border
    img = (100.0 * inside + rng.normal(0, 10, shape)).astype(numpy.float32)
    fg = numpy.zeros(shape, bool); fg[4:6, 3, 2:4] = True
    bg = numpy.zeros(shape, bool); bg[0] = bg[-1] = True; bg[:, 0] = bg[:, -1] = True; bg[:, :, 0] = bg[:, :, -1] = True
    prob = (1 / (1 + numpy.exp(-(img - 50.0) / 15.0))).astype(numpy.float32)
    return dict(image=img, fg=fg, bg=bg, prob=prob, sigma=14.0, boundary="difference_exponential", shape=shape)


def test_slab_bounds_cover_the_axis():
    from medpy_b200.distributed import slab_bounds
    for extent in (8, 10, 513):
        for world in (1, 2, 3, 4, 8):
            spans = [slab_bounds(extent, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == extent
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(b > a for a, b in spans)


@pytest.mark.parametrize("case,world", [("boundary", 2), ("regional", 2), ("regional", 3)])
def test_two_rank_gloo_matches_oracle(tmp_path, case, world):
    import torch.multiprocessing as mp
    from oracle import energy_terms as et, solvers
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(world, _free_port(), case, out), nprocs=world, join=True)
    got = numpy.load(out)
    vol = _volume(case)
    prob = et.build_problem(vol["fg"], vol["bg"], regional=(vol["prob"], 0.1) if case == "regional" else None,
                            boundary=(vol["boundary"], vol["image"], vol["sigma"], False))
    oflow, omask, _ = solvers.solve_port(prob)
    assert numpy.array_equal(got["mask"], omask)
    assert abs(float(got["energy"]) - oflow) <= 1e-9 * abs(oflow)
    assert 0 < omask.sum() < omask.size
