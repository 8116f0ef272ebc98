#!/usr/bin/env python
"""Timing of the region-graph path on one GPU: graph_from_labels with boundary_stawiaski on a
jittered supervoxel label image -> maxflow -> voxel mask, per stage, next to the CPU restatement of the reference
(oracle numpy terms + the real BK when oracle/_ref is built) on a smaller sample.  Prints one JSON object per size.

    python tools/bench_labels.py [--sizes 128,256] [--cell 4] [--cpu-size 64]

The reference's own term walks the border voxel pairs in a Python loop (energy_label.py:203-214, ~1 us per pair); the
oracle restatement timed here is vectorised numpy and therefore flatters the CPU side.
"""
import argparse
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def supervoxels(n, cell, seed):
    rng = numpy.random.default_rng(seed)
    shape = (n, n, n)
    lab = numpy.zeros(shape, numpy.int64)
    nb = -(-n // cell)
    for axis in range(3):
        idx = numpy.arange(n).reshape([-1 if a == axis else 1 for a in range(3)])
        jit = numpy.clip(idx + rng.integers(-1, 2, size=shape) * (rng.random(shape) < 0.15), 0, n - 1) // cell
        lab = lab * nb + jit
    _, inv = numpy.unique(lab, return_inverse=True)
    return (inv + 1).reshape(shape).astype(numpy.int32)


def volume(n, cell, seed):
    from medpy_b200 import synthetic
    vol = synthetic.two_blob_volume((n, n, n), seed, with_prob=False)
    img = vol["image"]
    grad = numpy.zeros_like(img)
    for d in range(3):
        g = numpy.gradient(img, axis=d)
        grad += g * g
    numpy.sqrt(grad, out=grad)
    return supervoxels(n, cell, seed), grad.astype(numpy.float32), vol["fg"], vol["bg"]


def gpu_run(n, cell, reps=3):
    import medpy_b200.graphcut as gc
    lab, grad, fg, bg = volume(n, cell, 1)
    t = {}
    best = None
    for rep in range(reps):
        t0 = time.perf_counter()
        g = gc.graph_from_labels(lab, fg, bg, boundary_term=gc.energy_label.boundary_stawiaski, boundary_term_args=grad)
        t1 = time.perf_counter()
        flow = g.maxflow()
        t2 = time.perf_counter()
        mask = gc.label_cut_mask(g)
        t3 = time.perf_counter()
        cur = dict(graph_s=t1 - t0, maxflow_s=t2 - t1, mask_s=t3 - t2, total_s=t3 - t0)
        if best is None or cur["total_s"] < best["total_s"]:
            best = cur
    st = g.stats()
    return dict(size=n, cell=cell, voxels=int(lab.size), regions=int(lab.max()), region_pairs=int(g.get_arc_num() // 2),
                energy=flow, fg_voxels=int(mask.sum()), mvox_per_s=lab.size / best["total_s"] / 1e6,
                solver=dict(global_relabels=st["global_relabels"], push_sweeps=st["push_sweeps"], kernel_launches=st["kernel_launches"],
                            ms_solve=st["ms_solve"]), **best)


def cpu_run(n, cell):
    from oracle import energy_label_terms as elt, solvers
    lab, grad, fg, bg = volume(n, cell, 1)
    t0 = time.perf_counter()
    i, j, w, wr = elt.stawiaski_calls(lab, grad)
    fgr, bgr = elt.marker_regions(lab, fg), elt.marker_regions(lab, bg)
    t1 = time.perf_counter()
    out = dict(size=n, cell=cell, voxels=int(lab.size), regions=int(lab.max()), border_pairs=int(i.size), terms_numpy_s=t1 - t0,
               reference_python_loop_s_estimate=1e-6 * i.size)
    tw = [(fgr, numpy.full(fgr.size, 65535.0), numpy.zeros(fgr.size)), (bgr, numpy.zeros(bgr.size), numpy.full(bgr.size, 65535.0))]
    t2 = time.perf_counter()
    flow, mask, secs = solvers.solve_sparse(int(lab.max()), i, j, w, wr, tw)
    out.update(bk_fill_and_maxflow_s=time.perf_counter() - t2, bk_maxflow_s=secs, energy=flow)
    out["mvox_per_s"] = lab.size / (out["terms_numpy_s"] + out.get("bk_fill_and_maxflow_s", 0.0)) / 1e6
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="128,256")
    ap.add_argument("--cell", type=int, default=4)
    ap.add_argument("--cpu-size", type=int, default=64, help="0 skips the CPU arm")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    if args.cpu_size > 0:
        print(json.dumps(dict(arm="cpu_oracle", **cpu_run(args.cpu_size, args.cell))), flush=True)
    for n in [int(s) for s in args.sizes.split(",") if s]:
        print(json.dumps(dict(arm="gpu", **gpu_run(n, args.cell, args.reps))), flush=True)


if __name__ == "__main__":
    main()
