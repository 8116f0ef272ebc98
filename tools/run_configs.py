#!/usr/bin/env python
"""Run the BASELINE.json configurations other than the bench headline once each on one GPU and print one JSON line per
config (device time to converged min cut, solver statistics, duality gap).  Usage: python tools/run_configs.py [2 4 5s ...]"""
import json
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(name, shape, kind, regional, fourd=False, e2e_api=True):
    import torch
    from medpy_b200 import synthetic
    import medpy_b200.graphcut as gc
    t0 = time.time()
    if fourd:
        vol = synthetic.multispectral_volume(shape, seed=0)
    else:
        vol = synthetic.two_blob_volume(shape, seed=0, with_prob=regional)
    gen = time.time() - t0
    term = getattr(gc.energy_voxel, "boundary_" + kind)
    kw = dict(boundary_term=term, boundary_term_args=(vol["image"], vol["sigma"], False))
    if regional:
        kw.update(regional_term=gc.energy_voxel.regional_probability_map, regional_term_args=(vol["prob"], vol["alpha"]))
    out = None
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = gc.graph_from_voxels(vol["fg"], vol["bg"], **kw)
        e = g.maxflow()
        m = g.get_mask()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = g.stats()
        n = int(numpy.prod(shape))
        out = dict(config=name, shape=list(shape), boundary=kind, regional=regional, n=n, energy=e, fg_voxels=int(m.sum()),
                   e2e_ms=1e3 * dt, e2e_mvox_s=n / dt / 1e6,
                   device_ms=st["ms_terms"] + st["ms_solve"] + st["ms_readout"],
                   device_mvox_s=n / ((st["ms_terms"] + st["ms_solve"] + st["ms_readout"]) * 1e-3) / 1e6,
                   stats={k: st[k] for k in ("push_sweeps", "global_relabels", "relabel_sweeps", "kernel_launches", "ms_terms",
                                             "ms_solve", "ms_readout", "ms_push", "ms_relabel", "device_bytes")},
                   gen_s=gen, sigma=vol["sigma"])
        del g
    print(json.dumps(out), flush=True)


CONFIGS = {
    "1": ("config1 64^3 difference_linear", (64, 64, 64), "difference_linear", False, False),
    "2": ("config2 256^3 difference_exponential boundary-only", (256, 256, 256), "difference_exponential", False, False),
    "3s": ("config3 256^3 regional + difference_exponential", (256, 256, 256), "difference_exponential", True, False),
    "4": ("config4 256x256x128x4 maximum_exponential (4-D lattice)", (256, 256, 128, 4), "maximum_exponential", False, True),
    "4s": ("config4 small 64x64x32x4 maximum_exponential (4-D lattice)", (64, 64, 32, 4), "maximum_exponential", False, True),
    "5s": ("config5 oracle: 1024^3 difference_exponential on ONE GPU", (1024, 1024, 1024), "difference_exponential", False, False),
    "2h": ("config2 at 512^3", (512, 512, 512), "difference_exponential", False, False),
}

if __name__ == "__main__":
    for key in (sys.argv[1:] or ["1", "2", "4s"]):
        name, shape, kind, regional, fourd = CONFIGS[key]
        if kind.endswith("linear"):
            import medpy_b200.graphcut as gc
            # the linear terms take (image, spacing)
            from medpy_b200 import synthetic
            import torch
            vol = synthetic.two_blob_volume(shape, seed=0, with_prob=False)
            t0 = time.perf_counter()
            g = gc.graph_from_voxels(vol["fg"], vol["bg"], boundary_term=gc.energy_voxel.boundary_difference_linear,
                                     boundary_term_args=(vol["image"], False))
            e = g.maxflow(); m = g.get_mask(); torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            st = g.stats()
            print(json.dumps(dict(config=name, energy=e, fg_voxels=int(m.sum()), e2e_ms=1e3 * dt, stats={k: st[k] for k in ("push_sweeps", "global_relabels", "relabel_sweeps", "ms_solve")})), flush=True)
            continue
        run(name, shape, kind, regional, fourd)
