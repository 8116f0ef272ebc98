"""Solver knobs on the hard (boundary-only) instances: device ms to converged min cut per variant.
Usage: python tools/tune_hard.py [2|4|5s ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
import torch
from medpy_b200 import synthetic
from medpy_b200.graphcut.device import graph_from_device_arrays

CONFIGS = {"3": ((512, 512, 512), "difference_exponential", False),"2": ((256, 256, 256), "difference_exponential", False), "4": ((256, 256, 128, 4), "maximum_exponential", True),
           "2h": ((512, 512, 512), "difference_exponential", False), "4s": ((64, 64, 32, 4), "maximum_exponential", True)}
VARIANTS = [
    {},
    {"MEDPY_GC_ITERS": 16}, {"MEDPY_GC_ITERS": 32},
    {"MEDPY_GC_SWEEP_ROUNDS": 2}, {"MEDPY_GC_SWEEP_ROUNDS": 1},
    {"MEDPY_GC_SWEEP_DONE_FRAC": 16}, {"MEDPY_GC_SWEEP_DONE_FRAC": 256},
    {"MEDPY_GC_PASSES0": 4}, {"MEDPY_GC_PASSES0": 8, "MEDPY_GC_ITERS": 16},
    {"MEDPY_GC_SWEEP": 0},
]
for key in (sys.argv[1:] or ["2"]):
    shape, boundary, fourd = CONFIGS[key]
    vol = synthetic.multispectral_volume(shape, seed=0) if fourd else synthetic.two_blob_volume(shape, seed=0, with_prob=(key == "3"))
    d_prob = torch.from_numpy(vol["prob"]).cuda() if key == "3" else None
    variants = VARIANTS if key != "3" else [{}, {"MEDPY_GC_PASSES0": 2}, {"MEDPY_GC_PASSES0": 3}, {"MEDPY_GC_PASSES0": 2, "MEDPY_GC_ITERS": 8}, {"MEDPY_GC_ITERS": 8}, {"MEDPY_GC_ITERS": 6}, {"MEDPY_GC_PASSES0": 4}]
    d_img = torch.from_numpy(vol["image"]).cuda()
    d_fg = torch.from_numpy(vol["fg"].view(numpy.uint8)).cuda()
    d_bg = torch.from_numpy(vol["bg"].view(numpy.uint8)).cuda()
    ref = None
    for var in variants:
        for k in ("MEDPY_GC_ITERS", "MEDPY_GC_SWEEP_ROUNDS", "MEDPY_GC_SWEEP_DONE_FRAC", "MEDPY_GC_PASSES0", "MEDPY_GC_SWEEP", "MEDPY_GC_SWEEP_FRAC"):
            os.environ.pop(k, None)
        for k, v in var.items():
            os.environ[k] = str(v)
        best = None
        for rep in range(3):
            g = graph_from_device_arrays(d_fg, d_bg, image=d_img, boundary=boundary, sigma=vol["sigma"], prob=d_prob, alpha=vol.get("alpha"))
            e = g.maxflow()
            st = g.stats()
            if best is None or st["ms_solve"] < best["ms_solve"]:
                best = st
            del g
        if ref is None:
            ref = e
        print(json.dumps({"config": key, "variant": var, "ms_solve": round(best["ms_solve"], 3), "ms_relabel": round(best["ms_relabel"], 3),
                          "ms_push": round(best["ms_push"], 3), "relabels": best["global_relabels"], "relabel_sweeps": best["relabel_sweeps"],
                          "push_sweeps": best["push_sweeps"], "energy_same": bool(e == ref)}), flush=True)
