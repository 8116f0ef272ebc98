#!/usr/bin/env python
"""One-off full-size parity runs: GPU mask/energy against the BK oracle port (bit-identical to the reference solver) on
the host.  usage: compare_fullsize.py c2 | c3 [size]"""
import json, os, sys, time
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medpy_b200 import synthetic
import medpy_b200.graphcut as gc
from oracle import energy_terms as et, solvers

which = sys.argv[1]
size = int(sys.argv[2]) if len(sys.argv) > 2 else (256 if which == "c2" else 512)
shape = (size,) * 3
vol = synthetic.two_blob_volume(shape, seed=0, with_prob=(which == "c3"))
kw = dict(boundary_term=gc.energy_voxel.boundary_difference_exponential, boundary_term_args=(vol["image"], vol["sigma"], False))
reg = None
if which == "c3":
    kw.update(regional_term=gc.energy_voxel.regional_probability_map, regional_term_args=(vol["prob"], vol["alpha"]))
    reg = (vol["prob"], vol["alpha"])
g = gc.graph_from_voxels(vol["fg"], vol["bg"], **kw)
e = g.maxflow(); m = g.get_mask().copy(); st = g.stats()
del g
t0 = time.time()
prob = et.build_problem(vol["fg"], vol["bg"], regional=reg, boundary=("difference_exponential", vol["image"], vol["sigma"], False))
t1 = time.time()
oflow, omask, tm = solvers.solve_port(prob)
diff = (m != omask)
extra = {}
if diff.any():
    sh = prob["shape"]
    extra["cut_energy_gpu_mask"] = et.cut_energy(sh, prob["wf"], prob["wb"], prob["tr"], prob["flow_const"], m)
    extra["cut_energy_bk_mask"] = et.cut_energy(sh, prob["wf"], prob["wb"], prob["tr"], prob["flow_const"], omask)
    det = []
    for idx in numpy.argwhere(diff)[:4]:
        idx = tuple(int(i) for i in idx)
        flat = int(numpy.ravel_multi_index(idx, sh))
        d = dict(idx=list(idx), gpu=int(m[idx]), bk=int(omask[idx]), tr=float(prob["tr"][flat]).hex(), image=float(vol["image"][idx]))
        ws, nb = [], []
        stride = int(numpy.prod(sh))
        for ax in range(len(sh)):
            stride //= sh[ax]
            for sg in (-1, 1):
                j = list(idx); j[ax] += sg
                if 0 <= j[ax] < sh[ax]:
                    w = prob["wf"][ax][flat if sg > 0 else flat - stride]
                    ws.append(float(w).hex()); nb.append([int(m[tuple(j)]), int(omask[tuple(j)])])
        d["weights"] = ws; d["nbr_masks_gpu_bk"] = nb
        # exact margin of moving this voxel: cost(in T) - cost(in S) with neighbours as in the BK mask
        import fractions
        F = fractions.Fraction
        trv = F(float(prob["tr"][flat]))
        cT = (trv if trv > 0 else F(0)); cS = (-trv if trv < 0 else F(0))
        k = 0
        stride = int(numpy.prod(sh))
        for ax in range(len(sh)):
            stride //= sh[ax]
            for sg in (-1, 1):
                j = list(idx); j[ax] += sg
                if 0 <= j[ax] < sh[ax]:
                    w = F(float(prob["wf"][ax][flat if sg > 0 else flat - stride]))
                    if omask[tuple(j)]: cT += w     # neighbour in S, voxel in T: arc nbr->v cut
                    else: cS += w                   # neighbour in T, voxel in S: arc v->nbr cut
        d["exact_margin_T_minus_S"] = float(cT - cS)
        det.append(d)
    extra["detail"] = det
print(json.dumps(dict(extra=extra, config=which, shape=list(shape), gpu_energy=e, bk_energy=oflow, rel_err=abs(e - oflow) / abs(oflow),
                      mask_hamming=int(diff.sum()), gpu_fg=int(m.sum()), bk_fg=int(omask.sum()), bk_terms_s=t1 - t0,
                      bk_setup_s=tm["setup_s"], bk_maxflow_s=tm["maxflow_s"], gpu_solve_ms=st["ms_solve"],
                      gpu_terms_ms=st["ms_terms"])), flush=True)
